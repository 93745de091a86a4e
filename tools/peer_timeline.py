"""Summarise the in-kernel timeline of the peer-memory training step.

The kernels of csrc/peer.inl stamp %globaltimer into a device log when AR_PEER_LOG=1; with
AR_PEER_LOG_DUMP=<prefix> the last ar_train_steps_peer call of every rank also writes its raw stamps to
<prefix>.rank<r>.bin ([steps][8] uint64: 0 fwd start, 1 fwd past its wait, 2 fwd last CTA done, 3 pull start,
4 pull past its wait, 5 pull done).  This tool prints per-rank stage statistics and, after aligning the
ranks' clocks on the barrier exits, the skew between the ranks at the first barrier.

    AR_PEER_LOG=1 AR_PEER_LOG_DUMP=gpurun_out/plog python -m torch.distributed.run --nproc-per-node 4 \
        --master-addr 127.0.0.1 bench.py --gpus 4
    python tools/peer_timeline.py gpurun_out/plog 4
"""
from __future__ import annotations

import sys

import numpy as np

STAGES = ["fwd wait", "fwd body", "gap to pull", "pull wait", "pull", "head+update+gaps", "step"]


def load(prefix, world):
    logs = [np.fromfile("%s.rank%d.bin" % (prefix, r), dtype=np.uint64).reshape(-1, 8)[:, :6].astype(np.int64)
            for r in range(world)]
    n = min(len(x) for x in logs)
    return np.stack([x[:n] for x in logs], 0)          # [rank][step][stamp]


def main(argv):
    if len(argv) != 3:
        print(__doc__)
        return 2
    prefix, world = argv[1], int(argv[2])
    S = load(prefix, world)[:, 2:]                       # drop the call's first steps (exposed catch-up)
    cur, nxt = S[:, :-1], S[:, 1:]
    d = np.stack([cur[..., 1] - cur[..., 0], cur[..., 2] - cur[..., 1], cur[..., 3] - cur[..., 2],
                  cur[..., 4] - cur[..., 3], cur[..., 5] - cur[..., 4], nxt[..., 0] - cur[..., 5],
                  nxt[..., 0] - cur[..., 0]], -1) / 1e3  # us
    print("%d ranks, %d steps; microseconds (mean / p10 / p90)" % (world, d.shape[1]))
    for r in range(world):
        print("rank %d: " % r + " | ".join("%s %.1f/%.1f/%.1f" % (name, d[r, :, i].mean(), np.percentile(d[r, :, i], 10),
                                                                  np.percentile(d[r, :, i], 90)) for i, name in enumerate(STAGES)))
    # %globaltimer is per GPU: align on the exit of the first barrier, which all ranks leave within a flag flight
    off = np.median(S[:, :, 1] - S[0:1, :, 1], axis=1)
    A = S - off[:, None, None]
    start = A[:, :, 0]
    print("clock offsets vs rank 0 [us]:", np.round(off / 1e3, 1).tolist())
    spread = (start.max(0) - start.min(0)) / 1e3
    print("forward start spread between ranks: mean %.1f us, p90 %.1f us" % (spread.mean(), np.percentile(spread, 90)))
    sig = (A[:, :, 1].min(0) - start.max(0)) / 1e3
    print("last rank starts -> first rank leaves the barrier: mean %.1f us, min %.1f us (flag latency)" % (sig.mean(), sig.min()))
    print("rank starting last, count per rank:", np.bincount(start.argmax(0), minlength=world).tolist())
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
