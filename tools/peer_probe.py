"""Developer probe of the multi-GPU (peer-memory) training step at cfg2 shapes:
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/peer_probe.py [steps] [zipf]
(N = 1 works on a single-GPU box: every pull is local.)  Prints us/step (max over ranks) and, for the persistent
kernel, CTA 0's phase timeline of the last chunk.  AR_PEER_STAGED=1 selects the per-step stage kernels."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import anime_recommendations_b200 as ar  # noqa: E402
from anime_recommendations_b200.dist import PeerTrainSession  # noqa: E402

N_USERS, N_ANIME, DIM, BATCH = 310_000, 16_500, 128, 10_000


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    zipf = len(sys.argv) > 2 and sys.argv[2] == "zipf"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    m = ar.EmbeddingDotModel((N_USERS + world - 1) // world, (N_ANIME + world - 1) // world, DIM, l2_reg_factor=1e-4,
                             seed=1 + rank, adam_mode="replay", dense_kernel=1.0)
    g = torch.Generator(device=dev)
    g.manual_seed(7 + rank)

    def synth(n):
        if zipf:
            rs = np.random.RandomState(11 + rank)
            pu = 1.0 / np.arange(1, N_USERS + 1)
            pa = 1.0 / np.arange(1, N_ANIME + 1)
            iu = torch.from_numpy(rs.choice(N_USERS, n, p=pu / pu.sum()).astype(np.int32)).to(dev)
            ia = torch.from_numpy(rs.choice(N_ANIME, n, p=pa / pa.sum()).astype(np.int32)).to(dev)
        else:
            iu = torch.randint(0, N_USERS, (n,), generator=g, device=dev, dtype=torch.int32)
            ia = torch.randint(0, N_ANIME, (n,), generator=g, device=dev, dtype=torch.int32)
        y = torch.randint(0, 11, (n,), generator=g, device=dev).float() / 10.0
        return iu, ia, y

    sess = PeerTrainSession(m, BATCH, total_steps=2 * steps + 8)
    wu = synth(steps * BATCH)
    tu = synth(steps * BATCH)
    sess.run(*wu, 1e-3, verify=False)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sess.run(*tu, 1e-3, verify=False)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    sess.verify()
    if rank == 0:
        us = float(ms.item()) * 1e3 / steps
        print("PEER_PROBE world %d persistent %s zipf %s: %.2f us/step, %.1f M samples/s" % (
            world, sess.persistent, zipf, us, world * BATCH / us))
        if sess.persistent:
            tl = sess.timeline()
            print("  phases us (CTA 0, mean): gate %.1f | fwd+barrier %.1f (own forward %.1f) | head %.1f | update %.1f | step %.1f" % tuple(
                float(np.mean(tl[k][4:])) for k in ("gate_us", "fwd_us", "fwd_own_us", "head_us", "update_us", "step_us")))
            print("  replay: %d items, %d element-steps, busy %.2f of warp-cycles" % (
                tl["replay_items"], tl["replay_element_steps"],
                tl["replay_busy_cycles"] / max(1, tl["replay_warps"] * tl["kernel_cycles"])))
    sess.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
