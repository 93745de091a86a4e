"""BASELINE configs[2] and [3] on N GPUs (launch under torchrun): all-pairs cosine top-10 over the 350 000 x 128
user table with the candidate rows sharded per GPU (local exact top-k -> NCCL all-gather -> merge) or the query
rows sharded (no data-path collective), and model_recs scoring of 65 000 users x 18 000 anime with the users
sharded.  One JSON line (rank 0); times are device times, max over ranks.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/cfg34_run.py
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import anime_recommendations_b200 as ar
    from anime_recommendations_b200 import similarity_dist as sd
    from anime_recommendations_b200.dist import Comm
    import bench
    comm = Comm()
    n, k = bench.N_USERS, 10
    g = torch.Generator(device=dev)
    g.manual_seed(7)                                             # the same table on every rank
    W = torch.randn((n, bench.DIM), generator=g, device=dev)
    out = dict(n_gpus=world)

    def timed(fn, reps=3):
        best = None
        for _ in range(reps):
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = float(t.item()) if best is None else min(best, float(t.item()))
        return best, r

    fl = 2.0 * n * n * bench.DIM
    for shard in ("candidates", "queries"):
        sd.allpairs_topk_sharded(W[:4096].contiguous(), k, comm, shard=shard)        # warm-up
        ms, (gi, gs) = timed(lambda: sd.allpairs_topk_sharded(W, k, comm, shard=shard))
        out["allpairs_users_" + shard] = dict(rows=n, k=k, ms=ms, rows_per_s=n / (ms / 1e3), tflops=fl / (ms / 1e3) / 1e12,
                                              checksum=int(gi.sum().item()))
    # cfg4: users sharded
    rng = np.random.RandomState(11)
    m = ar.EmbeddingDotModel(bench.N_USERS, bench.N_ANIME, bench.DIM, seed=5, dense_kernel=-0.8, device_init=True)
    m.U.copy_(torch.randn(m.U.shape, generator=g, device=dev))
    m.A.copy_(torch.randn(m.A.shape, generator=g, device=dev))
    nq = 65_000
    users = rng.choice(bench.N_USERS, nq, replace=False)
    counts = rng.randint(400, 1501, nq)
    indptr = np.r_[0, np.cumsum(counts)].astype(np.int64)
    start = rng.randint(0, bench.N_ANIME, nq)
    stride = np.array([7, 11, 13, 17, 19, 23, 29, 31])[rng.randint(0, 8, nq)]
    j = np.arange(indptr[-1], dtype=np.int64) - np.repeat(indptr[:-1], counts)
    widx = ((np.repeat(start, counts) + j * np.repeat(stride, counts)) % bench.N_ANIME).astype(np.int32)
    sd.score_topk_sharded(m, users[:512 * world], indptr[:512 * world + 1], widx[:indptr[512 * world]], 20, rank, world)
    best = None
    for _ in range(2):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        lo, hi, oi, pr = sd.score_topk_sharded(m, users, indptr, widx, 20, rank, world)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = float(t.item()) if best is None else min(best, float(t.item()))
    out["model_recs_scoring_users_sharded"] = dict(users=nq, anime=bench.N_ANIME, k=20, ms_host_to_host=best * 1e3,
                                                   users_per_s=nq / best, users_this_rank=hi - lo)
    comm.close()
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
