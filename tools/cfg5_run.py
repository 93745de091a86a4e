"""BASELINE configs[4] at scale: synthetic 10 M users x 1 M items, dim 256, batch 10 000 per GPU, row-sharded
tables with NCCL all-to-all (launch under torchrun with 2/4/8 ranks).  Prints one JSON line (rank 0).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/cfg5_run.py [steps] [n_users] [n_items] [dim]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    nu = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
    ni = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
    D = int(sys.argv[4]) if len(sys.argv) > 4 else 256
    B, W = 10_000, 10
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import anime_recommendations_b200 as ar
    from anime_recommendations_b200.dist import ShardedTrainSession
    m = ar.EmbeddingDotModel((nu + world - 1) // world, (ni + world - 1) // world, D, seed=1 + rank, adam_mode="replay",
                             dense_kernel=1.0, device_init=True)
    g = torch.Generator(device=dev)
    g.manual_seed(42 + rank)
    n = (W + steps) * B
    iu = torch.randint(0, nu, (n,), generator=g, device=dev, dtype=torch.int32)
    ia = torch.randint(0, ni, (n,), generator=g, device=dev, dtype=torch.int32)
    y = torch.randint(0, 11, (n,), generator=g, device=dev).float() / 10.0
    sess = ShardedTrainSession(m, B, total_steps=W + steps + 8)
    sess.run(iu[:W * B], ia[:W * B], y[:W * B], 1e-5)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sess.run(iu[W * B:], ia[W * B:], y[W * B:], 1e-5)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    loss = sess.metrics[m.iterations, 0].item()
    if rank == 0:
        row_b = D * 4
        per_gpu_exchange = 2 * (B * row_b + B * (row_b + 16))          # rows in + gradients out, both tables, upper bound
        print(json.dumps(dict(workload="cfg5: %d users x %d items, dim %d, batch %d/GPU, row-sharded tables" % (nu, ni, D, B),
                              n_gpus=world, steps=steps, ms_per_step=ms / steps, samples_per_s=world * steps * B / (ms / 1e3),
                              table_gb_per_gpu=3 * (m.U.numel() + m.A.numel()) * 4 / 1e9, exchange_cap=sess.caps[-1],
                              nvlink_bytes_per_step_per_gpu_upper=per_gpu_exchange, last_bce=loss)))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
