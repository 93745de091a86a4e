"""Timing of the tensor-core candidate pass and of the whole all-pairs pipeline (developer tool, not a bench line).

    python tools/perf_allpairs.py cand 350000 16      # candidate kernel only (honours AR_AP_DEBUG=0..3)
    python tools/perf_allpairs.py full 350000 16      # rownorm + candidates + re-rank + recovery
    python tools/perf_allpairs.py attrib 350000 16    # candidate kernel under AR_AP_DEBUG=3,2,1,0 (subprocesses)
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def table(n, dev):
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    return torch.randn((n, 128), generator=g, device=dev)


def cand(n, kp, reps=int(os.environ.get("AR_PERF_REPS", "3"))):
    import torch
    from anime_recommendations_b200 import similarity as sim
    dev = torch.device("cuda:0")
    W = table(n, dev)
    Wn = sim.normalize_rows_bf16(W)
    sim.allpairs_candidates(Wn, 0, min(n, 4096), Wn, 0, n, kp, exclude_self=True)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cl = sim.allpairs_candidates(Wn, 0, n, Wn, 0, n, kp, exclude_self=True)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    fl = 2.0 * n * n * 128
    print("cand n=%d kp=%d dbg=%s chunks=%d: %.3f ms  %.1f TFLOP/s  mean list %.1f" % (
        n, kp, os.environ.get("AR_AP_DEBUG", "0"), cl.idx.shape[0], ms, fl / ms / 1e9, float(cl.cnt.float().mean())),
        flush=True)


def full(n, kp, reps=2):
    import torch
    from anime_recommendations_b200 import similarity as sim
    dev = torch.device("cuda:0")
    W = table(n, dev)
    sim.allpairs_topk(W[:4096].contiguous(), k=10, kprime=kp)
    for _ in range(reps):
        st = {"time": True}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        oi, os_ = sim.allpairs_topk(W, k=10, kprime=kp, stats=st)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("full n=%d kp=%d: %.3f ms  %.1f TFLOP/s-equivalent  %.0f rows/s  %s" % (
            n, kp, dt * 1e3, 2.0 * n * n * 128 / dt / 1e12, n / dt, st), flush=True)


if __name__ == "__main__":
    mode, n, kp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    if mode == "cand":
        cand(n, kp)
    elif mode == "full":
        full(n, kp)
    else:
        for dbg in ("3", "2", "1", "0"):
            env = dict(os.environ, AR_AP_DEBUG=dbg)
            subprocess.run([sys.executable, os.path.abspath(__file__), "cand", str(n), str(kp)], env=env, check=False)
