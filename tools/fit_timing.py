"""Where does Model.fit spend its wall time?  (developer tool)  python tools/fit_timing.py [mode] [steps]"""
import os
import sys
import time

os.environ["AR_FIT_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import anime_recommendations_b200 as ar  # noqa: E402
import bench  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "replay"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 200
dev = torch.device("cuda:0")
m = ar.EmbeddingDotModel(bench.N_USERS, bench.N_ANIME, bench.DIM, l2_reg_factor=bench.L2, seed=1, adam_mode=mode,
                         dense_kernel=1.0)
m.lr = bench.LR
hu, ha, hy = (t.cpu().pin_memory() for t in bench.synth(K * bench.BATCH, 77, dev))
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m.fit([hu, ha], hy, batch_size=bench.BATCH, epochs=1, shuffle=False)
    torch.cuda.synchronize()
    print("fit %d: %.2f ms total; sections(ms): %s" % (
        rep, (time.perf_counter() - t0) * 1e3, {k: round(v * 1e3, 2) for k, v in m.timings["sections"][-1].items()}),
        flush=True)
