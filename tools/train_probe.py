"""Run a few chunks of the single-GPU training kernel at cfg2 shapes (developer tool; target of ncu captures).

    python tools/train_probe.py [mode] [steps] [depth]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mode = sys.argv[1] if len(sys.argv) > 1 else "replay"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 512
if len(sys.argv) > 3:
    os.environ["AR_REPLAY_DEPTH"] = sys.argv[3]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import anime_recommendations_b200 as ar  # noqa: E402
import bench  # noqa: E402
from anime_recommendations_b200.model import TrainSession  # noqa: E402

dev = torch.device("cuda:0")
m = ar.EmbeddingDotModel(bench.N_USERS, bench.N_ANIME, bench.DIM, l2_reg_factor=bench.L2, seed=1, adam_mode=mode,
                         dense_kernel=1.0)
iu, ia, y = bench.synth(K * bench.BATCH, 42, dev, zipf="--zipf" in sys.argv)
sess = TrainSession(m, bench.BATCH, total_steps=2 * K + 8)
sess.run(iu, ia, y, bench.LR)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
sess.run(iu, ia, y, bench.LR)
e1.record()
torch.cuda.synchronize()
sess.check_health()
tl = sess.timeline()
med = lambda v: float(np.median(v[8:])) if len(v) > 16 else float(np.median(v))
print("mode %s depth %d: %.2f us/step over %d steps" % (mode, sess.depth, e0.elapsed_time(e1) * 1e3 / K, K))
print("phases (median us):", {k: round(med(v), 2) for k, v in tl.items() if k.endswith("_us")})
if mode == "replay":
    print("replay: busy %.3f of warp time, %d items, %.1f M element-steps per step" % (
        tl["replay_busy_cycles"] / (tl["replay_warps"] * tl["kernel_cycles"]), tl["replay_items"],
        tl["replay_element_steps"] / tl["steps"] / 1e6))
