"""Workload for ncu captures of the Half-B kernels: a few single-query top-k calls over 350 000 users and one
all-pairs candidate pass (developer tool).  python tools/profile_similarity.py [n_allpairs]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from anime_recommendations_b200 import similarity as sim  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev)
g.manual_seed(7)
W = torch.randn((350000, 128), generator=g, device=dev)
for q in (5, 70000, 200000, 349999):
    sim.cosine_topk_query_device(W, q, 11, exclude=q)
torch.cuda.synchronize()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 350000
Wn = sim.normalize_rows_bf16(W[:n].contiguous())
sim.allpairs_candidates(Wn, 0, n, Wn, 0, n, 16, exclude_self=True)
torch.cuda.synchronize()
print("ok")
