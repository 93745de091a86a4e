"""Developer tool: repeat the small replay-mode fit of tests/test_gpu_train.py::test_fit_matches_reference_arithmetic many
times and report every trial whose tables differ from the oracle -- which rows, when they were touched, by how much.

    python tools/race_probe.py [dim=256] [trials=30] [mode=replay]        (AR_REPLAY_DEPTH=n to vary the look-ahead)

Round 2: 11-22 of 60 trials differed while a row's flag was published without its fence (a dangling else in
finish_loaded); dense mode and the previous build: 0 of 60."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import anime_recommendations_b200 as ar  # noqa: E402
import test_gpu_train as T  # noqa: E402
from oracle import train as ot  # noqa: E402

dim = int(sys.argv[1]) if len(sys.argv) > 1 else 256
trials = int(sys.argv[2]) if len(sys.argv) > 2 else 30
mode = sys.argv[3] if len(sys.argv) > 3 else "replay"
n_users, n_anime, n, B = 700, 90, 5300, 1000
iu, ia, y = T._problem(11, n_users, n_anime, n, False)
vu, va, vy = T._problem(12, n_users, n_anime, 500)
st = ot.init_state(n_users, n_anime, dim, seed=5, w=-1.3)
lr_kw = dict(start_lr=1e-3, min_lr=1e-3, max_lr=3e-3, rampup_epochs=2, sustain_epochs=0, exp_decay=0.8)
st0 = ot.init_state(n_users, n_anime, dim, seed=5, w=-1.3)
ot.fit(st, [iu, ia], y, B, 3, ([vu, va], vy), lr_kwargs=lr_kw, shuffle_seed=0, patience=99)
# touch pattern: step -> rows (the numpy shuffle of fit)
touch_u, touch_a = {}, {}
for e in range(3):
    perm = np.random.RandomState(0 + e).permutation(n)
    for s in range(6):
        idx = perm[s * B:(s + 1) * B]
        t = e * 6 + s + 1
        for r in np.unique(iu[idx]):
            touch_u.setdefault(int(r), []).append(t)
        for r in np.unique(ia[idx]):
            touch_a.setdefault(int(r), []).append(t)
bad = 0
for k in range(trials):
    m = T._model_from_state(st0, mode)
    sched = ar.LearningRateScheduler(lambda e: ar.lrfn(e, **lr_kw))
    m.fit([iu, ia], y, batch_size=B, epochs=3, validation_data=([vu, va], vy), callbacks=[sched], shuffle="numpy",
          shuffle_seed=0)
    w = m.get_weights()
    eu = np.abs(w[0] - st.U).max(axis=1)
    ea = np.abs(w[1] - st.A).max(axis=1)
    wu, wa = np.nonzero(eu > 2e-5)[0], np.nonzero(ea > 2e-5)[0]
    if len(wu) or len(wa):
        bad += 1
        print("TRIAL %d: %d user rows, %d anime rows off" % (k, len(wu), len(wa)))
        for r in wu[:4]:
            print("   user %4d err %.2e touched at %s" % (r, eu[r], touch_u.get(int(r), [])))
        for r in wa[:2]:
            print("   anime %3d err %.2e touched at %s" % (r, ea[r], touch_a.get(int(r), [])))
print("RACE_PROBE dim %d mode %s depth %s: %d of %d trials differ" % (dim, mode, os.environ.get("AR_REPLAY_DEPTH", "default"), bad, trials))
