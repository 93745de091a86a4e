"""Worker for tests/test_gpu_dist.py (one process per GPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sim_main(shard, dev, rank, world):
    """Sharded all-pairs top-k == the single-process oracle (every rank must hold the full, identical result)."""
    from anime_recommendations_b200 import similarity_dist as sd
    from anime_recommendations_b200.dist import Comm
    from oracle import similarity as osim
    from gpu_util import assert_topk_close
    n, k = 3001, 10
    W = np.random.RandomState(11).standard_normal((n, 128)).astype(np.float32)
    comm = Comm()
    gi, gs = sd.allpairs_topk_sharded(W, k, comm, shard=shard)
    gi, gs = gi.cpu().numpy(), gs.cpu().numpy()
    oi, os_ = osim.allpairs_topk_fast(W, k)
    np.testing.assert_allclose(gs, os_, rtol=0, atol=3e-6)
    Wn = osim.get_weights(W)
    for r in np.nonzero((gi != oi).any(axis=1))[0]:
        assert_topk_close(gi[r], gs[r], oi[r], os_[r], Wn @ Wn[r])
    t = torch.from_numpy(gi).to(dev)
    g = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(g, t)
    assert all(torch.equal(g[0], o) for o in g[1:]), "ranks disagree"
    comm.close()
    if rank == 0:
        print("DIST_OK", shard)


def score_main(dev, rank, world):
    """cfg4 scoring with the users sharded over the ranks == the single-process result (gathered on every rank)."""
    import anime_recommendations_b200 as ar
    from anime_recommendations_b200 import similarity as sim, similarity_dist as sd
    from anime_recommendations_b200.dist import Comm
    rng = np.random.RandomState(21)
    nu, na, k = 600, 1800, 20
    m = ar.EmbeddingDotModel(nu, na, 128, seed=3, dense_kernel=0.9)
    users = rng.choice(nu, 301, replace=False)
    counts = rng.randint(100, 700, len(users))
    indptr = np.r_[0, np.cumsum(counts)]
    widx = np.concatenate([rng.choice(na, c, replace=False) for c in counts]).astype(np.int32)
    comm = Comm()
    gi, gp = sd.score_topk_sharded(m, users, indptr, widx, k, rank, world, comm=comm, gather=True)
    wi, wp = sim.score_topk(m, users, indptr, widx, k)
    np.testing.assert_array_equal(gi, wi)
    np.testing.assert_array_equal(gp, wp)
    comm.close()
    if rank == 0:
        print("DIST_OK score")


def fit_main(dev, rank, world):
    """DistributedEmbeddingDotModel.fit on `world` GPUs == EmbeddingDotModel.fit on one GPU with batch world*B:
    history (loss incl. the exact L2 term, mse, val_loss, val_mse, lr), checkpoint callback, gathered tables."""
    import tempfile
    import anime_recommendations_b200 as ar
    from anime_recommendations_b200.dist_fit import DistributedEmbeddingDotModel
    nu, na, D, B = 3001, 403, 128, 500
    rng = np.random.RandomState(31)
    n = world * B * 5 + world * 123                                # 6 global steps / epoch, the last one partial
    iu, ia = rng.randint(0, nu, n), rng.randint(0, na, n)
    y = rng.randint(0, 11, n) / 10.0
    vu, va, vy = rng.randint(0, nu, 400), rng.randint(0, na, 400), rng.randint(0, 11, 400) / 10.0
    lr_kw = dict(start_lr=1e-3, min_lr=1e-3, max_lr=3e-3, rampup_epochs=2, sustain_epochs=0, exp_decay=0.8)
    tmp = tempfile.mkdtemp() if rank == 0 else None
    box = [tmp]
    dist.broadcast_object_list(box, src=0)
    ck = os.path.join(box[0], "best_weights.h5")
    dm = DistributedEmbeddingDotModel(nu, na, D, seed=7, dense_kernel=0.8)
    cbs = [ar.LearningRateScheduler(lambda e: ar.lrfn(e, **lr_kw)),
           ar.ModelCheckpoint(ck, save_weights_only=True, monitor="val_loss", mode="min", save_best_only=True),
           ar.EarlyStopping(patience=3, monitor="val_loss", mode="min", restore_best_weights=True)]
    h = dm.fit([iu, ia], y, batch_size=B, epochs=3, validation_data=([vu, va], vy), callbacks=cbs, shuffle_seed=4)
    w = dm.get_weights()
    p = dm.predict([vu, va])
    if rank == 0:
        m1 = ar.EmbeddingDotModel(nu, na, D, seed=7, dense_kernel=0.8)
        h1 = m1.fit([iu, ia], y, batch_size=world * B, epochs=3, validation_data=([vu, va], vy),
                    callbacks=[ar.LearningRateScheduler(lambda e: ar.lrfn(e, **lr_kw))], shuffle_seed=4)
        for k in ("loss", "mse", "val_loss", "val_mse", "lr"):
            np.testing.assert_allclose(h.history[k], h1.history[k], rtol=1e-4, atol=2e-6, err_msg=k)
        w1 = m1.get_weights()
        np.testing.assert_allclose(w[0], w1[0], rtol=1e-4, atol=3e-6)
        np.testing.assert_allclose(w[1], w1[1], rtol=1e-4, atol=3e-6)
        np.testing.assert_allclose(p, m1.predict([vu, va]), rtol=0, atol=2e-4)
        best = ar.load_model(ck)                                   # what the checkpoint callback wrote (rank 0)
        assert best.get_weights()[0].shape == (nu, D)
        print("DIST_OK fit", h.history["loss"])


def shard_main(mode, dev, rank, world, peer=False):
    """Row-sharded training (NCCL all-to-all, or NVLink peer memory) == single-GPU training on the concatenated batch."""
    import anime_recommendations_b200 as ar
    from anime_recommendations_b200.dist import PeerTrainSession, ShardedTrainSession, shard_rows
    from anime_recommendations_b200.model import TrainSession
    nu, na, D, B, steps = 5001, 703, 128, 1000, 6
    rng = np.random.RandomState(5)
    n_glob = world * B * steps - world * 300
    iu = rng.randint(0, nu, n_glob).astype(np.int32)
    ia = rng.randint(0, na, n_glob).astype(np.int32)
    ia[rng.rand(n_glob) < 0.2] = 3                                # heavy anime row
    y = (rng.randint(0, 11, n_glob) / 10.0).astype(np.float32)
    per = n_glob // world
    sl = slice(rank * per, (rank + 1) * per)
    full = ar.EmbeddingDotModel(nu, na, D, seed=2, adam_mode=mode, dense_kernel=-1.2)    # the single-GPU twin
    fw = full.get_weights()
    nul, nal = (nu + world - 1) // world, (na + world - 1) // world
    m = ar.EmbeddingDotModel(nul, nal, D, seed=0, adam_mode=mode, dense_kernel=-1.2)
    Us, As = np.zeros((nul, D), np.float32), np.zeros((nal, D), np.float32)
    mine_u, mine_a = shard_rows(fw[0], rank, world), shard_rows(fw[1], rank, world)
    Us[:len(mine_u)], As[:len(mine_a)] = mine_u, mine_a
    m.set_weights([Us, As] + fw[2:])
    if peer:   # 3 chunks of 2 steps: the double-buffered chunk planning is part of what is checked
        import anime_recommendations_b200.model as arm
        arm.PLAN_CHUNK = 2
    sess = (PeerTrainSession if peer else ShardedTrainSession)(m, B, total_steps=steps)
    sess.run(torch.from_numpy(iu[sl]).to(dev), torch.from_numpy(ia[sl]).to(dev), torch.from_numpy(y[sl]).to(dev), 2e-3)
    if peer:
        sess.verify()
    m._sync_tables()
    # assemble the global tables on rank 0
    parts = {}
    for name, t in (("U", m.U), ("A", m.A), ("mU", m.mU), ("vU", m.vU), ("mA", m.mA), ("vA", m.vA)):
        g = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(g, t.contiguous())
        parts[name] = [x.cpu().numpy() for x in g]
    heads = [torch.empty_like(m.head) for _ in range(world)]
    dist.all_gather(heads, m.head)
    assert all(torch.equal(heads[0], h) for h in heads[1:]), "head replicas diverged"
    if rank == 0:
        order = []
        for s in range(steps):
            for r in range(world):
                lo = r * per + s * B
                order.append(np.arange(lo, min(lo + B, (r + 1) * per)))
        order = np.concatenate(order)
        s1 = TrainSession(full, world * B, total_steps=steps)
        s1.run(torch.from_numpy(iu[order]).to(dev), torch.from_numpy(ia[order]).to(dev),
               torch.from_numpy(y[order]).to(dev), 2e-3)
        full._sync_tables()
        ref = dict(U=full.U, A=full.A, mU=full.mU, vU=full.vU, mA=full.mA, vA=full.vA)
        for name, tref in ref.items():
            tref = tref.cpu().numpy()
            got = np.zeros_like(tref)
            for r in range(world):
                rows = tref[r::world].shape[0]
                got[r::world] = parts[name][r][:rows]
            tol = dict(rtol=1e-4, atol=3e-6) if name[0] != "v" else dict(rtol=2e-4, atol=1e-11)
            np.testing.assert_allclose(got, tref, err_msg=name, **tol)
        np.testing.assert_allclose(np.delete(m.head.cpu().numpy(), 1), np.delete(full.head.cpu().numpy(), 1), rtol=1e-4, atol=3e-6)
        mt = sess.metrics[1:steps + 1].cpu().numpy()
        m1t = s1.metrics[1:steps + 1].cpu().numpy()
        np.testing.assert_allclose(mt[:, :3], m1t[:, :3], rtol=2e-6, atol=2e-6)
        if peer:
            assert B // 2 <= max(sess.counts) <= world * B
            print("DIST_OK peer", mode, "longest list", sess.counts)
        else:
            assert max(sess.caps) <= B and min(sess.caps) >= 4
            print("DIST_OK shard", mode, "cap", sess.caps)


def main(mode):
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    if mode.startswith("shard_") or mode.startswith("peer_"):
        shard_main(mode.split("_", 1)[1], dev, rank, world, peer=mode.startswith("peer_"))
        dist.barrier()
        dist.destroy_process_group()
        return
    if mode == "fit":
        fit_main(dev, rank, world)
        dist.barrier()
        dist.destroy_process_group()
        return
    if mode == "score":
        score_main(dev, rank, world)
        dist.barrier()
        dist.destroy_process_group()
        return
    if mode.startswith("sim_"):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        sim_main(mode[4:], dev, rank, world)
        dist.barrier()
        dist.destroy_process_group()
        return
    import anime_recommendations_b200 as ar
    from anime_recommendations_b200.dist import DistTrainSession
    from anime_recommendations_b200.model import TrainSession
    nu, na, D, B, steps = 5000, 700, 128, 1000, 6
    rng = np.random.RandomState(5)
    n_glob = world * B * steps - world * 300                      # last step partial (700 per rank)
    iu = rng.randint(0, nu, n_glob).astype(np.int32)
    ia = rng.randint(0, na, n_glob).astype(np.int32)
    ia[rng.rand(n_glob) < 0.2] = 3                                # heavy anime row
    y = (rng.randint(0, 11, n_glob) / 10.0).astype(np.float32)
    # global step s = concatenation over ranks of each rank's s-th local batch
    per = n_glob // world
    sl = slice(rank * per, (rank + 1) * per)
    m = ar.EmbeddingDotModel(nu, na, D, seed=2, adam_mode=mode, dense_kernel=-1.2)
    sess = DistTrainSession(m, B, total_steps=steps)
    sess.run(torch.from_numpy(iu[sl]).to(dev), torch.from_numpy(ia[sl]).to(dev), torch.from_numpy(y[sl]).to(dev), 2e-3)
    m._sync_tables()
    mine = [t.clone() for t in (m.U, m.A, m.mU, m.vU, m.mA, m.vA, m.head, m.bn_moving)]
    # replicas identical?
    for t in mine:
        g = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        for o in g[1:]:
            assert torch.equal(g[0], o), "replicas diverged"
    if rank == 0:
        # single-GPU reference: batch = world*B, samples ordered [rank0 batch s | rank1 batch s | ...]
        order = []
        for s in range(steps):
            for r in range(world):
                lo = r * per + s * B
                order.append(np.arange(lo, min(lo + B, (r + 1) * per)))
        order = np.concatenate(order)
        m1 = ar.EmbeddingDotModel(nu, na, D, seed=2, adam_mode=mode, dense_kernel=-1.2)
        s1 = TrainSession(m1, world * B, total_steps=steps)
        s1.run(torch.from_numpy(iu[order]).to(dev), torch.from_numpy(ia[order]).to(dev),
               torch.from_numpy(y[order]).to(dev), 2e-3)
        m1._sync_tables()
        ref = [m1.U, m1.A, m1.mU, m1.vU, m1.mA, m1.vA, m1.head, m1.bn_moving]
        names = ["U", "A", "mU", "vU", "mA", "vA", "head", "bn"]
        for n, a, b in zip(names, mine, ref):
            a, b = a.cpu().numpy(), b.cpu().numpy()
            if n == "head":
                a, b = np.delete(a, 1), np.delete(b, 1)          # Dense bias: noise walk
            tol = dict(rtol=1e-4, atol=3e-6) if n not in ("vU", "vA") else dict(rtol=2e-4, atol=1e-11)
            if n == "bn":
                tol = dict(rtol=1e-4, atol=1e-4)
            np.testing.assert_allclose(a, b, err_msg=n, **tol)
        mt = sess.metrics[1:steps + 1].cpu().numpy()
        m1t = s1.metrics[1:steps + 1].cpu().numpy()
        np.testing.assert_allclose(mt[:, :3], m1t[:, :3], rtol=2e-6, atol=2e-6)
        print("DIST_OK", mode)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
