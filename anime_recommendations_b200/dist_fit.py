"""Multi-GPU `fit`: the reference's data-parallel intent (tf.distribute strategy scope around model creation and
`model.fit`, neural_network.py:142-147,173-182,210-217) on one 8 x B200 box, one process per GPU (torchrun).

Every rank builds a DistributedEmbeddingDotModel and calls fit() with the SAME full arrays; the tables are
row-sharded (global row g lives on rank g % world with its Adam state) and trained through the NVLink peer-memory
path (dist.PeerTrainSession).  `batch_size` is the PER-REPLICA batch, as in the reference
(`batch_size * strategy.num_replicas_in_sync`, neural_network.py:176): global step s visits samples
[s*G*B, (s+1)*G*B) of the epoch order and rank r takes the r-th slice of B of them, so an N-GPU fit equals a
1-GPU fit with batch_size G*B on the same data, seed and shuffle rule (up to the order of the global sums).
The epoch order is cut to a multiple of G samples (at most G-1 samples of an epoch are not visited).
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np
import torch
import torch.distributed as dist

from . import _capi, weights_io
from ._capi import check, lib, ptr, stream_ptr
from .model import EmbeddingDotModel, History


def shard_of(table, rank, world):
    """Rows g with g % world == rank, padded to ceil(n / world) rows."""
    n = table.shape[0]
    rows = (n + world - 1) // world
    out = torch.zeros((rows,) + tuple(table.shape[1:]), dtype=table.dtype, device=table.device)
    mine = table[rank::world]
    out[:mine.shape[0]] = mine
    return out


def replica_slice(order, world, batch, rank):
    """Which samples of an epoch's visit order `order` replica `rank` trains on, in its own visit order.

    Global step s of the epoch covers order[s*G*B : (s+1)*G*B]; replica r takes the r-th run of B of them
    (`batch_size * strategy.num_replicas_in_sync`, neural_network.py:176).  The order is cut to a multiple of G samples;
    a last, shorter global step is split evenly.  Concatenating the replicas' slices of a step in rank order gives the
    step's global batch back -- which is what makes an N-GPU fit equal the 1-GPU fit with batch G*B."""
    order = np.asarray(order)
    G, B = int(world), int(batch)
    n_use = (len(order) // G) * G
    glob = order[:n_use]
    F = (n_use // (G * B)) * G * B
    head = glob[:F].reshape(-1, G, B)[:, rank, :].reshape(-1)
    tail = glob[F:].reshape(G, -1)[rank] if n_use > F else glob[:0]
    return np.concatenate([head, tail])


class DistributedEmbeddingDotModel:
    """The Keras-like facade of model.EmbeddingDotModel over row-sharded tables.  Collective methods (every rank
    must call them): fit, get_weights, get_layer(...).get_weights, predict, evaluate, save, save_weights."""

    def __init__(self, n_users, n_anime, embedding_size=128, l2_reg_factor=1e-4, seed=None, adam_mode="replay",
                 device=None, **kw):
        if not dist.is_initialized():
            raise _capi.AnimerecError("torch.distributed is not initialised: launch under torchrun "
                                      "(python -m torch.distributed.run --nproc-per-node N ...)")
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.n_users, self.n_anime, self.dim = int(n_users), int(n_anime), int(embedding_size)
        self.l2, self.adam_mode = float(l2_reg_factor), adam_mode
        self._kw = dict(l2_reg_factor=l2_reg_factor, adam_mode=adam_mode, device=device, **kw)
        # the same initial weights as the single-GPU model of this seed: build it once, keep this rank's rows
        full = EmbeddingDotModel(n_users, n_anime, embedding_size, seed=seed, **self._kw)
        G, r = self.world, self.rank
        self.shard = EmbeddingDotModel((n_users + G - 1) // G, (n_anime + G - 1) // G, embedding_size, seed=seed, **self._kw)
        self.shard.U.copy_(shard_of(full.U, r, G))
        self.shard.A.copy_(shard_of(full.A, r, G))
        self.shard.head.copy_(full.head)
        self.shard.bn_moving.copy_(full.bn_moving)
        self.names = full.names
        self.device = self.shard.device
        del full
        torch.cuda.empty_cache()
        self.lr = 1e-3
        self.stop_training = False
        self.history = None
        self.timings = {}

    # ------------------------------------------------------------------ assembling the replicated view
    def _gather(self, t, n_rows):
        g = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(g, t.contiguous())
        out = torch.empty((n_rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        for r in range(self.world):
            rows = out[r::self.world].shape[0]
            out[r::self.world] = g[r][:rows]
        return out

    def assemble(self, with_slots=False):
        """A replicated EmbeddingDotModel holding the full tables (and, with_slots, the Adam state) -- collective."""
        s = self.shard
        s._sync_tables()
        full = EmbeddingDotModel(self.n_users, self.n_anime, self.dim, seed=0, device_init=True, **self._kw)
        full.names = self.names
        full.U.copy_(self._gather(s.U, self.n_users))
        full.A.copy_(self._gather(s.A, self.n_anime))
        full.head.copy_(s.head)
        full.bn_moving.copy_(s.bn_moving)
        full.iterations = s.iterations
        if with_slots:
            full.mU.copy_(self._gather(s.mU, self.n_users))
            full.vU.copy_(self._gather(s.vU, self.n_users))
            full.mA.copy_(self._gather(s.mA, self.n_anime))
            full.vA.copy_(self._gather(s.vA, self.n_anime))
            full.head_m.copy_(s.head_m)
            full.head_v.copy_(s.head_v)
        full.lastU.fill_(full.iterations)
        full.lastA.fill_(full.iterations)
        full._t_flush = full.iterations
        return full

    def get_weights(self):
        return self.assemble().get_weights()

    def get_layer(self, name):
        return self.assemble().get_layer(name)

    def predict(self, x, verbose=0, batch_size=None):
        return self.assemble().predict(x, verbose=verbose, batch_size=batch_size)

    def evaluate(self, x, y, **kw):
        return self.assemble().evaluate(x, y, **kw)

    def save(self, path, include_optimizer=True):
        full = self.assemble(with_slots=include_optimizer)
        if self.rank == 0:
            weights_io.save_model(full, path, include_optimizer=include_optimizer)
        dist.barrier()

    def save_weights(self, path):
        full = self.assemble()
        if self.rank == 0:
            weights_io.save_model(full, path, include_optimizer=False, weights_only=True)
        dist.barrier()

    # the EarlyStopping callback keeps / restores the best weights: each rank its own shard
    def _device_weights(self):
        return self.shard._device_weights()

    def _restore_device_weights(self, w):
        self.shard._restore_device_weights(w)

    @property
    def iterations(self):
        return self.shard.iterations

    def _reg_total(self, local):
        t = torch.tensor([local], dtype=torch.float64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ------------------------------------------------------------------ training
    def fit(self, x, y, batch_size=10000, epochs=1, verbose=0, validation_data=None, callbacks=None,
            shuffle="numpy", shuffle_seed=0, initial_epoch=0):
        """model.fit of neural_network.py:210-217 under a data-parallel strategy; `batch_size` per replica."""
        from .dist import PeerTrainSession
        s, G, r, dev = self.shard, self.world, self.rank, self.device
        iu_all = np.asarray(x[0]).reshape(-1).astype(np.int64)
        ia_all = np.asarray(x[1]).reshape(-1).astype(np.int64)
        y_all = np.asarray(y, np.float64).reshape(-1).astype(np.float32)
        N = len(iu_all)
        if not (len(ia_all) == N == len(y_all)) or N < G:
            raise ValueError("x[0], x[1] and y must be of equal length, at least one sample per replica")
        if iu_all.min() < 0 or iu_all.max() >= self.n_users or ia_all.min() < 0 or ia_all.max() >= self.n_anime:
            raise IndexError("index outside the vocabulary")
        B = int(batch_size)
        GB = G * B
        n_use = (N // G) * G
        steps = (n_use + GB - 1) // GB
        full_steps = n_use // GB
        n_last_local = (n_use - full_steps * GB) // G if n_use > full_steps * GB else B
        sess = PeerTrainSession(s, B, total_steps=max(0, epochs - initial_epoch) * steps)
        callbacks = list(callbacks or [])
        for cb in callbacks:
            cb.set_model(self)
        hist = History()
        self.history = hist
        self.stop_training = False
        for cb in callbacks:
            cb.on_train_begin()
        val = None
        if validation_data is not None:
            vx, vy = validation_data[0], validation_data[1]
            val = (EmbeddingDotModel._as_idx(vx[0], dev), EmbeddingDotModel._as_idx(vx[1], dev),
                   EmbeddingDotModel._as_f32(vy, dev))
        try:
            for epoch in range(initial_epoch, epochs):
                t_epoch = time.perf_counter()
                for cb in callbacks:
                    cb.on_epoch_begin(epoch)
                lr = float(self.lr)
                t0 = s.iterations
                if shuffle in ("numpy", "device"):    # one rule on every rank: the seeded NumPy permutation
                    perm = np.random.RandomState(shuffle_seed + epoch).permutation(N)
                elif shuffle in (False, None, "none"):
                    perm = np.arange(N)
                else:
                    raise ValueError("shuffle must be 'numpy', 'device' or False")
                mine = replica_slice(perm, G, B, r)
                iu = torch.from_numpy(iu_all[mine].astype(np.int32)).to(dev)
                ia = torch.from_numpy(ia_all[mine].astype(np.int32)).to(dev)
                yy = torch.from_numpy(y_all[mine]).to(dev)
                reg0 = s._begin_reg(steps, B, n_last_local)
                sess.run(iu, ia, yy, lr)                       # verify() included
                acc = self._reg_total(s._end_reg())            # flushes this rank's shards, then all-reduces
                m = sess.metrics[t0 + 1:t0 + steps + 1].cpu().numpy().astype(np.float64)   # global-batch metrics
                w = m[:, 2]
                bce = float((m[:, 0] * w).sum() / n_use)
                mse = float((m[:, 1] * w).sum() / n_use)
                reg1 = self.l2 * self._reg_total(s.reg_sumsq())
                if self.adam_mode == "touched":
                    reg = 0.5 * (self.l2 * self._reg_total(reg0) + reg1)
                else:
                    reg = self.l2 * acc * GB / n_use
                logs = dict(loss=bce + reg, mse=mse)
                if val is not None:
                    full = self.assemble()
                    sums = torch.zeros(2, dtype=torch.float64, device=dev)
                    check(lib().ar_eval_sums(ptr(full.U), ptr(full.A), self.dim, ptr(full.head), ptr(full.bn_moving),
                                             ptr(val[0]), ptr(val[1]), ptr(val[2]), val[0].numel(), ptr(sums),
                                             stream_ptr()), "ar_eval_sums")
                    sv = sums.cpu().numpy()
                    nv = max(1, val[0].numel())
                    logs["val_loss"] = float(sv[0] / nv + reg1)
                    logs["val_mse"] = float(sv[1] / nv)
                    del full
                logs["lr"] = float(np.float32(lr))
                self.timings.setdefault("epoch_s", []).append(time.perf_counter() - t_epoch)
                for k, v in logs.items():
                    hist.history.setdefault(k, []).append(v)
                hist.epoch.append(epoch)
                if verbose and r == 0:
                    print("Epoch %d/%d - %.2fs - %s" % (epoch + 1, epochs, self.timings["epoch_s"][-1],
                                                         " - ".join("%s: %.6g" % kv for kv in logs.items())))
                for cb in callbacks:
                    cb.on_epoch_end(epoch, logs)
                if self.stop_training:
                    break
            for cb in callbacks:
                cb.on_train_end()
        finally:
            sess.close()
        return hist
