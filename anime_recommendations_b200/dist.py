"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for rendezvous, NCCL collectives
issued from libanimerec on the kernels' stream.

* DistTrainSession  -- replicated-table data-parallel training (BASELINE cfg2): every rank trains on its
  own shard of each global batch; SyncBN + all-gathered, merged row gradients keep replicas bit-identical.
* Comm              -- the library's own NCCL communicator (ncclCommInitRank with an id broadcast through
  torch.distributed), also used by the candidate-sharded top-k.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _capi
from ._capi import ArDistCtx, check, lib, ptr, stream_ptr
from .model import TrainSession


class Comm:
    """ncclComm_t owned by libanimerec; world = the default torch.distributed group."""

    def __init__(self):
        if not dist.is_initialized():
            raise _capi.AnimerecError("torch.distributed must be initialised (torchrun) before Comm()")
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        ident = (C.c_ubyte * 128)()
        if self.rank == 0:
            check(lib().ar_nccl_unique_id(ident), "ar_nccl_unique_id")
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.tensor(list(bytes(ident)), dtype=torch.uint8, device=dev)
        dist.broadcast(t, src=0)
        ident = (C.c_ubyte * 128)(*t.cpu().tolist())
        h = C.c_void_p()
        check(lib().ar_comm_init(ident, self.world, self.rank, C.byref(h)), "ar_comm_init")
        self.handle = h

    def allgather(self, t):
        """[world, *t.shape] tensor of every rank's `t` (equal shapes), on the current stream."""
        t = t.contiguous()
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        check(lib().ar_allgather_bytes(self.handle, ptr(t), ptr(out), t.numel() * t.element_size(), stream_ptr()),
              "ar_allgather_bytes")
        return out

    def close(self):
        if self.handle:
            check(lib().ar_comm_destroy(self.handle), "ar_comm_destroy")
            self.handle = None


class DistTrainSession(TrainSession):
    """TrainSession whose steps run ar_train_steps_dist: `batch` is the PER-RANK batch."""

    def __init__(self, model, batch, total_steps, comm=None):
        super().__init__(model, batch, total_steps)
        self.comm = comm or Comm()
        G, B, D, dev = self.comm.world, self.B, model.dim, model.device
        f = dict(dtype=torch.float32, device=dev)
        self.c_all, self.label_all, self.dy_all = (torch.empty(G * B, **f) for _ in range(3))
        self.fwd_part_all = torch.zeros(2 * ((G * B + 7) // 8), dtype=torch.float64, device=dev)
        self.head_part_all = torch.zeros(8 * ((G * B + 255) // 256), dtype=torch.float64, device=dev)
        words = 2 * B * (D + 2)
        self.send = torch.zeros(words, **f)
        self.recv = torch.zeros(G * words, **f)
        d = ArDistCtx()
        d.comm, d.n_ranks, d.rank = self.comm.handle, G, self.comm.rank
        d.c_all, d.label_all, d.dy_all = self.c_all.data_ptr(), self.label_all.data_ptr(), self.dy_all.data_ptr()
        d.fwd_part_all, d.head_part_all = self.fwd_part_all.data_ptr(), self.head_part_all.data_ptr()
        d.send, d.recv = self.send.data_ptr(), self.recv.data_ptr()
        # every rank's distinct-row lists of the planned chunk (look-ahead catch-up needs to know which rows
        # the OTHER ranks touch in the current step)
        S = self.n_slots
        self.uniq_all = [torch.zeros((G, S, B), dtype=torch.int32, device=dev) for _ in range(2)]
        self.meta_all = [torch.zeros((G, S, 4), dtype=torch.int32, device=dev) for _ in range(2)]
        self.dctx = d

    def run(self, iu, ia, y, lr, profile=None):
        m, B = self.model, self.B
        N = iu.numel()
        steps = (N + B - 1) // B
        t0 = m.iterations
        if t0 + steps > self.t_cap:
            raise _capi.AnimerecError("DistTrainSession sized for %d optimizer steps, %d requested" % (self.t_cap, t0 + steps))
        m._set_alpha(lr, t0 + 1, steps)
        ctx = self._ctx(iu, ia, y)
        st, L = stream_ptr(), lib()
        per_step = {"replay": 6, "dense": 7, "touched": 5}[m.adam_mode]
        for s0 in range(0, steps, self.n_slots):
            ns = min(self.n_slots, steps - s0)
            check(L.ar_plan_build(ptr(iu), N, B, s0, ns, C.byref(self.plan_u), st), "ar_plan_build(users)")
            check(L.ar_plan_build(ptr(ia), N, B, s0, ns, C.byref(self.plan_a), st), "ar_plan_build(anime)")
            for k, keep in enumerate((self._keep_u, self._keep_a)):
                for src, dst in ((keep["uniq"], self.uniq_all[k]), (keep["meta"], self.meta_all[k])):
                    check(L.ar_allgather_bytes(self.comm.handle, ptr(src), ptr(dst), src.numel() * 4, st), "ar_allgather_bytes")
            if m.adam_mode == "replay":
                for k, pl in enumerate((self.plan_u, self.plan_a)):
                    check(L.ar_plan_link(C.byref(pl), ns, ptr(self.uniq_all[k]), ptr(self.meta_all[k]), self.comm.world, st),
                          "ar_plan_link")
            check(L.ar_train_steps_dist(C.byref(ctx), C.byref(self.dctx), s0, 0, t0 + s0, ns, st), "ar_train_steps_dist")
            self.launches += 2 + ns * per_step
        m.iterations = t0 + steps
        return steps


def shard_rows(full, rank, world):
    """Rows of `full` owned by `rank` under the row % world rule, in local-index order (row // world)."""
    return full[rank::world]


class ShardedTrainSession(TrainSession):
    """Row-sharded data-parallel training (BASELINE cfg5): `model` holds THIS rank's shards -- an
    EmbeddingDotModel built with n_users = ceil(n_users_global / world) (same for anime) whose table row i is
    global row i * world + rank -- while the index arrays passed to run() hold GLOBAL row ids.  `batch` is
    the per-rank batch; BatchNorm statistics and the head update are global."""

    def __init__(self, model, batch, total_steps, comm=None):
        super().__init__(model, batch, total_steps)
        from ._capi import ArShardCtx
        self.comm = comm or Comm()
        G, B, D, dev, S = self.comm.world, self.B, model.dim, model.device, self.n_slots
        if G > 8:
            raise _capi.AnimerecError("row-sharded training supports up to 8 ranks")
        f = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        self._sh = dict(
            req_send=[torch.empty((G, S, B), **i32) for _ in range(2)],
            req_recv=[torch.empty((G, S, B), **i32) for _ in range(2)],
            emit_map=[torch.zeros((S, B), **i32) for _ in range(2)],
            cache_idx=[torch.zeros((S, B), **i32) for _ in range(2)],
            rows_out=[torch.zeros((G, B, D), **f) for _ in range(2)],
            rows_in=[torch.zeros((G, B, D), **f) for _ in range(2)],
            grad_send=[torch.zeros((G, B, D + 4), **f) for _ in range(2)],
            grad_recv=[torch.zeros((G, B, D + 4), **f) for _ in range(2)])
        self.max_count = torch.zeros(2, **i32)
        self.c_all, self.label_all, self.dy_all = (torch.empty(G * B, **f) for _ in range(3))
        self.fwd_part_all = torch.zeros(2 * ((G * B + 7) // 8), dtype=torch.float64, device=dev)
        self.head_part_all = torch.zeros(8 * ((G * B + 255) // 256), dtype=torch.float64, device=dev)
        h = ArShardCtx()
        h.comm, h.n_ranks, h.rank = self.comm.handle, G, self.comm.rank
        for name, bufs in self._sh.items():
            arr = getattr(h, name)
            for k in range(2):
                arr[k] = bufs[k].data_ptr()
        h.max_count = self.max_count.data_ptr()
        h.c_all, h.label_all, h.dy_all = self.c_all.data_ptr(), self.label_all.data_ptr(), self.dy_all.data_ptr()
        h.fwd_part_all, h.head_part_all = self.fwd_part_all.data_ptr(), self.head_part_all.data_ptr()
        self.hctx = h
        self.caps = []

    def run(self, iu, ia, y, lr, profile=None):
        m, B = self.model, self.B
        N = iu.numel()
        steps = (N + B - 1) // B
        t0 = m.iterations
        if t0 + steps > self.t_cap:
            raise _capi.AnimerecError("ShardedTrainSession sized for %d optimizer steps, %d requested" % (self.t_cap, t0 + steps))
        m._set_alpha(lr, t0 + 1, steps)
        ctx = self._ctx(iu, ia, y)
        ctx.sched_ws = None
        st, L = stream_ptr(), lib()
        for s0 in range(0, steps, self.n_slots):
            ns = min(self.n_slots, steps - s0)
            check(L.ar_plan_build(ptr(iu), N, B, s0, ns, C.byref(self.plan_u), st), "ar_plan_build(users)")
            check(L.ar_plan_build(ptr(ia), N, B, s0, ns, C.byref(self.plan_a), st), "ar_plan_build(anime)")
            check(L.ar_shard_plan(C.byref(self.plan_u), C.byref(self.plan_a), ns, C.byref(self.hctx), st), "ar_shard_plan")
            # one host read per chunk of steps: the exchange size every rank uses for the chunk (max over ranks)
            mc = self.max_count.max().reshape(1).to(torch.int64)
            dist.all_reduce(mc, op=dist.ReduceOp.MAX)
            cap = min(B, (int(mc.item()) + 3) // 4 * 4)
            self.caps.append(cap)
            check(L.ar_train_steps_sharded(C.byref(ctx), C.byref(self.hctx), s0, 0, t0 + s0, ns, max(cap, 4), st),
                  "ar_train_steps_sharded")
            self.launches += 4 + ns * (11 if m.adam_mode == "replay" else 9)
        m.iterations = t0 + steps
        return steps
