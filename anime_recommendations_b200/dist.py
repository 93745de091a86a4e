"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for rendezvous, NCCL collectives
issued from libanimerec on the kernels' stream.

* DistTrainSession  -- replicated-table data-parallel training (BASELINE cfg2): every rank trains on its
  own shard of each global batch; SyncBN + all-gathered, merged row gradients keep replicas bit-identical.
* Comm              -- the library's own NCCL communicator (ncclCommInitRank with an id broadcast through
  torch.distributed), also used by the candidate-sharded top-k.
"""
from __future__ import annotations

import ctypes as C
import os
import time

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


def peer_plan_cap(batch, world):
    """Capacity of one per-step selection list in peer mode: a rank lists the samples of the GLOBAL batch that
    touch its rows -- `batch` on average.  Twice that plus slack, bounded by the global batch and by what the
    plan sort holds; a chunk that overflows is refused (use the NCCL row-sharded session then)."""
    return min(_capi.AR_MAX_BATCH, batch * world, 2 * batch + 512)


class PeerTrainSession(TrainSession):
    """Row-sharded data-parallel training over NVLink peer memory (csrc/peer.inl): same model / index contract
    as ShardedTrainSession (`model` holds THIS rank's shards, index arrays hold GLOBAL row ids, `batch` is the
    per-rank batch), but no collective on the step's critical path: every rank processes the samples of the
    global batch that touch its rows and reads the other table's rows from the owners' HBM.  One node, up to 8
    ranks, all-to-all peer access (NVSwitch)."""

    def __init__(self, model, batch, total_steps, comm=None):
        from ._capi import ArPeerCtx, PEER_FLAG_WORDS, PEER_HANDLE_BYTES
        if not dist.is_initialized():
            raise _capi.AnimerecError("torch.distributed must be initialised (torchrun) before PeerTrainSession()")
        G, rank = dist.get_world_size(), dist.get_rank()
        if G > _capi.PEER_MAX_RANKS:
            raise _capi.AnimerecError("peer-memory training supports up to %d ranks" % _capi.PEER_MAX_RANKS)
        cap = peer_plan_cap(int(batch), G)
        super().__init__(model, batch, total_steps, plan_cap=cap)
        # replay mode: ONE persistent kernel per chunk (csrc/chunk.inl over peer memory); AR_PEER_STAGED=1 and the
        # other Adam modes chain the per-step stage kernels of csrc/peer.inl
        self.persistent = model.adam_mode == "replay" and os.environ.get("AR_PEER_STAGED") is None
        self.comm = comm or Comm()
        B, D, dev, S = self.B, model.dim, model.device, self.n_slots
        f = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        # everything the other ranks read or write lives in ONE allocation:
        # [U | A | published list (staged path) | flags | row words U | row words A | pair inbox | header inbox]
        nU, nA = model.U.numel(), model.A.numel()
        up4 = lambda n: (n + 3) // 4 * 4       # noqa: E731  (16-byte aligned regions)
        nFU, nFA = up4(model.lastU.numel()), up4(model.lastA.numel())
        nPairs = 2 * G * cap * 2               # [2 parities][G senders][cap] 64-bit words, in 32-bit units
        nHdr = 2 * G * 8 * 2                   # [2][G][8] 64-bit words
        words = nU + nA + 2 * cap + PEER_FLAG_WORDS + nFU + nFA + nPairs + nHdr
        self.arena = torch.zeros(words, **f)
        o = 0
        U = self.arena[o:o + nU].view_as(model.U); o += nU
        A = self.arena[o:o + nA].view_as(model.A); o += nA
        self.pub = self.arena[o:o + 2 * cap]; o += 2 * cap
        self.flags = self.arena[o:o + PEER_FLAG_WORDS].view(torch.int32); o += PEER_FLAG_WORDS
        # per row: the optimizer step the row is at, as the PEERS see it (written right behind the local flag, or behind
        # a system-scope fence with AR_PEER_STRICT=1; model.lastU / lastA stay the local flags).  Refreshed from the
        # local flags at the start of every run().
        self.rowflag = [self.arena[o:o + model.lastU.numel()].view(torch.int32),
                        self.arena[o + nFU:o + nFU + model.lastA.numel()].view(torch.int32)]
        o += nFU + nFA
        self.pair_inbox = self.arena[o:o + nPairs].view(torch.int32); o += nPairs
        self.pair_inbox.fill_(-1)              # every slot invalid
        self.hdr_inbox = self.arena[o:o + nHdr]; o += nHdr
        U.copy_(model.U)
        A.copy_(model.A)
        model.U, model.A = U, A                # the model's tables now live in the exported arena
        L = lib()
        handle = (C.c_ubyte * PEER_HANDLE_BYTES)()
        off = C.c_int64()
        check(L.ar_peer_export(ptr(self.arena), handle, C.byref(off)), "ar_peer_export")
        mine = torch.tensor(list(bytes(handle)) + [int(b) for b in int(off.value).to_bytes(8, "little")],
                            dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        every = self.comm.allgather(mine).cpu().numpy()
        h = ArPeerCtx()
        h.n_ranks, h.rank, h.sel_cap = G, rank, cap
        self.peer_base = []
        for r in range(G):
            if r == rank:
                base = self.arena.data_ptr()
            else:
                hb = (C.c_ubyte * PEER_HANDLE_BYTES)(*every[r, :PEER_HANDLE_BYTES].tolist())
                roff = int.from_bytes(bytes(every[r, PEER_HANDLE_BYTES:].tolist()), "little")
                p = C.c_void_p()
                check(L.ar_peer_open(hb, roff, C.byref(p)), "ar_peer_open(rank %d)" % r)
                base = p.value
            self.peer_base.append(base)
        # shard sizes differ by at most one row between ranks, so every rank reports its own layout
        lay = torch.tensor([nU, nA, nFU, nFA], dtype=torch.int64, device=dev)
        lays = self.comm.allgather(lay).cpu().tolist()
        for r in range(G):
            base, (rU, rA, rFU, rFA) = self.peer_base[r], lays[r]
            o = rU + rA
            h.W_peer[0][r] = base
            h.W_peer[1][r] = base + 4 * rU
            h.pub_peer[r] = base + 4 * o; o += 2 * cap
            h.flags_peer[r] = base + 4 * o; o += PEER_FLAG_WORDS
            h.rowflag_peer[0][r] = base + 4 * o
            h.rowflag_peer[1][r] = base + 4 * (o + rFU); o += rFU + rFA
            h.pairs_peer[r] = base + 4 * o; o += nPairs
            h.hdrin_peer[r] = base + 4 * o
        self.c_all = torch.zeros(G * B, **f)
        self.dy_all = torch.empty(G * B, **f)
        self.fwd_part_all = torch.zeros(2 * G * ((cap + 1023) // 1024), dtype=torch.float64, device=dev)
        self.head_part_all = torch.zeros(8 * ((G * B + 255) // 256), dtype=torch.float64, device=dev)
        h.c_all, h.dy_all = self.c_all.data_ptr(), self.dy_all.data_ptr()
        h.fwd_part_all, h.head_part_all = self.fwd_part_all.data_ptr(), self.head_part_all.data_ptr()
        # Two sets of per-chunk state (all-gathered samples, selection lists, plans): chunk i+1 is planned on a
        # side stream while the steps of chunk i run, so the all-gathers and the plan sorts stay off the steps'
        # critical path (in-stream they cost ~2 ms per 256-step chunk at 8 GPUs, ~7 us per step).
        self.sets = []
        for k in range(2):
            st = dict(plan_u=self.plan_u, plan_a=self.plan_a, keep=(self._keep_u, self._keep_a)) if k == 0 else None
            if st is None:
                pu, ku = self._make_plan(S, cap, dev)
                pa, ka = self._make_plan(S, cap, dev)
                st = dict(plan_u=pu, plan_a=pa, keep=(ku, ka))
            st["sel"] = dict(sel_key=[torch.zeros((S, cap), **i32) for _ in range(2)],
                             sel_samp=[torch.zeros((S, cap), **i32) for _ in range(2)],
                             sel_oth=[torch.zeros((S, cap), **i32) for _ in range(2)],
                             sel_cnt=[torch.zeros(S, **i32) for _ in range(2)],
                             sel_lab=[torch.zeros((S, cap), **f) for _ in range(2)])
            st["sched"], st["keep_sched"] = self._make_sched(S, cap, dev)   # replay schedule of the persistent kernel
            st["max_count"] = torch.zeros(2, **i32)
            st["label_step"] = torch.zeros((S, G * B), **f)
            st["gather"] = [torch.empty((G, S * B), **i32), torch.empty((G, S * B), **i32), torch.empty((G, S * B), **f)]
            hk = ArPeerCtx.from_buffer_copy(h)
            for name, bufs in st["sel"].items():
                arr = getattr(hk, name)
                for t in range(2):
                    arr[t] = bufs[t].data_ptr()
            hk.max_count, hk.label_step = st["max_count"].data_ptr(), st["label_step"].data_ptr()
            st["pctx"] = hk
            st["planned"], st["consumed"] = torch.cuda.Event(), torch.cuda.Event()
            self.sets.append(st)
        self.pctx = self.sets[0]["pctx"]
        self.plan_stream = torch.cuda.Stream(device=dev)
        self.counts = []
        self._max_seen = torch.zeros(2, **i32)             # longest selection list since the last verify()
        # nobody may signal a flag before every rank has zeroed and mapped its arena
        torch.cuda.synchronize()
        dist.barrier()

    def check_flags(self):
        """Raise if a flag barrier timed out (a rank died or fell out of step)."""
        bad = int(self.flags[_capi.PEER_ERR_WORD].item())
        if bad:
            raise _capi.AnimerecError("peer barrier %d timed out on rank %d" % (bad, self.pctx.rank))

    def close(self):
        """Unmap the other ranks' arenas (collective: every rank must have finished with them)."""
        torch.cuda.synchronize()
        dist.barrier()
        check(lib().ar_peer_close_all(), "ar_peer_close_all")

    def _plan_chunk(self, st, iu, ia, y, s0, ns, t0):
        """Queue the planning of steps [s0, s0+ns) (optimizer steps t0+s0+1 ..) into set `st` on the CURRENT stream."""
        B, S, L, sp = self.B, self.n_slots, lib(), stream_ptr()
        N = iu.numel()
        lo, hi = s0 * B, min(N, (s0 + ns) * B)
        for src, dst in zip((iu, ia, y), st["gather"]):
            # rank r's slice lands at dst[r, :hi-lo]: all-gather straight into the buffer when the chunk is full
            if hi - lo == S * B:
                check(L.ar_allgather_bytes(self.comm.handle, ptr(src[lo:hi]), ptr(dst), (hi - lo) * 4, sp), "ar_allgather_bytes")
            else:
                dst[:, :hi - lo].copy_(self.comm.allgather(src[lo:hi]))
        g = st["gather"]
        check(L.ar_peer_plan(ptr(g[0]), ptr(g[1]), ptr(g[2]), S * B, hi - lo, B, ns, C.byref(st["plan_u"]),
                             C.byref(st["plan_a"]), C.byref(st["pctx"]), sp), "ar_peer_plan")
        m = self.model
        if self.persistent:
            # the replay schedule over MY rows (local ids), exactly as on one GPU
            check(L.ar_plan_sched(C.byref(st["plan_u"]), C.byref(st["plan_a"]), ns, t0 + s0, m._t_flush,
                                  ptr(m.seenU), m.n_users, ptr(m.seenA), m.n_anime, self.depth, m.dim,
                                  C.byref(st["sched"]), sp), "ar_plan_sched")
        # no host round trip per chunk: the grids are sized for the list capacity, and the longest list of every
        # chunk is checked by verify() (an overflowing list is truncated on the device, nothing is corrupted)
        torch.maximum(self._max_seen, st["max_count"], out=self._max_seen)

    def run(self, iu, ia, y, lr, profile=None, verify=True):
        """Queue the steps; with verify=True (default) finish with verify() -- a synchronising host check that no
        selection list overflowed and no flag barrier timed out -- so that wrong results cannot leave this call
        unnoticed.  verify=False keeps the call asynchronous; the caller then owes a verify() before using the
        tables or the metrics."""
        m, B = self.model, self.B
        N = iu.numel()
        steps = (N + B - 1) // B
        t0 = m.iterations
        if t0 + steps > self.t_cap:
            raise _capi.AnimerecError("PeerTrainSession sized for %d optimizer steps, %d requested" % (self.t_cap, t0 + steps))
        m._set_alpha(lr, t0 + 1, steps)
        ctx = self._ctx(iu, ia, y)
        if self.persistent:
            # what happened to the rows outside this session (a flush, restored weights) reaches the peers' view here;
            # no rank polls these words between runs, and the values only grow
            self.rowflag[0].copy_(m.lastU)
            self.rowflag[1].copy_(m.lastA)
        # plan-order catch-up: with the rows sharded a row's replay is ~G times shorter than on one GPU, and the
        # longest-first schedule's two extra launches per step cost more than its balance gains (2 GPUs: 72.0 vs
        # 78.5 us/step)
        ctx.sched_ws = None
        main, L, S = torch.cuda.current_stream(), lib(), self.n_slots
        per_step = {"replay": 5, "dense": 6, "touched": 4}[m.adam_mode]
        chunks = [(s0, min(S, steps - s0)) for s0 in range(0, steps, S)]
        self.plan_stream.wait_stream(main)                 # the inputs (H2D copies) are queued on `main`
        with torch.cuda.stream(self.plan_stream):
            self._plan_chunk(self.sets[0], iu, ia, y, *chunks[0], t0)
            self.sets[0]["planned"].record()
        for i, (s0, ns) in enumerate(chunks):
            st = self.sets[i % 2]
            if i + 1 < len(chunks):                        # plan the next chunk while this one runs
                nxt = self.sets[(i + 1) % 2]
                with torch.cuda.stream(self.plan_stream):
                    if i >= 1:
                        self.plan_stream.wait_event(nxt["consumed"])   # chunk i-1 no longer reads that set
                    self._plan_chunk(nxt, iu, ia, y, *chunks[i + 1], t0)
                    nxt["planned"].record()
            main.wait_event(st["planned"])
            ctx.plan_u, ctx.plan_a = st["plan_u"], st["plan_a"]
            if self.persistent:
                ctx.sched = st["sched"]
            tq = time.perf_counter()
            check(L.ar_train_steps_peer(C.byref(ctx), C.byref(st["pctx"]), s0, 0, t0 + s0, ns, self.P, stream_ptr(main)),
                  "ar_train_steps_peer")
            self.enqueue_s += time.perf_counter() - tq
            st["consumed"].record(main)
            # per chunk: select, 2 plan sorts, 2 plan links (+ 5 schedule kernels and ONE step kernel when the
            # persistent kernel runs); else per step: forward, pull, head, row update and (replay) catch-up or
            # (dense) two table flushes
            self.launches += 5 + (6 if self.persistent else ns * per_step)
            self._last_ns = ns
        main.wait_stream(self.plan_stream)                 # nothing of this call is left on the side stream
        m.iterations = t0 + steps
        if verify:
            self.verify()
        return steps

    def verify(self):
        """Host check of everything run() queued so far (synchronises): list overflow and barrier time-outs.
        Called by the owner of the session at its own sync points (end of an epoch, before reading results)."""
        mc = self._max_seen.max().reshape(1).to(torch.int64)
        self._max_seen.zero_()
        dist.all_reduce(mc, op=dist.ReduceOp.MAX)
        mc = int(mc.item())
        self.counts.append(mc)
        if mc > self.P:
            raise _capi.AnimerecError("peer mode: %d samples of one step touch one rank's rows, capacity %d; the "
                                      "results of this run are invalid -- use ShardedTrainSession for this data" % (mc, self.P))
        self.check_flags()
        self.check_health()        # the persistent kernel's own time-outs and schedule checks
