"""Cosine-similarity top-k over embedding tables (Half B), host side.

Mirrors the NumPy code of similar_anime.py:136-171,399-468, similar_users.py:75-101,262-314,
user_recs.py:168-194,453-488 and the scoring of model_recs.py:373-456, with the arithmetic in
libanimerec.so.  Tables are torch CUDA tensors (or anything `as_table` can upload once).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _capi
from ._capi import check, lib, ptr, stream_ptr


def as_table(W, device=None):
    """float32 (n, D) contiguous CUDA tensor view/copy of W."""
    if not torch.cuda.is_available():
        raise _capi.AnimerecError("similarity needs a CUDA device; there is no CPU fallback")
    dev = torch.device(device or "cuda:%d" % torch.cuda.current_device())
    if isinstance(W, torch.Tensor):
        return W.to(device=dev, dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(W, np.float32)).to(dev)


def get_weights(model):
    """(anime_weights, user_weights) row-normalised float32 NumPy arrays -- similar_anime.py:136-171."""
    model._sync_tables()
    return normalize_rows(model.A).cpu().numpy(), normalize_rows(model.U).cpu().numpy()


def normalize_rows(W):
    W = as_table(W)
    out = torch.empty_like(W)
    check(lib().ar_rownorm(ptr(W), W.shape[0], W.shape[1], ptr(out), stream_ptr()), "ar_rownorm")
    return out


def pack_mask(mask, n, device):
    """bool (n,) -> uint32 bit words on device (bit r of word r>>5 set = row r is a candidate)."""
    if mask is None:
        return None
    m = np.asarray(mask, dtype=bool).reshape(-1)
    if m.shape[0] != n:
        raise ValueError("mask length %d != n_rows %d" % (m.shape[0], n))
    pad = (-n) % 32
    bits = np.packbits(np.concatenate([m, np.zeros(pad, bool)]), bitorder="little").view(np.uint32)
    return torch.from_numpy(bits.view(np.int32).copy()).to(device)


def cosine_topk_query(W, q, k, mask=None, exclude=None):
    """Top-k rows by cosine similarity to row q: (idx int32[<=k], score float32[<=k]), best first.

    `mask` restricts the candidates (Type/genre filter of similar_anime.py:438-463), `exclude`
    drops one row (the query itself, similar_anime.py:459)."""
    W = as_table(W)
    n, D = W.shape
    q = int(q)
    if not 0 <= q < n:
        raise KeyError("query row %d outside [0, %d)" % (q, n))
    if k <= 0:
        raise ValueError("k must be positive")
    if k > _capi.MAX_K:
        # the kernel ranks up to MAX_K rows per pass: take the best MAX_K, strike them from the candidates, repeat
        # (the reference returns Frame[:count] for any count, similar_anime.py:468)
        cand = np.ones(n, bool) if mask is None else np.asarray(mask, bool).copy()
        if exclude is not None:
            cand[int(exclude)] = False
        out_i, out_s = [], []
        while sum(len(x) for x in out_i) < k and cand.any():
            i, sc = cosine_topk_query(W, q, min(_capi.MAX_K, k - sum(len(x) for x in out_i)), mask=cand)
            if len(i) == 0:
                break
            cand[i] = False
            out_i.append(i)
            out_s.append(sc)
        if not out_i:
            return np.zeros(0, np.int32), np.zeros(0, np.float32)
        return np.concatenate(out_i), np.concatenate(out_s)
    L = lib()
    ws = torch.empty(max(8, L.ar_topk_query_workspace(n, k)), dtype=torch.uint8, device=W.device)
    oi = torch.empty(k, dtype=torch.int32, device=W.device)
    os_ = torch.empty(k, dtype=torch.float32, device=W.device)
    mbits = pack_mask(mask, n, W.device)
    check(L.ar_cosine_topk_query(ptr(W), n, D, q, ptr(mbits), -1 if exclude is None else int(exclude), k,
                                 ptr(oi), ptr(os_), ptr(ws), stream_ptr()), "ar_cosine_topk_query")
    oi, os_ = oi.cpu().numpy(), os_.cpu().numpy()
    keep = oi >= 0
    return oi[keep], os_[keep]


def cosine_topk_query_device(W, q, k, mask_bits=None, exclude=-1):
    """Same kernel, device tensors in and out (idx int32[k] with -1 padding, score float32[k]); no host copy."""
    n, D = W.shape
    L = lib()
    ws = torch.empty(max(8, L.ar_topk_query_workspace(n, k)), dtype=torch.uint8, device=W.device)
    oi = torch.empty(k, dtype=torch.int32, device=W.device)
    os_ = torch.empty(k, dtype=torch.float32, device=W.device)
    check(L.ar_cosine_topk_query(ptr(W), n, D, int(q), ptr(mask_bits), int(exclude), k, ptr(oi), ptr(os_), ptr(ws),
                                 stream_ptr()), "ar_cosine_topk_query")
    return oi, os_


def find_similar_users(W_users, q, n_users):
    """similar_users.py:293-312 / user_recs.py:475-488: top-(n+1) of ALL rows, then drop the query."""
    idx, sc = cosine_topk_query(W_users, q, n_users + 1)
    keep = idx != q
    return idx[keep], sc[keep]


def similar_anime(W_anime, q, count, mask=None):
    """similar_anime.py:404-468: rank all anime, keep the Type/genre candidates, drop the query."""
    return cosine_topk_query(W_anime, q, count, mask=mask, exclude=q)


# |bf16-operand score - fp32 cosine| <= resid_q + resid_c + resid_q*resid_c (rounding residual norms of the two
# unit rows, ar_rownorm_bf16) + the tensor core's fp32 accumulation error over dim terms of magnitude <= 1.
ACCUM_EPS = 3e-5
BF16_SCORE_EPS = 0.004   # blanket bound when no residuals are at hand: 2 * 2^-9 (+ accumulation slack)


def normalize_rows_bf16(W, with_resid=False):
    """Row-normalised bf16 copy of a table: the operand format of the tensor-core pass.
    with_resid -> (table, resid [n] float32 rounding-residual norms)."""
    W = as_table(W)
    out = torch.empty(W.shape, dtype=torch.bfloat16, device=W.device)
    res = torch.empty(W.shape[0], dtype=torch.float32, device=W.device) if with_resid else None
    check(lib().ar_rownorm_bf16(ptr(W), W.shape[0], W.shape[1], ptr(out), ptr(res), stream_ptr()), "ar_rownorm_bf16")
    return (out, res) if with_resid else out


class CandidateLists:
    """Output of the tensor-core pass: idx/score [n_chunks, nq, cap], cnt/thr [n_chunks, nq] (device)."""
    __slots__ = ("idx", "score", "cnt", "thr", "dump")

    def __init__(self, idx, score, cnt, thr, dump=None):
        self.idx, self.score, self.cnt, self.thr, self.dump = idx, score, cnt, thr, dump


def allpairs_candidates(Qn, q0, nq, Cn, c0, nc, kprime=16, exclude_self=False, watched=None, dump=False,
                        self_ids=None, thr_init=None, n_chunks=None):
    """Tensor-core candidate pass (ar_cosine_topk_allpairs): per query row and candidate chunk every
    candidate whose bf16 score beats a running threshold that rises to the kprime-th best seen."""
    L = lib()
    n_chunks = int(L.ar_allpairs_chunks(nq, nc)) if n_chunks is None else int(n_chunks)
    cap = int(L.ar_allpairs_list_cap(kprime))
    if cap <= 0:
        raise ValueError("kprime must be 16 or 24")
    dev = Qn.device
    oi = torch.empty((n_chunks, nq, cap), dtype=torch.int32, device=dev)
    os_ = torch.empty((n_chunks, nq, cap), dtype=torch.float32, device=dev)
    oc = torch.zeros((n_chunks, nq), dtype=torch.int32, device=dev)
    ot = torch.full((n_chunks, nq), float("-inf"), dtype=torch.float32, device=dev)
    dm = torch.zeros((nq, nc), dtype=torch.float32, device=dev) if dump else None
    stride = 0 if watched is None else watched.shape[1]
    check(L.ar_cosine_topk_allpairs(ptr(Qn), Qn.shape[0], q0, nq, ptr(Cn), Cn.shape[0], c0, nc, Qn.shape[1], kprime,
                                    1 if exclude_self else 0, ptr(self_ids), ptr(watched), stride, ptr(thr_init),
                                    n_chunks, ptr(oi), ptr(os_), ptr(oc), ptr(ot), ptr(dm), stream_ptr()),
          "ar_cosine_topk_allpairs")
    return CandidateLists(oi, os_, oc, ot, dm)


def rerank(Wq, q0, nq, Wc, cand, k, cand_cnt=None, cand_thr=None, eps=BF16_SCORE_EPS, q_eps=None):
    """Exact fp32 cosine re-rank of candidate lists cand [n_lists, nq, cap] (or [nq, cap]).
    -> (idx [nq,k], score [nq,k], certified [nq] uint8 or None), device tensors."""
    if cand.dim() == 2:
        cand = cand.unsqueeze(0)
    nl, _, cap = cand.shape
    oi = torch.empty((nq, k), dtype=torch.int32, device=Wq.device)
    os_ = torch.empty((nq, k), dtype=torch.float32, device=Wq.device)
    cert = torch.empty(nq, dtype=torch.uint8, device=Wq.device) if cand_thr is not None else None
    check(lib().ar_cosine_rerank(ptr(Wq), q0, nq, ptr(Wc), Wq.shape[1], ptr(cand.contiguous()),
                                 ptr(None if cand_cnt is None else cand_cnt.contiguous()),
                                 ptr(None if cand_thr is None else cand_thr.contiguous()), nl, cap, k, float(eps),
                                 ptr(q_eps), ptr(oi), ptr(os_), ptr(cert), stream_ptr()), "ar_cosine_rerank")
    return oi, os_, cert


TENSOR_DIM = 128            # embedding size the tcgen05 candidate kernel is built for (config.yaml:63 default)
N_SMS = 148                 # B200
QTILE = 256                 # query rows per work item of the tensor-core kernel
SAMPLE_FRACTION = 16        # the threshold-seeding pass scans 1/16 of the candidates
SAMPLE_MIN_ROWS = 65536     # ... and is skipped for candidate sets smaller than this


class _Timer:
    """CUDA-event stage timer, active only when the caller passes stats={'time': True}."""

    def __init__(self, on):
        self.on, self.ev = on, []

    def mark(self, name):
        if self.on:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.ev.append((name, e))

    def result(self):
        if not self.on or len(self.ev) < 2:
            return {}
        torch.cuda.synchronize()
        out = {}
        for (_, a), (nb, b) in zip(self.ev[:-1], self.ev[1:]):
            out[nb] = out.get(nb, 0.0) + a.elapsed_time(b)
        return out


def _sl(t, lo, hi):
    return None if t is None else t[lo:hi]


def _candidate_lists(Qn, nq, Cn, kprime, exclude_self, self_ids, watched, thr_init, seed=False, tm=None,
                     c_range=None):
    """Tensor-core candidate pass over query rows [0, nq) of Qn against all of Cn, scheduled for the machine:
      1. (seed=True, off by default) threshold seeding: a first pass over 1/16 of the candidates yields per row
         a valid lower bound on its (kprime+1)-th best score.  Measured on 350k x 350k it does not pay: the
         slow-path events are dominated by the steady-state rate 1024*kprime/m per warp-chunk, not by the
         warm-up, and the seeding pass itself runs entirely in warm-up mode (8.4 ms to save 5 ms);
      2. the query range is cut into a part whose 256-row tiles fill the 148 SMs an integral number of rounds
         (one candidate chunk) and a remainder whose candidate range is split so it also fills the machine.
    -> [(row_lo, row_hi, CandidateLists)]"""
    c0, c1 = c_range or (0, Cn.shape[0])
    nc = c1 - c0
    if seed and thr_init is None and nc >= SAMPLE_MIN_ROWS:
        ns = ((nc // SAMPLE_FRACTION + 127) // 128) * 128
        cl0 = allpairs_candidates(Qn, 0, nq, Cn, c0, ns, kprime, exclude_self=exclude_self, self_ids=self_ids,
                                  watched=watched, n_chunks=1)
        thr_init = cl0.thr[0]
        if tm:
            tm.mark("seed_pass")
    qtiles = (nq + QTILE - 1) // QTILE
    rounds = qtiles // N_SMS
    cuts = [0, nq]
    if rounds >= 2 and qtiles % N_SMS:
        cuts = [0, rounds * N_SMS * QTILE, nq]
    parts = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        cl = allpairs_candidates(Qn, lo, hi - lo, Cn, c0, nc, kprime, exclude_self=exclude_self,
                                 self_ids=_sl(self_ids, lo, hi), watched=_sl(watched, lo, hi),
                                 thr_init=_sl(thr_init, lo, hi))
        parts.append((lo, hi, cl))
    if tm:
        tm.mark("candidate_pass")
    return parts


def _rerank_parts(parts, Wq_f32, Wc_f32, k, q_eps, tm=None):
    nq = Wq_f32.shape[0]
    oi = torch.empty((nq, k), dtype=torch.int32, device=Wq_f32.device)
    os_ = torch.empty((nq, k), dtype=torch.float32, device=Wq_f32.device)
    cert = torch.empty(nq, dtype=torch.uint8, device=Wq_f32.device)
    L = lib()
    for lo, hi, cl in parts:
        nl, _, cap = cl.idx.shape
        check(L.ar_cosine_rerank(ptr(Wq_f32), lo, hi - lo, ptr(Wc_f32), Wq_f32.shape[1], ptr(cl.idx), ptr(cl.cnt),
                                 ptr(cl.thr), nl, cap, k, float(ACCUM_EPS), ptr(q_eps[lo:hi]), ptr(oi[lo:hi]),
                                 ptr(os_[lo:hi]), ptr(cert[lo:hi]), stream_ptr()), "ar_cosine_rerank")
    if tm:
        tm.mark("rerank")
    return oi, os_, cert


def _certified_topk(Wq_f32, Qn, q_res, Wc_f32, Cn, c_res_max, k, kprime, exclude_self, self_ids, watched, stats,
                    exact_row, c_range=None):
    """Tensor-core pass + fp32 re-rank + certification, with two recovery levels for the rows the first
    pass cannot certify: (1) the same kernel over just those rows with the provably sufficient fixed
    threshold (k-th fp32 score found so far - eps); (2) `exact_row(i)`, the fp32 single-query kernel.
    Qn may hold more rows than Wq_f32 (the whole table); the queries are its first Wq_f32.shape[0] rows
    unless self_ids says otherwise."""
    nq = Wq_f32.shape[0]
    if kprime <= k:
        raise ValueError("kprime (%d) must exceed k (%d)" % (kprime, k))
    tm = _Timer(bool(stats is not None and stats.get("time")))
    tm.mark("start")
    q_eps = (q_res + c_res_max + q_res * c_res_max).contiguous()
    parts = _candidate_lists(Qn, nq, Cn, kprime, exclude_self, self_ids, watched, None, tm=tm, c_range=c_range)
    oi, os_, cert = _rerank_parts(parts, Wq_f32, Wc_f32, k, q_eps, tm)
    bad = torch.nonzero(cert == 0).reshape(-1)
    n_bad1 = int(bad.numel())
    n_bad2 = 0
    if n_bad1:
        # every true top-k member has fp32 score >= the k-th fp32 score found so far, hence bf16 score above
        # that minus eps: list EVERYTHING above this fixed threshold (no compaction unless > cap such rows)
        kth = os_[bad, k - 1]
        thr0 = torch.where(torch.isfinite(kth), kth - (q_eps[bad] + ACCUM_EPS) * 1.0001 - 1e-6,
                           torch.full_like(kth, float("-inf"))).contiguous()
        Qb = Qn[bad].contiguous()
        Wb = Wq_f32[bad].contiguous()
        ids_b = None
        if exclude_self:
            ids_b = (self_ids[bad] if self_ids is not None else bad.to(torch.int32)).contiguous()
        wb = watched[bad].contiguous() if watched is not None else None
        parts2 = _candidate_lists(Qb, n_bad1, Cn, kprime, exclude_self, ids_b, wb, thr0, tm=None, c_range=c_range)
        oi2, os2, cert2 = _rerank_parts(parts2, Wb, Wc_f32, k, q_eps[bad].contiguous())
        oi[bad], os_[bad] = oi2, os2
        still = bad[cert2 == 0]
        n_bad2 = int(still.numel())
        for r in still.cpu().tolist():                 # exact fp32 path for the (near-)tied few
            fi, fs = exact_row(r)
            oi[r], os_[r] = fi, fs
        tm.mark("recovery")
    if stats is not None:
        cl = parts[0][2]
        stats.update(uncertified=n_bad1, uncertified_after_retry=n_bad2, n_chunks=[p[2].idx.shape[0] for p in parts],
                     kprime=kprime, mean_list=float(cl.cnt.float().mean().item()), ms=tm.result())
    return oi, os_


def allpairs_topk(W, k=10, kprime=16, q0=0, nq=None, stats=None):
    """BASELINE cfg3: for every query row in [q0, q0+nq) the k most cosine-similar OTHER rows of W.

    The bf16 tensor-core pass lists the candidates, the fp32 re-rank orders them exactly, and rows whose
    result cannot be certified exact (rigorous bf16 error bound) are retried / fall back to the fp32
    single-query kernel.  -> (idx [nq,k] int32, score [nq,k] float32) device tensors."""
    W = as_table(W)
    n = W.shape[0]
    nq = n - q0 if nq is None else nq
    if W.shape[1] < TENSOR_DIM:                               # smaller embeddings: zero columns change no cosine
        W = torch.nn.functional.pad(W, (0, TENSOR_DIM - W.shape[1])).contiguous()
    if W.shape[1] > TENSOR_DIM:                               # larger ones: exact fp32 queries, one C call for all rows
        L = lib()
        oi = torch.empty((nq, k), dtype=torch.int32, device=W.device)
        os_ = torch.empty((nq, k), dtype=torch.float32, device=W.device)
        ws = torch.empty(max(8, L.ar_topk_query_workspace(n, k)), dtype=torch.uint8, device=W.device)
        check(L.ar_cosine_topk_queries(ptr(W), n, W.shape[1], q0, nq, None, k, ptr(oi), ptr(os_), ptr(ws), stream_ptr()),
              "ar_cosine_topk_queries")
        return oi, os_
    Wn, res = normalize_rows_bf16(W, with_resid=True)
    res = torch.nan_to_num(res, nan=1.0)
    whole = q0 == 0                                      # queries are the leading rows: identity self ids
    Qn = Wn if whole else Wn[q0:q0 + nq].contiguous()
    Wq = W[:nq] if whole else W[q0:q0 + nq].contiguous()
    ids = None if whole else torch.arange(q0, q0 + nq, dtype=torch.int32, device=W.device)

    def exact_row(r):
        return cosine_topk_query_device(W, q0 + r, k, exclude=q0 + r)

    return _certified_topk(Wq, Qn, res[q0:q0 + nq], W, Wn, float(res.max().item()), k, kprime, True, ids, None,
                           stats, exact_row)


def watched_bits(indptr, idx, n_rows, n_cols, device, cand_mask=None):
    """[n_rows, ceil(n_cols/32)] uint32 rows (as int32 tensor): bit set = candidate is dropped."""
    words = (n_cols + 31) // 32
    if cand_mask is None:
        out = torch.zeros((n_rows, words), dtype=torch.int32, device=device)
    else:
        inv = pack_mask(~np.asarray(cand_mask, dtype=bool), n_cols, device)
        out = inv.reshape(1, words).repeat(n_rows, 1).contiguous()
    ip = torch.as_tensor(np.asarray(indptr, np.int64)).to(device)
    ix = torch.as_tensor(np.asarray(idx, np.int32)).to(device)
    check(lib().ar_bits_from_csr(ptr(ip), ptr(ix), n_rows, words, n_cols, ptr(out), stream_ptr()), "ar_bits_from_csr")
    return out


def score_topk(model, users, watched_indptr, watched_idx, k, cand_mask=None, kprime=24, stats=None):
    """model_recs over many users (BASELINE cfg4): predicted rating of every anime the user has NOT rated,
    top-k by Prediction (model_recs.py:132-192,373-456).  Prediction is a monotone map of cos(u, a)
    (Dense(1) -> BatchNorm(inference) -> sigmoid), so candidates are ranked by sign(w*gamma)*cos on the
    tensor cores, re-ranked in fp32, and only the winners go through the exact forward (ar_predict).
    -> (idx [n,k] int32 numpy, prediction [n,k] float32 numpy); -1 / -inf pad short lists."""
    model._sync_tables()
    dev = model.device
    users_t = torch.as_tensor(np.asarray(users, np.int64)).to(dev)
    nq, na = users_t.numel(), model.n_anime
    head = model.head.cpu().numpy()
    sign = -1.0 if float(head[0]) * float(head[2]) < 0 else 1.0
    Uq = (model.U[users_t] * sign).contiguous()               # query rows (negated when the map decreases)
    wb = watched_bits(watched_indptr, watched_idx, nq, na, dev, cand_mask)

    A_t = model.A
    if model.dim < TENSOR_DIM:                                # zero columns change no cosine: pad to the tensor-core width
        Uq = torch.nn.functional.pad(Uq, (0, TENSOR_DIM - model.dim)).contiguous()
        A_t = torch.nn.functional.pad(model.A, (0, TENSOR_DIM - model.dim)).contiguous()
    aug = []                                                  # the anime table + one scratch row, built on first use

    def exact_row(r):
        if not aug:
            aug.append(torch.cat([A_t, torch.zeros((1, A_t.shape[1]), dtype=A_t.dtype, device=dev)], dim=0).contiguous())
        return _query_vs_table(Uq[r], aug[0], k, wb[r])

    if model.dim <= TENSOR_DIM:
        (Qn, qres), (Cn, cres) = normalize_rows_bf16(Uq, True), normalize_rows_bf16(A_t, True)
        qres, cres = torch.nan_to_num(qres, nan=1.0), torch.nan_to_num(cres, nan=1.0)
        oi, os_ = _certified_topk(Uq, Qn, qres, A_t, Cn, float(cres.max().item()), k, kprime, False, None, wb,
                                  stats, exact_row)
    else:                                                     # larger embeddings: exact fp32 GEMV + top-k per user
        oi = torch.empty((nq, k), dtype=torch.int32, device=dev)
        os_ = torch.empty((nq, k), dtype=torch.float32, device=dev)
        for r in range(nq):
            oi[r], os_[r] = exact_row(r)
    valid = oi >= 0
    iu = users_t.to(torch.int32).reshape(-1, 1).expand(-1, k)[valid].contiguous()
    ia = oi[valid].contiguous()
    pred = torch.full((nq, k), float("-inf"), dtype=torch.float32, device=dev)
    if iu.numel():
        out = torch.empty(iu.numel(), dtype=torch.float32, device=dev)
        check(lib().ar_predict(ptr(model.U), ptr(model.A), model.dim, ptr(model.head), ptr(model.bn_moving), ptr(iu),
                               ptr(ia), iu.numel(), ptr(out), stream_ptr()), "ar_predict")
        pred[valid] = out
    if stats is not None:
        stats["sign"] = sign
    return oi.cpu().numpy(), pred.cpu().numpy()


def _query_vs_table(qrow, T, k, drop_bits):
    """fp32 fallback for one scoring row.  T is the candidate table with one scratch row appended (built once by the
    caller): the query is written there and ranked against the rest by the single-query kernel."""
    n = T.shape[0] - 1
    T[n].copy_(qrow)
    words = (n + 1 + 31) // 32
    keep = torch.zeros(words, dtype=torch.int32, device=T.device)
    keep[:drop_bits.numel()] = ~drop_bits
    return cosine_topk_query_device(T, n, k, mask_bits=keep, exclude=n)


def topk_merge(idx, score, k_out, lists_sorted=True):
    """Merge [n_lists, n_queries, k_in] partial lists (global row ids) into [n_queries, k_out].
    lists_sorted: every list is best-first (what the top-k entry points return) -> bound pruning."""
    nl, nq, kin = idx.shape
    oi = torch.empty((nq, k_out), dtype=torch.int32, device=idx.device)
    os_ = torch.empty((nq, k_out), dtype=torch.float32, device=idx.device)
    check(lib().ar_topk_merge(ptr(idx), ptr(score), nl, nq, kin, k_out, 1 if lists_sorted else 0, ptr(oi), ptr(os_), stream_ptr()),
          "ar_topk_merge")
    return oi, os_
