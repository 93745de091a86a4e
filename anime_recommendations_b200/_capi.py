"""ctypes binding of libanimerec.so (include/animerec.h).

The product path has no CPU fallback: if the library is missing or a call fails this module
raises.  PyTorch is only used by callers to own device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ANIMEREC_LIB") or os.path.join(PKG, "lib", "libanimerec.so")   # override: A/B builds
ABI_VERSION = 15

AR_MAX_BATCH = 16384
AR_HEAVY_LEN = 64
AR_SCHED_MAX_DEPTH = 4
AR_SCHED_SUB = AR_SCHED_MAX_DEPTH + 2
AR_SCHED_SPLIT_GAP = 256
AR_SCHED_PARTS = 296
ADAM_REPLAY, ADAM_DENSE, ADAM_TOUCHED = 0, 1, 2
ADAM_MODES = {"replay": ADAM_REPLAY, "dense": ADAM_DENSE, "touched": ADAM_TOUCHED}
MAX_K = 32

c_f32p = C.c_void_p  # device pointers travel as integers
c_i32p = C.c_void_p


class ArTable(C.Structure):
    _fields_ = [("n_rows", C.c_int32), ("dim", C.c_int32), ("W", C.c_void_p), ("m", C.c_void_p),
                ("v", C.c_void_p), ("last_step", C.c_void_p)]


class ArPlan(C.Structure):
    _fields_ = [("batch_cap", C.c_int32), ("heavy_cap", C.c_int32), ("n_slots", C.c_int32),
                ("order", C.c_void_p), ("uniq", C.c_void_p), ("off", C.c_void_p), ("meta", C.c_void_p),
                ("heavy", C.c_void_p), ("in_prev", C.c_void_p)]


class ArSched(C.Structure):
    _fields_ = [("cap", C.c_int32), ("n_slots", C.c_int32), ("codes", C.c_void_p), ("glen", C.c_void_p),
                ("sub", C.c_void_p), ("cursor", C.c_void_p),
                ("gap_u", C.c_void_p), ("gap_a", C.c_void_p), ("bounds", C.c_void_p)]


class ArTrainCtx(C.Structure):
    _fields_ = [("users", ArTable), ("anime", ArTable),
                ("head", C.c_void_p), ("head_m", C.c_void_p), ("head_v", C.c_void_p),
                ("bn_moving", C.c_void_p), ("alpha", C.c_void_p),
                ("iu", C.c_void_p), ("ia", C.c_void_p), ("label", C.c_void_p),
                ("n_samples", C.c_int64), ("batch", C.c_int32), ("l2", C.c_float), ("mode", C.c_int32),
                ("plan_u", ArPlan), ("plan_a", ArPlan),
                ("uh", C.c_void_p), ("ah", C.c_void_p), ("c", C.c_void_p), ("ru", C.c_void_p),
                ("ra", C.c_void_p), ("dy", C.c_void_p), ("fwd_part", C.c_void_p), ("head_part", C.c_void_p),
                ("stepc", C.c_void_p), ("ticket", C.c_void_p),
                ("metrics", C.c_void_p), ("reg_acc", C.c_void_p), ("stepw", C.c_void_p), ("reg_scale", C.c_float),
                ("sched_ws", C.c_void_p), ("sched", ArSched), ("depth", C.c_int32), ("chunk_ws", C.c_void_p),
                ("health", C.c_void_p)]


class ArDistCtx(C.Structure):
    _fields_ = [("comm", C.c_void_p), ("n_ranks", C.c_int32), ("rank", C.c_int32),
                ("c_all", C.c_void_p), ("label_all", C.c_void_p), ("dy_all", C.c_void_p),
                ("fwd_part_all", C.c_void_p), ("head_part_all", C.c_void_p),
                ("send", C.c_void_p), ("recv", C.c_void_p)]


class ArShardCtx(C.Structure):
    _fields_ = [("comm", C.c_void_p), ("n_ranks", C.c_int32), ("rank", C.c_int32),
                ("req_send", C.c_void_p * 2), ("req_recv", C.c_void_p * 2), ("emit_map", C.c_void_p * 2),
                ("cache_idx", C.c_void_p * 2), ("max_count", C.c_void_p),
                ("rows_out", C.c_void_p * 2), ("rows_in", C.c_void_p * 2),
                ("grad_send", C.c_void_p * 2), ("grad_recv", C.c_void_p * 2),
                ("c_all", C.c_void_p), ("label_all", C.c_void_p), ("dy_all", C.c_void_p),
                ("fwd_part_all", C.c_void_p), ("head_part_all", C.c_void_p)]


PEER_MAX_RANKS, PEER_HANDLE_BYTES, PEER_FLAG_WORDS = 8, 64, 64
PEER_ERR_WORD = 32   # flag word a timed-out barrier sets (csrc/peer.inl kPeerErrWord)


class ArPeerCtx(C.Structure):
    _fields_ = [("n_ranks", C.c_int32), ("rank", C.c_int32),
                ("W_peer", (C.c_void_p * PEER_MAX_RANKS) * 2), ("pub_peer", C.c_void_p * PEER_MAX_RANKS),
                ("flags_peer", C.c_void_p * PEER_MAX_RANKS), ("sel_cap", C.c_int32),
                ("sel_key", C.c_void_p * 2), ("sel_samp", C.c_void_p * 2), ("sel_oth", C.c_void_p * 2),
                ("sel_cnt", C.c_void_p * 2), ("max_count", C.c_void_p), ("label_step", C.c_void_p),
                ("c_all", C.c_void_p), ("dy_all", C.c_void_p), ("fwd_part_all", C.c_void_p), ("head_part_all", C.c_void_p),
                ("rowflag_peer", (C.c_void_p * PEER_MAX_RANKS) * 2), ("pairs_peer", C.c_void_p * PEER_MAX_RANKS),
                ("hdrin_peer", C.c_void_p * PEER_MAX_RANKS), ("sel_lab", C.c_void_p * 2)]


class AnimerecError(RuntimeError):
    pass


_P = C.c_void_p
_I32, _I64, _F = C.c_int32, C.c_int64, C.c_float

# name -> (restype, argtypes); every symbol include/animerec.h declares
SIGNATURES = {
    "ar_last_error": (C.c_char_p, []),
    "ar_abi_version": (C.c_int, []),
    "ar_check_device": (C.c_int, []),
    "ar_plan_build": (C.c_int, [_P, _I64, _I32, _I64, _I32, C.POINTER(ArPlan), _P]),
    "ar_plan_build_lists": (C.c_int, [_P, _I32, _P, _I32, C.POINTER(ArPlan), _P]),
    "ar_plan_link": (C.c_int, [C.POINTER(ArPlan), _I32, _P, _P, _I32, _P]),
    "ar_plan_sched": (C.c_int, [C.POINTER(ArPlan), C.POINTER(ArPlan), _I32, _I64, _I64, _P, _I32, _P, _I32, _I32, _I32,
                                C.POINTER(ArSched), _P]),
    "ar_train_steps": (C.c_int, [C.POINTER(ArTrainCtx), _I64, _I32, _I64, _I32, _P]),
    "ar_chunk_ws_info": (C.c_int, [_I32, _I32, _I32, C.POINTER(C.c_int64)]),
    "ar_nccl_unique_id": (C.c_int, [_P]),
    "ar_comm_init": (C.c_int, [_P, _I32, _I32, C.POINTER(C.c_void_p)]),
    "ar_comm_destroy": (C.c_int, [_P]),
    "ar_train_steps_dist": (C.c_int, [C.POINTER(ArTrainCtx), C.POINTER(ArDistCtx), _I64, _I32, _I64, _I32, _P]),
    "ar_allgather_bytes": (C.c_int, [_P, _P, _P, _I64, _P]),
    "ar_shard_plan": (C.c_int, [C.POINTER(ArPlan), C.POINTER(ArPlan), _I32, C.POINTER(ArShardCtx), _P]),
    "ar_train_steps_sharded": (C.c_int, [C.POINTER(ArTrainCtx), C.POINTER(ArShardCtx), _I64, _I32, _I64, _I32, _I32, _P]),
    "ar_peer_export": (C.c_int, [_P, _P, C.POINTER(C.c_int64)]),
    "ar_peer_open": (C.c_int, [_P, _I64, C.POINTER(C.c_void_p)]),
    "ar_peer_close_all": (C.c_int, []),
    "ar_peer_plan": (C.c_int, [_P, _P, _P, _I64, _I64, _I32, _I32, C.POINTER(ArPlan), C.POINTER(ArPlan),
                               C.POINTER(ArPeerCtx), _P]),
    "ar_train_steps_peer": (C.c_int, [C.POINTER(ArTrainCtx), C.POINTER(ArPeerCtx), _I64, _I32, _I64, _I32, _I32, _P]),
    "ar_table_flush": (C.c_int, [C.POINTER(ArTable), _P, _F, _I64, _P, _P, _F, _P]),
    "ar_embed_fwd": (C.c_int, [_P, _P, _I32, _P, _P, _I32, _P, _P, _P, _P, _P, _P]),
    "ar_head_step": (C.c_int, [_P, _P, _I32, _P, _P, _P, _P, _P, _I64, _P, _P, _P]),
    "ar_rows_catchup": (C.c_int, [C.POINTER(ArTable), C.POINTER(ArPlan), _I32, _P, _F, _I64, _P]),
    "ar_rows_update": (C.c_int, [C.POINTER(ArTable), C.POINTER(ArPlan), _I32, _P, _P, _P, _P, _P, _P, _F,
                                 _I64, _I32, _P]),
    "ar_predict": (C.c_int, [_P, _P, _I32, _P, _P, _P, _P, _I64, _P, _P]),
    "ar_eval_sums": (C.c_int, [_P, _P, _I32, _P, _P, _P, _P, _P, _I64, _P, _P]),
    "ar_sumsq": (C.c_int, [_P, _I64, _P, _P]),
    "ar_bench_sfu": (C.c_int, [_P, _I32, _I32, _I32, _P]),
    "ar_user_favourites": (C.c_int, [_P, _P, _I32, C.c_double, _P, _P, _P]),
    "ar_user_recs": (C.c_int, [_P, _P, _P, _I32, _P, _I32, _P, _I32, _I32, _P, _P, _P]),
    "ar_rownorm": (C.c_int, [_P, _I64, _I32, _P, _P]),
    "ar_topk_query_workspace": (C.c_int64, [_I64, _I32]),
    "ar_cosine_topk_queries": (C.c_int, [_P, _I64, _I32, _I64, _I64, _P, _I32, _P, _P, _P, _P]),
    "ar_cosine_topk_query": (C.c_int, [_P, _I64, _I32, _I64, _P, _I64, _I32, _P, _P, _P, _P]),
    "ar_topk_merge": (C.c_int, [_P, _P, _I32, _I64, _I32, _I32, _I32, _P, _P, _P]),
    "ar_cosine_rerank": (C.c_int, [_P, _I64, _I64, _P, _I32, _P, _P, _P, _I32, _I32, _I32, _F, _P, _P, _P, _P, _P]),
    "ar_bits_from_csr": (C.c_int, [_P, _P, _I64, _I64, _I64, _P, _P]),
    "ar_rownorm_bf16": (C.c_int, [_P, _I64, _I32, _P, _P, _P]),
    "ar_allpairs_chunks": (C.c_int32, [_I64, _I64]),
    "ar_allpairs_list_cap": (C.c_int32, [_I32]),
    "ar_cosine_topk_allpairs": (C.c_int, [_P, _I64, _I64, _I64, _P, _I64, _I64, _I64, _I32, _I32, _I32, _P, _P,
                                          _I64, _P, _I32, _P, _P, _P, _P, _P, _P]),
}

_lib = None


def lib():
    """Load libanimerec.so (once).  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AnimerecError(
                "libanimerec.so is missing (%s); run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `python -m anime_recommendations_b200.build` -- there is no CPU fallback" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        if l.ar_abi_version() != ABI_VERSION:
            raise AnimerecError("libanimerec.so ABI %d != binding ABI %d; rebuild" % (l.ar_abi_version(), ABI_VERSION))
        _lib = l
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().ar_last_error().decode("utf-8", "replace")
        raise AnimerecError("%s failed (%d): %s" % (what or "libanimerec call", rc, msg))


def ptr(t):
    """Device pointer of a torch tensor (must be contiguous) or None."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise AnimerecError("tensor passed to libanimerec must be contiguous")
    return C.c_void_p(t.data_ptr())


def stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)
