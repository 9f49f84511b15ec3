"""ctypes binding of libcleverrec_b200.so (the C ABI declared in include/cleverrec_b200.h).

There is deliberately no fallback: if the shared library is missing or the machine has no B200 the first
call raises.  torch is used only to own device memory and the current stream."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CRB_LIB_PATH") or os.path.join(_HERE, "libcleverrec_b200.so")  # CRB_LIB_PATH: A/B builds of the same library

OPT_SGD, OPT_ADAGRAD, OPT_ADAM = 0, 1, 2
ADAM_TF1, ADAM_LAZY = 0, 1
LOSS_BPR, LOSS_CROSS_ENTROPY, LOSS_SQUARE, LOSS_HINGE = 0, 1, 2, 3
SCORE_DOT, SCORE_GMF, SCORE_SQDIST, SCORE_DOT_BIAS = 0, 1, 2, 3

EXPORTS = [
    "crb_abi_version", "crb_last_error", "crb_create", "crb_destroy", "crb_set_history", "crb_sample_pairwise",
    "crb_sample_pointwise", "crb_sample_cml", "crb_epoch_rows", "crb_train_step_bpr", "crb_train_epoch_bpr",
    "crb_train_step_pointwise", "crb_adam_flush", "crb_score_pairs", "crb_score_pairs_topk", "crb_topk_segments", "crb_score_topk",
    "crb_score_topk_stats", "crb_eval_cache_invalidate", "crb_launch_count", "crb_profile_enable", "crb_profile_read", "crb_profile_read_tag",
    "crb_train_step_cml", "crb_set_history_lists", "crb_train_step_fism", "crb_fism_user_vectors", "crb_clip_rows",
    "crb_train_step_neumf", "crb_score_pairs_neumf", "crb_mask_seen",
    "crb_sample_nais", "crb_train_step_nais", "crb_train_epoch_nais", "crb_score_nais",
    "crb_shard_step_compute", "crb_shard_step_compute_pointwise", "crb_shard_apply_dense", "crb_shard_step_prepare", "crb_shard_apply_inbox", "crb_shard_barrier", "crb_shard_check", "crb_sampler_errors", "crb_malloc", "crb_free", "crb_ipc_export",
    "crb_ipc_open", "crb_ipc_close", "crb_build_history", "crb_prep_filter_reindex", "crb_prep_split_loo", "crb_prep_eval_negatives", "crb_set_item_lists", "crb_train_step_transcf", "crb_transcf_neighbourhood", "crb_score_pairs_transcf", "crb_np_seed", "crb_np_set_state", "crb_np_get_state", "crb_sample_epoch_numpy", "crb_sample_epoch_numpy_sbpr",
    "crb_train_step_lrml", "crb_score_pairs_lrml", "crb_set_social", "crb_sample_sbpr", "crb_train_step_sbpr", "crb_train_epoch_bpr_feeds", "crb_train_epoch_pointwise",
]


class CrbTable(C.Structure):
    _fields_ = [("w", C.c_void_p), ("s1", C.c_void_p), ("s2", C.c_void_p), ("last", C.c_void_p), ("rows", C.c_int64),
                ("dim", C.c_int32), ("_pad", C.c_int32)]


class CrbOpt(C.Structure):
    _fields_ = [("kind", C.c_int32), ("adam_mode", C.c_int32), ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double),
                ("eps", C.c_double), ("step", C.c_int64)]


MAX_RANKS = 8
SHARD_FLAGS, SHARD_ERR, SHARD_DENSE = 16, 8, 512


class CrbShard(C.Structure):
    _fields_ = [("n_ranks", C.c_int32), ("rank", C.c_int32), ("rows_cap", C.c_int64), ("q", CrbTable * MAX_RANKS),
                ("inbox_grad", C.c_void_p * MAX_RANKS), ("inbox_stamp", C.c_void_p * MAX_RANKS), ("flags", C.c_void_p * MAX_RANKS),
                ("dense_inbox", C.c_void_p * MAX_RANKS)]


class CrbError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "cleverrec_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load():
    """Load the shared library (building it is `python -m cleverrec_b200.build`).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libcleverrec_b200.so not built: run `python -m cleverrec_b200.build` (needs nvcc); "
                          "cleverrec_b200 has no CPU / PyTorch fallback path")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u32, u64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_float
    lib.crb_last_error.restype = C.c_char_p
    lib.crb_abi_version.restype = C.c_int
    lib.crb_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.crb_destroy.argtypes = [vp]
    lib.crb_set_history.argtypes = [vp, i64, i64, i64, vp, vp, vp, vp, vp]
    lib.crb_build_history.argtypes = [vp, vp, vp, i64, i64, i64, vp, vp, vp, vp, C.POINTER(i64), vp, vp, vp]
    lib.crb_prep_filter_reindex.argtypes = [vp, vp, vp, i64, i32, i32, vp, vp, vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), vp, vp, vp]
    lib.crb_prep_split_loo.argtypes = [vp, vp, vp, i64, i64, vp, vp, vp]
    lib.crb_prep_eval_negatives.argtypes = [vp, u64, vp, i64, i32, vp, vp]
    lib.crb_sample_pairwise.argtypes = [vp, u64, u32, i64, i64, i32, vp, vp, vp, vp, vp]
    lib.crb_sample_pointwise.argtypes = [vp, u64, u32, i64, i64, i32, vp, vp, vp, vp, vp]
    lib.crb_sample_cml.argtypes = [vp, u64, u32, i64, i64, i32, vp, vp, vp, vp]
    lib.crb_epoch_rows.argtypes = [vp, i32, i32]
    lib.crb_epoch_rows.restype = i64
    lib.crb_train_step_bpr.argtypes = [vp, C.POINTER(CrbTable), C.POINTER(CrbTable), C.POINTER(CrbOpt), vp, vp, vp, i64, f32, vp, vp]
    lib.crb_train_epoch_bpr.argtypes = [vp, C.POINTER(CrbTable), C.POINTER(CrbTable), C.POINTER(CrbOpt), u64, u32, i64, i64, i64,
                                        i32, f32, vp, vp]
    lib.crb_train_epoch_bpr_feeds.argtypes = [vp, C.POINTER(CrbTable), C.POINTER(CrbTable), C.POINTER(CrbOpt), vp, vp, vp, i64, i64, f32, vp, vp]
    lib.crb_train_step_pointwise.argtypes = [vp, i32, C.POINTER(CrbTable), C.POINTER(CrbTable), vp, vp, vp, C.POINTER(CrbOpt), i32,
                                             vp, vp, vp, i64, f32, vp, vp]
    lib.crb_train_epoch_pointwise.argtypes = [vp, i32, C.POINTER(CrbTable), C.POINTER(CrbTable), vp, vp, vp, C.POINTER(CrbOpt), i32, u64, u32, i64, i64, i64,
                                              i32, f32, vp, vp]
    lib.crb_adam_flush.argtypes = [vp, C.POINTER(CrbTable), C.POINTER(CrbOpt), vp]
    lib.crb_score_pairs.argtypes = [vp, i32, vp, vp, vp, i32, vp, vp, i64, vp, vp]
    lib.crb_topk_segments.argtypes = [vp, vp, vp, i64, i32, i32, vp, vp]
    lib.crb_score_pairs_topk.argtypes = [vp, i32, vp, vp, vp, i32, vp, vp, vp, i64, i32, i32, vp, vp]
    lib.crb_score_topk.argtypes = [vp, i32, vp, vp, vp, i64, i32, vp, vp, i64, i32, i32, vp, vp, vp]
    lib.crb_score_topk_stats.argtypes = [vp, C.POINTER(C.c_int64 * 4)]
    lib.crb_eval_cache_invalidate.argtypes = [vp]
    lib.crb_launch_count.argtypes = [vp]
    lib.crb_launch_count.restype = i64
    lib.crb_profile_enable.argtypes = [vp, i32]
    lib.crb_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.crb_profile_read_tag.argtypes = [vp, i32, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    T, O = C.POINTER(CrbTable), C.POINTER(CrbOpt)
    lib.crb_train_step_cml.argtypes = [vp, T, T, vp, vp, O, vp, vp, vp, i64, i32, f32, f32, i64, vp, vp]
    lib.crb_set_history_lists.argtypes = [vp, vp, vp]
    lib.crb_train_step_fism.argtypes = [vp, T, T, T, vp, vp, vp, O, vp, vp, vp, vp, i64, f32, f32, f32, i64, vp, vp]
    lib.crb_fism_user_vectors.argtypes = [vp, vp, i32, vp, vp, i64, f32, vp, vp]
    lib.crb_clip_rows.argtypes = [vp, vp, vp, i64, i32, f32, vp]
    lib.crb_train_step_neumf.argtypes = [vp, T, T, T, T, vp, vp, vp, vp, vp, vp, vp, i32, O, i32, vp, vp, vp, i64, f32, f32, vp, vp]
    lib.crb_score_pairs_neumf.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, i64, vp, vp]
    lib.crb_train_step_lrml.argtypes = [vp, T, T, vp, vp, vp, vp, vp, i32, O, vp, vp, vp, i64, f32, f32, vp, vp]
    lib.crb_score_pairs_lrml.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, i64, vp, vp]
    lib.crb_set_social.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, vp]
    lib.crb_sample_sbpr.argtypes = [vp, u64, u32, i64, i64, i32, vp, vp, vp, vp, vp, vp]
    lib.crb_train_step_sbpr.argtypes = [vp, T, T, T, vp, vp, vp, O, vp, vp, vp, vp, vp, i64, f32, vp, vp]
    lib.crb_mask_seen.argtypes = [vp, vp, vp, i64, i64, f32, vp]
    lib.crb_sample_nais.argtypes = [vp, u64, u32, i64, i32, i32, vp, vp, vp]
    lib.crb_train_step_nais.argtypes = [vp, T, T, T, vp, vp, vp, vp, vp, vp, i32, i32, O, vp, i32, vp, vp, i32, f32, f32, vp, vp]
    lib.crb_train_epoch_nais.argtypes = [vp, T, T, T, vp, vp, vp, vp, vp, vp, i32, i32, O, u64, u32, vp, vp, i64, i32, f32, f32, vp, vp]
    lib.crb_score_nais.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, vp, i32, vp, i32, f32, vp, vp]
    S = C.POINTER(CrbShard)
    lib.crb_shard_step_compute.argtypes = [vp, T, S, O, vp, vp, vp, u64, u32, i64, i32, i64, f32, vp, vp]
    lib.crb_set_item_lists.argtypes = [vp, vp, vp, vp]
    lib.crb_train_step_transcf.argtypes = [vp, T, T, vp, vp, O, vp, vp, vp, i64, f32, f32, f32, vp, vp]
    lib.crb_transcf_neighbourhood.argtypes = [vp, i32, vp, i32, vp, i64, vp, vp]
    lib.crb_score_pairs_transcf.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp, i64, vp, vp]
    lib.crb_shard_step_prepare.argtypes = [vp, T, u64, u32, i64, i32, i64, i64, vp, vp, vp, vp]
    lib.crb_shard_apply_inbox.argtypes = [vp, S, O, vp]
    lib.crb_shard_step_compute_pointwise.argtypes = [vp, i32, T, S, vp, O, i32, vp, vp, vp, u64, u32, i64, i32, i64, f32, vp, vp]
    lib.crb_shard_apply_dense.argtypes = [vp, S, O, vp, vp, vp, i32, vp]
    lib.crb_shard_barrier.argtypes = [vp, S, u32, i32, vp]
    lib.crb_shard_check.argtypes = [vp, S, vp]
    lib.crb_sampler_errors.argtypes = [vp, C.POINTER(u32), vp]
    lib.crb_malloc.argtypes = [vp, i64, C.POINTER(vp)]
    lib.crb_free.argtypes = [vp, vp]
    lib.crb_ipc_export.argtypes = [vp, vp, C.c_char_p]
    lib.crb_ipc_open.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    lib.crb_ipc_close.argtypes = [vp, vp]
    lib.crb_np_seed.argtypes = [vp, u32]
    lib.crb_np_set_state.argtypes = [vp, vp, i32]
    lib.crb_np_get_state.argtypes = [vp, vp, C.POINTER(i32)]
    lib.crb_sample_epoch_numpy.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp]
    lib.crb_sample_epoch_numpy_sbpr.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp]
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise CrbError(rc, load().crb_last_error().decode("utf-8", "replace"))


def ptr(x):
    """Raw address of a torch tensor (device or host), a NumPy array (host) or None."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return x.ctypes.data
    if not x.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return x.data_ptr()


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
