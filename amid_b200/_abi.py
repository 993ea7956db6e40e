"""ctypes binding of include/amid_b200.h (the C ABI of libamid_b200.so).

The library is built in-tree by amid_b200/build.py (nvcc, sm_100a).  There is no CPU
fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from ctypes import POINTER, c_float, c_int32, c_int64, c_uint32, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libamid_b200.so")

BLOCKS = 2


class AmidError(RuntimeError):
    pass


class Dropout(C.Structure):
    _fields_ = [("train", c_int32), ("p", c_float), ("seed", c_uint64), ("site_base", c_uint32), ("batch_offset", c_int32)]


class EncoderTensors(C.Structure):
    _fields_ = [("pos_emb", c_void_p),
                ("ln1_w", c_void_p * BLOCKS), ("ln1_b", c_void_p * BLOCKS),
                ("in_w", c_void_p * BLOCKS), ("in_b", c_void_p * BLOCKS),
                ("out_w", c_void_p * BLOCKS), ("out_b", c_void_p * BLOCKS),
                ("ln2_w", c_void_p * BLOCKS), ("ln2_b", c_void_p * BLOCKS),
                ("c1_w", c_void_p * BLOCKS), ("c1_b", c_void_p * BLOCKS),
                ("c2_w", c_void_p * BLOCKS), ("c2_b", c_void_p * BLOCKS),
                ("ln3_w", c_void_p), ("ln3_b", c_void_p)]


class EncoderSaved(C.Structure):
    _fields_ = [(n, c_void_p * BLOCKS) for n in
                ("qn", "q", "k", "v", "o", "lse", "x1", "y", "h", "xout", "st1", "st2")] + [("st3", c_void_p)]


class HeadTensors(C.Structure):
    _fields_ = [("w0", c_void_p), ("b0", c_void_p), ("w2", c_void_p), ("b2", c_void_p)]


class BatchSource(C.Structure):
    _fields_ = [("hist_d1_vals", c_void_p), ("hist_d1_offs", c_void_p), ("hist_d2_vals", c_void_p), ("hist_d2_offs", c_void_p),
                ("excl_vals", c_void_p), ("excl_offs", c_void_p), ("target", c_void_p), ("user", c_void_p),
                ("domain", c_void_p), ("overlap", c_void_p), ("pool_d1", c_void_p), ("pool_d2", c_void_p),
                ("n_pool_d1", c_int64), ("n_pool_d2", c_int64), ("n_rows", c_int64)]


class BatchOut(C.Structure):
    _fields_ = [("seq_d1", c_void_p), ("seq_d2", c_void_p), ("i_node", c_void_p), ("user_node", c_void_p),
                ("domain_id", c_void_p), ("overlap_label", c_void_p), ("long_tail_mask_d1", c_void_p),
                ("long_tail_mask_d2", c_void_p), ("neg_samples", c_void_p)]


P = c_void_p
_SIGS = {
    "amid_version": (c_int32, []),
    "amid_launch_count": (c_int64, []),
    "amid_profile_enable": (c_int32, [c_int32]),
    "amid_profile_report_host_sync": (c_int64, [C.c_char_p, c_int64]),
    "amid_emb_gather_fwd": (c_int32, [P, c_int64, P, c_int64, P, P]),
    "amid_gather_error_host_sync": (c_int32, []),
    "amid_seq_embed_fwd": (c_int32, [P, c_int64, P, P, P, c_int32, c_int32, P, P, POINTER(Dropout), P]),
    "amid_embed_all_fwd": (c_int32, [P, c_int64, P, c_int64, P, P, P, P, c_int32, c_int32, P, P, P, P, P,
                                     POINTER(Dropout), P]),
    "amid_seq_embed_bwd": (c_int32, [P, P, c_int32, c_int32, P, POINTER(Dropout), P]),
    "amid_encoder_fwd_workspace_bytes": (c_int64, [c_int32, c_int32]),
    "amid_encoder_fwd": (c_int32, [POINTER(EncoderTensors), P, P, c_int32, c_int32, POINTER(Dropout),
                                   POINTER(EncoderSaved), P, P, c_int64, P]),
    "amid_encoder_bwd_workspace_bytes": (c_int64, [c_int32, c_int32]),
    "amid_encoder_bwd": (c_int32, [POINTER(EncoderTensors), P, P, c_int32, c_int32, POINTER(Dropout),
                                   POINTER(EncoderSaved), P, P, POINTER(EncoderTensors), P, P, c_int64, P]),
    "amid_encoder_fwd_tc": (c_int32, [POINTER(EncoderTensors), P, P, c_int32, c_int32, POINTER(Dropout),
                                      POINTER(EncoderSaved), P, P, c_int64, P]),
    "amid_encoder_bwd_tc": (c_int32, [POINTER(EncoderTensors), P, P, c_int32, c_int32, POINTER(Dropout),
                                      POINTER(EncoderSaved), P, P, POINTER(EncoderTensors), P, P, c_int64, P]),
    "amid_encoder_fwd_bf16": (c_int32, [POINTER(EncoderTensors), P, P, c_int32, c_int32, POINTER(Dropout),
                                        POINTER(EncoderSaved), P, P, c_int64, P]),
    "amid_encoder_bwd_bf16": (c_int32, [POINTER(EncoderTensors), P, P, c_int32, c_int32, POINTER(Dropout),
                                        POINTER(EncoderSaved), P, P, POINTER(EncoderTensors), P, P, c_int64, P]),
    "amid_encoder_fwd_x3": (c_int32, [POINTER(EncoderTensors), P, P, c_int32, c_int32, POINTER(Dropout),
                                      POINTER(EncoderSaved), P, P, c_int64, P]),
    "amid_encoder_bwd_x3": (c_int32, [POINTER(EncoderTensors), P, P, c_int32, c_int32, POINTER(Dropout),
                                      POINTER(EncoderSaved), P, P, POINTER(EncoderTensors), P, P, c_int64, P]),
    "amid_mim_scores": (c_int32, [P, P, c_int32, c_int32, P, P]),
    "amid_mim_scores_tc": (c_int32, [P, P, c_int32, c_int32, P, P]),
    "amid_mim_gate": (c_int32, [P, P, c_int32, c_float, P, P, P, P, P, P, P]),
    "amid_mim_aggregate": (c_int32, [P, P, P, P, c_int32, c_int32, c_int32, P, P]),
    "amid_mim_project": (c_int32, [P, P, P, P, P, c_int32, P, P, P]),
    "amid_mim_concat": (c_int32, [P, P, c_int32, c_int32, P, P]),
    "amid_mim_bwd": (c_int32, [P, P, P, P, P, P, P, P, P, P, c_int32, c_int32, c_int32, P, P, P, P, P, P, P]),
    "amid_meanpool_fwd": (c_int32, [P, P, c_int32, c_int32, c_float, P, P]),
    "amid_meanpool_bwd": (c_int32, [P, c_int32, c_int32, c_float, c_int32, P, P, P]),
    "amid_score_fwd": (c_int32, [P, P, P, POINTER(HeadTensors), c_int32, c_int32, c_int32, c_int32, P, P]),
    "amid_score_bwd_workspace_bytes": (c_int64, [c_int32, c_int32, c_int32, c_int32]),
    "amid_score_bwd": (c_int32, [P, P, P, POINTER(HeadTensors), c_int32, c_int32, c_int32, c_int32, P, P, P, P, P,
                                 POINTER(HeadTensors), P, c_int64, P]),
    "amid_loss_fwd_bwd": (c_int32, [P, c_int32, c_int32, c_int32, P, P, P, c_int32, c_float, c_float, P, P, P]),
    "amid_embgrad_workspace_bytes": (c_int64, [c_int64]),
    "amid_embgrad_segreduce": (c_int32, [P, P, c_int64, c_int64, P, P, P, P, c_int64, P]),
    "amid_embgrad_scatter_dense": (c_int32, [P, P, P, c_int64, P, c_int64, P]),
    "amid_embgrad_scatter_add": (c_int32, [P, P, c_int64, c_int64, P, c_int64, P]),
    "amid_adam_dense": (c_int32, [P, P, P, P, c_int64, c_int32, c_float, c_float, c_float, c_float, P]),
    "amid_adam_rows_lazy": (c_int32, [P, P, P, P, P, P, P, c_int64, c_int32, c_float, c_float, c_float, c_float, P]),
    "amid_adam_rows_flush": (c_int32, [P, P, P, P, c_int64, c_int32, c_float, c_float, c_float, c_float, P]),
    "amid_shard_plan_workspace_bytes": (c_int64, [c_int64]),
    "amid_shard_plan": (c_int32, [P, c_int64, c_int64, c_int32, P, P, P, P, P, c_int64, P]),
    "amid_rank_counts": (c_int32, [P, c_int64, c_int32, c_float, P, P, P]),
    "amid_batch_build": (c_int32, [POINTER(BatchSource), P, c_int32, c_int32, c_int32, c_int32, c_int64, c_uint64,
                                  POINTER(BatchOut), P]),
    "amid_catalogue_item_proj": (c_int32, [P, c_int64, P, c_int64, P, P, c_int32, P, P]),
    "amid_catalogue_user_proj": (c_int32, [P, P, c_int32, P, c_int32, P, P]),
    "amid_catalogue_rank": (c_int32, [P, P, c_int32, c_int32, P, c_int32, c_int32, P, P, P, c_float, P, P, P]),
    "amid_catalogue_scores": (c_int32, [P, P, c_int32, c_int32, P, c_int32, c_int32, P, P, P, P, P, P]),
    "amid_tc_linear_test": (c_int32, [P, P, P, c_int32, P, P]),
    "amid_tc_linear16_test": (c_int32, [P, P, P, c_int32, P, P]),
    "amid_tc_wgrad16_test": (c_int32, [P, P, c_int32, P, c_int32, P]),
    "amid_x3_linear_test": (c_int32, [P, P, P, c_int32, P, P, P]),
    "amid_attn_fwd_test": (c_int32, [P, P, P, P, P, c_int32, c_int32, POINTER(Dropout), c_uint32, c_int32, P]),
    "amid_attn_bwd_test": (c_int32, [P, P, P, P, P, P, P, P, P, c_int32, c_int32, POINTER(Dropout), c_uint32, c_int32, P]),
    "amid_x3_wgrad_test": (c_int32, [P, P, c_int32, P, P, c_int32, P]),
    "amid_dropout_mask_feature": (c_int32, [POINTER(Dropout), c_uint32, c_int64, P, P]),
    "amid_dropout_mask_attn": (c_int32, [POINTER(Dropout), c_uint32, c_int32, c_int32, P, P]),
}
EXPORTS = sorted(list(_SIGS) + ["amid_last_error"])

_lib = None
launches = 0   # number of kernel-launching ABI calls made by this process (bench.py reports it)


def lib():
    """Load libamid_b200.so (raises AmidError if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AmidError(f"{LIB_PATH} is missing: run `python amid_b200/build.py` (there is no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        l.amid_last_error.restype = C.c_char_p
        l.amid_last_error.argtypes = []
        _lib = l
    return _lib


def call(name: str, *args):
    """Call an int-returning entry point and raise on a negative return code."""
    global launches
    l = lib()
    rc = getattr(l, name)(*args)
    if rc < 0:
        raise AmidError(f"{name} failed ({rc}): {l.amid_last_error().decode()}")
    launches += 1
    return rc


def kernel_launches() -> int:
    """Kernels launched by libamid_b200.so so far in this process."""
    return int(lib().amid_launch_count())


def profile(on: bool) -> None:
    lib().amid_profile_enable(1 if on else 0)


def profile_report() -> dict:
    """{kernel: (count, total_ms)} recorded since profile(True); synchronises the device."""
    buf = C.create_string_buffer(1 << 16)
    lib().amid_profile_report_host_sync(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.split()
        out[name] = (int(cnt), float(ms))
    return out
