"""ctypes binding of libmcgm.so (the C ABI declared in include/mcgm.h).

There is no CPU fallback and no alternative backend: if the shared library is
missing, or a call returns non-zero, a Python exception is raised.
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_long, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# MCGM_LIB_PATH: developer override for A/B runs of kernel variants (tools/); the product loads the in-tree build
LIB_PATH = os.environ.get("MCGM_LIB_PATH") or os.path.join(_HERE, "libmcgm.so")

F32, BF16 = 0, 1
GATE_BN, GATE_RELU, GATE_UP2, GATE_MASK, GATE_AFFINE = 1, 2, 4, 8, 16
EPI_BIAS, EPI_RES, EPI_RES_UP2, EPI_RELU, EPI_MASK, EPI_TANH, EPI_POSMASK = 1, 2, 4, 8, 16, 32, 64
CONV_S1, CONV_S2K4, CONV_PHASE2, CONV_TAPS = 0, 1, 2, 3

# name -> (restype, argtypes); every symbol of include/mcgm.h
SIGNATURES = {
    "mcgm_last_error": (c_char_p, []),
    "mcgm_version": (c_int, []),
    "mcgm_sm_count": (c_int, []),
    "mcgm_pack_codebook": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "mcgm_mc_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "mcgm_bn_stats_workspace_bytes": (c_size_t, [c_int]),
    "mcgm_bn_stats": (c_int, [c_void_p, c_int, c_long, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_float, c_float, c_void_p, c_void_p]),
    "mcgm_gate_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                              c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_int, c_void_p]),
    "mcgm_gate_bwd_workspace_bytes": (c_size_t, [c_int]),
    "mcgm_gate_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                              c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "mcgm_pack_weight": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mcgm_unpack_wgrad": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "mcgm_sn_power_iter": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                   c_float, c_void_p]),
    "mcgm_sn_weight_bwd_workspace_bytes": (c_size_t, []),
    "mcgm_sn_weight_bwd": (c_int, [c_void_p, c_long, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                   c_void_p, c_void_p]),
    "mcgm_pack_weight_batched": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p]),
    "mcgm_sn_weight_bwd_batched_workspace_bytes": (c_size_t, []),
    "mcgm_sn_weight_bwd_batched": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_void_p, c_void_p, c_void_p]),
    "mcgm_pack_weight_padded": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "mcgm_conv_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "mcgm_conv_fwd_ex": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "mcgm_conv_wgrad_ex_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "mcgm_conv_wgrad_ex": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                   c_int, c_int, c_int, c_void_p, c_void_p]),
    "mcgm_conv_wgrad_grouped_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "mcgm_conv_wgrad_grouped": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                        c_int, c_int, c_void_p, c_void_p]),
    "mcgm_conv_wgrad_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "mcgm_conv_wgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                c_int, c_int, c_void_p, c_void_p]),
    "mcgm_conv_fwd_taps": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "mcgm_conv_wgrad_taps_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "mcgm_conv_wgrad_taps": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_void_p, c_void_p, c_void_p, c_void_p]),
    "mcgm_sgate_workspace_bytes": (c_size_t, [c_int]),
    "mcgm_sgate_stats": (c_int, [c_void_p, c_long, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float,
                                 c_void_p, c_void_p]),
    "mcgm_sgate_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p]),
    "mcgm_sgate_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "mcgm_thin_im2col": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_int, c_int, c_void_p]),
    "mcgm_thin_conv_fwd": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int,
                                   c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mcgm_thin_conv_wgrad_workspace_bytes": (c_size_t, [c_int]),
    "mcgm_thin_conv_wgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                     c_void_p]),
    "mcgm_split_tf32": (c_int, [c_void_p, c_void_p, c_void_p, c_long, c_void_p]),
    "mcgm_split_bf16x3": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_void_p]),
    "mcgm_thin_col2im": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                 c_void_p]),
    "mcgm_pool2": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "mcgm_unpool2": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "mcgm_conv_fwd_bnstats_bytes": (c_size_t, [c_int]),
    "mcgm_conv_fwd_bnstats": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                      c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "mcgm_bn_finalize": (c_int, [c_void_p, c_int, c_int, c_long, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float,
                                 c_void_p]),
    "mcgm_flow_front_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_int, c_int,
                                    c_void_p]),
    "mcgm_flow_front_bwd_workspace_bytes": (c_size_t, [c_int]),
    "mcgm_flow_front_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_long, c_int, c_int, c_void_p, c_void_p]),
    "mcgm_flow_tail_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                   c_void_p]),
    "mcgm_flow_tail_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mcgm_flow_tail_rev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mcgm_grad_clip_workspace_bytes": (c_size_t, []),
    "mcgm_grad_clip_coef": (c_int, [c_void_p, c_long, c_float, c_void_p, c_void_p, c_void_p]),
    "mcgm_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_void_p, c_float, c_float, c_float,
                               c_float, c_float, c_float, c_void_p, c_void_p]),
    "mcgm_adam_step_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mcgm_debug_set": (c_int, [c_int, c_uint64]),
}

_lib = None
launch_count = 0          # number of libmcgm entry-point calls that launch kernels (bench: gpu_launches)


def load():
    """Load libmcgm.so once; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` at the repo root. "
            "There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    # developer A/B knobs of the kernels (tools/, bring-up): MCGM_DEBUG_SET="key=value,key=value" -> mcgm_debug_set
    for kv in filter(None, os.environ.get("MCGM_DEBUG_SET", "").split(",")):
        k, v = kv.split("=")
        if lib.mcgm_debug_set(int(k), int(v)) != 0:
            raise RuntimeError(f"mcgm_debug_set({k}, {v}) failed")
    return lib


def call(name, *args):
    """Call an int-returning entry point; non-zero -> RuntimeError(mcgm_last_error())."""
    global launch_count
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.mcgm_last_error()
        raise RuntimeError(f"{name} failed ({rc}): {msg.decode() if msg else '?'}")
    launch_count += 1
    return rc


def query(name, *args):
    """Call a size/version query (no error convention)."""
    return getattr(load(), name)(*args)
