"""ctypes binding of libselavi_b200.so (the C ABI declared in include/selavi_b200.h).

The product path has NO CPU fallback: if the shared library is missing or a call fails, this module
raises.  PyTorch is only used by callers for device memory, streams and torch.distributed.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libselavi_b200.so")

_lib = None

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_ll = ctypes.c_longlong
c_double = ctypes.c_double
c_float = ctypes.c_float
c_size_t = ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol of include/selavi_b200.h (tests check this)
SIGNATURES = {
    "selavi_version": (c_int, []),
    "selavi_last_error": (ctypes.c_char_p, []),
    "selavi_sk_softmax_product": (c_int, [c_void_p, c_void_p, c_ll, c_int, c_void_p, c_void_p]),
    "selavi_sk_softmax64": (c_int, [c_void_p, c_ll, c_int, c_void_p, c_void_p]),
    "selavi_l1_cost_workspace_bytes": (c_size_t, [c_ll, c_int]),
    "selavi_l1_cost_matrix": (c_int, [c_void_p, c_void_p, c_ll, c_int, c_void_p, c_void_p, c_void_p]),
    "selavi_sk_workspace_bytes": (c_size_t, [c_int]),
    "selavi_sk_kp": (c_int, [c_int]),
    "selavi_sk_solve": (c_int, [c_void_p, c_ll, c_ll, c_int, c_double, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_int, c_int, c_double, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "selavi_conv_tiles": (c_int, [c_int, c_void_p, c_void_p]),
    "selavi_conv_wpack_bytes": (c_size_t, [c_int, c_int]),
    "selavi_conv_wpack_bytes_f16": (c_size_t, [c_int, c_int]),
    "selavi_conv_pack_weights": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "selavi_conv_gemm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int,
                                 c_int, c_void_p]),
    "selavi_wgrad_workspace_bytes": (c_size_t, [c_void_p]),
    "selavi_conv_wgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p,
                                  c_int, c_int, c_void_p]),
    "selavi_bn_reduce_partials": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "selavi_bn_finalize": (c_int, [c_void_p, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_int,
                                   c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "selavi_bn_eval_affine": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_void_p, c_void_p,
                                      c_void_p]),
    "selavi_bn_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_ll, c_int,
                                c_void_p]),
    "selavi_bn_bwd_blocks": (c_int, [c_ll]),
    "selavi_bn_bwd_reduce": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_ll,
                                     c_int, c_void_p, c_void_p, c_void_p]),
    "selavi_bn_bwd_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_double, c_ll, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p]),
    "selavi_conv_wgrad_bf16_planes": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int,
                                              c_int, c_void_p]),
    "selavi_split_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_ll, c_int, c_void_p]),
    "selavi_conv_wgrad_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                       c_void_p, c_int, c_int, c_void_p]),
    "selavi_dgrad_wpack_bytes": (c_size_t, [c_void_p]),
    "selavi_dgrad_pack_weights": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "selavi_conv_dgrad_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "selavi_conv_halo_plan": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "selavi_conv_halo_pack_weights": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "selavi_conv_halo_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int,
                                     c_void_p]),
    "selavi_conv_halo_dgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "selavi_conv_halo_dgrad_bnstats": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                               c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "selavi_relu_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_ll, c_int, c_void_p]),
    "selavi_maxpool3x3s2_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "selavi_maxpool3x3s2_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                        c_void_p]),
    "selavi_avgpool_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "selavi_avgpool_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "selavi_nchw_to_cl": (c_int, [c_void_p, c_void_p, c_int, c_int, c_ll, c_int, c_void_p]),
    "selavi_sgd_step": (c_int, [c_void_p, c_int, c_float, c_float, c_float, c_int, c_void_p]),
    "selavi_bgemm": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_ll, c_ll, c_ll, c_void_p, c_ll, c_void_p,
                             c_void_p, c_ll, c_ll, c_ll, c_void_p, c_void_p, c_ll, c_void_p, c_ll, c_ll, c_ll, c_int,
                             c_void_p]),
    "selavi_heads_bn_stats": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "selavi_heads_bn_finalize": (c_int, [c_void_p, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float,
                                         c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "selavi_heads_bn_eval_affine": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_void_p,
                                            c_void_p, c_void_p]),
    "selavi_heads_act": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "selavi_heads_bn_bwd_reduce": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                           c_int, c_int, c_void_p, c_void_p]),
    "selavi_heads_bn_bwd_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_double, c_int, c_int, c_int, c_void_p, c_void_p]),
    "selavi_heads_sum_masked": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_ll, c_int, c_void_p]),
    "selavi_heads_colsum": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "selavi_ce_loss": (c_int, [c_void_p, c_void_p, c_ll, c_void_p, c_ll, c_ll, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                               c_void_p, c_void_p]),
    "selavi_sgd_step_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_float, c_int, c_void_p]),
    "selavi_mel_logfbank": (c_int, [c_void_p, c_int, c_ll, c_int, c_int, c_int, c_void_p, c_int, c_int, c_double, c_int,
                                    c_void_p, c_void_p]),
    "selavi_conv_wgrad_plan": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "selavi_clip_augment": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "selavi_symm_alloc": (c_int, [c_size_t, c_void_p, c_void_p]),
    "selavi_symm_open": (c_int, [c_void_p, c_void_p]),
    "selavi_symm_close": (c_int, [c_void_p]),
    "selavi_symm_free": (c_int, [c_void_p]),
    "selavi_symm_memset": (c_int, [c_void_p, c_int, c_size_t, c_void_p]),
    "selavi_p2p_allreduce_f64": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_ll, c_int, c_ll, c_void_p]),
}


class SelaviError(RuntimeError):
    pass


# kernels launched per C-ABI call (for bench.py's `gpu_launches` claim); host-only queries launch none
KERNELS_PER_CALL = {
    "selavi_sk_solve": 1, "selavi_sk_softmax_product": 1, "selavi_sk_softmax64": 1, "selavi_l1_cost_matrix": 2, "selavi_conv_pack_weights": 1, "selavi_conv_gemm": 1, "selavi_conv_wgrad": 4,
    "selavi_bn_reduce_partials": 1, "selavi_bn_finalize": 1, "selavi_bn_eval_affine": 1, "selavi_bn_apply": 1,
    "selavi_bn_bwd_reduce": 2, "selavi_bn_bwd_apply": 1, "selavi_relu_bwd": 1, "selavi_maxpool3x3s2_fwd": 1,
    "selavi_maxpool3x3s2_bwd": 1, "selavi_avgpool_fwd": 1, "selavi_avgpool_bwd": 1, "selavi_nchw_to_cl": 1,
    "selavi_sgd_step": 1, "selavi_sgd_step_host": 6, "selavi_bgemm": 1, "selavi_heads_bn_stats": 1, "selavi_heads_bn_finalize": 1,
    "selavi_heads_bn_eval_affine": 1, "selavi_heads_act": 1, "selavi_heads_bn_bwd_reduce": 1,
    "selavi_heads_bn_bwd_apply": 1, "selavi_heads_sum_masked": 1, "selavi_heads_colsum": 1, "selavi_ce_loss": 2,
    "selavi_mel_logfbank": 1, "selavi_split_bf16": 1, "selavi_p2p_allreduce_f64": 1, "selavi_conv_wgrad_bf16": 3, "selavi_conv_wgrad_bf16_planes": 2,
    "selavi_dgrad_pack_weights": 1, "selavi_conv_dgrad_bf16": 1,
    "selavi_conv_halo_pack_weights": 1, "selavi_conv_halo_fwd": 1, "selavi_conv_halo_dgrad": 1, "selavi_conv_halo_dgrad_bnstats": 1, "selavi_clip_augment": 1,
}
COUNT_CALLS = False
CALLS = {}


def kernel_launches():
    """Number of OUR kernels launched since CALLS was cleared (counted at the C-ABI boundary)."""
    return sum(KERNELS_PER_CALL.get(k, 0) * v for k, v in CALLS.items())


class _Handle:
    """Attribute access to the typed ctypes functions, with optional call counting."""

    def __init__(self, h):
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, self._wrap(name, fn))

    @staticmethod
    def _wrap(name, fn):
        def call(*a):
            if COUNT_CALLS:
                CALLS[name] = CALLS.get(name, 0) + 1
            return fn(*a)
        return call


def lib():
    """Load (once) and return the ctypes handle; raises if the CUDA extension is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SelaviError(
                f"{LIB_PATH} not found: build it with `python -m selavi_b200.build` "
                "(no CPU fallback exists for the selavi_b200 hot path)")
        _lib = _Handle(ctypes.CDLL(LIB_PATH))
    return _lib


def check(code, what):
    if code != 0:
        msg = lib().selavi_last_error()
        raise SelaviError(f"{what} failed with code {code}: {msg.decode() if msg else ''}")


def ptr(t):
    """Device/host pointer of a torch tensor (or None)."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
