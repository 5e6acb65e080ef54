"""ctypes binding of libb200dit.so (C ABI: include/b200dit.h).

The library is built in-tree by `__graft_entry__.build()` / `csrc/build.sh`.  There is no fallback:
if the shared object is missing the import fails loudly, and every compute entry point fails when
no CUDA device is present.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200DIT_LIB selects another build of the same ABI (developer A/B runs); the default is the in-tree build
LIB_PATH = os.environ.get("B200DIT_LIB") or os.path.join(_HERE, "libb200dit.so")

DTYPE_F32, DTYPE_F16, DTYPE_BF16 = 0, 1, 2
MAX_ITEMS = 16


class DitConfig(C.Structure):
    _fields_ = [("dim", C.c_int32), ("ffn_dim", C.c_int32), ("num_heads", C.c_int32), ("num_layers", C.c_int32),
                ("in_dim", C.c_int32), ("out_dim", C.c_int32), ("text_dim", C.c_int32), ("text_len", C.c_int32),
                ("freq_dim", C.c_int32), ("i2v", C.c_int32), ("eps", C.c_float)]


class B200Error(RuntimeError):
    pass


# name -> (restype, argtypes); kept in one table so tests can check it against include/b200dit.h
_P, _I, _L, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
_PP = C.POINTER(C.c_void_p)
_IP = C.POINTER(C.c_int32)
SIGNATURES = {
    "b200dit_create": (_I, [C.POINTER(DitConfig), _PP]),
    "b200dit_destroy": (None, [_P]),
    "b200dit_load_weight": (_I, [_P, C.c_char_p, _P, _I, _I, C.POINTER(C.c_int64)]),
    "b200dit_finalize": (_I, [_P]),
    "b200dit_forward": (_I, [_P, _I, _PP, _PP, _I, _P, _PP, _IP, _I, _PP, _I, _I, _I, _I, _PP, _P]),
    "b200dit_forward_cfg": (_I, [_P, _I, _PP, _PP, _I, _P, _PP, _IP, _PP, _IP, _I, _PP, _I, _I, _I, _I, _F, _PP, _P]),
    "b200dit_context_hint": (_I, [_P, C.c_uint64]),
    "b200dit_weight_names": (C.c_int64, [C.POINTER(DitConfig), C.c_char_p, C.c_int64]),
    "b200dit_set_tap": (_I, [_P, _I, _P, C.c_int64]),
    "b200dit_set_taps": (_I, [_P, _I, _IP, _PP, C.c_int64]),
    "b200dit_set_graphs": (_I, [_P, _I]),
    "b200dit_set_pad_to_seq_len": (_I, [_P, _I]),
    "b200dit_last_flops": (C.c_double, [_P]),
    "b200dit_nonfinite_rows": (_I, [_P, _P, C.POINTER(C.c_uint32)]),
    "b200dit_train_forward": (_I, [_P, _I, _PP, _P, _PP, _IP, _I, _I, _I, _I, _I, _PP, _P]),
    "b200dit_backward": (_I, [_P, _PP, _F, _I, _PP, _P]),
    "b200dit_zero_grad": (_I, [_P, _P]),
    "b200dit_grad_buffers": (_I, [_P, _PP, C.POINTER(C.c_int64), _PP, C.POINTER(C.c_int64)]),
    "b200dit_read_grad": (_I, [_P, C.c_char_p, _P, _L, _F, _I, _P]),
    "b200vae_create": (_I, [_I, _I, _PP]),
    "b200vae_destroy": (None, [_P]),
    "b200vae_load_weight": (_I, [_P, C.c_char_p, _P, _I, _I, C.POINTER(C.c_int64)]),
    "b200vae_finalize": (_I, [_P]),
    "b200vae_decode": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "b200vae_encode": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "b200vae_pipe_prepare": (_I, [_P, _I, _I, _P]),
    "b200vae_pipe_connect": (_I, [_P, _P]),
    "b200vae_decode_pipelined": (_I, [_P, _P, _I, _I, _I, _P, _I, _I, _I, _I, _P]),
    "b200vae_pipe_chunks": (_I, [_I, _I]),
    "b200disc_create": (_I, [_I, _I, _I, _F, _PP]),
    "b200disc_destroy": (None, [_P]),
    "b200disc_load_weight": (_I, [_P, C.c_char_p, _P, _I, _I, C.POINTER(C.c_int64)]),
    "b200disc_finalize": (_I, [_P]),
    "b200disc_forward": (_I, [_P, _PP, _I, _I, _P, _P, _P]),
    "b200omni_audio_tokens": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _L, _P]),
    "b200_flash_attention": (_I, [_P, _P, _P, _IP, _I, _I, _I, _I, _F, _P, _P]),
    "b200_flash_attention_backward": (_I, [_P, _P, _P, _P, _IP, _I, _I, _I, _I, _F, _P, _P, _P, _P]),
    "b200_linear": (_I, [_P, _L, _P, _L, _P, _I, _I, _I, _I, _P, _L, _I, _P]),
    "b200_solver_lincomb": (_I, [_I, _PP, _I, _PP, C.POINTER(C.c_float), _L, _P]),
    "b200_last_error": (C.c_char_p, []),
    "b200_kernel_launches": (C.c_int64, []),
    "b200_version": (C.c_char_p, []),
    "b200_profile_enable": (_I, [_I]),
    "b200_profile_collect": (_I, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                  C.POINTER(C.c_int64)]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise B200Error(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU / PyTorch fallback for the engine)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(status):
    if status != 0:
        msg = lib().b200_last_error().decode("utf-8", "replace")
        if "exceeds limit" in msg:            # the reference raises AssertionError here (model.py:521)
            raise AssertionError(msg)
        raise B200Error(msg)


def ptr_array(ptrs):
    arr = (C.c_void_p * len(ptrs))(*ptrs)
    return arr


def int_array(vals):
    return (C.c_int32 * len(vals))(*[int(v) for v in vals])
