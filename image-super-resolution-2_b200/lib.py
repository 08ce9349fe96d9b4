"""ctypes binding of csrc/libffb200.so (the C ABI of include/ffb200.h).  Fails loudly when the
library is missing or a call returns an error -- there is no fallback path."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FFB200_LIB") or os.path.join(_HERE, "csrc", "libffb200.so")      # FFB200_LIB: development override

ACT_NONE, ACT_GELU, ACT_RELU, ACT_LRELU, ACT_SIGMOID, ACT_CLAMP01 = range(6)
CONV_1X1, CONV_3X3, CONV_2X2S2 = range(3)


class FFConvGemm(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("x_ld", C.c_int), ("cin", C.c_int),
        ("kind", C.c_int), ("w", C.c_void_p), ("n_pad", C.c_int), ("n_store", C.c_int), ("bias", C.c_void_p),
        ("act", C.c_int), ("alpha", C.c_float), ("col_scale", C.c_void_p), ("mul", C.c_void_p), ("mul_ld", C.c_int),
        ("aux", C.c_void_p), ("aux_ld", C.c_int), ("aux_chan", C.c_void_p), ("aux_chan_ld", C.c_int),
        ("aux_alpha", C.c_float), ("res", C.c_void_p), ("res_ld", C.c_int), ("res_is_f32", C.c_int),
        ("post_act", C.c_int), ("out_bf16", C.c_void_p), ("out_ld", C.c_int), ("out_f32", C.c_void_p),
        ("out_f32_ld", C.c_int), ("pixel_shuffle", C.c_int), ("gate_pairs", C.c_int), ("w_batch_rows", C.c_int),
        ("debug_simt", C.c_int), ("col_sums", C.c_void_p), ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p),
        ("ln_eps", C.c_float), ("ln_cols", C.c_int), ("ln_out", C.c_void_p), ("ln_out_ld", C.c_int),
        ("out_crop_h", C.c_int), ("out_crop_w", C.c_int), ("x2", C.c_void_p), ("x2_ld", C.c_int), ("cin2", C.c_int),
    ]


class FFWinAttn(C.Structure):
    _fields_ = [
        ("qkv", C.c_void_p), ("ld", C.c_int), ("q_off", C.c_int), ("k_off", C.c_int), ("v_off", C.c_int),
        ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("wh", C.c_int), ("ww", C.c_int), ("kh", C.c_int),
        ("kw", C.c_int), ("kpad_y", C.c_int), ("kpad_x", C.c_int), ("shift_y", C.c_int), ("shift_x", C.c_int),
        ("heads", C.c_int), ("head_off", C.c_int), ("bias_table", C.c_void_p), ("T", C.c_int),
        ("bias_heads", C.c_int), ("bias_head_off", C.c_int), ("rel_sign", C.c_int), ("rel_off_y", C.c_int),
        ("rel_off_x", C.c_int), ("rel_stride", C.c_int), ("out", C.c_void_p), ("out_ld", C.c_int),
        ("out_off", C.c_int), ("Hp", C.c_int), ("Wp", C.c_int),
    ]


class FFMlpFused(C.Structure):
    _fields_ = [
        ("t", C.c_void_p), ("t_ld", C.c_int), ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("w1", C.c_void_p), ("b1", C.c_void_p),
        ("w2", C.c_void_p), ("b2", C.c_void_p), ("x", C.c_void_p), ("x_ld", C.c_int), ("out_bf16", C.c_void_p), ("out_ld", C.c_int),
        ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p), ("ln_eps", C.c_float), ("ln_cols", C.c_int), ("ln_out", C.c_void_p),
        ("ln_out_ld", C.c_int),
    ]


class FFHabTail(C.Structure):
    _fields_ = [
        ("a0", C.c_void_p), ("a0_ld", C.c_int), ("a1", C.c_void_p), ("a1_ld", C.c_int), ("a1_diag", C.c_void_p), ("a1_diag_ld", C.c_int), ("a1_alpha", C.c_float), ("B", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("wp", C.c_void_p), ("wp_batch_rows", C.c_int), ("bp", C.c_void_p), ("res", C.c_void_p), ("res_ld", C.c_int),
        ("ln2_gamma", C.c_void_p), ("ln2_beta", C.c_void_p), ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
        ("x", C.c_void_p), ("x_ld", C.c_int), ("out_bf16", C.c_void_p), ("out_ld", C.c_int), ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p),
        ("ln_out", C.c_void_p), ("ln_out_ld", C.c_int), ("ln_eps", C.c_float), ("ln_cols", C.c_int),
    ]


class FFNafTail(C.Structure):
    _fields_ = [
        ("a0", C.c_void_p), ("a0_ld", C.c_int), ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("w3", C.c_void_p), ("w3_batch_rows", C.c_int),
        ("b3", C.c_void_p), ("res", C.c_void_p), ("res_ld", C.c_int), ("ln2_gamma", C.c_void_p), ("ln2_beta", C.c_void_p),
        ("w4", C.c_void_p), ("b4", C.c_void_p), ("w5", C.c_void_p), ("b5", C.c_void_p), ("x", C.c_void_p), ("x_ld", C.c_int),
        ("out_bf16", C.c_void_p), ("out_ld", C.c_int), ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p), ("ln_eps", C.c_float),
    ]


class FFError(RuntimeError):
    pass


_lib = None


def load():
    """Load libffb200.so; raises if it has not been built (python -m isr2_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FFError(f"{LIB_PATH} is missing: build it with `python image-super-resolution-2_b200/build.py` "
                      "(there is no CPU/PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.ff_last_error.restype = C.c_char_p
    lib.ff_launch_count.restype = C.c_longlong
    lib.ff_ssim_y_scratch_bytes.restype = C.c_size_t
    lib.ff_eval_scratch_bytes.restype = C.c_size_t
    if lib.ff_abi_version() != 6:
        raise FFError("libffb200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        raise FFError(f"{what} failed (rc={rc}): {load().ff_last_error().decode()}")


def launch_count():
    return int(load().ff_launch_count())
