"""ctypes binding of liblas_b200.so (the C-ABI in include/las_b200.h).

The product path has no CPU or eager-PyTorch fallback: if the shared library is missing, or a
compute entry point is called without a CUDA device, this module raises.
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblas_b200.so")

_lib = None


class LasError(RuntimeError):
    pass


class DecArgs(ctypes.Structure):
    """Mirror of `struct las_dec_args` (include/las_b200.h)."""
    _fields_ = (
        [(n, c_int32) for n in ("B", "L", "Te", "Hd", "O", "A", "V", "E", "H", "C", "K", "mode")]
        + [("att_scaling", c_float), ("smooth_scaling", c_float), ("denc_accumulate", c_int32), ("_pad", c_int32)]
        + [(n, c_void_p) for n in (
            "enc_h", "P", "embx", "cell_bias", "wr_pk", "we_pk", "mlp_dec_pk", "mlp_o_pk", "mlp_o_b", "out_pk",
            "out_b", "emb_w", "conv_w", "mlp_att", "gvec",
            "ws", "zc", "ctx", "c_state", "emb_op", "logits", "pred", "e_buf", "dzf", "gates_save", "c_save",
            "wrT_pk", "mlp_oT_pk", "mlp_decT_pk", "dzc_all", "dcz_tot", "dcz_all", "dctx_all", "dw_buf",
            "dattc_all", "ddz_all", "dP", "att_part", "dc_state", "dgates", "dmlp_att", "dgvec", "dconv_w",
            "denc", "Q", "wr2_pk", "cpre", "conv_save", "wrT2_pk", "mlp_decT2_pk", "de_all", "dc_all", "cbias", "weT_pk", "outT_pk", "dlogits", "dl_tot", "demb_buf")]
        + [("drop_p", c_float), ("drop_site", ctypes.c_uint32)]
        + [(n, c_void_p) for n in ("seed_dev", "zcd", "pbar")]
        + [("t_begin", c_int32), ("t_end", c_int32), ("mlp_dec_pk_p", c_void_p)]
        + [("out_bf", c_void_p), ("bos_token", c_int32), ("stop_token", c_int32)]
        + [("tok_teacher", c_void_p), ("tf_mask", c_void_p), ("sample", c_int32), ("_pad3", c_int32)]
    )


P, I, L, F = c_void_p, c_int, c_int64, c_float
# name -> (restype, argtypes); every symbol include/las_b200.h declares
SIGNATURES = {
    "las_last_error": (c_char_p, []),
    "las_version": (c_int, []),
    "las_num_sms": (c_int, []),
    "las_launch_count": (ctypes.c_ulonglong, []),
    "las_path_counters": (c_int, [P, I]),
    "las_gemm_bf16": (c_int, [P, L, I, P, L, I, P, L, I, P, I, I, I, I, I, P]),
    "las_gemm_bf16_ws": (c_int, [P, L, I, P, L, I, P, L, I, P, I, I, I, I, I, P, L, P]),
    "las_cvt_pad_bf16": (c_int, [P, L, L, I, P, L, P]),
    "las_add2": (c_int, [P, P, P, L, P]),
    "las_dropout": (c_int, [P, I, L, L, I, L, L, I, F, P, ctypes.c_uint32, P]),
    "las_relu_bwd": (c_int, [P, P, I, P, L, P]),
    "las_colsum": (c_int, [P, I, L, L, I, P, P]),
    "las_gather_rows_bf16": (c_int, [P, I, P, L, P, L, P]),
    "las_scatter_add_rows": (c_int, [P, L, I, P, L, L, P, P]),
    "las_ce_ls_fwd": (c_int, [P, L, L, L, L, I, P, P, F, P, P, P, P]),
    "las_ce_ls_bwd": (c_int, [P, L, L, L, L, I, P, P, F, P, P, P, P]),
    "las_grad_norm": (c_int, [P, L, P, P, P]),
    "las_adam_step": (c_int, [P, P, P, P, P, L, F, F, F, F, F, P, F, P, F, P]),
    "las_pack_afrag": (c_int, [P, L, I, I, I, I, I, I, P, P]),
    "las_afrag_bytes": (c_int64, [I, I, I, I]),
    "las_smallmm": (c_int, [P, I, I, P, I, L, I, P, P, L, P, L, P, L, P]),
    "las_lstm_cell_step": (c_int, [P, P, P, P, L, P, L, I, P, P, L, I, I, P]),
    "las_lstm_ws_bytes": (c_int64, [I, I, I]),
    "las_lstm_seq_fwd": (c_int, [P, P, P, I, I, I, I, P, L, L, I, P, L, L, P, P, P, P]),
    "las_lstm_seq_bwd": (c_int, [P, L, L, I, P, I, P, I, I, I, I, P, P, P, L, L, P, P]),
    "las_lstm_persist_fwd": (c_int, [P, P, P, I, I, I, I, P, L, L, I, P, L, L, P, P]),
    "las_lstm_persist_bwd": (c_int, [P, L, L, I, P, P, I, I, I, I, P, P, L, L, P]),
    "las_lstm_persistent_geometry": (c_int, [I, P, P]),
    "las_lstm_persist_max_clusters": (c_int, [I, I]),
    "las_set_persistent": (c_int, [I]),
    "las_set_debug_buffer": (c_int, [P]),
    "las_whhT_owner_bytes": (c_int64, [I]),
    "las_pack_whhT_owner": (c_int, [P, I, P, P]),
    "las_pyramid_lens": (c_int, [P, I, I, P, P]),
    "las_att_init": (c_int, [P, I, I, P, L, P]),
    "las_dec_persistent_supported": (c_int, [ctypes.POINTER(DecArgs)]),
    "las_dec_persistent_pack_bytes": (c_int64, [I, I, I, I]),
    "las_dec_persistent_pack": (c_int, [I, P, L, I, I, I, P, P]),
    "las_att_dq": (c_int, [P, P, I, I, I, I, P, P]),
    "las_att_dconv": (c_int, [P, P, I, I, I, I, I, P, P, P]),
    "las_att_scratch_floats": (c_int64, [I, I, I, I, I, I]),
    "las_att_bwd_lean_supported": (c_int, [I, I]),
    "las_att_param_grads": (c_int, [P, P, P, P, P, P, I, I, I, I, I, P, P, P, P, P]),
    "las_att_param_grads_part": (c_int, [P, P, P, P, P, P, I, I, I, I, I, I, P, P, P, P, P]),
    "las_dec_fwd": (c_int, [ctypes.POINTER(DecArgs), P]),
    "las_dec_bwd": (c_int, [ctypes.POINTER(DecArgs), P]),
    "las_att_step": (c_int, [ctypes.POINTER(DecArgs), I, P]),
}


def build(verbose=False):
    """Compile csrc/*.cu for sm_100a into liblas_b200.so (nvcc cross-compiles without a GPU)."""
    import subprocess

    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise LasError("building liblas_b200.so failed")
    return LIB_PATH


def lib():
    """Load the library once; fail loudly when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LasError(
                f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no fallback path)"
            )
        _lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
    return _lib


def check(rc):
    if rc != 0:
        raise LasError(lib().las_last_error().decode("utf-8", "replace"))


PATH_NAMES = ("lstm_persist_fwd", "lstm_persist_bwd", "lstm_step_fwd", "lstm_step_bwd", "dec_persist_fwd",
              "dec_persist_bwd", "dec_step_fwd", "dec_step_bwd")


def path_counters(reset=False):
    """{path name: calls} since load / the last reset (las_path_counters): which kernels the entry points chose."""
    buf = (ctypes.c_ulonglong * len(PATH_NAMES))()
    n = lib().las_path_counters(buf, int(bool(reset)))
    assert n == len(PATH_NAMES)
    return dict(zip(PATH_NAMES, (int(v) for v in buf)))


def ptr(t):
    """Device pointer of a torch tensor (or None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr():
    import torch

    if not torch.cuda.is_available():
        raise LasError("liblas_b200 needs a CUDA device (sm_100a); there is no CPU path")
    return torch.cuda.current_stream().cuda_stream


# Per-call device timing (bench.py's per-kernel roofline): set TIMER to a list and every entry-point call made through
# call() appends (name, start_event, end_event), both recorded on the stream the kernels are launched on.
TIMER = None


def call(name, *args):
    """Invoke an int-returning C-ABI function on torch's current stream; raise LasError on failure."""
    fn = getattr(lib(), name)
    if TIMER is None:
        check(fn(*args, stream_ptr()))
        return
    import torch
    s = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    check(fn(*args, s.cuda_stream))
    e1.record(s)
    TIMER.append((name, e0, e1))
