"""ctypes binding of liblas_b200.so (the C-ABI in include/las_b200.h).

The product path has no CPU or eager-PyTorch fallback: if the shared library is missing, or a
compute entry point is called without a CUDA device, this module raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblas_b200.so")

_lib = None


class LasError(RuntimeError):
    pass


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
        _lib.las_last_error.restype = ctypes.c_char_p
    return _lib


def check(rc):
    if rc != 0:
        raise LasError(lib().las_last_error().decode("utf-8", "replace"))


def ptr(t):
    """Device pointer of a torch tensor (or None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def i64(v):
    return ctypes.c_int64(int(v))


def i32(v):
    return ctypes.c_int(int(v))


def f32(v):
    return ctypes.c_float(float(v))


def call(name, *args):
    """Invoke an int-returning C-ABI function and raise LasError on a non-zero return."""
    fn = getattr(lib(), name)
    check(fn(*args))
