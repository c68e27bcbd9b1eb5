"""ctypes binding of libsella_b200.so (C ABI declared in include/sella_b200.h).

There is no CPU fallback: if the shared object is missing, or a call is made
without a CUDA device, this raises.  PyTorch is used only as the owner of device
memory and streams; the C ABI itself takes raw device pointers.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libsella_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "sella_b200.h")

_lib = None


class SellaB200Error(RuntimeError):
    pass


def declared_symbols():
    """Names of every function declared in include/sella_b200.h."""
    with open(HEADER_PATH) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(sb_\w+)\s*\(", text)))


def get_lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise SellaB200Error(
            "libsella_b200.so is not built (%s). Run `python -m sella_b200._build` "
            "(or __graft_entry__.build()); there is no CPU fallback." % LIB_PATH)
    try:
        import torch  # noqa: F401  (loads the CUDA runtime the library links against)
    except Exception:  # pragma: no cover
        pass
    _lib = ctypes.CDLL(LIB_PATH)
    for name in declared_symbols():
        fn = getattr(_lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = ctypes.c_int
    return _lib


def _p(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise SellaB200Error("sella_b200 needs a CUDA device (sm_100a); no CPU fallback exists")


def check_f64(*tensors):
    import torch
    for t in tensors:
        if t is None:
            continue
        if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
            raise SellaB200Error("expected contiguous CUDA float64 tensors")


def call(name, *args):
    """Invoke a C-ABI entry point; non-zero return codes become exceptions."""
    rc = getattr(get_lib(), name)(*args)
    if rc != 0:
        raise SellaB200Error("%s failed with code %d" % (name, rc))


I = ctypes.c_int
D = ctypes.c_double
LL = ctypes.c_longlong
