"""ctypes binding of libmagnet_b200.so.  Prototypes are parsed from include/magnet_b200.h so the
Python side cannot drift from the C ABI.  There is NO fallback: if the library is missing or a
call fails, this raises — the product path never degrades to PyTorch/CPU code.
"""
import ctypes
import os
import re
from typing import Dict, List, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), "include", "magnet_b200.h")
_VARIANT = os.environ.get("MGB_VARIANT", "")      # developer builds (magnet_b200/build.py)
LIB_PATH = os.path.join(_HERE, "lib" + ("_" + _VARIANT if _VARIANT else ""), "libmagnet_b200.so")

_CTYPES = {
    "int": ctypes.c_int, "int64_t": ctypes.c_int64, "size_t": ctypes.c_size_t, "double": ctypes.c_double,
    "float": ctypes.c_float, "void": None,
}


def parse_header(path: str = HEADER) -> Dict[str, Tuple[object, List[object]]]:
    """{name: (restype, [argtypes])} for every `mgb_*` prototype in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    protos = {}
    for m in re.finditer(r"(const\s+char\s*\*|size_t|int64_t|int)\s+(mgb_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        restype = ctypes.c_char_p if "char" in ret else _CTYPES[ret]
        argtypes = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    base = a.replace("const", "").split()[0]
                    argtypes.append(_CTYPES[base])
        protos[name] = (restype, argtypes)
    return protos


_lib = None
_protos = None


def lib():
    global _lib, _protos
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m magnet_b200.build` "
                "(magnet_b200 has no CPU or PyTorch fallback)")
        l = ctypes.CDLL(LIB_PATH)
        _protos = parse_header()
        for name, (restype, argtypes) in _protos.items():
            fn = getattr(l, name)          # AttributeError => header/library mismatch, fail loudly
            fn.restype = restype
            fn.argtypes = argtypes
        if l.mgb_abi_version() != 1:
            raise RuntimeError("libmagnet_b200.so ABI version mismatch")
        _lib = l
    return _lib


def last_error() -> str:
    return lib().mgb_last_error().decode(errors="replace")


def check(rc: int, what: str = ""):
    if rc != 0:
        raise RuntimeError(f"magnet_b200 {what} failed (code {rc}): {last_error()}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("magnet_b200 kernels need CUDA tensors (there is no CPU fallback)")


def workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def f32c(t: torch.Tensor) -> torch.Tensor:
    """contiguous fp32 view/copy"""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def ver(t: torch.Tensor) -> int:
    """Version counter of a tensor for cache keys.  Inference tensors (created under ``torch.inference_mode()``, which
    Lightning >= 1.8 uses for validate/test/predict) do not track one: they cannot be modified in place outside
    inference mode either, so (data_ptr, shape, identity) identifies their contents and -1 stands in for the counter."""
    try:
        return t._version
    except RuntimeError:
        return -1


def on(t: torch.Tensor):
    """Context manager: make the device of ``t`` current for the launches inside (kernels run on the current device's
    current stream; tensors on another device than the current one would otherwise be launched against foreign pointers)."""
    return torch.cuda.device(t.device)


def same_device(*tensors):
    dev = None
    for t in tensors:
        if t is None:
            continue
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"magnet_b200: operands live on different devices ({dev} and {t.device})")


def guard(fn):
    """Decorator for the forward entry points: run ``fn`` with the device of its first CUDA tensor argument current, and
    refuse operands spread over several devices (kernels launch on the current device's current stream)."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        dev = None
        for a in list(args) + list(kwargs.values()):
            if torch.is_tensor(a) and a.is_cuda:
                if dev is None:
                    dev = a.device
                elif a.device != dev:
                    raise RuntimeError(f"magnet_b200.{fn.__name__}: operands live on different devices ({dev} and {a.device})")
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapped
