"""ctypes binding of libunetb200.so.  Signatures are parsed from include/unetb200.h, so the header is the single
source of truth for the C ABI.  Raises ImportError if the library has not been built (no fallback)."""
from __future__ import annotations

import ctypes
import os
import re

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, "libunetb200.so")
HEADER_PATH = os.path.join(_ROOT, "include", "unetb200.h")

UB_BF16, UB_F32 = 0, 1

_CTYPES = {
    "int": ctypes.c_int, "long long": ctypes.c_longlong, "float": ctypes.c_float, "double": ctypes.c_double,
    "unsigned long long": ctypes.c_ulonglong, "cudaStream_t": ctypes.c_void_p,
}


def parse_header(path=HEADER_PATH):
    """-> ({name: (restype, [(ctype, argname), ...])}, {macro: int})"""
    src = open(path).read()
    macros = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(UB_\w+)\s+\(?(-?\d+)\)?", src)}
    body = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"(const char\*|long long|int)\s+(ub_\w+)\s*\(([^)]*)\)\s*;", body):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = {"const char*": ctypes.c_char_p, "long long": ctypes.c_longlong, "int": ctypes.c_int}[ret]
        argl = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    argl.append((ctypes.c_void_p, a.split("*")[-1].strip()))
                else:
                    ty, nm = a.rsplit(" ", 1)
                    argl.append((_CTYPES[ty], nm))
        decls[name] = (restype, argl)
    return decls, macros


DECLS, MACROS = parse_header()
UB_STATS_ROWS = MACROS["UB_STATS_ROWS"]
UB_MAX_CLASSES = MACROS["UB_MAX_CLASSES"]
UB_ZSCORE_BLOCKS = MACROS["UB_ZSCORE_BLOCKS"]

if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(there is no CPU or library fallback for the U-Net hot path)")
lib = ctypes.CDLL(LIB_PATH)
for _name, (_res, _args) in DECLS.items():
    _fn = getattr(lib, _name)      # AttributeError here = header/library mismatch
    _fn.restype = _res
    _fn.argtypes = [t for t, _ in _args]


class UBError(RuntimeError):
    pass


def last_error() -> str:
    return (lib.ub_last_error() or b"").decode()


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, int):
        return t
    return t.data_ptr()


def call(name, *args):
    """Invoke an int-returning entry point; tensors are passed as device pointers; raises UBError on failure."""
    fn = getattr(lib, name)
    conv = [(_ptr(a) if (a is None or hasattr(a, "data_ptr")) else a) for a in args]
    rc = fn(*conv)
    if rc != 0:
        raise UBError(f"{name} failed ({rc}): {last_error()}")
    return rc


launch_count = 0


def counted_call(name, *args):
    global launch_count
    launch_count += 1
    return call(name, *args)
