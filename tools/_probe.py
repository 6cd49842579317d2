"""ctypes loader of libunetb200_probe.so (include/unetb200_probe.h): the hardware probes live outside the product library.
Build with `make -C semantic-segmentation-unet_b200/csrc probes`."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unetb200._C as C  # noqa: E402

LIB = os.path.join(ROOT, "semantic-segmentation-unet_b200", "libunetb200_probe.so")
if not os.path.exists(LIB):
    raise ImportError(f"{LIB} not found: make -C semantic-segmentation-unet_b200/csrc probes")
DECLS, _ = C.parse_header(os.path.join(ROOT, "include", "unetb200_probe.h"))
lib = ctypes.CDLL(LIB)
for name, (res, args) in DECLS.items():
    fn = getattr(lib, name)
    fn.restype = res
    fn.argtypes = [t for t, _ in args]


def call(name, *args):
    rc = getattr(lib, name)(*[C._ptr(a) if (a is None or hasattr(a, "data_ptr")) else a for a in args])
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc})")
    return rc
