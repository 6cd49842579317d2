"""CPU (-m "not gpu"): the C-ABI shared library loads, exports every symbol include/unetb200.h declares, and its
argument checking works without a GPU (no compute calls here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "unetb200.h")
LIB = os.path.join(ROOT, "semantic-segmentation-unet_b200", "libunetb200.so")


@pytest.fixture(scope="module")
def C():
    import __graft_entry__ as g
    if not os.path.exists(LIB):
        g.build()
    import unetb200._C as C
    return C


def test_every_declared_symbol_is_exported(C):
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    declared = set(re.findall(r"\b(ub_\w+)\s*\(", src))
    assert len(declared) >= 35
    assert declared == set(C.DECLS), "header parser and header disagree"
    lib = ctypes.CDLL(LIB)
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing


def test_no_undeclared_ub_exports(C):
    out = subprocess.run(["nm", "-D", "--defined-only", LIB], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("ub_")}
    internal = {"ub_set_error", "ub_num_sms", "ub_tmap_act4d", "ub_tmap_mat2d"}
    extra = {e for e in exported if not e.startswith("_Z")} - set(C.DECLS) - internal
    assert not extra, extra


def test_version_and_constants(C):
    assert C.lib.ub_version() == C.MACROS["UB_VERSION"]
    assert C.UB_STATS_ROWS == 592 and C.UB_MAX_CLASSES == 8


def test_argument_errors_are_reported_not_thrown(C):
    # null pointers / bad shapes are rejected before any CUDA call: negative status + message, no abort
    rc = C.lib.ub_conv3x3_fwd(None, 64, None, 0, None, None, None, None, 1, 16, 16, 64, 1, None)
    assert rc == C.MACROS["UB_ERR_INVALID_ARG"]
    assert "null pointer" in C.last_error()
    rc = C.lib.ub_conv3x3_fwd(ctypes.c_void_p(256), 48, None, 0, ctypes.c_void_p(256), None, ctypes.c_void_p(256), None, 1, 16, 16, 64, 1, None)
    assert rc == C.MACROS["UB_ERR_UNSUPPORTED_SHAPE"]
    assert "multiples of 64" in C.last_error()
    rc = C.lib.ub_head_fwd(ctypes.c_void_p(256), ctypes.c_void_p(256), ctypes.c_void_p(256), ctypes.c_void_p(256), None, 10, 256, 0, None)
    assert rc == C.MACROS["UB_ERR_UNSUPPORTED_SHAPE"]          # number_classes > UB_MAX_CLASSES_ANY (labels and masks are uint8)
    assert "255" in C.last_error()
    with pytest.raises(C.UBError):
        C.call("ub_adam", None, None, None, None, None, 0, 0.0, 0.0, 0.0, 0.0, 1.0, None)


def test_product_package_has_no_oracle_or_cpu_fallback():
    """the product path must never import oracle/ and must fail loudly without CUDA"""
    pkg = os.path.join(ROOT, "semantic-segmentation-unet_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("load_oracle_params", ""), fn
    import torch
    if not torch.cuda.is_available():
        from unetb200.model import UNet
        with pytest.raises(RuntimeError):
            UNet(2, 1, 1)
