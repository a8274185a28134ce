"""The C-ABI library loads and exports every symbol include/mntr_gpu.h declares (no compute)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mntr_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mntr_gpu_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    from minotaur_b200 import build, engine
    build.build()
    lib = ctypes.CDLL(engine.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in mntr_gpu.h but not exported"
    assert sorted(engine.ABI_SYMBOLS) == syms
    assert lib.mntr_gpu_abi_version() == 1


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    from minotaur_b200 import engine
    lib = engine.load_library()
    if lib.mntr_gpu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(engine.EngineError):
        engine.GpuBoundEngine(0)


def test_product_never_imports_oracle():
    """Nothing under minotaur_b200/ may import, link or execute oracle/."""
    pkg = os.path.join(ROOT, "minotaur_b200")
    for dp, _, files in os.walk(pkg):
        if "build" in dp.split(os.sep)[-1:]:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".txt")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "fbbt_oracle" not in txt, os.path.join(dp, f)
                assert "libminotaur_ref" not in txt, os.path.join(dp, f)


def test_reference_side_patch_applies_and_compiles():
    """minotaur_b200/handler/{strong,weak}_brancher_prefetch.patch (all strong-branching candidates of a node
    in one device call) apply to the reference's StrongBrancher.cpp / WeakBrancher.cpp as they are and compiles against the reference's headers
    and GpuBoundHandler.h.  Only where the reference tree is present."""
    import subprocess
    if not os.path.exists("/root/reference/src/base/StrongBrancher.cpp"):
        pytest.skip("reference sources not present")
    res = subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "patch_check"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "strong_brancher_prefetch.patch applies" in res.stdout and "weak_brancher_prefetch.patch applies" in res.stdout
    assert "quad_handler_gpu.patch applies" in res.stdout
