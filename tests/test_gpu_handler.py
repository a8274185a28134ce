"""GpuBoundHandler (the Minotaur::Handler adapter, minotaur_b200/handler/) against the reference's own
LinearHandler + NlPresHandler through Handler::presolveNode on real Relaxation objects.  The test binary is
built by `make -C oracle handler_test` where the reference sources exist and travels prebuilt to the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "handler_test")

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def test_handler_matches_reference_handlers():
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/handler_test not built (needs /root/reference at build time)")
    res = subprocess.run([BIN, os.path.join(ROOT, "tests", "golden", "tls4_flat.txt")], capture_output=True, text=True, timeout=500)
    print(res.stdout[-2000:], res.stderr[-4000:])
    assert res.returncode == 0, res.stderr[-4000:]
    assert "0 failures" in res.stdout
