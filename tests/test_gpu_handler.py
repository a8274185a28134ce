"""GpuBoundHandler (the Minotaur::Handler adapter, minotaur_b200/handler/) against the reference's own
LinearHandler + NlPresHandler through Handler::presolveNode on real Relaxation objects.  The test binary is
built by `make -C oracle handler_test` where the reference sources exist and travels prebuilt to the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "handler_test")
QUAD_BIN = os.path.join(ROOT, "oracle", "_ref", "quad_patch_test")

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def test_handler_matches_reference_handlers():
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/handler_test not built (needs /root/reference at build time)")
    res = subprocess.run([BIN, os.path.join(ROOT, "tests", "golden", "tls4_flat.txt")], capture_output=True, text=True, timeout=500)
    print(res.stdout[-2000:], res.stderr[-4000:])
    assert res.returncode == 0, res.stderr[-4000:]
    assert "0 failures" in res.stdout


def test_patched_quad_handler_on_the_device_matches_itself_on_the_host():
    """The reference's QuadHandler with minotaur_b200/handler/quad_handler_gpu.patch applied (built from the reference's
    own sources by `make -C oracle quad_patch_test`): presolveNode with an engine context attached -- its propagation
    loop on the device through mntr_gpu_quad_presolve_node -- against the same handler without one, on the node boxes of
    tests/golden/quad_node_case.txt: verdicts equal, every bound of the feasible boxes bit for bit, the relaxation's
    bounds moved with the problem's."""
    if not os.path.exists(QUAD_BIN):
        pytest.skip("oracle/_ref/quad_patch_test not built (needs /root/reference at build time)")
    res = subprocess.run([QUAD_BIN, os.path.join(ROOT, "tests", "golden", "quad_node_case.txt")], capture_output=True, text=True,
                         timeout=300)
    print(res.stdout[-2000:], res.stderr[-2000:])
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert " 0 failures" in res.stdout
