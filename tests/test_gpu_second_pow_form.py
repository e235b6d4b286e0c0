"""The parity suite on the SECOND build of the library.

ryujin_b200/libryujin_b200.so calls the reference's pow in the form VCL has on an FMA host (what every
AVX2 / AVX-512 build of the reference executes); ryujin_b200/libryujin_b200_vcl.so is the same source
built with the non-FMA form (SSE2 ... AVX builds of the reference), whose parity target is
oracle/_ref/liboracle_vcl.so -- the oracle build VERDICT r1 names.  This test re-runs the shock-flow
parity tests (BASELINE config 1 at 256 x 256 for 100 cycles, config 3, single updates with every
intermediate, non-default parameters) in a subprocess with RYUJIN_B200_LIB pointing at that build:
both forms must meet north_star's tolerance against their oracle build."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_non_fma_build_against_liboracle_vcl():
    import ryujin_b200 as rj

    if os.environ.get("RYUJIN_B200_LIB"):
        pytest.skip("already running on an alternative library build")
    second = os.path.join(os.path.dirname(rj.LIB_PATH), "libryujin_b200_vcl.so")
    assert os.path.exists(second), "build it with make -C ryujin_b200/csrc"
    env = dict(os.environ, RYUJIN_B200_LIB=second)
    select = "config1 or config3 or single_substep or stage_weights or limiter_iterations or gamma or restart"
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-q", "-x", "-s",
                          "-m", "gpu", "-k", select], cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    sys.stdout.write(out.stdout[-4000:])
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert " passed" in out.stdout
