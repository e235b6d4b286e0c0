"""Device-side Postprocessor (source/postprocessor.template.h:113-281): schlieren of the
density and vorticity of the momentum on the exponential scale against the oracle's restatement."""
import numpy as np
import pytest

from helpers import make_oracle, make_product, matching_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case,n_cells", [("sod2d", [48, 20]), ("vortex2d", [32, 32]), ("vortex3d", [16, 16, 6]), ("sod1d", [100])])
def test_schlieren_and_vorticity(case, n_cells):
    import ryujin_b200 as rj

    oracle = matching_oracle()[0]
    od, module = make_product(case, n_cells)
    off, sim = make_oracle(oracle, case, od.host)
    ti = rj.TimeIntegrator(module)
    state = module.state_vector()
    state.interpolate(0.0)
    t = 0.0
    for _ in range(5):
        t += ti.step(state, t)
    U = state.download()
    dim = od.host.dim
    post = rj.Postprocessor(module, schlieren_beta=10.0, schlieren_quantities=["rho"],
                            vorticity_quantities=["m"] if dim > 1 else [])
    result = post.compute(state)
    ref, bounds = off.postprocess(U, 0, 0, 10.0)
    assert np.allclose(post.bounds["schlieren_rho"], bounds, rtol=1e-13, atol=0)
    assert np.max(np.abs(result["schlieren_rho"] - ref)) < 1e-13
    assert 0.0 <= result["schlieren_rho"].min() and result["schlieren_rho"].max() <= 1.0
    if case == "sod2d":
        assert result["schlieren_rho"].max() > 0.9  # the shock shows up
    if dim > 1:
        ref, bounds = off.postprocess(U, 1, 1, 10.0)
        assert np.allclose(post.bounds["vorticity_m"], bounds, rtol=1e-12, atol=1e-300)
        assert np.max(np.abs(result["vorticity_m"] - ref)) < 1e-12
