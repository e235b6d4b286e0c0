"""Randomised single-update comparison of the product (GPU, or the host emulation of the CUDA
sources selected through RYUJIN_B200_LIB) with the oracle: shallow water, dim 1-2: random masked meshes, all boundary ids, random bathymetry, friction, dry patches (h = 0), both limiter flags; step<0> and step<1>.

    python tests/fuzz/fuzz_shallow_water.py <seed> <cases>      -> last line: JSON of the worst deviations
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import ryujin_b200 as rj
from oracle import orc
from helpers import matching_oracle, oracle_offline_from_host, rel_linf, rel_linf_per_component

oracle = matching_oracle()[0]  # the oracle build with the pow the product was built for
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
B = rj.Boundary
ids_pool = [B.slip, B.dirichlet, B.do_nothing, B.no_slip, B.dynamic, B.dirichlet_momentum]
worst = {}
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 50
for case in range(n_cases):
    dim = int(rng.integers(1, 3))
    n = [int(rng.integers(3, 10)) for _ in range(dim)]
    mask = None
    if dim == 2 and rng.random() < 0.4:
        mask = (rng.random((n[1], n[0])) > 0.25).astype(np.uint8); mask[0, 0] = 1
    ids = [int(rng.choice(ids_pool)) for _ in range(2 * dim)]
    mesh = rj.CartesianMesh(n, [0.0] * dim, [1.0 + rng.random() for _ in range(dim)], ids, cell_mask=mask,
                            masked_boundary_id=int(rng.choice([B.slip, B.no_slip])))
    od = rj.OfflineData.from_mesh(mesh)
    host = od.host
    eq = dict(gravity=9.81, manning_friction_coefficient=float(rng.choice([0.0, 0.03])), reference_water_depth=float(10 ** rng.uniform(-1, 1)),
              dry_state_relaxation_factor=float(rng.choice([2e-1, 1e-3, 0.0])))
    lim = dict(limit_on_kinetic_energy=int(rng.integers(0, 2)), limit_on_square_velocity=int(rng.integers(0, 2)))
    swp = [1.0 + rng.random(), rng.normal()]
    module = rj.HyperbolicModule(od, rj.HyperbolicSystem(rj.Equation.shallow_water, **eq),
                                 rj.InitialValues(rj.InitialState.sw_uniform, direction=[1.0] + [0.0] * (dim - 1), sw_params=swp),
                                 rj.ModuleParameters(**lim))
    off = oracle_offline_from_host(oracle, host)
    prm = dict(eq); prm.update(lim)
    sim = off.sw_sim(kind=orc.SW_UNIFORM, direction=tuple([1.0] + [0.0] * (dim - 1)), position=tuple([1.0] + [0.0] * (dim - 1)), p=tuple(swp + [0.0, 0.0]), **prm)
    N = host.n_owned
    z = 0.3 * rng.random(N) * float(rng.choice([0.0, 1.0]))
    module.set_initial_precomputed(z); sim.set_initial_precomputed(z)
    h = 10.0 ** rng.uniform(-3, 0.5, N)
    dry = rng.random(N) < 0.15
    h[dry] = 0.0
    vel = rng.normal(size=(N, dim)) * 10.0 ** rng.uniform(-2, 0.5, (N, 1))
    U = np.zeros((N, dim + 1)); U[:, 0] = h; U[:, 1:] = h[:, None] * vel
    S = [module.state_vector() for _ in range(4)]
    S[0].upload(U); sim.set_U(U, 0)
    cfl = 0.2 + 0.5 * rng.random(); module.cfl = cfl; sim.set_cfl(cfl)
    module.prepare_state_vector(S[0], 0.1); sim.prepare_state_vector(0, 0.1)
    try:
        tau_ref = sim.step(0, [], [], 1)
    except Exception as e:
        try:
            module.step(S[0], [], [], S[1]); print("case", case, "oracle raised but product did not:", e)
        except Exception:
            pass
        continue
    tau = module.step(S[0], [], [], S[1])
    k = dim + 1
    def rel(a, b):
        s = np.max(np.abs(b)); return float(np.max(np.abs(a - b)) / (s if s > 0 else 1.0))
    errs = dict(
        tau=abs(tau - tau_ref) / tau_ref,
        alpha=np.max(np.abs(module.scratch("alpha") - sim.scratch("alpha"))),
        dij=rel(module.scratch("dij"), sim.scratch("dij")),
        bounds=rel(module.scratch("bounds"), sim.scratch("bounds")),
        r=rel(module.scratch("r"), sim.scratch("r")),
        lij=np.max(np.abs(module.scratch("lij") - sim.scratch("lij"))),
        U1=rel_linf(S[1].download(), sim.get_U(1)),
        counters=float((module.n_restarts(), module.n_warnings()) != sim.counters()),
    )
    module.prepare_state_vector(S[1], 0.1 + tau); sim.prepare_state_vector(1, 0.1 + tau_ref)
    try:
        sim.step(1, [0], [-1.0], 2, tau_ref); module.step(S[1], [S[0]], [-1.0], S[2], tau)
        errs["U2"] = rel_linf(S[2].download(), sim.get_U(2))
    except Exception as e:
        errs["U2"] = -1.0
    for key, v in errs.items():
        worst[key] = max(worst.get(key, 0.0), float(v))
    if max(v for kk, v in errs.items() if kk not in ("U2", "lij")) > 1e-9 or errs["lij"] > 1e-6:
        print("case", case, "dim", dim, n, ids, eq, lim, {a: "%.2e" % b for a, b in errs.items()})
import json
print(json.dumps(worst))
