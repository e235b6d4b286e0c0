"""Randomised single-update comparison of the product (GPU, or the host emulation of the CUDA
sources selected through RYUJIN_B200_LIB) with the oracle: Euler, dim 1-3: random masked meshes, random boundary ids, states with density / pressure contrasts up to 1e6, random cfl; step<0> and step<1>.

    python tests/fuzz/fuzz_euler.py <seed> <cases>      -> last line: JSON of the worst deviations
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
ids_pool = [B.slip, B.dirichlet, B.do_nothing, B.no_slip, B.dynamic]
worst = {}
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 50
for case in range(n_cases):
    dim = int(rng.integers(1, 4))
    n = [int(rng.integers(3, 9)) for _ in range(dim)]
    mask = None
    if dim == 2 and rng.random() < 0.5:
        mask = (rng.random((n[1], n[0])) > 0.25).astype(np.uint8)
        mask[0, 0] = 1
    ids = [int(rng.choice(ids_pool)) for _ in range(2 * dim)]
    mesh = rj.CartesianMesh(n, [0.0] * dim, [1.0 + rng.random() for _ in range(dim)], ids, cell_mask=mask,
                            masked_boundary_id=int(rng.choice([B.slip, B.no_slip])))
    try:
        od = rj.OfflineData.from_mesh(mesh)
    except Exception as e:
        print("mesh rejected", e); continue
    host = od.host
    prim = [1.0 + rng.random(), rng.normal(), 0.5 + rng.random()]
    module = rj.HyperbolicModule(od, rj.HyperbolicSystem(), rj.InitialValues(rj.InitialState.uniform, primitive=prim,
                                 direction=[1.0] + [0.0] * (dim - 1)))
    off = oracle_offline_from_host(oracle, host)
    sim = off.euler_sim(kind=orc.UNIFORM, primitive=tuple(prim), direction=tuple([1.0] + [0.0] * (dim - 1)))
    # random admissible state with strong contrasts (up to 1e6 in density and pressure)
    N = host.n_owned
    rho = 10.0 ** rng.uniform(-3, 3, N)
    vel = rng.normal(size=(N, dim)) * 10.0 ** rng.uniform(-2, 1, (N, 1))
    p = 10.0 ** rng.uniform(-3, 3, N)
    U = np.zeros((N, dim + 2)); U[:, 0] = rho; U[:, 1:1 + dim] = rho[:, None] * vel
    U[:, -1] = p / 0.4 + 0.5 * rho * np.sum(vel * vel, axis=1)
    S = [module.state_vector() for _ in range(4)]
    S[0].upload(U); sim.set_U(U, 0)
    cfl = 0.2 + 0.6 * rng.random(); module.cfl = cfl; sim.set_cfl(cfl)
    module.prepare_state_vector(S[0], 0.1); sim.prepare_state_vector(0, 0.1)
    tau = module.step(S[0], [], [], S[1]); tau_ref = sim.step(0, [], [], 1)
    k = dim + 2
    errs = dict(
        tau=abs(tau - tau_ref) / tau_ref,
        alpha=np.max(np.abs(module.scratch("alpha") - sim.scratch("alpha"))),
        dij=np.max(np.abs(module.scratch("dij") - sim.scratch("dij"))) / np.max(np.abs(sim.scratch("dij"))),
        bounds=np.max(rel_linf_per_component(module.scratch("bounds").reshape(-1, 3), sim.scratch("bounds").reshape(-1, 3))),
        r=np.max(rel_linf_per_component(module.scratch("r").reshape(-1, k), sim.scratch("r").reshape(-1, k))),
        U1=rel_linf(S[1].download(), sim.get_U(1)),
        counters=float((module.n_restarts(), module.n_warnings()) != sim.counters()),
    )
    # second update with a stage
    module.prepare_state_vector(S[1], 0.1 + tau); sim.prepare_state_vector(1, 0.1 + tau_ref)
    try:
        module.step(S[1], [S[0]], [-1.0], S[2], tau); sim.step(1, [0], [-1.0], 2, tau_ref)
        errs["U2"] = rel_linf(S[2].download(), sim.get_U(2))
    except Exception as e:
        errs["U2"] = -1.0
    for key, v in errs.items():
        worst[key] = max(worst.get(key, 0.0), float(v))
    if max(v for kk, v in errs.items() if kk not in ("U2",)) > 1e-9:
        print("case", case, "dim", dim, n, ids, errs)
import json
print(json.dumps(worst))
