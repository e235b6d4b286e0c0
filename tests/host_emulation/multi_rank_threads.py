#!/usr/bin/env python
"""N > 1 path of the module on a CPU box: the ranks are host threads of this process, each with
its own emulated device (tests/host_emulation/), the halo exchanges C1-C9 travel through the
in-process NCCL emulation.  Same checks as tests/multi_gpu_parity.py on real GPUs: the gathered
multi-rank state against the single-rank run, and overlap on / off bitwise.

    RYUJIN_B200_LIB=tests/host_emulation/libryujin_b200_hostemu.so \
        python tests/host_emulation/multi_rank_threads.py [n_ranks]
"""
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ryujin_b200 as rj  # noqa: E402
from helpers import CASES, SW_CASES, make_mesh, make_sw_product, rel_linf, sw_bathymetry  # noqa: E402


def build(case, n_cells, rank, world, comm):
    if case in SW_CASES:
        od, module = make_sw_product(case, n_cells, comm=comm, rank=rank, n_ranks=world)
        module.set_initial_precomputed(sw_bathymetry(od.host))
        return od, module
    c = CASES[case]
    iv = dict(c["iv"])
    kind = getattr(rj.InitialState, iv.pop("kind"))
    od = rj.OfflineData.from_mesh(make_mesh(case, n_cells), rank=rank, n_ranks=world)
    return od, rj.HyperbolicModule(od, rj.HyperbolicSystem(), rj.InitialValues(kind, **iv), comm=comm)


def run_rank(case, n_cells, cycles, rank, world, uid, out, ti_kw):
    try:
        comm = rj.Communicator(uid, rank, world, 0) if world > 1 else None
        od, module = build(case, n_cells, rank, world, comm)
        ti = rj.TimeIntegrator(module, rj.TimeSteppingScheme.erk_33, **ti_kw)
        state = module.state_vector()
        state.interpolate(0.0)
        t, taus = 0.0, []
        for _ in range(cycles):
            tau = ti.step(state, t)
            taus.append(tau)
            t += tau
        host = od.host
        out[rank] = (host.global_index[: host.n_owned].astype(np.int64), state.download(), np.array(taus),
                     (module.n_restarts(), module.n_warnings()), int(host.n_global_dofs))
    except BaseException as e:  # noqa: BLE001 -- reported in the main thread
        out[rank] = e


def run(case, n_cells, cycles, world, **ti_kw):
    uid = rj.Communicator.unique_id() if world > 1 else None
    out = [None] * world
    threads = [threading.Thread(target=run_rank, args=(case, n_cells, cycles, r, world, uid, out, ti_kw)) for r in range(world)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(600)
    for r, o in enumerate(out):
        if isinstance(o, BaseException):
            raise RuntimeError("rank %d failed: %r" % (r, o))
        if o is None:
            raise RuntimeError("rank %d did not finish" % r)
    n_global = out[0][4]
    full = np.zeros((n_global, out[0][1].shape[1]))
    for idx, U, _, _, _ in out:
        full[idx] = U
    for o in out[1:]:
        assert np.array_equal(o[2], out[0][2]), "ranks disagree on tau"
        assert o[3] == out[0][3], "ranks disagree on the restart / warning counters"
    return full, out[0][2], out[0][3]


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    none = dict(cfl_min=0.2, cfl_max=0.2, cfl_recovery_strategy=rj.CFLRecoveryStrategy.none)
    ok = True
    for case, n_cells, cycles, kw, tol in (
        ("vortex3d", [10, 10, 4 * world], 4, none, 2e-10),
        ("sod2d", [24, 8 * world], 10, dict(), 1e-7),
        ("sod3d", [12, 5, 4 * world], 5, dict(), 1e-7),
        ("sw_dam2d", [16, 8 * world], 8, dict(), 1e-7),
    ):
        os.environ["RYUJIN_B200_OVERLAP"] = "1"
        U, taus, counters = run(case, n_cells, cycles, world, **kw)
        os.environ["RYUJIN_B200_OVERLAP"] = "0"
        U_serial, taus_serial, counters_serial = run(case, n_cells, cycles, world, **kw)
        same = np.array_equal(U, U_serial) and np.array_equal(taus, taus_serial) and counters == counters_serial
        U1, taus1, counters1 = run(case, n_cells, cycles, 1, **kw)
        err = rel_linf(U, U1)
        dtau = float(np.max(np.abs(taus - taus1) / taus1))
        print("%s %s x%d ranks, %d cycles: vs single rank %.3e (tau %.2e), overlap on/off bit-identical %s, counters %s"
              % (case, n_cells, world, cycles, err, dtau, same, counters == counters1), flush=True)
        ok = ok and same and err < tol and dtau < max(tol, 1e-12) and counters == counters1
    print("multi-rank parity ok" if ok else "multi-rank parity FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
