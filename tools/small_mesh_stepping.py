#!/usr/bin/env python
"""Meshes that do not fill the GPU are bound by launch / dependency latency: time one ERK33 cycle with
the ways of driving it -- reference control flow (host), device-resident cycle, CUDA graph replay -- with all
14 kernels of an update launched separately (RYUJIN_B200_FOLD=0) and with the scalar kernels folded into the
sweeps and d_ij + alpha as one launch (the default), over a range of mesh sizes.  (A third way, one
persistent cooperative kernel per update with grid-wide barriers, was built and measured in round 2 --
profiles/r2i_small_mesh_fused_kernel.txt, r2j_small_mesh_fused8.txt: 464 us against 392 us per cycle on
config 1, slower at every size -- and removed.)  Checksums must agree between all modes of a mesh (bit-identical paths)."""
import json
import os
import sys
import time

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np  # noqa: E402

import ryujin_b200 as rj  # noqa: E402
from helpers import make_product  # noqa: E402

MESHES = [("sod2d", [256, 256]), ("sod2d", [512, 512]), ("sod2d", [1024, 1024]), ("vortex3d", [32, 32, 32]),
          ("vortex3d", [48, 48, 48]), ("vortex3d", [64, 64, 64]), ("vortex3d", [100, 100, 100])]
if len(sys.argv) > 1 and sys.argv[1] == "--quick":
    MESHES = MESHES[:1] + MESHES[4:5]

CONFIGS = [  # label, environment, stepping modes
    ("", dict(RYUJIN_B200_FOLD="0"), (rj.Stepping.host, rj.Stepping.device, rj.Stepping.graph)),
    (" folded", dict(RYUJIN_B200_FOLD="1"), (rj.Stepping.device, rj.Stepping.graph)),  # the default
]
out = {}
for case, n in MESHES:
    sums = set()
    for label, env, modes in CONFIGS:
        os.environ.update(env)
        for mode in modes:
            od, module = make_product(case, n)
            ti = rj.TimeIntegrator(module, rj.TimeSteppingScheme.erk_33, 0.2, 0.2, rj.CFLRecoveryStrategy.none)
            ti.stepping = mode
            state = module.state_vector()
            state.interpolate(0.0)
            t = 0.0
            for _ in range(20):
                t += ti.step(state, t)
            module.synchronize()
            cycles = 200 if od.host.n_owned < 500000 else 60
            launches0 = module.kernel_launches()
            t0 = time.perf_counter()
            for _ in range(cycles):
                t += ti.step(state, t)
            module.synchronize()
            dt = (time.perf_counter() - t0) / cycles
            key = "%s %s %s%s" % (case, "x".join(map(str, n)), mode.name, label)
            checksum = float(np.sum(state.download()))
            sums.add(checksum)
            out[key] = dict(n_dofs=od.host.n_owned, us_per_cycle=dt * 1e6, mdof_updates_per_s=od.host.n_owned * 3 / dt / 1e6,
                            launches_per_cycle=(module.kernel_launches() - launches0) / cycles, checksum=checksum)
            print("%-36s %8d dofs %9.1f us/cycle %8.1f MDoF-updates/s %5.1f launches/cycle"
                  % (key, od.host.n_owned, dt * 1e6, out[key]["mdof_updates_per_s"], out[key]["launches_per_cycle"]), flush=True)
            del ti, state, module, od
    assert len(sums) == 1, ("modes disagree", case, n, sums)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/small_mesh_stepping.json", "w"), indent=1)
