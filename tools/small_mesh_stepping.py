#!/usr/bin/env python
"""Config 1 (2D Sod 256x256) is launch bound: time one ERK33 cycle with the three ways of
driving it (reference control flow / device-resident cycle / CUDA graph replay)."""
import json
import sys
import time

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np  # noqa: E402

import ryujin_b200 as rj  # noqa: E402
from helpers import make_product  # noqa: E402

out = {}
for case, n in (("sod2d", [256, 256]), ("vortex3d", [48, 48, 48])):
    for mode in (rj.Stepping.host, rj.Stepping.device, rj.Stepping.graph):
        od, module = make_product(case, n)
        ti = rj.TimeIntegrator(module, rj.TimeSteppingScheme.erk_33, 0.2, 0.2, rj.CFLRecoveryStrategy.none)
        ti.stepping = mode
        state = module.state_vector()
        state.interpolate(0.0)
        t = 0.0
        for _ in range(20):
            t += ti.step(state, t)
        module.synchronize()
        t0 = time.perf_counter()
        cycles = 200
        for _ in range(cycles):
            t += ti.step(state, t)
        module.synchronize()
        dt = (time.perf_counter() - t0) / cycles
        out["%s %s" % (case, mode.name)] = dict(
            us_per_cycle=dt * 1e6, mdof_updates_per_s=od.host.n_owned * 3 / dt / 1e6, checksum=float(np.sum(state.download())))
        print(case, mode.name, out["%s %s" % (case, mode.name)], flush=True)
json.dump(out, open("gpurun_out/small_mesh_stepping.json", "w"), indent=1)
