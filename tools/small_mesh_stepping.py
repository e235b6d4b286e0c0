#!/usr/bin/env python
"""Meshes that do not fill the GPU are bound by launch / dependency latency: time one ERK33 cycle with
the ways of driving it -- reference control flow (host), device-resident cycle, CUDA graph replay, each
with the separate kernels (RYUJIN_B200_FUSED=0) and with one persistent kernel per update
(kernels_fused.cuh, RYUJIN_B200_FUSED=1) -- over a range of mesh sizes (where does the fused kernel stop
paying?).  Checksums must agree between all modes of a mesh (bit-identical paths)."""
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

out = {}
for case, n in MESHES:
    sums = set()
    for fused in ("0", "1"):
        os.environ["RYUJIN_B200_FUSED"] = fused
        for mode in (rj.Stepping.host, rj.Stepping.device, rj.Stepping.graph):
            if fused == "1" and mode == rj.Stepping.host:
                continue  # the reference control flow always takes the separate kernels
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
            t0 = time.perf_counter()
            for _ in range(cycles):
                t += ti.step(state, t)
            module.synchronize()
            dt = (time.perf_counter() - t0) / cycles
            key = "%s %s %s%s" % (case, "x".join(map(str, n)), mode.name, " fused" if fused == "1" else "")
            checksum = float(np.sum(state.download()))
            sums.add(checksum)
            out[key] = dict(n_dofs=od.host.n_owned, us_per_cycle=dt * 1e6, mdof_updates_per_s=od.host.n_owned * 3 / dt / 1e6,
                            checksum=checksum)
            print("%-36s %8d dofs %9.1f us/cycle %8.1f MDoF-updates/s" % (key, od.host.n_owned, dt * 1e6, out[key]["mdof_updates_per_s"]),
                  flush=True)
            del ti, state, module, od
    assert len(sums) == 1, ("modes disagree", case, n, sums)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/small_mesh_stepping.json", "w"), indent=1)
