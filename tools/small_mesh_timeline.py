#!/usr/bin/env python
"""Where does a cycle of a SMALL mesh go (BASELINE config 1: 2D Sod tube, 257^2 = 66k DoFs, ERK33)?

  python tools/small_mesh_timeline.py             # us per cycle: graph replay / device-resident stepping
  ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv \
      --log-file gpurun_out/small_mesh_launches.csv python tools/small_mesh_timeline.py --cycles 6 --stepping device

The second form lists every kernel of a few cycles with warm caches (no cache flush between the replays):
sum of the kernel durations of one cycle vs the measured time per cycle = the share a cooperative
one-kernel-per-update path could remove at best (launch gaps), see DESIGN.md section 7.
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))


def main():
    import torch

    import ryujin_b200 as rj
    from helpers import make_product

    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=256)
    ap.add_argument("--cycles", type=int, default=300)
    ap.add_argument("--stepping", default="graph,device")
    args = ap.parse_args()

    for name in args.stepping.split(","):
        od, module = make_product("sod2d", [args.cells, args.cells])
        ti = rj.TimeIntegrator(module, rj.TimeSteppingScheme.erk_33, 0.45, 0.9)
        ti.stepping = getattr(rj.Stepping, name)
        state = module.state_vector()
        state.interpolate(0.0)
        t = 0.0
        for _ in range(10):
            t += ti.step(state, t)
        torch.cuda.synchronize()
        wall = time.perf_counter()
        for _ in range(args.cycles):
            t += ti.step(state, t)
        torch.cuda.synchronize()
        wall = time.perf_counter() - wall
        print("%s: %d DoFs, %d cycles, %.1f us per ERK33 cycle (wall clock around the loop incl. the tau read-back), "
              "graphs captured / launched %s" % (name, module.n_owned, args.cycles, 1e6 * wall / args.cycles, ti.graph_stats()),
              flush=True)


if __name__ == "__main__":
    main()
