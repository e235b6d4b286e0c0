#!/usr/bin/env python
"""bench.py -- MDoF-updates/s of the explicit hyperbolic update on B200.

    python bench.py --gpus N --steps K --warmup W            (our arm, CUDA)
    python bench.py --impl reference --gpus N --steps K ...  (CPU arm: the oracle on host cores)

Workload (BASELINE.json configs[1]): 3D Euler isentropic vortex on a uniform hex mesh of
200^3 = 8.0M DoFs per GPU, time dependent Dirichlet data on all faces, ERK33.  With N > 1
GPUs the mesh is 200 x 200 x (200 N) nodes in N z-slabs (weak scaling; N = 8 is
configs[3], 64M DoFs), ghost exchange over NCCL.

One "step" is one ERK33 cycle = 3 passes of the hot path (prepare_state_vector + step), so
that the stage-dependent cost of step<0>, step<1>, step<2> is averaged exactly as in the
reference's MQ/s counter (source/time_loop.template.h:1156-1162):
    value = n_dofs_global * 3 * K / seconds / 1e6     [MDoF-updates/s]

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

NODES = 200  # nodes per direction and GPU: 200^3 = 8.0M DoFs (configs[1])
SUBSTEPS = 3  # ERK33
CFL = 0.2  # verification prms use cfl min = cfl max = 0.2, recovery none


def algorithmic_bytes(dim, s=None):
    """Per-DoF compulsory bytes of each pass; s = interior row length."""
    k, p, b, w = dim + 2, 2, 3, 8
    d = dim
    s = s if s is not None else 3**dim
    per_pass = {
        "precompute": w * (k + p),
        # step 2 of the reference is one pass (1044 B in 3D); here it is two kernels, the split
        # re-reads cols and c_ij once more -- that extra traffic is NOT counted as algorithmic
        "dij": s * (d * w + 4 + w) + w * k,
        "alpha": w * (p + 2),
        "dii_tau": s * (2 * w + 4) + w,
        "low_order": s * (w * (d + 2 + k) + 4) + w * (3 * k + p + b + 3),
        "pij_lij": s * (2 * k * w + 2 * w + 4) + w * (b + 2 * k + 1),
        "high_order_next_lij": s * (2 * w + 4 + k * w) + w * (2 * k + b),
        "high_order_final": s * (w + 4 + k * w) + 2 * k * w,
    }
    total = sum(per_pass.values())
    assert total == w * (s * (2 * d + 5 * k + 10) + 11 * k + 3 * p + 3 * b + 7) + 24 * s
    return per_pass, total


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in self.samples:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(mx)) if mx else None,
            "reasons": sorted(reasons),
            "samples": len(sm),
        }


def workload_mesh(n_gpus, nodes=NODES):
    import ryujin_b200 as rj

    B = rj.Boundary
    cells = [nodes - 1, nodes - 1, nodes * n_gpus - 1]
    h = 10.0 / (nodes - 1)
    return rj.CartesianMesh(cells, [-5.0, -5.0, 0.0], [h * c for c in cells], [B.dirichlet] * 6)


def initial_values():
    import ryujin_b200 as rj

    return rj.InitialValues(rj.InitialState.isentropic_vortex, direction=[1.0, 1.0, 0.0], position=[-1.0, -1.0, 0.0],
                            mach_number=1.0, beta=5.0)


def cpu_baseline(nodes, cycles, threads=None, variant="auto"):
    """The oracle (CPU restatement of the reference path) on a bounded sample of the same
    workload: the same vortex on a nodes^3 mesh, `cycles` ERK33 cycles after one warm-up."""
    import ryujin_b200 as rj
    from oracle import orc

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oracle_offline_from_host

    oracle = orc.Oracle(variant)
    if threads:
        oracle.set_threads(threads)
    available = oracle.max_threads()
    host = rj.HostOfflineData(workload_mesh(1, nodes))
    off = oracle_offline_from_host(oracle, host)
    sim = off.euler_sim(kind=orc.ISENTROPIC_VORTEX, direction=(1.0, 1.0, 0.0), position=(-1.0, -1.0, 0.0), mach=1.0, beta=5.0)
    sim.configure_time_integrator("erk 33", CFL, CFL, "none")
    # initial state from the oracle's own InitialValues at the generator's positions
    from helpers import initial_state_host

    sim.set_U(initial_state_host(oracle, "vortex3d", host, 0.0), 0)
    t = sim.time_step(0.0)  # warm-up
    # The GPU boxes expose many hardware threads that are shared with other tenants: more OpenMP
    # threads is not always faster.  Time one cycle per candidate count and keep the fastest, so
    # that the CPU arm is the best the host can do, not an artefact of oversubscription.
    cores, sweep = available, {}
    if not threads:
        for cand in sorted({available, max(available // 2, 1), max(available // 4, 1), min(16, available), min(8, available)}):
            oracle.set_threads(cand)
            t0 = time.perf_counter()
            t += sim.time_step(t)
            sweep[cand] = time.perf_counter() - t0
        cores = min(sweep, key=sweep.get)
        oracle.set_threads(cores)
    t0 = time.perf_counter()
    for _ in range(cycles):
        t += sim.time_step(t)
    dt = time.perf_counter() - t0
    value = host.n_owned * SUBSTEPS * cycles / dt / 1e6
    return {
        "value": value,
        "unit": "MDoF-updates/s",
        "cores": cores,
        "kind": "port",
        "pow": oracle.kind,
        "sample": "same workload on %d^3 = %d DoFs, %d ERK33 cycles (%d updates) in %.2f s" % (nodes, host.n_owned, cycles, cycles * SUBSTEPS, dt),
        "per_core": value / cores,
        "threads_available": available,
        "thread_sweep_s_per_cycle": {str(c): round(v, 4) for c, v in sorted(sweep.items())},
    }, dt / cycles


def run_reference(args):
    """CPU arm: the reference's path on the host cores (oracle; the reference itself needs
    deal.II + MPI + p4est and cannot be built here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nodes = args.cpu_nodes
    base, sec_per_cycle = cpu_baseline(nodes, max(args.steps, 1))
    line = {
        "impl": "reference",
        "metric": "MDoF-updates/s",
        "value": base["value"],
        "unit": "MDoF-updates/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": sec_per_cycle * 1e3,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": "3D Euler isentropic vortex (mach 1, beta 5), uniform hex mesh, dirichlet on all faces, ERK33; "
                        "CPU arm runs a bounded sample of it",
            "sample_n_dofs": nodes**3,
            "substeps_per_step": SUBSTEPS,
        },
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "MDoF-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist

    import ryujin_b200 as rj

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    if rj.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; ryujin_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = rj.Communicator.from_torch_distributed(local_rank)

    mesh = workload_mesh(world, args.nodes)
    t_setup = time.perf_counter()
    od = rj.OfflineData.from_mesh(mesh, rank=rank, n_ranks=world, device=local_rank)
    module = rj.HyperbolicModule(od, rj.HyperbolicSystem(), initial_values(), rj.ModuleParameters(), comm=comm)
    ti = rj.TimeIntegrator(module, rj.TimeSteppingScheme.erk_33, CFL, CFL, rj.CFLRecoveryStrategy.none)
    state = module.state_vector()
    state.interpolate(0.0)
    module.synchronize()
    t_setup = time.perf_counter() - t_setup
    n_global = od.host.n_global_dofs
    n_local = od.host.n_owned

    stream = torch.cuda.ExternalStream(module.stream, device=torch.device("cuda", local_rank))

    def barrier():
        module.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn, steps):
        """K calls of fn bracketed by barrier + synchronize; device time, max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    t = [0.0]

    def cycle():
        t[0] += ti.step(state, t[0])

    for _ in range(args.warmup):
        cycle()

    # ---- device resident throughput ("value") ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = module.kernel_launches()
    ms = timed(cycle, args.steps)
    launches = module.kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    value = n_global * SUBSTEPS * args.steps / (ms * 1e-3) / 1e6

    # ---- per-kernel profile for the roofline (CUDA events around every launch) ----
    module.set_profiling(True)
    prof_steps = max(1, min(args.steps, 5))
    timed(cycle, prof_steps)
    prof = module.profile()
    module.set_profiling(False)

    # ---- end to end: host buffers in / out every step through the public API ----
    k = module.k
    host_in = torch.empty((n_local, k), dtype=torch.float64).pin_memory()
    host_out = torch.empty((n_local, k), dtype=torch.float64).pin_memory()
    state.download_to_pointer(host_in.data_ptr())

    def cycle_e2e():
        state.upload_from_pointer(host_in.data_ptr())
        t[0] += ti.step(state, t[0])
        state.download_to_pointer(host_out.data_ptr())

    e2e_steps = max(1, min(args.steps, 5))
    cycle_e2e()
    ms_e2e = timed(cycle_e2e, e2e_steps)
    e2e_value = n_global * SUBSTEPS * e2e_steps / (ms_e2e * 1e-3) / 1e6
    finite = bool(torch.isfinite(host_out).all().item())

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    per_pass, b_alg = algorithmic_bytes(3)
    peak, peak_kind = measured_peak_gbs()
    kernel_ms = {name: v for name, v in prof.items() if v[1] > 0}
    total_prof_ms = sum(v[0] for v in kernel_ms.values())
    candidates = [n for n in kernel_ms if n in per_pass and n != "alpha"]
    dominant = max(candidates, key=lambda n: kernel_ms[n][0])
    dom_ms, dom_count = kernel_ms[dominant]
    # low_order reads one extra state vector per ERK stage: (0 + 1 + 2)/3 = 1 on average
    extra = 8 * k * 1.0 if dominant == "low_order" else 0.0
    bytes_per_launch = (per_pass[dominant] + extra) * n_local
    achieved = bytes_per_launch / (dom_ms / dom_count * 1e-3) / 1e9
    traffic = None
    tfile = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tfile):
        with open(tfile) as f:
            traffic = json.load(f).get(dominant)
    roofline = {
        "bound": "hbm",
        "kernel": dominant,
        "achieved": achieved,
        "peak": peak,
        "peak_kind": peak_kind,
        "unit": "GB/s",
        "frac": achieved / peak,
        "traffic": traffic,
        "algorithmic_bytes_per_dof": per_pass[dominant] + extra,
        "avg_launch_ms": dom_ms / dom_count,
        "share_of_step": dom_ms / total_prof_ms,
        "whole_update": {
            "algorithmic_bytes_per_dof": b_alg,
            "achieved": value * 1e6 / world * b_alg / 1e9,
            "frac": value * 1e6 / world * b_alg / 1e9 / peak,
            "note": "MDoF-updates/s per GPU x B_alg (BASELINE.md section 3)",
        },
        "kernel_ms_per_update": {n: v[0] / (prof_steps * SUBSTEPS) for n, v in kernel_ms.items()},
    }

    base = None
    if not args.no_cpu_baseline:
        base, _ = cpu_baseline(args.cpu_nodes, args.cpu_cycles)

    line = {
        "metric": "MDoF-updates/s",
        "value": value,
        "unit": "MDoF-updates/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": "BASELINE configs[%d]: 3D Euler isentropic vortex (mach 1, beta 5, gamma 1.4), uniform hex mesh "
                        "%d x %d x %d nodes, time dependent dirichlet on all faces, ERK33 (cfl %.1f), limiter iterations 2"
                        % (1 if world == 1 else 3, args.nodes, args.nodes, args.nodes * world, CFL),
            "n_dofs": n_global,
            "n_dofs_per_gpu": n_local,
            "substeps_per_step": SUBSTEPS,
            "parallelism": "z-slabs x%d, NCCL ghost exchange" % world if world > 1 else "single GPU",
            "l2": "inputs (%.1f GB of stencil + state per GPU) are far larger than the 126 MB L2; no flush needed"
                  % (od.n_entries_owned * (8 * (3 + 1 + 1 + 2 + k) + 8) / 1e9),
            "setup_seconds": t_setup,
        },
        "clocks": clocks,
        "e2e": {
            "value": e2e_value,
            "unit": "MDoF-updates/s",
            "h2d_bytes_per_step": n_local * k * 8,
            "d2h_bytes_per_step": n_local * k * 8,
            "steps": e2e_steps,
            "ms_per_step": ms_e2e / e2e_steps,
            "result_finite": finite,
        },
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": base,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _quiet_stdout():
    """Libraries (NCCL prints its version banner) write to fd 1: keep the contract of ONE
    JSON line on stdout by routing fd 1 to stderr and returning a handle on the real stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    # torchrun exports OMP_NUM_THREADS=1; the host-side mesh generator and the CPU baseline are
    # OpenMP parallel: give every rank its share of the cores before the libraries load
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if "--impl" in sys.argv and "reference" in sys.argv:
        world = 1  # the CPU arm runs on rank 0 alone (the other ranks exit): it may use every core
    if os.environ.get("OMP_NUM_THREADS", "1") == "1" or world > 1:
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // max(world, 1)))
    real_stdout = _quiet_stdout()
    import builtins

    def emit(*a, **k):
        k.setdefault("file", real_stdout)
        builtins.print(*a, **k)
        real_stdout.flush()

    globals()["print"] = emit
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nodes", type=int, default=NODES, help="nodes per direction and GPU (default 200: 8.0M DoFs)")
    ap.add_argument("--cpu-nodes", type=int, default=64, help="edge of the CPU sample mesh")
    ap.add_argument("--cpu-cycles", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
