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


def block_partition(n):
    """ranks per direction for n GPUs: 1x1x2, 1x2x2, 2x2x2 (BASELINE config 4: 2 x 2 x 2 blocks)"""
    p = [1, 1, 1]
    d = 2
    while n > 1:
        if n % 2:
            raise SystemExit("block partition: the number of GPUs must be a power of two")
        p[d] *= 2
        n //= 2
        d = (d - 1) % 3
    return p


def workload_mesh(n_gpus, nodes=NODES, scaling="weak", partition="slabs"):
    """weak: `nodes`^3 DoFs PER GPU (slabs: nodes x nodes x nodes*N; blocks: nodes*p per direction);
    strong: `nodes`^3 DoFs in total, split into slabs or blocks"""
    import ryujin_b200 as rj

    B = rj.Boundary
    part = block_partition(n_gpus) if partition == "blocks" else [1, 1, n_gpus]
    if scaling == "weak":
        n_nodes = [nodes * p for p in part]
    else:
        n_nodes = [nodes] * 3
    cells = [n - 1 for n in n_nodes]
    h = 10.0 / (nodes - 1)
    return rj.CartesianMesh(cells, [-5.0, -5.0, 0.0], [h * c for c in cells], [B.dirichlet] * 6,
                            partition=part if n_gpus > 1 else None)


def initial_values():
    import ryujin_b200 as rj

    return rj.InitialValues(rj.InitialState.isentropic_vortex, direction=[1.0, 1.0, 0.0], position=[-1.0, -1.0, 0.0],
                            mach_number=1.0, beta=5.0)


def host_thread_policy():
    """Fixed thread policy of the CPU arm: the CPUs this process may actually use -- the cgroup quota
    (the GPU boxes show 128 hardware threads but grant 16 CPUs), else the affinity mask -- capped at 64.
    No sweep: the same box gives the same count on every run."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(round(int(quota) / int(period)))))
    except (OSError, ValueError):
        pass
    return max(1, min(n, 64))


class CpuArm:
    """The reference path on the host cores: the CPU restatement (oracle/) linked with the reference's
    own pow (VCL, FMA form: oracle/_ref/liboracle_vcl_fma.so, built -march=x86-64-v3), FTZ/DAZ as
    source/main.cc:26-36, OpenMP over rows.  The mesh comes from the oracle's own assembly
    (orc_mesh.h): this arm never loads the product library.  Rows are processed one at a time
    (scalar port, not the reference's SIMD groups of 4-8 rows) -- a reported baseline only.

    Workload = the bench workload (same vortex, same boundary data, ERK33 at cfl 0.2) on a bounded
    sample mesh of `nodes`^3 DoFs; a step is ONE forward-Euler update of it (the three updates of an
    ERK33 cycle in rotation, so stage costs average as in MQ/s)."""

    def __init__(self, nodes, timing=True):
        from oracle import orc

        self.orc = orc
        # timing: the -O3 / -march / contraction-on build (what a production build of the reference
        # gets); checking: the contraction-free build the product is bit-comparable with
        variant = "native" if timing else "vcl_fma"
        if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "liboracle_%s.so" % variant)):
            variant = "auto"
        self.oracle = orc.Oracle(variant)
        self.threads_available = self.oracle.max_threads()
        self.threads = min(host_thread_policy(), self.threads_available) if self.threads_available > 1 else 1
        self.oracle.set_threads(self.threads)
        self.ftz_threads = self.oracle.flush_denormals_to_zero(timing)
        t0 = time.perf_counter()
        h = 10.0 / (nodes - 1)
        self.off = self.oracle.assemble([nodes - 1] * 3, [-5.0, -5.0, 0.0], [h] * 3, [orc.DIRICHLET] * 6)
        self.sim = self.off.euler_sim(kind=orc.ISENTROPIC_VORTEX, direction=(1.0, 1.0, 0.0), position=(-1.0, -1.0, 0.0),
                                      mach=1.0, beta=5.0)
        self.sim.set_cfl(CFL)
        self.sim.interpolate(0.0, 0)
        self.sim_initial = self.sim.get_U(0).reshape(self.off.n, -1).copy()
        self.n_dofs = self.off.n
        self.nodes = nodes
        self.setup_s = time.perf_counter() - t0
        self.t, self.tau, self.phase = 0.0, 0.0, 0
        self.updates = 0

    def update(self):
        """one forward-Euler update: the substeps of ERK33 (time_integrator.template.h:373-403) in
        rotation on the oracle's state slots 0..3"""
        sim = self.sim
        if self.phase == 0:
            sim.prepare_state_vector(0, self.t)
            self.tau = sim.step(0, [], [], 1)
        elif self.phase == 1:
            sim.prepare_state_vector(1, self.t + self.tau)
            sim.step(1, [0], [-1.0], 2, self.tau)
        else:
            sim.prepare_state_vector(2, self.t + 2.0 * self.tau)
            sim.step(2, [0, 1], [0.75, -2.0], 3, self.tau)
            sim.set_U(sim.get_U(3), 0)
            self.t += 3.0 * self.tau
        self.phase = (self.phase + 1) % 3
        self.updates += 1

    def run(self, warmup, steps):
        for _ in range(warmup):
            self.update()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.update()
        dt = time.perf_counter() - t0
        value = self.n_dofs * steps / dt / 1e6
        return {
            "value": value,
            "unit": "MDoF-updates/s",
            "cores": self.threads,
            "kind": "port",
            "pow": self.oracle.kind + (" (FMA form)" if "_ref" in self.oracle.path and "vcl.so" not in self.oracle.path else ""),
            "build": "-O3 -march=x86-64-v3, FP contraction on" if "native" in self.oracle.path else "-O2, contraction off",
            "library": os.path.relpath(self.oracle.path, ROOT),
            "simd": "scalar rows (the reference processes SIMD groups of 4-8 rows)",
            "ftz_daz": self.ftz_threads > 0,
            "sample": "same workload on a %d^3 = %d DoF block of the mesh, %d forward-Euler updates (ERK33 substeps in rotation) "
                      "after %d warm-up updates, in %.2f s" % (self.nodes, self.n_dofs, steps, warmup, dt),
            "per_core": value / self.threads,
            "reference_prm_comment_per_core": 0.28,
            "threads_available": self.threads_available,
            "thread_policy": "cgroup quota / affinity, capped at 64; no sweep",
            "setup_seconds": self.setup_s,
        }, dt / steps


def workload_config(nodes, world, scaling="weak", partition="slabs"):
    part = block_partition(world) if partition == "blocks" else [1, 1, world]
    n_nodes = [nodes * p for p in part] if scaling == "weak" else [nodes] * 3
    n_dofs = n_nodes[0] * n_nodes[1] * n_nodes[2]
    return {
        "workload": "BASELINE configs[%d]: 3D Euler isentropic vortex (mach 1, beta 5, gamma 1.4), uniform hex mesh "
                    "%d x %d x %d nodes, time dependent dirichlet on all faces, ERK33 (cfl %.1f), limiter iterations 2"
                    % (1 if world == 1 else 3, n_nodes[0], n_nodes[1], n_nodes[2], CFL),
        "n_dofs": n_dofs,
        "n_dofs_per_gpu": n_dofs // world,
        "substeps_per_step": SUBSTEPS,
        "parallelism": ("%d x %d x %d blocks (%s), NCCL ghost exchange" % (part[0], part[1], part[2], partition)) if world > 1
                       else "single GPU",
    }


def run_reference(args):
    """CPU arm: the reference's path on the host cores (oracle; the reference itself needs
    deal.II + MPI + p4est and cannot be built here).  Same config / metric / unit as our
    arm; each step is a bounded sample of the workload (CpuArm)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arm = CpuArm(args.cpu_nodes)
    base, sec_per_step = arm.run(args.warmup, max(args.steps, 1))
    line = {
        "impl": "reference",
        "metric": "MDoF-updates/s",
        "value": base["value"],
        "unit": "MDoF-updates/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": sec_per_step * 1e3,
        "higher_is_better": True,
        "scaling": args.scaling,
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args.nodes, args.gpus, args.scaling, args.partition),
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

    mesh = workload_mesh(world, args.nodes, args.scaling, args.partition)
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
    ms_e2e_serial = timed(cycle_e2e, e2e_steps)
    e2e_serial_value = n_global * SUBSTEPS * e2e_steps / (ms_e2e_serial * 1e-3) / 1e6

    # the same work through the streaming transfers of the public API (rj_state_upload_begin/_end,
    # rj_state_download_begin/_end): every step still uploads its input and downloads its result, but
    # the upload of step n+1 and the download of step n-1 run on the copy streams while step n computes
    host_out2 = torch.empty((n_local, k), dtype=torch.float64).pin_memory()
    outs = (host_out, host_out2)

    def pipeline_e2e(steps):
        def run():
            state.upload_begin(host_in.data_ptr())
            for n in range(steps):
                state.upload_end()
                if n + 1 < steps:
                    state.upload_begin(host_in.data_ptr())
                t[0] += ti.step(state, t[0])
                if n > 0:
                    state.download_end()
                state.download_begin(outs[n % 2].data_ptr())
            state.download_end()  # orders the module's stream behind the last copy: inside the timed region

        return run

    pipeline_e2e(2)()
    ms_e2e = timed(pipeline_e2e(e2e_steps), 1)
    e2e_value = n_global * SUBSTEPS * e2e_steps / (ms_e2e * 1e-3) / 1e6
    host_out = outs[(e2e_steps - 1) % 2]
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
    # per UPDATE, not per launch: with N > 1 every exchanged sweep runs as two launches (export rows,
    # then the interior rows), so bytes of ALL owned rows go with the summed time of both launches
    n_updates = prof_steps * SUBSTEPS
    bytes_per_update = (per_pass[dominant] + extra) * n_local
    achieved = bytes_per_update / (dom_ms / n_updates * 1e-3) / 1e9
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
        "ms_per_update": dom_ms / n_updates,
        "launches_per_update": dom_count / n_updates,
        "share_of_step": dom_ms / total_prof_ms,
        "whole_update": {
            "algorithmic_bytes_per_dof": b_alg,
            "achieved": value * 1e6 / world * b_alg / 1e9,
            "frac": value * 1e6 / world * b_alg / 1e9 / peak,
            "note": "MDoF-updates/s per GPU x B_alg (BASELINE.md section 3)",
        },
        "kernel_ms_per_update": {n: v[0] / (prof_steps * SUBSTEPS) for n, v in kernel_ms.items()},
    }
    # second roofline: the two FP64-issue bound sweeps (steps 2 and 4) against the FP64 pipe.
    # Peak = 148 SMs x 64 FP64 lanes x clock (DFMA counted as 2 flop); the kernels execute mul and add
    # separately (-fmad=false, bit parity with the reference), so the instruction-issue ceiling is the
    # meaningful one: warp-instructions/s = 148 x 2 per clock.
    fp64_file = os.path.join(ROOT, "profiles", "fp64_instructions.json")
    if os.path.exists(fp64_file) and clocks and clocks.get("sm_mhz"):
        with open(fp64_file) as f:
            counts = json.load(f)
        issue_peak = 148 * 2 * clocks["sm_mhz"] * 1e6  # FP64 warp-instructions per second
        fp64 = {}
        for name in ("dij", "low_order"):
            if name in counts and name in kernel_ms:
                per_dof = counts[name]["fp64_warp_instructions_per_dof"]
                rate = per_dof * n_local / (kernel_ms[name][0] / n_updates * 1e-3)
                fp64[name] = {"bound": "fp64-issue", "achieved": rate / 1e9, "peak": issue_peak / 1e9, "unit": "G warp-inst/s",
                              "frac": rate / issue_peak, "source": counts[name].get("source")}
        roofline["fp64"] = fp64

    base, parity = None, None
    if not args.no_cpu_baseline:
        n_cpu = 3 * max(1, args.cpu_cycles)
        base, _ = CpuArm(args.cpu_nodes).run(3, n_cpu)
        # the result of the GPU path is CHECKED against the CPU oracle (north_star: 1e-12), not merely
        # finite: the same ERK33 cycles from the same initial state on a block of the mesh, oracle build
        # with contraction off (the parity target of the test suite)
        arm = CpuArm(args.check_nodes, timing=False)
        arm.run(0, 3 * args.check_cycles)
        n_cycles = args.check_cycles
        mesh_s = workload_mesh(1, args.check_nodes)
        od_s = rj.OfflineData.from_mesh(mesh_s, device=local_rank)
        module_s = rj.HyperbolicModule(od_s, rj.HyperbolicSystem(), initial_values(), rj.ModuleParameters())
        ti_s = rj.TimeIntegrator(module_s, rj.TimeSteppingScheme.erk_33, CFL, CFL, rj.CFLRecoveryStrategy.none)
        st = module_s.state_vector()
        st.upload(arm.sim_initial)
        t_s = 0.0
        for _ in range(n_cycles):
            t_s += ti_s.step(st, t_s)
        U_gpu, U_cpu = st.download(), arm.sim.get_U(0).reshape(-1, k)
        scale = float(np.max(np.abs(U_cpu)))
        parity = {"mesh": "%d^3" % args.check_nodes, "erk33_cycles": n_cycles, "rel_linf_vs_oracle": float(np.max(np.abs(U_gpu - U_cpu)) / scale),
                  "t_gpu": t_s, "t_oracle": arm.t, "oracle": os.path.relpath(arm.oracle.path, ROOT),
                  "tolerance": 1e-12}
        parity["ok"] = bool(parity["rel_linf_vs_oracle"] <= 1e-12)

    line = {
        "metric": "MDoF-updates/s",
        "value": value,
        "unit": "MDoF-updates/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms / args.steps,
        "higher_is_better": True,
        "scaling": args.scaling,
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args.nodes, world, args.scaling, args.partition),
        "run_info": {
            "l2": "inputs (%.1f GB of stencil + state per GPU) are far larger than the 126 MB L2; no flush needed"
                  % (od.n_entries_owned * (8 * (3 + 1 + 1 + 2 + k) + 8) / 1e9),
            "setup_seconds": t_setup,
            "pow": {0: "accurate", 1: "vcl", 2: "vcl_fma"}[int(rj.lib.rj_pow_implementation())],
            "tile_mask": os.environ.get("RYUJIN_B200_TILE", "default"),
        },
        "clocks": clocks,
        "e2e": {
            "value": e2e_value,
            "unit": "MDoF-updates/s",
            "h2d_bytes_per_step": n_local * k * 8,
            "d2h_bytes_per_step": n_local * k * 8,
            "steps": e2e_steps,
            "ms_per_step": ms_e2e / e2e_steps,
            "transfers": "streaming (rj_state_upload_begin/_end, rj_state_download_begin/_end): every step uploads "
                         "its input and downloads its result; the copies of neighbouring steps overlap with the sweeps",
            "blocking_transfers": {"value": e2e_serial_value, "ms_per_step": ms_e2e_serial / e2e_steps,
                                   "note": "upload, step, download in sequence (rj_state_upload / rj_state_download)"},
            "result_finite": finite,
            "result_check": parity,
        },
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": base,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        raise SystemExit("bench.py: the GPU result deviates from the oracle: %r" % parity)


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
    ap.add_argument("--nodes", type=int, default=NODES,
                    help="nodes per direction: per GPU with --scaling weak (default 200: 8.0M DoFs per GPU), of the whole mesh with "
                         "--scaling strong")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--partition", default="slabs", choices=["slabs", "blocks"],
                    help="slabs along z (default) or blocks 1x1x2 / 1x2x2 / 2x2x2 (BASELINE config 4: 400^3 in 2x2x2 blocks)")
    ap.add_argument("--cpu-nodes", type=int, default=128, help="edge of the CPU arm's sample block of the mesh")
    ap.add_argument("--cpu-cycles", type=int, default=2, help="timed ERK33 cycles of the cpu_baseline leg of our arm")
    ap.add_argument("--check-nodes", type=int, default=96, help="edge of the mesh block of the in-bench parity check")
    ap.add_argument("--check-cycles", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
