"""CI coverage of the CUDA sources without a GPU.

tests/host_emulation/ compiles the product's own kernel and host sources (ryujin_b200/csrc/*.cu,
*.cuh -- not a copy, not a port) with g++ against an emulated CUDA runtime: a kernel launch runs
its blocks sequentially, the threads of a block are coroutines, so __syncthreads / __shfl_xor_sync
keep their meaning.  The resulting library is loaded through RYUJIN_B200_LIB in a subprocess and
the `-m gpu` parity tests (GPU vs oracle, reference goldens, prm decks) run against it.

This is test infrastructure only: the product library is built by nvcc for sm_100a, has no host
path for the kernels and fails with RJ_ERR_NO_DEVICE without a GPU
(tests/test_offline_data.py::test_no_cpu_fallback).  On the GPU box the same tests run against
the real library.  Stream capture / graph replay and NCCL (ranks as host threads) are emulated too.
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "host_emulation")
EMU_LIB = os.path.join(EMU_DIR, "libryujin_b200_hostemu.so")

# minutes-long on one host thread (66k..450k DoFs for hundreds of updates): left to the GPU
HEAVY = ("conservation_and_symmetry or config1 or vortex_3d_cycles or golden_on_gpu "
         "or verification_goldens or decks_reproduce or config2_vortex or benchmark_size or config5_dam "
         "or non_fma_build or device_path_equals_host or cylinder_o_grid or (euler_aeos and test_cycles and sod3d)")


def build():
    subprocess.check_call(["make", "-s", "-C", EMU_DIR], stdout=subprocess.DEVNULL)
    assert os.path.exists(EMU_LIB)


def run_gpu_tests(selection, timeout=900):
    env = dict(os.environ, RYUJIN_B200_LIB=EMU_LIB)  # default stepping: CUDA graph replay (capture is emulated)
    cmd = [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "-q", "-m", "gpu", "-k", selection,
           "-p", "no:cacheprovider"]
    try:  # the emulated tests are independent processes' worth of work: spread them over a few cores
        import xdist  # noqa: F401

        workers = max(1, min(4, (os.cpu_count() or 2) // 2))
        cmd += ["-n", str(workers)]
        env["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 2) // workers))  # the oracle is OpenMP parallel
    except ImportError:
        pass
    out = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    return out.returncode, out.stdout + out.stderr


def test_parity_suite_against_the_host_emulation_of_the_cuda_sources():
    build()
    rc, text = run_gpu_tests("not (%s)" % HEAVY)
    summary = text.strip().splitlines()[-1] if text.strip() else ""
    assert rc == 0, text[-4000:]
    m = re.search(r"(\d+) passed", summary)
    assert m and int(m.group(1)) >= 30, summary
    assert "failed" not in summary


def test_one_reference_deck_and_golden_against_the_host_emulation():
    """the smallest of the reference's end-to-end goldens (33^2 dofs, 180 cycles) through the prm
    front-end, and the l5 verification golden through the TimeIntegrator"""
    build()
    rc, text = run_gpu_tests("decks_reproduce and 0] or (verification_goldens and 5-erk)")
    assert rc == 0, text[-4000:]
    assert re.search(r"2 passed", text), text[-500:]


def run_fuzz(script, seed, cases, order=""):
    import json

    env = dict(os.environ, RYUJIN_B200_LIB=EMU_LIB, RJ_EMU_ORDER=order)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fuzz", script), str(seed), str(cases)], cwd=ROOT, env=env,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    lines = out.stdout.strip().splitlines()
    assert not [l for l in lines[:-1] if l.startswith("case")], "\n".join(lines[:-1])
    return json.loads(lines[-1])


def test_fuzz_euler_kernels_against_the_oracle():
    """random meshes (masked, dim 1-3), random boundary ids, 1e6 contrasts: every intermediate of
    one update and the states after step<0>, step<1> (tests/fuzz/fuzz_euler.py).  Kernel source and
    oracle execute the same IEEE operations in the same order, with the same pow (the reference's
    VCL; libm-rounded where the reference itself calls std::pow): bit-identical under the emulation."""
    build()
    worst = run_fuzz("fuzz_euler.py", 11, 16)
    assert all(v == 0.0 for v in worst.values()), worst


def test_fuzz_shallow_water_kernels_against_the_oracle():
    """dry patches, bathymetry, friction, all boundary ids, both limiter flags: the kernel source and
    the oracle execute the same operations in the same order -- under the host emulation (same
    compiler, same libm) the results are bit-identical"""
    build()
    worst = run_fuzz("fuzz_shallow_water.py", 12, 16)
    assert all(v == 0.0 for v in worst.values()), worst


def test_multi_rank_module_path_with_ranks_as_threads():
    """SURVEY 8(e) on a CPU box: 3 ranks (one of them interior, two neighbours) as host threads,
    each with its own emulated device, NCCL emulated in process -- the gathered state against the
    single-rank run (the reference's partition round-off), overlap on / off bitwise, identical tau
    and restart counters on every rank (tests/host_emulation/multi_rank_threads.py)."""
    build()
    env = dict(os.environ, RYUJIN_B200_LIB=EMU_LIB)  # graph stepping: the NCCL calls are captured and replayed
    out = subprocess.run([sys.executable, os.path.join(EMU_DIR, "multi_rank_threads.py"), "3"], cwd=ROOT, env=env,
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "multi-rank parity ok" in out.stdout
    assert out.stdout.count("overlap on/off bit-identical True") == 5
    assert out.stdout.count("vs oracle with the same partition") == 5 and "bit-identical False" not in out.stdout


def test_block_decomposition_with_ranks_as_threads():
    """BASELINE config 4 asks for 2 x 2 x 2 blocks: the generator's block partition (face, edge and
    corner neighbours) with 8 ranks as host threads, every case bit-identical to the oracle run on
    the same partition"""
    build()
    env = dict(os.environ, RYUJIN_B200_LIB=EMU_LIB)
    out = subprocess.run([sys.executable, os.path.join(EMU_DIR, "multi_rank_threads.py"), "8", "blocks"], cwd=ROOT, env=env,
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "multi-rank parity ok" in out.stdout
    assert out.stdout.count("vs oracle with the same partition") == 4 and "bit-identical False" not in out.stdout


def test_euler_aeos_with_ranks_as_threads():
    """the euler_aeos module on 3 ranks: the second precomputation cycle runs on owned rows and the
    precomputed vector (gamma_min, s, eta of the ghost rows) travels through the halo exchange --
    slabs and a block partition, strict and non-strict bounds, bit-identical to the oracle run on the
    same partition"""
    build()
    env = dict(os.environ, RYUJIN_B200_LIB=EMU_LIB)
    out = subprocess.run([sys.executable, os.path.join(EMU_DIR, "multi_rank_threads.py"), "3", "aeos"], cwd=ROOT, env=env,
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "multi-rank parity ok" in out.stdout
    assert out.stdout.count("vs oracle with the same partition") == 3 and "bit-identical False" not in out.stdout


def test_results_do_not_depend_on_the_block_and_thread_order():
    """a poor man's race detector: the emulator runs the blocks of every launch last-to-first with
    the threads of a block started from the highest index (RJ_EMU_ORDER=reverse), and in a shuffled
    block order; a kernel in which one thread consumes what another thread of the same launch
    produces would give different bits.  Shallow water: still bit-identical to the oracle; Euler:
    the same deviations to the last digit as in forward order."""
    build()
    forward = run_fuzz("fuzz_euler.py", 21, 10)
    for order in ("reverse", "shuffle"):
        assert all(v == 0.0 for v in run_fuzz("fuzz_shallow_water.py", 22, 10, order).values()), order
        assert run_fuzz("fuzz_euler.py", 21, 10, order) == forward, order
