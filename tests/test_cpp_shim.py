"""The header-only C++ shim (include/ryujin_b200.hpp) compiles against the C ABI and links
with the library (no GPU needed: nothing is executed beyond rj_device_count / rj_version)."""
import os
import subprocess
import sys

import ryujin_b200 as rj

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SOURCE = r"""
#include <ryujin_b200.hpp>
#include <cstdio>
using namespace ryujin_b200;
// instantiate the reference call sequence of TimeIntegrator::step_erk_33
double erk33(HyperbolicModule &hm, StateVector &s, StateVector &T0, StateVector &T1, StateVector &T2, double t, double tau_max)
{
  hm.prepare_state_vector(s, t);
  double tau = hm.step<0>(s, {}, {}, T0, 0., tau_max / 3.);
  hm.prepare_state_vector(T0, t + 1.0 * tau);
  hm.step<1>(T0, {{s}}, {{-1.}}, T1, tau);
  hm.prepare_state_vector(T1, t + 2.0 * tau);
  hm.step<2>(T1, {{s, T0}}, {{0.75, -2.}}, T2, tau);
  s.swap(T2);
  return 3. * tau;
}
// the shallow water additions and the stepping switch are part of the surface
void configure(HyperbolicModule &hm, TimeIntegrator &ti, const double *bathymetry)
{
  hm.set_initial_precomputed(bathymetry);
  ti.stepping(RJ_STEPPING_GRAPH);
  if (ti.stepping() != RJ_STEPPING_GRAPH)
    std::printf("unexpected\n");
}
int main()
{
  std::printf("%s devices=%d\n", rj_version(), rj_device_count());
  try { check(RJ_RESTART); } catch (const Restart &) { std::printf("restart ok\n"); }
  return 0;
}
"""


def test_cpp_shim_compiles_and_links(tmp_path):
    src = tmp_path / "shim.cc"
    src.write_text(SOURCE)
    exe = tmp_path / "shim"
    libdir = os.path.dirname(rj.LIB_PATH)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-lryujin_b200", "-Wl,-rpath," + libdir])
    out = subprocess.check_output([str(exe)], text=True)
    assert "ryujin_b200" in out and "restart ok" in out


# -----------------------------------------------------------------------------------------------
# the shim EXECUTED on the device: a C++ host program drives the reference's
# call sequence of TimeIntegrator::step_erk_33 (time_integrator.template.h:373-403) through
# ryujin_b200::HyperbolicModule::step<stages>(), and the shim's TimeIntegrator; both must give the
# bits of the Python host path.
GPU_SOURCE = r"""
#include <ryujin_b200.hpp>
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace ryujin_b200;

static double erk33(HyperbolicModule &hm, StateVector &s, StateVector &T0, StateVector &T1, StateVector &T2, double t)
{
  hm.prepare_state_vector(s, t);
  double tau = hm.step<0>(s, {}, {}, T0);
  hm.prepare_state_vector(T0, t + 1.0 * tau);
  hm.step<1>(T0, {{s}}, {{-1.}}, T1, tau);
  hm.prepare_state_vector(T1, t + 2.0 * tau);
  hm.step<2>(T1, {{s, T0}}, {{0.75, -2.}}, T2, tau);
  s.swap(T2);
  return 3. * tau;
}

int main(int argc, char **argv)
{
  const int cycles = std::atoi(argv[1]);
  rj_mesh_desc mesh{};
  mesh.dim = 2;
  mesh.n_cells[0] = 48; mesh.n_cells[1] = 12;
  mesh.h[0] = 1. / 48; mesh.h[1] = 1. / 12;
  const int ids[4] = {RJ_BOUNDARY_DIRICHLET, RJ_BOUNDARY_DIRICHLET, RJ_BOUNDARY_SLIP, RJ_BOUNDARY_SLIP};
  for (int f = 0; f < 4; ++f) mesh.box_boundary_ids[f] = ids[f];
  mesh.masked_boundary_id = RJ_BOUNDARY_SLIP;
  rj_host_offline *host = nullptr;
  check(rj_mesh_generate(&host, &mesh, 0, 1));
  rj_offline_desc desc{};
  check(rj_host_offline_desc(host, &desc));
  OfflineData offline_data(desc, 0);

  rj_equation_desc eq; rj_equation_desc_default(&eq, RJ_EQUATION_EULER);
  rj_module_params mp; rj_module_params_default(&mp);
  rj_initial_values_desc iv; rj_initial_values_desc_default(&iv, RJ_INITIAL_CONTRAST);
  iv.direction[0] = 1.; iv.direction[1] = 0.; iv.position[0] = 0.5; iv.position[1] = 0.;
  iv.primitive[0] = 1.; iv.primitive[1] = 0.; iv.primitive[2] = 1.;
  iv.primitive_right[0] = 0.125; iv.primitive_right[1] = 0.; iv.primitive_right[2] = 0.1;

  std::vector<double> out(size_t(desc.n_owned) * 4);
  for (int variant = 0; variant < 2; ++variant) {
    HyperbolicModule hm(offline_data, eq, iv, mp);
    StateVector s(hm), T0(hm), T1(hm), T2(hm);
    s.interpolate(0.);
    double t = 0.;
    if (variant == 0) {
      hm.cfl(0.5);
      for (int c = 0; c < cycles; ++c)
        t += erk33(hm, s, T0, T1, T2, t);
    } else {
      TimeIntegrator ti(hm, RJ_ERK_33, 0.5, 0.5, RJ_CFL_RECOVERY_NONE);
      for (int c = 0; c < cycles; ++c)
        t += ti.step(s, t);
    }
    s.download(out.data());
    std::FILE *f = std::fopen(argv[2 + variant], "wb");
    std::fwrite(&t, sizeof(double), 1, f);
    std::fwrite(out.data(), sizeof(double), out.size(), f);
    std::fclose(f);
    std::printf("variant %d: t = %.17g, restarts %u\n", variant, t, hm.n_restarts());
  }
  // a violated invariant domain surfaces as ryujin_b200::Restart, like ryujin::Restart
  {
    HyperbolicModule hm(offline_data, eq, iv, mp);
    StateVector s(hm), T0(hm);
    s.interpolate(0.);
    hm.prepare_state_vector(s, 0.);
    hm.cfl(8.0);
    hm.id_violation_strategy(IDViolationStrategy::raise_exception);
    try { hm.step<0>(s, {}, {}, T0); std::printf("no restart?\n"); }
    catch (const Restart &) { std::printf("Restart thrown, n_restarts = %u\n", hm.n_restarts()); }
  }
  rj_host_offline_destroy(host);
  return 0;
}
"""


import pytest  # noqa: E402


@pytest.mark.gpu
def test_cpp_shim_runs_the_reference_call_sequence_on_the_device(tmp_path):
    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import make_product

    src = tmp_path / "shim_gpu.cc"
    src.write_text(GPU_SOURCE)
    exe = tmp_path / "shim_gpu"
    libdir = os.path.dirname(rj.LIB_PATH)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-l:" + os.path.basename(rj.LIB_PATH), "-Wl,-rpath," + libdir])
    cycles = 12
    files = [str(tmp_path / "manual.bin"), str(tmp_path / "integrator.bin")]
    out = subprocess.check_output([str(exe), str(cycles)] + files, text=True)
    assert "Restart thrown, n_restarts = 1" in out, out

    od, module = make_product("sod2d", [48, 12])
    ti = rj.TimeIntegrator(module, rj.TimeSteppingScheme.erk_33, 0.5, 0.5, rj.CFLRecoveryStrategy.none)
    state = module.state_vector()
    state.interpolate(0.0)
    t = 0.0
    for _ in range(cycles):
        t += ti.step(state, t)
    U = state.download()
    for path in files:
        raw = np.fromfile(path, dtype=np.float64)
        assert raw[0] == t, (path, raw[0], t)
        assert np.array_equal(raw[1:].reshape(U.shape), U), path
