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
