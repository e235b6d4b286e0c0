"""TimeLoop: drives the B200 path from one of the reference's parameter files and prints what
the reference's time loop prints (source/time_loop.template.h:218-432 run(), :694-842
compute_error(), :1108-1372 throughput block), so that the output can be compared with the
reference's golden `.output` files.  SURVEY.md 8(f) row 2.

Only what the hot path needs is honoured: sections A (time loop control, error norms),
B (equation), C (rectangular domain / step geometry as uniform Q1 grids), E (initial values that
exist on the device), F (indicator / limiter / Riemann solver parameters) and H (explicit time
integrators).  VTU output, checkpointing, mesh adaptation, the parabolic module and quantities of
interest are outside the path; entries that request them raise, all other unknown entries are
reported by `unused_entries()`.
"""
import sys
import time

import numpy as np

from . import prm as _prm
from .hyperbolic_module import (CFLRecoveryStrategy, Equation, EquationOfState, HyperbolicModule, HyperbolicSystem, InitialState,
                                InitialValues, ModuleParameters, TimeIntegrator, TimeSteppingScheme)
from .offline_data import Boundary, CartesianMesh, OfflineData

BOUNDARY = {
    "do nothing": Boundary.do_nothing, "periodic": Boundary.periodic, "slip": Boundary.slip,
    "no slip": Boundary.no_slip, "dirichlet": Boundary.dirichlet, "dynamic": Boundary.dynamic,
    "dirichlet momentum": Boundary.dirichlet_momentum,
}

SCHEMES = {
    "ssprk 22": TimeSteppingScheme.ssprk_22, "ssprk 33": TimeSteppingScheme.ssprk_33,
    "erk 11": TimeSteppingScheme.erk_11, "erk 22": TimeSteppingScheme.erk_22,
    "erk 33": TimeSteppingScheme.erk_33, "erk 43": TimeSteppingScheme.erk_43,
    "erk 54": TimeSteppingScheme.erk_54,
}

COMPONENT_NAMES = {  # View::component_names, euler/hyperbolic_system.h:240-249, shallow_water/...:207-216
    Equation.euler: {1: ["rho", "m", "E"], 2: ["rho", "m_1", "m_2", "E"], 3: ["rho", "m_1", "m_2", "m_3", "E"]},
    Equation.shallow_water: {1: ["h", "m"], 2: ["h", "m_1", "m_2"]},
    Equation.euler_aeos: {1: ["rho", "m", "E"], 2: ["rho", "m_1", "m_2", "E"], 3: ["rho", "m_1", "m_2", "m_3", "E"]},
}


class NotOnThePath(NotImplementedError):
    """the parameter file asks for something outside the hot path this library replaces"""


def _boundary(section, key, default="dirichlet"):
    name = section.get_str(key, default)
    if name not in BOUNDARY:
        raise _prm.ParameterError("unknown boundary condition »%s«" % name)
    if BOUNDARY[name] == Boundary.periodic:
        raise NotOnThePath("periodic boundary conditions need constrained rows (not generated here)")
    return BOUNDARY[name]


def q1_norms(values, n_cells, h):
    """Linf of the nodal values and L1 / L2 norms of the Q1 interpolant over the uniform grid with
    QGauss(3) per direction -- VectorTools::integrate_difference against zero, as used by
    TimeLoop::compute_error (time_loop.template.h:743-795).  values: nodal array in
    lexicographic order (x fastest)."""
    dim = len(n_cells)
    shape = tuple(n + 1 for n in reversed(n_cells))
    v = np.asarray(values, dtype=np.float64).reshape(shape)
    linf = float(np.max(np.abs(v))) if v.size else 0.0
    a = 0.5 * np.sqrt(3.0 / 5.0)
    qx = np.array([0.5 - a, 0.5, 0.5 + a])
    qw = np.array([5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0])
    # corner values per cell: list over corners c (bit d of c = upper node in direction d)
    corners = []
    for c in range(1 << dim):
        sl = []
        for axis in range(dim):  # axis 0 of v is the slowest direction (dim-1)
            d = dim - 1 - axis
            sl.append(slice(1, None) if (c >> d) & 1 else slice(0, -1))
        corners.append(v[tuple(sl)])
    cell_measure = float(np.prod(h))
    l1 = 0.0
    l2sq = 0.0
    for q in range(3**dim):
        idx = [(q // 3**d) % 3 for d in range(dim)]
        w = cell_measure * float(np.prod([qw[i] for i in idx]))
        val = 0.0
        for c in range(1 << dim):
            phi = 1.0
            for d in range(dim):
                phi *= qx[idx[d]] if (c >> d) & 1 else 1.0 - qx[idx[d]]
            val = val + phi * corners[c]
        l1 += w * float(np.sum(np.abs(val)))
        l2sq += w * float(np.sum(val * val))
    return linf, l1, float(np.sqrt(l2sq))


class TimeLoop:
    """source/time_loop.h: owns mesh, offline data, module, time integrator; `run()` is the
    reference's main loop."""

    def __init__(self, parameters, device=0, comm=None, rank=0, n_ranks=1, out=None):
        self.prm = _prm.load(parameters) if isinstance(parameters, str) else parameters
        self.out = out if out is not None else sys.stdout
        self.rank, self.n_ranks, self.comm, self.device = rank, n_ranks, comm, device
        p = self.prm

        # ---- A - TimeLoop
        A = p.subsection("A - TimeLoop")
        self.base_name = A.get_str("basename", "test")
        self.t_final = A.get_float("final time", 5.0)
        self.enforce_t_final = A.get_bool("enforce final time", False)
        self.timer_granularity = A.get_float("timer granularity", 0.01)
        self.enable_compute_error = A.get_bool("enable compute error", False)
        self.error_normalize = A.get_bool("error normalize", True)
        self.terminal_update_interval = A.get_float("terminal update interval", 5.0)
        for key in ("enable checkpointing", "enable output full", "enable output levelsets",
                    "enable mesh adaptivity", "resume"):
            if A.get_bool(key, False):
                raise NotOnThePath("A - TimeLoop/%s = true is outside the explicit hyperbolic update" % key)
        # Quantities (source/quantities.template.h): only the space averaged time series of interior
        # manifolds that select every node (level set "0."), the form the reference's regression
        # decks use; the series is printed at the end like the reference's `debug filename` dump
        self.enable_compute_quantities = A.get_bool("enable compute quantities", False)
        self.debug_filename = A.get_str("debug filename", "")
        self.time_series = []
        K = p.subsection("K - Quantities")
        manifolds = [m.strip() for m in K.get_str("interior manifolds", "").split(";") if m.strip()]
        if K.get_str("boundary manifolds", "").strip():
            raise NotOnThePath("K - Quantities/boundary manifolds")
        self.manifold_name = None
        if self.enable_compute_quantities:
            for m in manifolds:
                name, level_set, options = [x.strip() for x in m.split(":")]
                try:
                    everywhere = float(level_set) == 0.0
                except ValueError:
                    everywhere = False
                if not everywhere or "space_averaged" not in options:
                    raise NotOnThePath("K - Quantities/interior manifolds »%s«: only level set 0 with space_averaged" % m)
                self.manifold_name = name

        # ---- B - Equation
        B = p.subsection("B - Equation")
        self.dim = B.get_int("dimension", 2)
        eq_name = B.get_str("equation", "euler")
        if eq_name == "euler":
            self.equation = Equation.euler
            system = HyperbolicSystem(
                Equation.euler, gamma=B.get_float("gamma", 7.0 / 5.0), reference_density=B.get_float("reference density", 1.0),
                vacuum_state_relaxation_small=B.get_float("vacuum state relaxation small", 1.0e2),
                vacuum_state_relaxation_large=B.get_float("vacuum state relaxation large", 1.0e4))
        elif eq_name == "shallow water":
            self.equation = Equation.shallow_water
            system = HyperbolicSystem(
                Equation.shallow_water, gravity=B.get_float("gravity", 9.81),
                manning_friction_coefficient=B.get_float("manning friction coefficient", 0.0),
                reference_water_depth=B.get_float("reference water depth", 1.0),
                dry_state_relaxation_factor=B.get_float("dry state relaxation factor", 2.0e-1),
                dry_state_relaxation_small=B.get_float("dry state relaxation small", 1.0e2),
                dry_state_relaxation_large=B.get_float("dry state relaxation large", 1.0e4))
        elif eq_name == "euler aeos":
            # euler_aeos/hyperbolic_system.h:770-812 + the subsection of the selected equation of state
            # (equation_of_state_{polytropic_gas,noble_abel_stiffened_gas,van_der_waals}.h)
            self.equation = Equation.euler_aeos
            eos_name = B.get_str("equation of state", "polytropic gas")
            eos_kinds = {"polytropic gas": EquationOfState.polytropic_gas,
                         "noble abel stiffened gas": EquationOfState.noble_abel_stiffened_gas,
                         "van der waals": EquationOfState.van_der_waals}
            if eos_name not in eos_kinds:
                raise NotOnThePath("equation of state »%s« is not available on the device (polytropic gas, "
                                   "noble abel stiffened gas, van der waals)" % eos_name)
            S = B.subsection(eos_name)
            eos = dict(gamma=S.get_float("gamma", 7.0 / 5.0))
            S.get_float("gas constant R", 287.052874)  # temperature only: not on the path
            if eos_name == "noble abel stiffened gas":
                eos.update(covolume_b=S.get_float("covolume b", 0.0), reference_pressure=S.get_float("reference pressure", 0.0),
                           reference_specific_internal_energy=S.get_float("reference specific internal energy", 0.0))
            elif eos_name == "van der waals":
                eos.update(covolume_b=S.get_float("covolume b", 0.0), vdw_a=S.get_float("vdw a", 0.0))
            system = HyperbolicSystem(
                Equation.euler_aeos, equation_of_state=int(eos_kinds[eos_name]),
                compute_strict_bounds=int(B.get_bool("compute strict bounds", True)),
                reference_density=B.get_float("reference density", 1.0),
                vacuum_state_relaxation_small=B.get_float("vacuum state relaxation small", 1.0e2),
                vacuum_state_relaxation_large=B.get_float("vacuum state relaxation large", 1.0e4), **eos)
        else:
            raise NotOnThePath("equation »%s« has no B200 plug-in (euler, euler aeos, shallow water)" % eq_name)
        if self.dim not in COMPONENT_NAMES[self.equation]:
            raise NotOnThePath("%s is not instantiated for dim = %d" % (eq_name, self.dim))
        self.component_names = COMPONENT_NAMES[self.equation][self.dim]
        self.error_quantities = A.get_list("error quantities", self.component_names, convert=str)

        # ---- C - Discretization
        Cs = p.subsection("C - Discretization")
        if Cs.get_str("finite element ansatz", "cG Q1") != "cG Q1":
            raise NotOnThePath("only the continuous Q1 ansatz is on the path")
        if Cs.get_float("mesh distortion", 0.0) != 0.0:
            raise NotOnThePath("distorted meshes need the general assembly (stencil importer, SURVEY 8f)")
        refinement = Cs.get_int("mesh refinement", 5)
        geometry = Cs.get_str("geometry", "cylinder")
        if geometry == "rectangular domain":
            G = Cs.subsection("rectangular domain")
            lo = G.get_list("position bottom left", [0.0] * self.dim)
            hi = G.get_list("position top right", [20.0] * self.dim)
            if len(lo) != self.dim or len(hi) != self.dim:
                raise _prm.ParameterError("rectangular domain: corner points do not have %d components" % self.dim)
            for key, default in (("grading push forward", None), ("grading pull back", None)):
                g = G.get_str(key, default)
                if g is not None and "".join(g.split()).rstrip(";") not in ("x", "x;y", "x;y;z"):
                    raise NotOnThePath("graded meshes need the general assembly")
            names = [("x", "left", "right"), ("y", "bottom", "top"), ("z", "back", "front")][: self.dim]
            cells, ids = [], []
            for axis, lower, upper in names:
                cells.append(G.get_int("subdivisions " + axis, 1) * 2**refinement)
                ids += [_boundary(G, "boundary condition " + lower), _boundary(G, "boundary condition " + upper)]
            self.mesh = CartesianMesh.rectangular_domain(lo, hi, cells, ids)
        elif geometry == "step":
            if self.dim != 2:
                raise NotOnThePath("the step geometry is two dimensional")
            G = Cs.subsection("step")
            self.mesh = CartesianMesh.forward_facing_step(
                refinement, length=G.get_float("length", 3.0), height=G.get_float("height", 1.0),
                step_position=G.get_float("step position", 0.6), step_height=G.get_float("step height", 0.2))
        else:
            raise NotOnThePath("geometry »%s«: only uniform Q1 grids are generated here (rectangular domain, step)" % geometry)

        # ---- E - InitialValues
        E = p.subsection("E - InitialValues")
        configuration = E.get_str("configuration", "uniform")
        direction = E.get_list("direction", [1.0] + [0.0] * (self.dim - 1))
        position = E.get_list("position", [1.0] + [0.0] * (self.dim - 1))
        if E.get_float("perturbation", 0.0) != 0.0:
            raise NotOnThePath("random perturbations of the initial state are not generated on the device")
        S = E.subsection(configuration)
        if self.equation in (Equation.euler, Equation.euler_aeos):
            if configuration == "uniform":
                iv = InitialValues(InitialState.uniform, direction, position, primitive=S.get_list("primitive state", [1.4, 3.0, 1.0]))
            elif configuration == "contrast":
                iv = InitialValues(InitialState.contrast, direction, position,
                                   primitive=S.get_list("primitive state left", [1.4, 0.0, 1.0]),
                                   primitive_right=S.get_list("primitive state right", [1.4, 0.0, 1.0]))
            elif configuration == "isentropic vortex":
                iv = InitialValues(InitialState.isentropic_vortex, direction, position,
                                   mach_number=S.get_float("mach number", 2.0), beta=S.get_float("beta", 5.0))
            else:
                raise NotOnThePath("initial state »%s« is not available on the device" % configuration)
        else:
            if configuration == "uniform":
                iv = InitialValues(InitialState.sw_uniform, direction, position, sw_params=S.get_list("primitive state", [1.0, 1.0]))
            elif configuration == "circular dam break":
                iv = InitialValues(InitialState.sw_circular_dam_break, direction, position,
                                   sw_params=[S.get_float("still water depth", 0.5), S.get_float("radius", 2.5),
                                              S.get_float("dam amplitude", 2.5)])
            elif configuration == "ritter dam break":
                iv = InitialValues(InitialState.sw_ritter_dam_break, direction, position,
                                   sw_params=[S.get_float("time initial", 0.1), S.get_float("left water depth", 0.005)])
            elif configuration == "smooth vortex":
                if S.get_bool("with bathymetry", False):
                    raise NotOnThePath("smooth vortex with bathymetry: supply Z through set_initial_precomputed")
                iv = InitialValues(InitialState.sw_smooth_vortex, direction, position,
                                   sw_params=[S.get_float("reference depth", 1.0), S.get_float("mach number", 2.0),
                                              S.get_float("beta", 0.1)])
            else:
                raise NotOnThePath("initial state »%s« is not available on the device" % configuration)

        # ---- F - HyperbolicModule
        F = p.subsection("F - HyperbolicModule")
        ind, lim, rie = F.subsection("indicator"), F.subsection("limiter"), F.subsection("riemann solver")
        mp = dict(indicator_evc_factor=ind.get_float("evc factor", 1.0), limiter_iterations=lim.get_int("iterations", 2),
                  limiter_newton_tolerance=lim.get_float("newton tolerance", 1.0e-10),
                  limiter_newton_max_iterations=lim.get_int("newton max iterations", 2),
                  limiter_relaxation_factor=lim.get_float("relaxation factor", 1.0))
        if self.equation == Equation.euler:
            mp.update(riemann_newton_tolerance=rie.get_float("newton tolerance", 1.0e-10),
                      riemann_newton_max_iterations=rie.get_int("newton max iterations", 0))
        elif self.equation == Equation.shallow_water:
            mp.update(limit_on_kinetic_energy=int(lim.get_bool("limit on kinetic energy", False)),
                      limit_on_square_velocity=int(lim.get_bool("limit on square velocity", True)))

        # ---- H - TimeIntegrator
        H = p.subsection("H - TimeIntegrator")
        scheme = H.get_str("time stepping scheme", "erk 33")
        if scheme not in SCHEMES:
            raise NotOnThePath("time stepping scheme »%s« (the explicit schemes ssprk 22 ... erk 54 are on the path)" % scheme)
        self.scheme_name = scheme
        recovery = H.get_str("cfl recovery strategy", "bang bang control")
        if recovery not in ("none", "bang bang control"):
            raise _prm.ParameterError("unknown cfl recovery strategy »%s«" % recovery)

        self._info("initiating flux capacitor")
        self._info("dispatching to driver »%s« with dim=%d" % (eq_name, self.dim))
        self._info("initializing data structures")
        self._info("creating mesh and interpolating initial values")
        self.offline_data = OfflineData.from_mesh(self.mesh, rank=rank, n_ranks=n_ranks, device=device)
        self._info("preparing compute kernels")
        self.module = HyperbolicModule(self.offline_data, system, iv, ModuleParameters(**mp), comm=comm)
        self.time_integrator = TimeIntegrator(
            self.module, SCHEMES[scheme], H.get_float("cfl min", 0.45), H.get_float("cfl max", 0.90),
            CFLRecoveryStrategy.none if recovery == "none" else CFLRecoveryStrategy.bang_bang_control)
        self.state = self.module.state_vector()
        self.n_dofs = int(self.offline_data.host.n_global_dofs)
        self.cycle = 0
        self.t = 0.0
        self.wall_time = 0.0

    # -------------------------------------------------------------------------------------
    def _info(self, text):
        if self.rank == 0:
            print("[INFO] " + text, file=self.out)

    def unused_entries(self):
        return self.prm.unused()

    def run(self):
        """TimeLoop::run(), time_loop.template.h:218-432"""
        self.state.interpolate(0.0)
        t, cycle, timer_cycle = 0.0, 1, 0
        self._info("entering main loop")
        eps = np.finfo(np.float64).eps
        relax = 1.0 - 10.0 * eps if self.enforce_t_final else 1.0
        t_start = time.perf_counter()
        last_terminal = float("inf") if self.terminal_update_interval == 0.0 else float("-inf")
        while True:
            if self.manifold_name is not None:
                self.time_series.append((t, self.space_averaged_primitive_state()))
            if t >= timer_cycle * self.timer_granularity:
                timer_cycle += 1  # output / postprocessing ticks of the reference: nothing to do here
            if t >= relax * self.t_final:
                break
            tau = self.time_integrator.step(self.state, t, self.t_final if self.enforce_t_final else None)
            t += tau
            if self.terminal_update_interval != 0.0:
                wall = time.perf_counter() - t_start
                if t >= timer_cycle * self.timer_granularity or wall >= last_terminal + self.terminal_update_interval:
                    self.print_cycle_statistics(cycle, t, wall)
                    last_terminal = wall
            cycle += 1
        cycle -= 1
        self.module.synchronize()
        self.wall_time = time.perf_counter() - t_start
        self.cycle, self.t = cycle, t
        if self.terminal_update_interval != 0.0:
            self.print_cycle_statistics(cycle, t, self.wall_time, final=True)
        result = dict(cycle=cycle, t=t, n_dofs=self.n_dofs, wall_time=self.wall_time)
        if self.enable_compute_error:
            result.update(self.compute_error(t))
        if self.manifold_name is not None:
            result["time_series"] = self.time_series
            if self.rank == 0 and self.debug_filename.endswith("space_averaged_time_series.dat"):
                self.print_time_series()
        return result

    def primitive_state(self, U):
        """View::to_primitive_state: (rho, v, p) / (h, v)"""
        v = U[:, 1 : 1 + self.dim] / U[:, :1]
        if self.equation == Equation.euler:
            gamma = self.module.hyperbolic_system.desc.gamma
            p = (gamma - 1.0) * (U[:, -1] - 0.5 * U[:, 0] * np.sum(v * v, axis=1))
            return np.column_stack([U[:, 0], v, p])
        if self.equation == Equation.euler_aeos:
            # (rho, v, e) as euler_aeos/hyperbolic_system.h:1481-1500
            rho = U[:, 0]
            e = (U[:, -1] - 0.5 * rho * np.sum(v * v, axis=1)) / rho
            return np.column_stack([rho, v, e])
        return np.column_stack([U[:, 0], v])

    def space_averaged_primitive_state(self):
        """Quantities::internal_accumulate (quantities.template.h:369-423): lumped-mass weighted
        average of the primitive state and of its second moments over all (unconstrained) nodes"""
        host = self.offline_data.host
        U = self.gather_global(self.state.download())
        m = self.gather_global(np.asarray(host.mi[: host.n_owned], dtype=np.float64).reshape(-1, 1))[:, 0]
        prim = self.primitive_state(U)
        w = m / np.sum(m)
        return np.concatenate([w @ prim, w @ (prim * prim)])

    def print_time_series(self):
        """<basename>-<manifold>-R0000-space_averaged_time_series.dat as the reference writes it"""
        names = {Equation.euler: ["rho"] + ["v_%d" % (d + 1) for d in range(self.dim)] + ["p"],
                 Equation.euler_aeos: ["rho"] + ["v_%d" % (d + 1) for d in range(self.dim)] + ["e"],
                 Equation.shallow_water: ["h"] + ["v_%d" % (d + 1) for d in range(self.dim)]}[self.equation]
        if self.dim == 1:
            names[1] = "v"
        print("# time t\tprimitive state (%s)\t and 2nd moments" % ", ".join(names), file=self.out)
        k = len(names)
        for t, values in self.time_series:
            print("%.14e\t%s\t%s" % (t, " ".join("%.14e" % x for x in values[:k]), " ".join("%.14e" % x for x in values[k:])),
                  file=self.out)

    def throughput(self, cycle, wall):
        """MQ/s of print_throughput (:1156-1162): cycles * n_dofs * efficiency / wall time"""
        return cycle * self.n_dofs / 1.0e6 / max(wall, 1e-300) * self.time_integrator.efficiency

    def print_cycle_statistics(self, cycle, t, wall, final=False):
        if self.rank != 0:
            return
        mq = self.throughput(cycle, wall)
        lines = [
            "",
            "Information: (HYP) %s" % {Equation.euler: "Euler", Equation.euler_aeos: "Euler AEOS",
                                         Equation.shallow_water: "ShallowWater"}[self.equation],
            "  [%s %dD] with %d Qdofs on %d ranks / 1 GPU each" % (self.base_name, self.dim, self.n_dofs, self.n_ranks),
            "",
            "Job status:%s" % (" (final)" if final else ""),
            "  current cycle  %d and time t = %.6e (%.1f%%)" % (cycle, t, 100.0 * t / self.t_final if self.t_final else 0.0),
            "",
            "Throughput:",
            "  WALL: %.4f MQ/s  (%.4e s/Qdof/substep)  (%.2f cycles/s)" % (mq, 1.0e-6 / max(mq, 1e-300), cycle / max(wall, 1e-300)),
            "        [ %s with CFL = %.2f (%d/0 rsts) (%d/0 warn) ]"
            % (self.scheme_name, self.module.cfl, self.module.n_restarts(), self.module.n_warnings()),
            "        [ dt = %.2e ( %.2e dt/s) ]" % (t / max(cycle, 1), t / max(wall, 1e-300)),
        ]
        print("\n".join(lines), file=self.out)

    def gather_global(self, array):
        """owned rows of all ranks in global (lexicographic) numbering, on every rank"""
        host = self.offline_data.host
        k = array.shape[1]
        full = np.zeros((self.n_dofs, k))
        full[host.global_index[: host.n_owned].astype(np.int64)] = array
        if self.n_ranks > 1:
            import torch
            import torch.distributed as dist

            t = torch.from_numpy(full)
            if dist.get_backend() == "nccl":
                t = t.cuda(self.device)
            dist.all_reduce(t)  # disjoint ownership: sum == gather
            full = t.cpu().numpy()
        return full

    def compute_error(self, t):
        """TimeLoop::compute_error(), time_loop.template.h:694-842"""
        if self.mesh.cell_mask is not None:
            raise NotOnThePath("error norms are implemented for the rectangular domain")
        self.module.prepare_state_vector(self.state, t)
        U = self.gather_global(self.state.download())
        analytic_state = self.module.state_vector()
        analytic_state.interpolate(t)
        A = self.gather_global(analytic_state.download())
        linf = l1 = l2 = 0.0
        for name in self.error_quantities:
            if name not in self.component_names:
                raise _prm.ParameterError("Unknown component name »%s«" % name)
            c = self.component_names.index(name)
            e_linf, e_l1, e_l2 = q1_norms(U[:, c] - A[:, c], self.mesh.n_cells, self.mesh.h)
            if self.error_normalize:
                a_linf, a_l1, a_l2 = q1_norms(A[:, c], self.mesh.n_cells, self.mesh.h)
                linf += e_linf / a_linf
                l1 += e_l1 / a_l1
                l2 += e_l2 / a_l2
            else:
                linf += e_linf
                l1 += e_l1
                l2 += e_l2
        if self.rank == 0:
            description = "Normalized consolidated" if self.error_normalize else "Consolidated"
            print(description + " Linf, L1, and L2 errors at final time ", file=self.out)
            print("#dofs = %d" % self.n_dofs, file=self.out)
            for label, value in (("t    ", t), ("Linf ", linf), ("L1   ", l1), ("L2   ", l2)):
                print("%s = %s" % (label, _format16(value)), file=self.out)
        return dict(Linf=linf, L1=l1, L2=l2)


def _format16(x):
    """std::setprecision(16) in the default float format (%.16g)"""
    return "%.16g" % x


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) != 1:
        print("usage: python -m ryujin_b200 <parameter file.prm>", file=sys.stderr)
        return 2
    loop = TimeLoop(argv[0])
    loop.run()
    unused = loop.unused_entries()
    if unused:
        print("[INFO] entries of the parameter file that the B200 path does not read: " + ", ".join(unused), file=sys.stderr)
    return 0
