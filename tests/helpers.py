"""Shared builders for the parity tests: the same case set up twice, once through the
product (ryujin_b200, CUDA) and once through the CPU oracle fed with the product's own
stencil arrays, so that both sides see bit-identical inputs in the same DoF numbering."""
import numpy as np

from oracle import orc

B_DIRICHLET, B_SLIP, B_DO_NOTHING = 4, 2, 0

POW_IMPL_NAMES = {0: "accurate", 1: "vcl", 2: "vcl_fma"}


def product_pow_impl():
    """which pow() the loaded product library was built with (rj_pow_implementation)"""
    import ryujin_b200 as rj

    return int(rj.lib.rj_pow_implementation())


def matching_oracle():
    """The oracle build whose pow() is the one the product library was built for:
      vcl      -> oracle/_ref/liboracle_vcl.so      (the reference's own VCL sources, non-FMA form;
                  shares no pow code with the product -- this is the independent parity target)
      vcl_fma  -> oracle/_ref/liboracle_vcl_fma.so  (VCL as it compiles on an FMA host)
      accurate -> oracle/liboracle.so               (libm pow; parity only up to pow rounding)
    -> (oracle, exact) where `exact` says that every operation of the path, pow included, is
    the same IEEE operation on both sides, i.e. that north_star's 1e-12 must hold outright."""
    impl = product_pow_impl()
    if impl == 1:
        o = orc.Oracle("vcl")
        assert o.lib.orc_uses_vcl(), "oracle/_ref/liboracle_vcl.so is not a VCL build"
        return o, True
    if impl == 2:
        o = orc.Oracle("vcl_fma")
        assert o.lib.orc_uses_vcl()
        return o, True
    return orc.Oracle("port"), False


def oracle_offline_from_host(oracle, host):
    """Feed the product generator's arrays (local numbering) to the oracle."""
    return oracle.offline_from_arrays(
        host.dim,
        host.row_starts.astype(np.uint32),
        host.cols,
        host.mij,
        host.cij,
        host.mi[: host.n_owned],
        host.measure_of_omega,
        host.boundary_dof,
        host.boundary_normal,
        host.boundary_id,
        host.boundary_position,
        host.coupling_pairs.ravel(),
    )


CASES = {
    # config 1 of BASELINE.json at reduced size by default: 2D Sod shock tube
    "sod2d": dict(
        dim=2, bottom_left=[0.0, 0.0], top_right=[1.0, 1.0],
        boundary=[B_DIRICHLET, B_DIRICHLET, B_SLIP, B_SLIP],
        iv=dict(kind="contrast", direction=[1.0, 0.0], position=[0.5, 0.0], primitive=[1.0, 0.0, 1.0],
                primitive_right=[0.125, 0.0, 0.1]),
    ),
    # config 2: 3D isentropic vortex, all faces (time dependent) dirichlet
    "vortex3d": dict(
        dim=3, bottom_left=[-5.0, -5.0, -1.0], top_right=[5.0, 5.0, 1.0],
        boundary=[B_DIRICHLET] * 6,
        iv=dict(kind="isentropic_vortex", direction=[1.0, 1.0, 0.0], position=[-1.0, -1.0, 0.0], mach_number=1.0, beta=5.0),
    ),
    # 2D vortex of tests/euler/verification-isentropic_vortex-2d-*.prm
    "vortex2d": dict(
        dim=2, bottom_left=[-5.0, -5.0], top_right=[5.0, 5.0],
        boundary=[B_DIRICHLET] * 4,
        iv=dict(kind="isentropic_vortex", direction=[1.0, 1.0], position=[-1.0, -1.0], mach_number=1.0, beta=5.0),
    ),
    # 3D shock tube with slip walls: exercises slip normals on faces / edges / corners
    "sod3d": dict(
        dim=3, bottom_left=[0.0, 0.0, 0.0], top_right=[1.0, 0.5, 0.5],
        boundary=[B_DIRICHLET, B_DIRICHLET, B_SLIP, B_SLIP, B_SLIP, B_SLIP],
        iv=dict(kind="contrast", direction=[1.0, 0.0, 0.0], position=[0.5, 0.0, 0.0], primitive=[1.0, 0.0, 1.0],
                primitive_right=[0.125, 0.0, 0.1]),
    ),
}

# euler_aeos: parameters of the equation of state under the product's names (rj_equation_desc)
AEOS_EOS = {
    "polytropic": dict(equation_of_state=0, gamma=1.4),
    "nasg": dict(equation_of_state=1, gamma=1.6, covolume_b=0.12, reference_pressure=0.4,
                 reference_specific_internal_energy=0.3),
    "vdw": dict(equation_of_state=2, gamma=1.3, covolume_b=0.1, vdw_a=0.002),
}


def aeos_parameters(eos, strict=True):
    """-> (keyword arguments of rj.HyperbolicSystem, keyword arguments of the oracle's aeos_sim)"""
    eq = dict(AEOS_EOS[eos], compute_strict_bounds=int(strict))
    prm = dict(eq)
    prm["eos"] = prm.pop("equation_of_state")
    prm["compute_strict_bounds"] = bool(strict)
    return eq, prm


def aeos_case(name):
    """'aeos:<eos>:<strict|loose>:<case>' -> (case, eos, strict) or None"""
    if not name.startswith("aeos:"):
        return None
    _, eos, mode, case = name.split(":")
    return case, eos, mode == "strict"


KIND_ORACLE = {"uniform": orc.UNIFORM, "contrast": orc.CONTRAST, "isentropic_vortex": orc.ISENTROPIC_VORTEX}


def make_mesh(case, n_cells, partition=None):
    import ryujin_b200 as rj

    if aeos_case(case):
        case = aeos_case(case)[0]
    c = CASES[case] if case in CASES else SW_CASES[case]
    return rj.CartesianMesh.rectangular_domain(c["bottom_left"], c["top_right"], n_cells, c["boundary"], partition=partition)


def make_product(case, n_cells, device=0, comm=None, rank=0, n_ranks=1, module_params=None, equation="euler", partition=None,
                 **eq):
    """-> (offline_data, module) on the GPU"""
    import ryujin_b200 as rj

    c = CASES[case]
    mesh = make_mesh(case, n_cells, partition=partition)
    od = rj.OfflineData.from_mesh(mesh, rank=rank, n_ranks=n_ranks, device=device)
    iv = dict(c["iv"])
    kind = getattr(rj.InitialState, iv.pop("kind"))
    module = rj.HyperbolicModule(
        od,
        rj.HyperbolicSystem(getattr(rj.Equation, equation), **eq),
        rj.InitialValues(kind, **iv),
        rj.ModuleParameters(**(module_params or {})),
        comm=comm,
    )
    return od, module


def make_oracle(oracle, case, host, off=None, equation="euler", **prm):
    """-> (offline, sim) of the oracle on the product's arrays (or on a prepared Offline `off`)"""
    c = CASES[case]
    if off is None:
        off = oracle_offline_from_host(oracle, host)
    iv = c["iv"]
    sim = (off.aeos_sim if equation == "euler_aeos" else off.euler_sim)(
        kind=KIND_ORACLE[iv["kind"]],
        direction=iv.get("direction", (1.0, 0.0, 0.0)),
        position=iv.get("position", (1.0, 0.0, 0.0)),
        primitive=iv.get("primitive", (1.4, 3.0, 1.0)),
        primitive_right=iv.get("primitive_right", (1.4, 0.0, 1.0)),
        mach=iv.get("mach_number", 2.0),
        beta=iv.get("beta", 5.0),
        **prm,
    )
    return off, sim


def initial_state_host(oracle, case, host, t=0.0):
    """Initial U (AoS, local numbering) evaluated by the oracle's InitialValues at the
    product's node positions."""
    c = CASES[case]
    iv = c["iv"]
    n = host.n_owned
    # an oracle mesh is not needed: evaluate through a tiny sim on a 1-cell mesh? Simpler:
    # use the oracle sim of make_oracle with positions supplied through the boundary map is
    # not general, so evaluate in numpy following the same formulas.
    x = host.positions
    dim = host.dim
    gamma = 1.4
    d = np.array(iv.get("direction", [1.0] + [0.0] * (dim - 1)), dtype=float)[:dim]
    d = d / np.linalg.norm(d)
    pos = np.array(iv.get("position", [1.0] + [0.0] * (dim - 1)), dtype=float)[:dim]
    y = x - pos
    if dim == 3:
        n_x, n_z = d[0], d[2]
        nr = np.hypot(n_x, n_z)
        if nr > 1e-14:
            n_x, n_z = n_x / nr, n_z / nr
            y = np.stack([n_x * y[:, 0] + n_z * y[:, 2], y[:, 1], -n_z * y[:, 0] + n_x * y[:, 2]], axis=1)
    if dim >= 2:
        n_x, n_y = d[0], d[1]
        nr = np.hypot(n_x, n_y)
        if nr > 1e-14:
            n_x, n_y = n_x / nr, n_y / nr
            y0 = n_x * y[:, 0] + n_y * y[:, 1]
            y1 = -n_y * y[:, 0] + n_x * y[:, 1]
            y = np.concatenate([y0[:, None], y1[:, None], y[:, 2:]], axis=1)
    k = dim + 2
    U = np.zeros((n, k))

    def from_prim(prim):
        rho, u, p = prim
        s = np.zeros(k)
        s[0] = rho
        s[1] = rho * u
        s[dim + 1] = p / (gamma - 1.0) + 0.5 * rho * u * u
        return s

    if iv["kind"] == "uniform":
        U[:] = from_prim(iv["primitive"])
    elif iv["kind"] == "contrast":
        left, right = from_prim(iv["primitive"]), from_prim(iv["primitive_right"])
        U[:] = np.where((y[:, 0] > 0.0)[:, None], right[None, :], left[None, :])
    else:
        mach, beta = iv["mach_number"], iv["beta"]
        pb0 = y[:, 0] - mach * t
        pb1 = y[:, 1]
        r2 = pb0 * pb0 + pb1 * pb1
        factor = beta / (2.0 * np.pi) * np.exp(0.5 - 0.5 * r2)
        T = 1.0 - (gamma - 1.0) / (2.0 * gamma) * factor * factor
        u = mach - factor * pb1
        v = factor * pb0
        rho = T ** (1.0 / (gamma - 1.0))
        p = rho**gamma
        U[:, 0] = rho
        U[:, 1] = rho * u
        U[:, 2] = rho * v
        U[:, dim + 1] = p / (gamma - 1.0) + 0.5 * rho * (u * u + v * v)
    # Galilei transform of the momentum
    m = U[:, 1 : 1 + dim].copy()
    if dim >= 2:
        n_x, n_y = d[0], d[1]
        nr = np.hypot(n_x, n_y)
        if nr > 1e-14:
            n_x, n_y = n_x / nr, n_y / nr
            m0 = n_x * m[:, 0] - n_y * m[:, 1]
            m1 = n_y * m[:, 0] + n_x * m[:, 1]
            m[:, 0], m[:, 1] = m0, m1
    if dim == 3:
        n_x, n_z = d[0], d[2]
        nr = np.hypot(n_x, n_z)
        if nr > 1e-14:
            n_x, n_z = n_x / nr, n_z / nr
            m0 = n_x * m[:, 0] - n_z * m[:, 2]
            m2 = n_z * m[:, 0] + n_x * m[:, 2]
            m[:, 0], m[:, 2] = m0, m2
    U[:, 1 : 1 + dim] = m
    return U


def rel_linf(a, b):
    """relative L-infinity distance over all components (north_star's parity norm)"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / scale) if scale > 0 else float(np.max(np.abs(a - b)))


def rel_linf_per_component(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(len(b), -1)
    b = np.asarray(b, dtype=np.float64).reshape(len(b), -1)
    out = []
    # a component that vanishes identically (e.g. the y-momentum of a 1D-in-2D shock tube)
    # only carries round-off noise: floor its scale at 1e-8 of the largest component
    floor = 1e-8 * np.max(np.abs(b)) if b.size else 0.0
    for c in range(b.shape[1]):
        scale = max(np.max(np.abs(b[:, c])), floor)
        out.append(np.max(np.abs(a[:, c] - b[:, c])) / (scale if scale > 0 else 1.0))
    return np.array(out)


# ---------------------------------------------------------------------------------------
# shallow water (BASELINE config 5)
B_DYNAMIC, B_DIRICHLET_MOMENTUM, B_NO_SLIP = 5, 6, 3

SW_CASES = {
    # tests/shallow_water/verification-smooth_vortex-erk33-*.prm
    "sw_vortex2d": dict(
        dim=2, bottom_left=[-6.0, -6.0], top_right=[6.0, 6.0], boundary=[B_DIRICHLET] * 4,
        iv=dict(kind="sw_smooth_vortex", direction=[1.0, 1.0], position=[-1.0, -1.0], sw_params=[2.0, 1.0, 2.0]),
        eq=dict(gravity=9.81, reference_water_depth=2.0, dry_state_relaxation_factor=0.0),
        mp=dict(limit_on_kinetic_energy=0, limit_on_square_velocity=1),
    ),
    # config 5 of BASELINE.json: 2D circular dam break, slip walls; h jumps 2.5 -> 0.5 at |x|^2 = 2.5
    "sw_dam2d": dict(
        dim=2, bottom_left=[-5.0, -5.0], top_right=[5.0, 5.0], boundary=[B_SLIP] * 4,
        iv=dict(kind="sw_circular_dam_break", direction=[1.0, 0.0], position=[0.0, 0.0], sw_params=[0.5, 2.5, 2.5]),
        eq=dict(gravity=9.81, reference_water_depth=1.0),
        mp=dict(limit_on_kinetic_energy=0, limit_on_square_velocity=1),
    ),
    # both limiter branches + friction + mixed boundary ids on the same dam break
    "sw_dam2d_friction": dict(
        dim=2, bottom_left=[-5.0, -5.0], top_right=[5.0, 5.0],
        boundary=[B_DYNAMIC, B_DIRICHLET_MOMENTUM, B_NO_SLIP, B_SLIP],
        iv=dict(kind="sw_circular_dam_break", direction=[1.0, 0.0], position=[0.3, -0.2], sw_params=[0.5, 2.5, 2.5]),
        eq=dict(gravity=9.81, reference_water_depth=1.0, manning_friction_coefficient=0.05),
        mp=dict(limit_on_kinetic_energy=1, limit_on_square_velocity=1),
    ),
    # tests/shallow_water/verification-ritter_dam_break-erk33-*.prm: 1D, dry bed, kinetic-energy limiter
    "sw_ritter1d": dict(
        dim=1, bottom_left=[0.0], top_right=[10.0], boundary=[B_DIRICHLET] * 2,
        iv=dict(kind="sw_ritter_dam_break", direction=[1.0], position=[5.0], sw_params=[1.0, 0.005]),
        eq=dict(gravity=9.81, reference_water_depth=0.005, dry_state_relaxation_factor=1.0e-3),
        mp=dict(limit_on_kinetic_energy=1, limit_on_square_velocity=0),
    ),
}

SW_KIND_ORACLE = {
    "sw_circular_dam_break": orc.SW_CIRCULAR_DAM_BREAK, "sw_uniform": orc.SW_UNIFORM,
    "sw_ritter_dam_break": orc.SW_RITTER_DAM_BREAK, "sw_smooth_vortex": orc.SW_SMOOTH_VORTEX,
}


def make_sw_product(case, n_cells, device=0, comm=None, rank=0, n_ranks=1, **mp_extra):
    import ryujin_b200 as rj

    c = SW_CASES[case]
    mesh = rj.CartesianMesh.rectangular_domain(c["bottom_left"], c["top_right"], n_cells, c["boundary"])
    od = rj.OfflineData.from_mesh(mesh, rank=rank, n_ranks=n_ranks, device=device)
    iv = dict(c["iv"])
    kind = getattr(rj.InitialState, iv.pop("kind"))
    mp = dict(c["mp"])
    mp.update(mp_extra)
    module = rj.HyperbolicModule(
        od, rj.HyperbolicSystem(rj.Equation.shallow_water, **c["eq"]), rj.InitialValues(kind, **iv),
        rj.ModuleParameters(**mp), comm=comm,
    )
    return od, module


def make_sw_oracle(oracle, case, host, off=None, **prm_extra):
    c = SW_CASES[case]
    if off is None:
        off = oracle_offline_from_host(oracle, host)
    iv = c["iv"]
    prm = dict(c["eq"])
    prm.update(c["mp"])
    prm.update(prm_extra)
    sim = off.sw_sim(kind=SW_KIND_ORACLE[iv["kind"]], direction=iv["direction"], position=iv["position"],
                     p=iv["sw_params"], **prm)
    return off, sim


def sw_bathymetry(host, amplitude=0.2):
    """a smooth, non-trivial bathymetry at the product's node positions"""
    x = host.positions
    z = amplitude * np.cos(0.9 * x[:, 0])
    if host.dim >= 2:
        z = z * np.sin(0.7 * x[:, 1] + 0.3) + 0.05 * x[:, 1] / 5.0
    return np.ascontiguousarray(z)


# ---------------------------------------------------------------------------------------
# cases added after a line-coverage run of the kernels (tests/host_emulation, gcov): Euler in 1D and
# the `dynamic` (Riemann characteristic) boundary condition were not reached by any GPU test
CASES.update({
    "sod1d": dict(
        dim=1, bottom_left=[0.0], top_right=[1.0], boundary=[B_DIRICHLET, B_DIRICHLET],
        iv=dict(kind="contrast", direction=[1.0], position=[0.5], primitive=[1.0, 0.0, 1.0], primitive_right=[0.125, 0.0, 0.1]),
    ),
    # uniform Mach 0.5 flow at an angle through a box with `dynamic` boundaries on all sides:
    # subsonic inflow on two sides, subsonic outflow on the other two
    "dynamic2d": dict(
        dim=2, bottom_left=[0.0, 0.0], top_right=[1.0, 1.0], boundary=[B_DYNAMIC] * 4,
        iv=dict(kind="uniform", direction=[1.0, 0.5], position=[0.0, 0.0], primitive=[1.4, 0.5, 1.0]),
    ),
    # the same box with a Mach 3 flow: supersonic inflow / outflow branches
    "dynamic2d_supersonic": dict(
        dim=2, bottom_left=[0.0, 0.0], top_right=[1.0, 1.0], boundary=[B_DYNAMIC] * 4,
        iv=dict(kind="uniform", direction=[1.0, 0.3], position=[0.0, 0.0], primitive=[1.4, 3.0, 1.0]),
    ),
})


# ---------------------------------------------------------------------------------------
# N-rank runs against the oracle: the reference's result depends on the partition at
# round-off level -- every rank computes d_ij towards its ghost columns itself and sums its rows in
# its LOCAL column order (ghost columns last).  The oracle reproduces an N-rank run bit for bit
# inside one address space when it is handed (a) the concatenation of the ranks' local stencils,
# rows and columns in each owner's local order, and (b) the owner rank of every row
# (orc_module.h: OfflineData::partition).
def merged_partition_stencil(mesh, n_ranks):
    """-> dict of arrays of the concatenated per-rank stencils (numbering G = rank offset + local
    owned index) + 'partition' + 'hosts' + 'offsets'"""
    import ryujin_b200 as rj

    hosts = [rj.HostOfflineData(mesh, r, n_ranks) for r in range(n_ranks)]
    offsets = np.concatenate([[0], np.cumsum([h.n_owned for h in hosts])]).astype(np.int64)
    n_global = int(offsets[-1])
    assert n_global == hosts[0].n_global_dofs
    G_of_mesh_global = np.full(n_global, -1, dtype=np.int64)
    for r, h in enumerate(hosts):
        G_of_mesh_global[h.global_index[: h.n_owned].astype(np.int64)] = offsets[r] + np.arange(h.n_owned)
    assert (G_of_mesh_global >= 0).all()
    row_starts, cols, mij, cij, mi, b_dof, b_normal, b_id, b_pos, pairs, part = [np.zeros(1, dtype=np.int64)], [], [], [], [], [], [], [], [], [], []
    nnz0 = 0
    for r, h in enumerate(hosts):
        local_to_G = G_of_mesh_global[h.global_index.astype(np.int64)]  # owned and ghost rows
        assert np.array_equal(local_to_G[: h.n_owned], offsets[r] + np.arange(h.n_owned))
        row_starts.append(h.row_starts[1:].astype(np.int64) + nnz0)
        nnz0 += h.nnz
        cols.append(local_to_G[h.cols.astype(np.int64)])
        mij.append(h.mij)
        cij.append(h.cij)
        mi.append(h.mi[: h.n_owned])
        b_dof.append(h.boundary_dof.astype(np.int64) + offsets[r])
        b_normal.append(h.boundary_normal)
        b_id.append(h.boundary_id)
        b_pos.append(h.boundary_position)
        if len(h.coupling_pairs):
            p = h.coupling_pairs.astype(np.int64)
            pairs.append(np.stack([p[:, 0] + offsets[r], p[:, 1], local_to_G[p[:, 2]]], axis=1))
        part.append(np.full(h.n_owned, r, dtype=np.int32))
    return dict(
        dim=hosts[0].dim, hosts=hosts, offsets=offsets,
        row_starts=np.concatenate(row_starts).astype(np.uint32), cols=np.concatenate(cols).astype(np.uint32),
        mij=np.concatenate(mij), cij=np.concatenate(cij), mi=np.concatenate(mi),
        measure_of_omega=hosts[0].measure_of_omega,
        boundary_dof=np.concatenate(b_dof).astype(np.uint32), boundary_normal=np.concatenate(b_normal),
        boundary_id=np.concatenate(b_id).astype(np.int32), boundary_position=np.concatenate(b_pos),
        coupling_pairs=(np.concatenate(pairs) if pairs else np.zeros((0, 3))).astype(np.uint32),
        partition=np.concatenate(part), positions=np.concatenate([h.positions for h in hosts]),
    )


def oracle_offline_from_partition(oracle, merged):
    m = merged
    off = oracle.offline_from_arrays(m["dim"], m["row_starts"], m["cols"], m["mij"], m["cij"], m["mi"], m["measure_of_omega"],
                                     m["boundary_dof"], m["boundary_normal"], m["boundary_id"], m["boundary_position"],
                                     m["coupling_pairs"].ravel())
    off.set_partition(m["partition"])
    return off
