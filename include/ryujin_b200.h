/*
 * ryujin_b200 -- C ABI of the B200-native explicit hyperbolic update.
 *
 * This header is the drop-in boundary for ONE path of ryujin: the explicit
 * invariant-domain-preserving update HyperbolicModule::step() with its pre-step
 * prepare_state_vector() (reference: source/hyperbolic_module.h:110-278,
 * source/hyperbolic_module.template.h:97-193,236-1211).  The reference has no
 * FFI; its boundary is a C++ template surface.  Every entry point below names
 * the reference interface it replaces.  The C++ shim that re-creates the
 * reference classes on top of this ABI is include/ryujin_b200.hpp; the binding
 * a ryujin maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types cross this boundary;
 *  - "host" arrays are ordinary host memory and are copied during the call;
 *    "device" pointers are CUDA device pointers on the handle's device;
 *  - all floating point data is IEEE double (the reference's default NUMBER);
 *  - return code 0 = ok, RJ_RESTART (1) = invariant domain violation with
 *    strategy raise_exception (the shim throws ryujin::Restart), < 0 = fatal
 *    (rj_last_error() has the message);
 *  - a handle is used from one host thread at a time (the reference's module
 *    is not re-entrant either: its scratch is mutable, hyperbolic_module.h:319-333);
 *  - there is NO CPU fallback: every compute entry point fails with
 *    RJ_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef RYUJIN_B200_H
#define RYUJIN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RJ_OK 0
#define RJ_RESTART 1
#define RJ_ERR_INVALID (-1)
#define RJ_ERR_CUDA (-2)
#define RJ_ERR_NCCL (-3)
#define RJ_ERR_NUMERIC (-4) /* NaN/inf/<=0 tau_max, hyperbolic_module.template.h:573-576 */
#define RJ_ERR_NO_DEVICE (-5)

/* Boundary ids: enum Boundary, source/discretization.h:48-123 */
enum rj_boundary_id {
  RJ_BOUNDARY_DO_NOTHING = 0,
  RJ_BOUNDARY_PERIODIC = 1,
  RJ_BOUNDARY_SLIP = 2,
  RJ_BOUNDARY_NO_SLIP = 3,
  RJ_BOUNDARY_DIRICHLET = 4,
  RJ_BOUNDARY_DYNAMIC = 5,
  RJ_BOUNDARY_DIRICHLET_MOMENTUM = 6
};

/* Equation modules ("Description"), source/<equation>/description.h */
enum rj_equation { RJ_EQUATION_EULER = 0, RJ_EQUATION_SHALLOW_WATER = 1, RJ_EQUATION_EULER_AEOS = 2 };

/* EquationOfStateLibrary of the euler_aeos module: the analytic equations of state of
 * source/euler_aeos/equation_of_state_{polytropic_gas,noble_abel_stiffened_gas,van_der_waals}.h
 * (the tabulated "sesame" and the Jones-Wilkins-Lee equations of state are not offered) */
enum rj_equation_of_state {
  RJ_EOS_POLYTROPIC_GAS = 0,
  RJ_EOS_NOBLE_ABEL_STIFFENED_GAS = 1,
  RJ_EOS_VAN_DER_WAALS = 2
};

/* IDViolationStrategy, source/hyperbolic_module.h:37-52 */
enum rj_id_violation_strategy { RJ_STRATEGY_WARN = 0, RJ_STRATEGY_RAISE_EXCEPTION = 1 };

/* TimeSteppingScheme, source/time_integrator.h (explicit schemes only) */
enum rj_time_stepping_scheme {
  RJ_SSPRK_22 = 0,
  RJ_SSPRK_33 = 1,
  RJ_ERK_11 = 2,
  RJ_ERK_22 = 3,
  RJ_ERK_33 = 4,
  RJ_ERK_43 = 5,
  RJ_ERK_54 = 6
};

/* CFLRecoveryStrategy, source/time_integrator.h */
enum rj_cfl_recovery_strategy { RJ_CFL_RECOVERY_NONE = 0, RJ_CFL_RECOVERY_BANG_BANG = 1 };

/* Initial state / Dirichlet data evaluated on the device
 * (source/initial_values.template.h:65-219 + source/<equation>/initial_state_*.h). */
enum rj_initial_state_kind {
  RJ_INITIAL_UNIFORM = 0,           /* euler/initial_state_uniform.h */
  RJ_INITIAL_CONTRAST = 1,          /* euler/initial_state_contrast.h */
  RJ_INITIAL_ISENTROPIC_VORTEX = 2, /* euler/initial_state_isentropic_vortex.h */
  RJ_INITIAL_SW_CIRCULAR_DAM_BREAK = 16, /* shallow_water/initial_state_circular_dam_break.h */
  RJ_INITIAL_SW_UNIFORM = 17,            /* shallow_water/initial_state_uniform.h */
  RJ_INITIAL_SW_RITTER_DAM_BREAK = 18,   /* shallow_water/initial_state_ritter_dam_break.h */
  RJ_INITIAL_SW_SMOOTH_VORTEX = 19       /* shallow_water/initial_state_smooth_vortex.h */
};

const char *rj_last_error(void);
/* number of usable CUDA devices (0 on a CPU-only box); never fails */
int rj_device_count(void);
/* library version string */
const char *rj_version(void);

/* pow() of the path.  The reference routes every pow through ryujin::pow
 * (source/simd.template.h:196-231), which is Agner Fog's VCL pow_template_d
 * (simd-math/vectormath_exp.h:1508-1688).  The library is built with ONE implementation
 * (csrc/Makefile: RJ_POW=vcl|vcl_fma|accurate); rj_pow_implementation() names it. */
enum rj_pow_impl {
  RJ_POW_ACCURATE = 0, /* accurate_pow.h: 0.50x ulp, ~libm rounding */
  RJ_POW_VCL = 1,      /* vcl_pow.h: the reference's VCL pow, non-FMA form (default) */
  RJ_POW_VCL_FMA = 2   /* vcl_pow.h: VCL as compiled on an FMA host */
};
int rj_pow_implementation(void);
/* out[i] = pow(x[i], y[i]) with implementation `impl` (any of the three, independent of the
 * build default), evaluated on the host (on_device = 0; works without a GPU: this is plain
 * arithmetic, not a fallback of the sweep) or by a CUDA kernel (on_device = 1).  Test hook for
 * tests/test_pow.py: device bits == host bits == reference VCL bits. */
int rj_pow_evaluate(int impl, int on_device, size_t n, const double *x, const double *y, double *out);

/* ------------------------------------------------------------------------ *
 * Offline data (replaces the read-only accessors of OfflineData,
 * source/offline_data.h:109-264, the SparsityPatternSIMD arrays,
 * source/sparse_matrix_simd.h:62-148, and the Partitioner exchange lists).
 * ------------------------------------------------------------------------ */

/*
 * Host description of one rank's stencil data in LOCAL indexing.
 * Ranges: n_export <= n_owned <= n_relevant (offline_data.h:136-169; the
 * reference's n_internal SIMD boundary is a property of the host layout only,
 * see simd_length below).
 *
 * Stencil: rows [0, n_owned) of the sparsity graph; each row lists the
 * diagonal first, then the remaining columns in ascending local index
 * (tests/common/sparse_matrix_simd.output).  Two host layouts are accepted:
 *  - simd_length == 1: plain CSR, row_starts has n_owned + 1 entries (in units
 *    of matrix entries), column_indices / mass_matrix have nnz entries,
 *    cij_matrix has nnz * dim entries, entry-major ((c_ij)_d at e*dim + d);
 *  - simd_length  > 1: the reference's SparsityPatternSIMD/SparseMatrixSIMD
 *    layout for rows [0, n_internal) in groups of simd_length rows followed by
 *    a CSR tail (sparse_matrix_simd.h:40-61,389-448): row_starts as described
 *    at sparse_matrix_simd.template.h:94-146, matrix data
 *    data[(row_starts[g] + col*simd_length) * n_comp + c*simd_length + lane].
 * The library decodes either into its own warp-sliced device layout.
 */
typedef struct rj_offline_desc {
  int dim;
  uint32_t n_export;   /* owned rows that are ghosts elsewhere come first */
  uint32_t n_internal; /* only used for simd_length > 1 */
  uint32_t n_owned;
  uint32_t n_relevant; /* owned + ghost */
  uint32_t simd_length;

  const uint64_t *row_starts;
  const uint32_t *column_indices;
  const double *mass_matrix;        /* m_ij, offline_data.h mass_matrix() */
  const double *cij_matrix;         /* c_ij, offline_data.h cij_matrix()  */
  const double *lumped_mass_matrix; /* m_i for rows [0, n_relevant)       */
  double measure_of_omega;          /* global |Omega|                      */

  /* boundary_map(), offline_data.h:66-72,222: one entry per (dof, merged
   * face normal), sorted by dof, entries of one dof in application order */
  uint32_t n_boundary;
  const uint32_t *boundary_dof;
  const double *boundary_normal;   /* n_boundary * dim */
  const int32_t *boundary_id;      /* rj_boundary_id   */
  const double *boundary_position; /* n_boundary * dim */

  /* coupling_boundary_pairs(), offline_data.h:78-80,230: (i, col_idx, j) */
  uint32_t n_coupling_pairs;
  const uint32_t *coupling_pairs; /* 3 * n_coupling_pairs */
  /* optional: c_ji of every pair (n_coupling_pairs * dim).  NULL = read it
   * through the transposed index (cij_matrix.get_transposed_tensor,
   * hyperbolic_module.template.h:479), which needs j to be an owned row; the
   * reference obtains ghost rows of c_ij by an exchange at setup
   * (offline_data.template.h:756-762). */
  const double *coupling_cji;

  /* Exchange lists (dealii::Utilities::MPI::Partitioner as used at
   * multicomponent_vector.cc:11-41 and sparse_matrix_simd.template.h:152-267).
   * Ghost rows [n_owned, n_relevant) are grouped by owner rank in ascending
   * rank order, in the owner's local order. */
  uint32_t n_neighbors;
  const int32_t *neighbor_rank;    /* n_neighbors                         */
  const uint32_t *ghost_count;     /* rows received from each neighbor    */
  const uint32_t *send_offsets;    /* n_neighbors + 1 into send_indices   */
  const uint32_t *send_indices;    /* local owned rows sent to neighbor p */

  /* optional: positions of rows [0, n_owned) (n_owned * dim), used to
   * interpolate initial values on the device */
  const double *positions;
} rj_offline_desc;

typedef struct rj_offline rj_offline; /* device resident stencil data */

/* Upload and re-layout (warp-sliced ELL, C = 32). */
int rj_offline_create(rj_offline **out, const rj_offline_desc *desc, int device);
void rj_offline_destroy(rj_offline *od);

/* Host-only variant: computes the same layout and exchange lists without touching
 * CUDA (no device arrays; cannot be used to create a module).  For tests of the
 * host logic on CPU boxes. */
int rj_offline_create_host_layout(rj_offline **out, const rj_offline_desc *desc);
/* sizes = {n_neighbors, n_send_rows, n_mat_send} */
int rj_offline_exchange_sizes(const rj_offline *od, uint64_t sizes[3]);
/* neighbor_info: 9 values per neighbour {rank, ghost_start, ghost_count, send_offset,
 * send_count, mat_send_offset, mat_send_count, mat_recv_offset, mat_recv_count};
 * send_rows: local rows packed for the vector exchanges; mat_send_entries: device-layout
 * entry indices packed for the matrix ghost-row exchanges (cf. entries_to_be_sent,
 * sparse_matrix_simd.template.h:214-266); the receiver stores them contiguously at
 * entry index mat_recv_offset (cf. receive_targets, :164-188). */
int rj_offline_exchange_get(const rj_offline *od, uint32_t *neighbor_info, uint32_t *send_rows, uint32_t *mat_send_entries);

/* Device layout introspection for parity tests ("bit-exact stencil indices"):
 * sizes = {n_slices, n_entries_owned, n_entries_total, ld (vector stride)} */
int rj_offline_layout_sizes(const rj_offline *od, uint64_t sizes[4]);
/* copies back slice_offsets[n_slices+1], row_length[n_owned], and per entry
 * cols / transposed index (n_entries_total each); any pointer may be NULL */
int rj_offline_layout_get(const rj_offline *od,
                          uint64_t *slice_offsets,
                          uint32_t *row_length,
                          uint32_t *cols,
                          uint32_t *transposed);

/* ------------------------------------------------------------------------ *
 * Mesh generator (host): uniform continuous Q1 on a Cartesian cell grid with
 * an optional cell mask, partitioned into slabs along the last axis.  Replaces,
 * for these meshes only, OfflineData::prepare() = setup + assemble
 * (source/offline_data.template.h:58-1105) and Discretization::prepare()
 * (geometry "rectangular domain", source/geometry_rectangular_domain.h).
 * ------------------------------------------------------------------------ */
typedef struct rj_mesh_desc {
  int dim;
  uint32_t n_cells[3];
  double origin[3];
  double h[3];
  const uint8_t *cell_mask; /* NULL = all cells active; x fastest */
  int32_t box_boundary_ids[6]; /* faces x-,x+,y-,y+,z-,z+ of the bounding box */
  int32_t masked_boundary_id;  /* faces bordering an inactive cell */
  /* ranks per direction (x fastest): rank r owns block (r % p0, (r / p0) % p1, r / (p0 p1)) of the
   * node grid, e.g. {2, 2, 2} for BASELINE config 4.  All zero = slabs along the last axis. */
  uint32_t partition[3];
} rj_mesh_desc;

typedef struct rj_host_offline rj_host_offline;

int rj_mesh_generate(rj_host_offline **out, const rj_mesh_desc *mesh, int rank, int n_ranks);
void rj_host_offline_destroy(rj_host_offline *h);
/* view of the generated arrays (pointers stay valid until destroy) */
int rj_host_offline_desc(const rj_host_offline *h, rj_offline_desc *desc);
/* global lexicographic dof index of local rows [0, n_relevant) */
int rj_host_offline_global_indices(const rj_host_offline *h, uint64_t *global);
uint64_t rj_host_offline_n_global_dofs(const rj_host_offline *h);

/* Decode the reference's SIMD host layout to CSR (exposed for tests). */
int rj_decode_simd_pattern(uint32_t n_rows,
                           uint32_t n_internal,
                           uint32_t simd_length,
                           const uint64_t *row_starts,
                           const uint32_t *column_indices,
                           uint64_t *csr_row_ptr /* n_rows+1 */,
                           uint32_t *csr_cols);

/* ------------------------------------------------------------------------ *
 * Communicator: one process per GPU.  The unique id is created on rank 0 and
 * distributed by the host program (MPI_Bcast in ryujin, torch.distributed in
 * bench.py).  Replaces the MPI_Comm argument of the module constructor.
 * ------------------------------------------------------------------------ */
typedef struct rj_comm rj_comm;
#define RJ_UNIQUE_ID_BYTES 128
int rj_comm_unique_id(uint8_t id[RJ_UNIQUE_ID_BYTES]);
int rj_comm_create(rj_comm **out, const uint8_t id[RJ_UNIQUE_ID_BYTES], int rank, int n_ranks, int device);
void rj_comm_destroy(rj_comm *c);

/* ------------------------------------------------------------------------ *
 * Equation module parameters (compile-time "Description" + runtime
 * ParameterAcceptor values of the reference).
 * ------------------------------------------------------------------------ */
typedef struct rj_equation_desc {
  int equation; /* rj_equation */
  /* B - Equation, euler/hyperbolic_system.h:668-684 */
  double gamma;
  double reference_density;
  double vacuum_state_relaxation_small;
  double vacuum_state_relaxation_large;
  /* shallow_water/hyperbolic_system.h:643-674 */
  double gravity;
  double manning_friction_coefficient;
  double reference_water_depth;
  double dry_state_relaxation_factor;
  double dry_state_relaxation_small;
  double dry_state_relaxation_large;
  /* euler_aeos/hyperbolic_system.h:456-490 and the parameters of the selected equation of state:
   * gamma (above) is its "gamma"; covolume b (noble abel stiffened gas, van der waals), reference
   * pressure and reference specific internal energy (noble abel stiffened gas), vdw a.  The
   * interpolatory covolume / pinfty / q of the surrogate follow from them as in the
   * parse_parameters_call_back of each equation of state. */
  int32_t equation_of_state; /* rj_equation_of_state */
  int32_t compute_strict_bounds;
  double covolume_b;
  double reference_pressure;
  double reference_specific_internal_energy;
  double vdw_a;
} rj_equation_desc;

typedef struct rj_module_params {
  /* F - HyperbolicModule/indicator, euler/indicator.h:28-31 */
  double indicator_evc_factor;
  /* F - HyperbolicModule/limiter, euler/limiter.h:26-49 */
  uint32_t limiter_iterations;
  double limiter_newton_tolerance;
  uint32_t limiter_newton_max_iterations;
  double limiter_relaxation_factor;
  /* F - HyperbolicModule/riemann solver, euler/riemann_solver.h:26-38 */
  double riemann_newton_tolerance;
  uint32_t riemann_newton_max_iterations;
  /* shallow water limiter, shallow_water/limiter.h:42-58 */
  int32_t limit_on_kinetic_energy;
  int32_t limit_on_square_velocity;
} rj_module_params;

typedef struct rj_initial_values_desc {
  int kind; /* rj_initial_state_kind */
  /* E - InitialValues direction / position, initial_values.template.h:39-49 */
  double direction[3];
  double position[3];
  double primitive[3];       /* uniform / contrast left: (rho, u, p)          */
  double primitive_right[3]; /* contrast right                                */
  double mach_number, beta;  /* isentropic vortex                             */
  /* shallow water states: circular dam break (still water depth, radius, dam
   * amplitude); uniform (h, u); ritter (time initial, left depth); smooth
   * vortex (reference depth, mach number, beta) */
  double sw_params[4];
} rj_initial_values_desc;

void rj_equation_desc_default(rj_equation_desc *eq, int equation);
void rj_module_params_default(rj_module_params *mp);
void rj_initial_values_desc_default(rj_initial_values_desc *iv, int kind);

/* ------------------------------------------------------------------------ *
 * State vectors (replace StateVector = tuple<U, precomputed, V>,
 * source/state_vector.h:47-51; MultiComponentVector, multicomponent_vector.h).
 * Device layout: component-major ("SoA"), (U_i)_c at U[c*ld + i].
 * Host layout at this boundary: the reference's AoS, (U_i)_c at i*k + c.
 * ------------------------------------------------------------------------ */
typedef struct rj_hyperbolic_module rj_hyperbolic_module;
typedef struct rj_state rj_state;

int rj_state_create(rj_state **out, rj_hyperbolic_module *hm);
void rj_state_destroy(rj_state *s);
/* host AoS [n_owned * k] -> device (copy + transpose on the module's stream) */
int rj_state_upload(rj_state *s, const double *host_U);
int rj_state_download(const rj_state *s, double *host_U);
int rj_state_download_precomputed(const rj_state *s, double *host_prec /* n_owned * n_prec */);
/* Streaming transfers: the same copies, asynchronous, on two private copy streams (one per direction)
 * so that they overlap with the sweeps of the module's stream; the host buffers must be page locked.
 * (The reference keeps its vectors in host memory, multicomponent_vector.h:266 -- these calls are what
 * a caller that owns host-side vectors uses to keep the device busy while it moves them.)
 *   upload_begin    starts host -> staging; returns at once (waits, on the device, for the previous
 *                   upload_end's transposition to have left the staging buffer)
 *   upload_end      module's stream: waits for that copy, transposes into U; no host synchronisation
 *   download_begin  module's stream: transposes U into a second staging buffer; copy stream: staging -> host
 *   download_end    blocks the host until the data has arrived; also orders the module's stream behind it
 * One upload and one download may be in flight per state; a second *_begin before the matching *_end is
 * RJ_ERR_INVALID. */
int rj_state_upload_begin(rj_state *s, const double *pinned_host_U);
int rj_state_upload_end(rj_state *s);
int rj_state_download_begin(rj_state *s, double *pinned_host_U);
int rj_state_download_end(rj_state *s);
/* U <- initial_state(x_i, t) for all owned rows (InitialValues::interpolate_hyperbolic_vector) */
int rj_state_interpolate(rj_state *s, double t);
double *rj_state_device_U(rj_state *s);           /* k * ld doubles  */
double *rj_state_device_precomputed(rj_state *s); /* n_prec * ld     */
/* dst = s * dst + b * src on U (sadd, time_integrator.template.h:18-25) */
int rj_state_sadd(rj_state *dst, double s, double b, const rj_state *src);
int rj_state_swap(rj_state *a, rj_state *b);

/* ------------------------------------------------------------------------ *
 * HyperbolicModule (source/hyperbolic_module.h:110-278)
 * ------------------------------------------------------------------------ */

/* ctor + prepare(): hyperbolic_module.template.h:28-86.  comm may be NULL
 * (single rank).  stream is a cudaStream_t or NULL for an internal stream. */
int rj_hyperbolic_module_create(rj_hyperbolic_module **out,
                                const rj_offline *od,
                                const rj_equation_desc *eq,
                                const rj_module_params *mp,
                                const rj_initial_values_desc *iv,
                                rj_comm *comm,
                                void *stream);
void rj_hyperbolic_module_destroy(rj_hyperbolic_module *hm);

/* initial_precomputed_ (hyperbolic_module.h:309; filled by
 * InitialValues::interpolate_initial_precomputed_vector in the reference): for the
 * shallow water equations one value per owned row, the bathymetry Z_i
 * (shallow_water/hyperbolic_system.h:710-728).  Zero until set.  Ghost rows are
 * exchanged here.  Returns RJ_ERR_INVALID for equations without such a vector. */
int rj_hyperbolic_module_set_initial_precomputed(rj_hyperbolic_module *hm, const double *host_values);

/* prepare_state_vector(state, t): hyperbolic_module.template.h:97-193 */
int rj_hyperbolic_module_prepare_state_vector(rj_hyperbolic_module *hm, rj_state *state, double t);

/* step<stages>(old, stage_states, stage_weights, new, tau, tau_max) -> tau:
 * hyperbolic_module.template.h:236-1211.  stages in [0, 4]
 * (hyperbolic_module.cc:26-42).  Returns RJ_RESTART when an invariant domain
 * violation occurred and the strategy is raise_exception. */
int rj_hyperbolic_module_step(rj_hyperbolic_module *hm,
                              const rj_state *old_state,
                              int stages,
                              const rj_state *const *stage_states,
                              const double *stage_weights,
                              rj_state *new_state,
                              double tau,
                              double tau_max,
                              double *tau_out);

/* cfl(Number) / cfl(): hyperbolic_module.h:232-240 */
int rj_hyperbolic_module_set_cfl(rj_hyperbolic_module *hm, double cfl);
double rj_hyperbolic_module_cfl(const rj_hyperbolic_module *hm);
/* public mutable id_violation_strategy_: hyperbolic_module.h:276-278 */
int rj_hyperbolic_module_set_id_violation_strategy(rj_hyperbolic_module *hm, int strategy);
/* n_restarts() / n_warnings(): hyperbolic_module.h:253-262 */
int rj_hyperbolic_module_counters(const rj_hyperbolic_module *hm, uint32_t *n_restarts, uint32_t *n_warnings);
/* alpha(): hyperbolic_module.h:245-248 (device pointer, ld doubles) */
const double *rj_hyperbolic_module_device_alpha(const rj_hyperbolic_module *hm);
/* block until the module's stream is idle */
int rj_hyperbolic_module_synchronize(rj_hyperbolic_module *hm);
void *rj_hyperbolic_module_stream(rj_hyperbolic_module *hm);

/* problem sizes: out = {k (problem_dimension), n_precomputed, n_bounds, n_owned, ld, n_entries_owned} */
int rj_hyperbolic_module_sizes(const rj_hyperbolic_module *hm, uint64_t out[6]);

/* Scratch introspection for per-kernel parity tests.  what: 0 alpha[n_owned],
 * 1 d_ij, 2 l_ij, 3 l_ij_next (CSR entry order, nnz), 4 p_ij (nnz*k, AoS),
 * 5 r (n_owned*k AoS), 6 bounds (n_owned*n_bounds AoS).  Host output. */
int rj_hyperbolic_module_get_scratch(rj_hyperbolic_module *hm, int what, double *host_out);
/* number of kernels launched by this module since creation */
uint64_t rj_hyperbolic_module_kernel_launches(const rj_hyperbolic_module *hm);

/* Per-kernel device time, measured with CUDA events on the module's stream
 * around every launch while profiling is on (bench.py roofline).  Slots:
 * 0 boundary conditions, 1 precompute, 2 d_ij (edge parallel), 3 alpha,
 * 4 boundary d_ij, 5 d_ii + tau, 6 low-order update, 7 P_ij + l_ij,
 * 8 high-order + next l_ij, 9 final high-order update, 10 plumbing,
 * 11 main stream blocked on the communication stream (exposed halo exchange),
 * 12 scalar all-reduces. */
#define RJ_N_PROFILE_SLOTS 13
int rj_hyperbolic_module_set_profiling(rj_hyperbolic_module *hm, int on);
int rj_hyperbolic_module_get_profile(rj_hyperbolic_module *hm, double ms[RJ_N_PROFILE_SLOTS], uint64_t counts[RJ_N_PROFILE_SLOTS]);

/* ------------------------------------------------------------------------ *
 * Postprocessor (source/postprocessor.h, postprocessor.template.h:113-281) on the device: the
 * schlieren quantity |sum_j c_ij q_j| / m_i of conserved component `first_component` (mode 0) or
 * the vorticity -- curl -- of the vector quantity starting at `first_component` (mode 1; the
 * momentum: 1), normalised like the reference: bounds over all ranks (NCCL), then
 * copysign(1 - exp(-beta ratio), q).  host_out: n_owned values in local row order;
 * bounds_out (optional): {q_max, q_min} before the normalisation.
 * ------------------------------------------------------------------------ */
int rj_postprocessor_compute(rj_hyperbolic_module *hm, const rj_state *state, int mode, int first_component, double beta,
                             double *host_out, double *bounds_out);

/* ------------------------------------------------------------------------ *
 * TimeIntegrator (source/time_integrator.h, .template.h:207-443): host-side
 * driver kept so that callers of TimeIntegrator::step() are drop-in.
 * ------------------------------------------------------------------------ */
typedef struct rj_time_integrator rj_time_integrator;

int rj_time_integrator_create(rj_time_integrator **out,
                              rj_hyperbolic_module *hm,
                              int scheme,
                              double cfl_min,
                              double cfl_max,
                              int cfl_recovery_strategy);
void rj_time_integrator_destroy(rj_time_integrator *ti);
/* step(state, t, t_final) -> tau; t_final >= DBL_MAX (the reference's default,
 * numeric_limits::max()) or NaN means "no final time"; any other value, negative ones included,
 * is enforced (tau <= t_final - t) */
int rj_time_integrator_step(rj_time_integrator *ti, rj_state *state, double t, double t_final, double *tau_out);
double rj_time_integrator_efficiency(const rj_time_integrator *ti);

/* How a cycle is driven (results are bit-identical):
 *   RJ_STEPPING_HOST    the reference's control flow, tau and the violation flag return to the host
 *                       after every step() (time_integrator.template.h:302-443);
 *   RJ_STEPPING_DEVICE  the whole cycle is enqueued back to back; tau, t + c tau and the flags stay
 *                       in device memory; one synchronisation per cycle;
 *   RJ_STEPPING_GRAPH   as DEVICE, captured into CUDA graphs (cached per state-pointer permutation)
 *                       and replayed (default; falls back to DEVICE while profiling is on).
 * The environment variable RYUJIN_B200_STEPPING = host | device | graph overrides the default. */
enum rj_stepping { RJ_STEPPING_HOST = 0, RJ_STEPPING_DEVICE = 1, RJ_STEPPING_GRAPH = 2 };
int rj_time_integrator_set_stepping(rj_time_integrator *ti, int stepping);
int rj_time_integrator_stepping(const rj_time_integrator *ti);
/* stats = {graphs captured, graph launches} */
int rj_time_integrator_graph_stats(const rj_time_integrator *ti, uint64_t stats[2]);

#ifdef __cplusplus
}
#endif

#endif /* RYUJIN_B200_H */
