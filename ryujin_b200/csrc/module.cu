// Host side of the B200 hyperbolic module: owns the device scratch, launches
// the sweep kernels of kernels.cuh in the order of HyperbolicModule::step()
// (source/hyperbolic_module.template.h:236-1211) and runs the ghost exchanges
// C1-C9 of SURVEY.md section 2b over NCCL.  Also: state vectors, the
// TimeIntegrator shim and the extern "C" surface declared in
// include/ryujin_b200.h.

#include "device_offline.h"
#include "euler.cuh"
#include "host_offline.h"
#include "initial_values.cuh"
#include "kernels.cuh"
#include "kernels_tile.cuh"
#include "shallow_water.cuh"
#include "euler_aeos.cuh"

#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

namespace rjb
{
  thread_local std::string g_last_error;

  inline void nccl_check(ncclResult_t r, const char *what, const char *file, int line)
  {
    if (r != ncclSuccess)
      throw NcclError(std::string(what) + " failed: " + ncclGetErrorString(r) + " (" + file + ":" +
                      std::to_string(line) + ")");
  }
#define RJ_NCCL(expr) ::rjb::nccl_check((expr), #expr, __FILE__, __LINE__)

  template <typename F>
  int guarded(const F &f)
  {
    try {
      return f();
    } catch (const InvalidArgument &e) {
      g_last_error = e.what();
      return RJ_ERR_INVALID;
    } catch (const CudaError &e) {
      g_last_error = e.what();
      return RJ_ERR_CUDA;
    } catch (const NcclError &e) {
      g_last_error = e.what();
      return RJ_ERR_NCCL;
    } catch (const NumericError &e) {
      g_last_error = e.what();
      return RJ_ERR_NUMERIC;
    } catch (const NoDevice &e) {
      g_last_error = e.what();
      return RJ_ERR_NO_DEVICE;
    } catch (const std::exception &e) {
      g_last_error = e.what();
      return RJ_ERR_INVALID;
    }
  }
} // namespace rjb

using namespace rjb;

struct rj_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, n_ranks = 1, device = 0;
};

struct rj_state {
  rj_hyperbolic_module *hm = nullptr;
  int device = 0; /* destruction must not depend on the module still being alive */
  double *U = nullptr;    /* k * ld */
  double *prec = nullptr; /* n_prec * ld */
  double *aux = nullptr;  /* n_aux * ld: per-node part of the flux contribution (1/rho, p) */
  /* streaming transfers (rj_state_upload_begin / _end, rj_state_download_begin / _end): private AoS
   * staging buffers and two copy streams, created on first use; not part of what swap() exchanges */
  struct Streaming {
    double *stage_in = nullptr, *stage_out = nullptr; /* n_owned * k each */
    cudaStream_t h2d = nullptr, d2h = nullptr;
    cudaEvent_t h2d_done = nullptr, in_consumed = nullptr, out_ready = nullptr, d2h_done = nullptr;
    bool upload_pending = false, download_pending = false;
  } *streaming = nullptr;
};

/* ------------------------------------------------------------------------ */

struct rj_hyperbolic_module {
  const rj_offline *od = nullptr;
  int device = 0;
  rj_comm *comm = nullptr;
  cudaStream_t stream = nullptr;
  bool owns_stream = false;
  /* Halo exchanges run on a second stream so that they overlap with the interior rows of the
   * kernel that produced the data (the reference's SynchronizationDispatch, openmp.h:141-182:
   * "fire the exchange once the export rows are done").  `cur` is the stream launch() and the
   * exchanges currently target. */
  cudaStream_t comm_stream = nullptr;
  int overlap_mode = RJ_OVERLAP_MODE_DEFAULT; /* 1: export rows, then exchange || interior rows; 2: (export rows, exchange) || interior rows */
  cudaStream_t cur = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool overlap = true;

  rj_equation_desc eq{};
  rj_module_params mp{};
  InitialValuesDevice iv{};

  double cfl = 0.2; /* hyperbolic_module.template.h:45 */
  int strategy = RJ_STRATEGY_WARN;
  uint32_t n_restarts = 0, n_warnings = 0;
  uint64_t launches = 0;

  int k = 0, n_prec = 0, n_bounds = 0, n_aux = 0;
  uint32_t max_row_length = 0; /* longest stencil row (sizes the shared memory tiles of the slice-tiled kernels) */
  /* which sweeps run as slice-tiled kernels with bulk-copied tiles (kernels_tile.cuh); bit 0: step 5,
   * bit 1: step 6.  RYUJIN_B200_TILE overrides the default (A/B measurements, bit-identical results) */
  int tile_mask = RJ_TILE_DEFAULT_MASK;

  /* scratch, hyperbolic_module.h:319-333 */
  double *alpha = nullptr, *bounds = nullptr, *r = nullptr;
  double *dij = nullptr, *lij = nullptr, *lij_next = nullptr, *pij = nullptr;
  bool fold_scalar_kernels = true; /* RYUJIN_B200_FOLD=0: the separate scalar kernels (A/B, bit-identical) */
  uint32_t small_mesh_rows = 262144; /* at most: d_ij and alpha_i as one launch (k_dij_alpha); RYUJIN_B200_SMALL_MESH_ROWS */
  double *edge_cache = nullptr; /* E_owned doubles if the equation caches a symmetric edge quantity (RJ_EDGE_CACHE) */
  bool want_edge_cache = false;
  StepScalars *scalars = nullptr;
  StepScalars *h_scalars = nullptr; /* pinned */
  double *send_buffer = nullptr, *recv_buffer = nullptr;
  size_t send_buffer_size = 0;
  double *staging = nullptr; /* n_owned * max(k, n_prec, n_bounds) */
  /* initial_precomputed_ of the reference (hyperbolic_module.h:309): bathymetry for the
   * shallow water equations, ld entries incl. ghosts, zero until set; null for Euler */
  double *initial_precomputed = nullptr;
  int n_initial_precomputed = 0;
  uint32_t *csr_to_sliced = nullptr;
  size_t nnz = 0;

  virtual ~rj_hyperbolic_module();
  /* cycle != nullptr: t = cycle->t + tau_factor * cycle->tau, evaluated on the device */
  virtual void prepare_state_vector(rj_state &state, double t, const CycleScalars *cycle = nullptr, double tau_factor = 0.) = 0;
  /* enqueue all kernels and exchanges of one update on the stream; no host synchronisation */
  virtual void enqueue_step(const rj_state &old_state,
                            int stages,
                            const rj_state *const *stage_states,
                            const double *stage_weights,
                            rj_state &new_state,
                            TauSource tau_source,
                            CycleScalars *cycle,
                            int substep) = 0;

  /* the reference call: enqueue, wait, evaluate the flags (hyperbolic_module.template.h:571-578,1190-1210) */
  int step(const rj_state &old_state,
           int stages,
           const rj_state *const *stage_states,
           const double *stage_weights,
           rj_state &new_state,
           double tau,
           double tau_max,
           double *tau_out)
  {
    TauSource src{0, tau, tau_max, 1.};
    enqueue_step(old_state, stages, stage_states, stage_weights, new_state, src, nullptr, 0);
    RJ_CUDA(cudaMemcpyAsync(h_scalars, scalars, sizeof(StepScalars), cudaMemcpyDeviceToHost, stream));
    launch(k_reset_scalars, 1, 1, scalars); /* device-resident cycles rely on clean scalars between substeps */
    RJ_CUDA(cudaStreamSynchronize(stream));
    if (profiling)
      collect_profile();
    if (h_scalars->numeric_error)
      throw NumericError("I'm sorry, Dave. I'm afraid I can't do that.\nWe crashed. "
                         "(tau_max is nan, inf or not positive)");
    *tau_out = h_scalars->tau;
    return account_violation(h_scalars->restart_needed != 0);
  }

  /* IDViolationStrategy, hyperbolic_module.template.h:1190-1210 */
  int account_violation(bool restart_needed)
  {
    if (restart_needed) {
      if (strategy == RJ_STRATEGY_WARN)
        n_warnings++;
      else {
        n_restarts++;
        return RJ_RESTART;
      }
    }
    return RJ_OK;
  }
  virtual void interpolate(rj_state &state, double t) = 0;

  StencilView view() const
  {
    StencilView v;
    v.n_owned = od->n_owned;
    v.ld = od->ld;
    v.E_owned = od->E_owned;
    v.slice_offsets = od->slice_offsets;
    v.row_length = od->row_length;
    v.cols = od->cols;
    v.transposed = od->transposed;
    v.mij = od->mij;
    v.cij = od->cij;
    v.mi = od->mi;
    v.mi_inv = od->mi_inv;
    v.measure_of_omega_inverse = 1. / od->measure_of_omega;
    return v;
  }

  void allocate()
  {
    RJ_CUDA(cudaSetDevice(od->device));
    const size_t ld = od->ld;
    auto alloc = [&](double *&p, size_t n) {
      RJ_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(double)));
      RJ_CUDA(cudaMemsetAsync(p, 0, std::max<size_t>(n, 1) * sizeof(double), stream));
    };
    alloc(alpha, ld);
    alloc(bounds, size_t(n_bounds) * ld);
    alloc(r, size_t(k) * ld);
    alloc(dij, od->E_owned);
    alloc(lij, od->E_total);
    alloc(lij_next, od->E_total);
    alloc(pij, size_t(k) * od->E_owned);
    if (want_edge_cache)
      alloc(edge_cache, od->E_owned);
    RJ_CUDA(cudaMalloc(&scalars, sizeof(StepScalars)));
    cur = stream;
    launch(k_reset_scalars, 1, 1, scalars);
    launches = 0;
    RJ_CUDA(cudaMallocHost(&h_scalars, sizeof(StepScalars)));
    send_buffer_size =
        std::max<size_t>(size_t(od->n_send_rows) * std::max(k, std::max(n_prec, n_bounds)),
                         od->n_mat_send);
    alloc(send_buffer, send_buffer_size);
    alloc(recv_buffer, size_t(od->n_relevant - od->n_owned) * std::max(k, std::max(n_prec, n_bounds)));
    alloc(staging, size_t(od->n_owned) * std::max(k, std::max(n_prec, n_bounds)));
    if (n_initial_precomputed > 0)
      alloc(initial_precomputed, size_t(n_initial_precomputed) * ld);
    RJ_CUDA(cudaStreamSynchronize(stream));
  }

  /* -------- launches -------- */
  template <typename Kernel, typename... Args>
  void launch(Kernel kernel, size_t n_threads, int block, Args... args)
  {
    if (n_threads == 0)
      return;
    const unsigned int grid = (unsigned int)((n_threads + block - 1) / block);
    if (profiling) {
      /* per-kernel device time on the launching stream (bench.py roofline) */
      if (prof_used + 2 > prof_events.size()) {
        prof_events.resize(prof_used + 2);
        RJ_CUDA(cudaEventCreate(&prof_events[prof_used]));
        RJ_CUDA(cudaEventCreate(&prof_events[prof_used + 1]));
      }
      RJ_CUDA(cudaEventRecord(prof_events[prof_used], cur));
      RJ_LAUNCH_KERNEL(kernel, grid, block, cur, args...);
      RJ_CUDA(cudaEventRecord(prof_events[prof_used + 1], cur));
      prof_slots.push_back(prof_slot);
      prof_used += 2;
    } else
      RJ_LAUNCH_KERNEL(kernel, grid, block, cur, args...);
    RJ_CUDA(cudaGetLastError());
    ++launches;
  }

  /* profiling: device time of whatever `f` enqueues on the current stream, booked on `slot` */
  template <typename F>
  void timed_region(int slot, const F &f)
  {
    if (!profiling) {
      f();
      return;
    }
    if (prof_used + 2 > prof_events.size()) {
      prof_events.resize(prof_used + 2);
      RJ_CUDA(cudaEventCreate(&prof_events[prof_used]));
      RJ_CUDA(cudaEventCreate(&prof_events[prof_used + 1]));
    }
    const size_t first = prof_used;
    prof_used += 2;
    prof_slots.push_back(slot); /* slot n <-> events 2n, 2n + 1 (collect_profile) */
    RJ_CUDA(cudaEventRecord(prof_events[first], cur));
    f();
    RJ_CUDA(cudaEventRecord(prof_events[first + 1], cur));
  }

  /* slice-tiled kernels (kernels_tile.cuh): one block of `block` threads per slice and `smem`
   * bytes of dynamic shared memory; above 48 kB the kernel needs the opt-in attribute (once) */
  template <typename Kernel, typename... Args>
  void launch_tile(Kernel kernel, uint32_t n_blocks, int block, size_t smem, Args... args)
  {
    if (n_blocks == 0)
      return;
    if (smem > 48 * 1024) {
      const void *key = reinterpret_cast<const void *>(kernel);
      if (std::find(smem_opt_in.begin(), smem_opt_in.end(), key) == smem_opt_in.end()) {
        RJ_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        smem_opt_in.push_back(key);
      }
    }
    if (profiling) {
      if (prof_used + 2 > prof_events.size()) {
        prof_events.resize(prof_used + 2);
        RJ_CUDA(cudaEventCreate(&prof_events[prof_used]));
        RJ_CUDA(cudaEventCreate(&prof_events[prof_used + 1]));
      }
      RJ_CUDA(cudaEventRecord(prof_events[prof_used], cur));
      RJ_LAUNCH_KERNEL_SMEM(kernel, n_blocks, block, smem, cur, args...);
      RJ_CUDA(cudaEventRecord(prof_events[prof_used + 1], cur));
      prof_slots.push_back(prof_slot);
      prof_used += 2;
    } else
      RJ_LAUNCH_KERNEL_SMEM(kernel, n_blocks, block, smem, cur, args...);
    RJ_CUDA(cudaGetLastError());
    ++launches;
  }
  std::vector<const void *> smem_opt_in;

  /* profiling: slots name the kernels of one update */
  enum ProfSlot {
    P_BC = 0,
    P_PRECOMPUTE,
    P_DIJ,
    P_ALPHA,
    P_BOUNDARY_DIJ,
    P_DII_TAU,
    P_LOW_ORDER,
    P_PIJ_LIJ,
    P_HIGH_ORDER_NEXT,
    P_HIGH_ORDER_FINAL,
    P_OTHER,
    P_COMM_WAIT, /* main stream blocked on the communication stream (exposed exchange time) */
    P_ALLREDUCE, /* scalar all-reduces (tau_max; restart flags) */
    P_N_SLOTS
  };
  bool profiling = false;
  int prof_slot = P_OTHER;
  std::vector<cudaEvent_t> prof_events;
  std::vector<int> prof_slots;
  size_t prof_used = 0;
  double prof_ms[P_N_SLOTS] = {};
  uint64_t prof_count[P_N_SLOTS] = {};

  void tag(int slot) { prof_slot = slot; }

  /* -------- two-stream overlap -------- */
  bool overlapping() const
  {
    return overlap && od->export_rows_first && comm && comm->n_ranks > 1 && !od->neighbors.empty() && comm_stream;
  }
  /* rows [0, n_export_rows()) have ghost columns and are ghosts elsewhere; slice aligned */
  uint32_t n_export_rows() const { return std::min<uint32_t>(od->n_owned, (od->n_export + 31u) & ~31u); }
  /* what follows on the communication stream starts after everything enqueued on the main stream */
  void fork_comm()
  {
    RJ_CUDA(cudaEventRecord(ev_fork, stream));
    RJ_CUDA(cudaStreamWaitEvent(comm_stream, ev_fork, 0));
    cur = comm_stream;
  }
  /* back to the main stream without waiting for the communication stream */
  void continue_main() { cur = stream; }
  /* the main stream waits for everything enqueued on the communication stream */
  void join_comm()
  {
    RJ_CUDA(cudaEventRecord(ev_join, comm_stream));
    cur = stream;
    if (profiling) {
      /* time the main stream spends blocked in this join = exposed communication */
      if (prof_used + 2 > prof_events.size()) {
        prof_events.resize(prof_used + 2);
        RJ_CUDA(cudaEventCreate(&prof_events[prof_used]));
        RJ_CUDA(cudaEventCreate(&prof_events[prof_used + 1]));
      }
      RJ_CUDA(cudaEventRecord(prof_events[prof_used], stream));
      RJ_CUDA(cudaStreamWaitEvent(stream, ev_join, 0));
      RJ_CUDA(cudaEventRecord(prof_events[prof_used + 1], stream));
      prof_slots.push_back(P_COMM_WAIT);
      prof_used += 2;
    } else
      RJ_CUDA(cudaStreamWaitEvent(stream, ev_join, 0));
  }

  /* call after a stream synchronize */
  void collect_profile()
  {
    for (size_t n = 0; n < prof_slots.size(); ++n) {
      float ms = 0.f;
      RJ_CUDA(cudaEventElapsedTime(&ms, prof_events[2 * n], prof_events[2 * n + 1]));
      prof_ms[prof_slots[n]] += ms;
      prof_count[prof_slots[n]]++;
    }
    prof_slots.clear();
    prof_used = 0;
  }

  /* -------- ghost exchanges (NCCL send/recv groups, cf.
   * multicomponent_vector / Partitioner::export_to_ghosted_array and
   * SparseMatrixSIMD::update_ghost_rows, sparse_matrix_simd.h:651-763) -------- */
  void exchange_vector(double *v, int n_comp)
  {
    if (!comm || od->neighbors.empty())
      return;
    const size_t ld = od->ld;
    if (od->neighbors.size() > size_t(max_neighbors))
      throw InvalidArgument("more than 26 neighbour ranks");
    /* one pack launch, ONE send and ONE receive per neighbour, one unpack launch */
    ExchangePlan send_plan{}, recv_plan{};
    send_plan.n_neighbors = recv_plan.n_neighbors = uint32_t(od->neighbors.size());
    uint32_t n_ghost = 0;
    for (size_t p = 0; p < od->neighbors.size(); ++p) {
      const auto &nb = od->neighbors[p];
      send_plan.row_offset[p] = nb.send_offset;
      send_plan.row_offset[p + 1] = nb.send_offset + nb.send_count;
      recv_plan.row_offset[p] = n_ghost;
      recv_plan.ghost_start[p] = nb.ghost_start;
      n_ghost += nb.ghost_count;
      recv_plan.row_offset[p + 1] = n_ghost;
    }
    launch(k_pack_blocks, size_t(od->n_send_rows) * n_comp, 256, send_plan, uint32_t(n_comp), uint32_t(ld), od->send_rows, v,
           send_buffer);
    RJ_NCCL(ncclGroupStart());
    for (size_t p = 0; p < od->neighbors.size(); ++p) {
      const auto &nb = od->neighbors[p];
      if (nb.send_count)
        RJ_NCCL(ncclSend(send_buffer + size_t(nb.send_offset) * n_comp, size_t(nb.send_count) * n_comp, ncclDouble, nb.rank,
                         comm->comm, cur));
      if (nb.ghost_count)
        RJ_NCCL(ncclRecv(recv_buffer + size_t(recv_plan.row_offset[p]) * n_comp, size_t(nb.ghost_count) * n_comp, ncclDouble,
                         nb.rank, comm->comm, cur));
    }
    RJ_NCCL(ncclGroupEnd());
    launch(k_unpack_blocks, size_t(n_ghost) * n_comp, 256, recv_plan, uint32_t(n_comp), uint32_t(ld), recv_buffer, v);
  }

  void exchange_matrix(double *m)
  {
    if (!comm || od->neighbors.empty())
      return;
    for (const auto &nb : od->neighbors)
      launch(k_pack_entries,
             nb.mat_send_count,
             256,
             nb.mat_send_count,
             od->mat_send_entries + nb.mat_send_offset,
             m,
             send_buffer + nb.mat_send_offset);
    RJ_NCCL(ncclGroupStart());
    for (const auto &nb : od->neighbors) {
      if (nb.mat_send_count)
        RJ_NCCL(ncclSend(send_buffer + nb.mat_send_offset,
                         nb.mat_send_count,
                         ncclDouble,
                         nb.rank,
                         comm->comm,
                         cur));
      if (nb.mat_recv_count)
        RJ_NCCL(ncclRecv(
            m + nb.mat_recv_offset, nb.mat_recv_count, ncclDouble, nb.rank, comm->comm, cur));
    }
    RJ_NCCL(ncclGroupEnd());
  }

  void ensure_csr_map()
  {
    if (csr_to_sliced)
      return;
    std::vector<uint32_t> map;
    for (uint32_t i = 0; i < od->n_owned; ++i)
      for (uint32_t c = 0; c < od->h_row_length[i]; ++c)
        map.push_back(od->entry(i, c));
    nnz = map.size();
    RJ_CUDA(cudaMalloc(&csr_to_sliced, std::max<size_t>(nnz, 1) * sizeof(uint32_t)));
    RJ_CUDA(cudaMemcpy(csr_to_sliced, map.data(), nnz * sizeof(uint32_t), cudaMemcpyHostToDevice));
  }
};

rj_hyperbolic_module::~rj_hyperbolic_module()
{
  /* garbage collected hosts may destroy the offline data first: do not touch od */
  cudaSetDevice(device);
  auto f = [](void *p) {
    if (p)
      cudaFree(p);
  };
  f(alpha);
  f(bounds);
  f(r);
  f(dij);
  f(lij);
  f(lij_next);
  f(pij);
  f(edge_cache);
  f(scalars);
  f(send_buffer);
  f(recv_buffer);
  f(staging);
  f(initial_precomputed);
  f(csr_to_sliced);
  for (auto ev : prof_events)
    cudaEventDestroy(ev);
  if (h_scalars)
    cudaFreeHost(h_scalars);
  if (ev_fork)
    cudaEventDestroy(ev_fork);
  if (ev_join)
    cudaEventDestroy(ev_join);
  if (comm_stream)
    cudaStreamDestroy(comm_stream);

  if (owns_stream && stream)
    cudaStreamDestroy(stream);
}

namespace rjb
{
  template <class Eq>
  typename Eq::Params make_params(const rj_equation_desc &eq, const rj_module_params &mp);

  template <int dim>
  EulerParams make_euler_params(const rj_equation_desc &eq, const rj_module_params &mp)
  {
    EulerParams p;
    p.gamma = eq.gamma;
    p.gamma_inverse = 1. / eq.gamma;
    p.gamma_plus_one_inverse = 1. / (eq.gamma + 1.);
    p.gamma_minus_one_inverse = 1. / (eq.gamma - 1.);
    p.reference_density = eq.reference_density;
    p.vacuum_relaxation_small = eq.vacuum_state_relaxation_small;
    p.vacuum_relaxation_large = eq.vacuum_state_relaxation_large;
    p.evc_factor = mp.indicator_evc_factor;
    p.limiter_newton_tolerance = mp.limiter_newton_tolerance;
    p.limiter_relaxation_factor = mp.limiter_relaxation_factor;
    p.limiter_newton_max_iterations = mp.limiter_newton_max_iterations;
    p.riemann_newton_tolerance = mp.riemann_newton_tolerance;
    p.riemann_newton_max_iterations = mp.riemann_newton_max_iterations;
    return p;
  }

  template <>
  EulerParams make_params<Euler<1>>(const rj_equation_desc &eq, const rj_module_params &mp)
  {
    return make_euler_params<1>(eq, mp);
  }
  template <>
  EulerParams make_params<Euler<2>>(const rj_equation_desc &eq, const rj_module_params &mp)
  {
    return make_euler_params<2>(eq, mp);
  }
  template <>
  EulerParams make_params<Euler<3>>(const rj_equation_desc &eq, const rj_module_params &mp)
  {
    return make_euler_params<3>(eq, mp);
  }


  template <int dim>
  ShallowWaterParams make_sw_params(const rj_equation_desc &eq, const rj_module_params &mp)
  {
    ShallowWaterParams p;
    p.gravity = eq.gravity;
    p.manning_friction_coefficient = eq.manning_friction_coefficient;
    p.reference_water_depth = eq.reference_water_depth;
    p.dry_state_relaxation_factor = eq.dry_state_relaxation_factor;
    p.dry_state_relaxation_small = eq.dry_state_relaxation_small;
    p.dry_state_relaxation_large = eq.dry_state_relaxation_large;
    p.evc_factor = mp.indicator_evc_factor;
    p.limiter_newton_tolerance = mp.limiter_newton_tolerance;
    p.limiter_relaxation_factor = mp.limiter_relaxation_factor;
    p.limiter_newton_max_iterations = mp.limiter_newton_max_iterations;
    p.limit_on_kinetic_energy = mp.limit_on_kinetic_energy;
    p.limit_on_square_velocity = mp.limit_on_square_velocity;
    return p;
  }
  template <>
  ShallowWaterParams make_params<ShallowWater<1>>(const rj_equation_desc &eq, const rj_module_params &mp)
  {
    return make_sw_params<1>(eq, mp);
  }
  template <>
  ShallowWaterParams make_params<ShallowWater<2>>(const rj_equation_desc &eq, const rj_module_params &mp)
  {
    return make_sw_params<2>(eq, mp);
  }


  /* the constructor and parse_parameters_call_back of the selected equation of state
   * (equation_of_state_noble_abel_stiffened_gas.h:40-75, equation_of_state_van_der_waals.h:40-80) */
  inline EulerAeosParams make_aeos_params(const rj_equation_desc &eq, const rj_module_params &mp)
  {
    EulerAeosParams p;
    p.eos = eq.equation_of_state;
    p.gamma = eq.gamma;
    p.eos_b = p.eos_q = p.eos_pinf = p.vdw_a = 0.;
    p.b = p.pinf = p.q = 0.;
    if (eq.equation_of_state == RJ_EOS_NOBLE_ABEL_STIFFENED_GAS) {
      p.eos_b = p.b = eq.covolume_b;
      p.eos_pinf = p.pinf = eq.reference_pressure;
      p.eos_q = p.q = eq.reference_specific_internal_energy;
    } else if (eq.equation_of_state == RJ_EOS_VAN_DER_WAALS) {
      p.eos_b = p.b = eq.covolume_b;
      p.vdw_a = eq.vdw_a;
      if (eq.covolume_b > 0.)
        p.pinf = eq.vdw_a / (eq.covolume_b * eq.covolume_b);
    } else if (eq.equation_of_state != RJ_EOS_POLYTROPIC_GAS)
      throw InvalidArgument("rj_hyperbolic_module_create: equation of state not available");
    p.compute_strict_bounds = eq.compute_strict_bounds;
    p.reference_density = eq.reference_density;
    p.vacuum_relaxation_small = eq.vacuum_state_relaxation_small;
    p.vacuum_relaxation_large = eq.vacuum_state_relaxation_large;
    p.evc_factor = mp.indicator_evc_factor;
    p.limiter_newton_tolerance = mp.limiter_newton_tolerance;
    p.limiter_relaxation_factor = mp.limiter_relaxation_factor;
    p.limiter_newton_max_iterations = mp.limiter_newton_max_iterations;
    return p;
  }
  template <>
  EulerAeosParams make_params<EulerAeos<1>>(const rj_equation_desc &eq, const rj_module_params &mp)
  {
    return make_aeos_params(eq, mp);
  }
  template <>
  EulerAeosParams make_params<EulerAeos<2>>(const rj_equation_desc &eq, const rj_module_params &mp)
  {
    return make_aeos_params(eq, mp);
  }
  template <>
  EulerAeosParams make_params<EulerAeos<3>>(const rj_equation_desc &eq, const rj_module_params &mp)
  {
    return make_aeos_params(eq, mp);
  }


  template <class Eq>
  struct ModuleImpl : rj_hyperbolic_module {
    typename Eq::Params prm;
    static constexpr int block = 128;

    ModuleImpl()
    {
      k = Eq::k;
      n_prec = Eq::n_prec;
      n_bounds = Eq::n_bounds;
      n_aux = Eq::n_aux;
      n_initial_precomputed = Eq::shallow_water ? 1 : 0;
      want_edge_cache = use_edge_cache;
    }
    static constexpr bool use_edge_cache = Eq::edge_cache && (RJ_EDGE_CACHE != 0);

    /* rows [begin, end), begin multiple of 32 */
    template <typename Kernel, typename... Args>
    void launch_rows(Kernel kernel, uint32_t begin, uint32_t end, Args... args)
    {
      if (end > begin)
        launch(kernel, end - begin, block, args...);
    }

    void interpolate(rj_state &state, double t) override
    {
      if (!od->positions)
        throw InvalidArgument("rj_state_interpolate: offline data has no positions");
      launch(k_interpolate<Eq, InitialValuesDevice>,
             od->n_owned,
             256,
             od->n_owned,
             od->ld,
             od->positions,
             prm,
             iv,
             t,
             state.U);
    }

    /* hyperbolic_module.template.h:97-193 */
    void prepare_state_vector(rj_state &state, double t, const CycleScalars *cycle, double tau_factor) override
    {
      RJ_CUDA(cudaSetDevice(od->device));
      const StencilView v = view();
      cur = stream;
      tag(P_BC);
      launch(k_apply_boundary_conditions<Eq, InitialValuesDevice>,
             od->n_bgroups,
             128,
             od->n_bgroups,
             od->bgroup_dof,
             od->bgroup_start,
             od->bnormal,
             od->bposition,
             od->bid,
             od->ld,
             prm,
             iv,
             t,
             cycle,
             tau_factor,
             state.U);
      /* C1, then precomputed values for owned AND ghost rows from the exchanged state (replaces
       * C2).  With neighbours: the exchange and the ghost rows run on the communication stream
       * while the owned rows are precomputed on the main stream. */
      auto precompute = [&](uint32_t begin, uint32_t end) {
        tag(P_PRECOMPUTE);
        if (end > begin)
          launch(k_precompute<Eq>, end - begin, 256, v, prm, begin, end, state.U, initial_precomputed, state.prec, state.aux);
      };
      if (overlapping()) {
        fork_comm();
        tag(P_OTHER);
        exchange_vector(state.U, Eq::k);
        precompute(od->n_owned, od->n_relevant);
        continue_main();
        precompute(0, od->n_owned);
        join_comm();
      } else {
        tag(P_OTHER);
        exchange_vector(state.U, Eq::k);
        precompute(0, od->n_relevant);
      }
      if constexpr (Eq::aeos) {
        /* second precomputation cycle over the stencil of the owned rows, then the ghost rows of
         * the precomputed vector from their owners (euler_aeos/hyperbolic_system.h:943-973,
         * hyperbolic_module.template.h:157-160) */
        tag(P_PRECOMPUTE);
        launch_rows(k_precompute_stencil<Eq>, 0, od->n_owned, v, prm, 0u, od->n_owned, state.U, state.prec);
        tag(P_OTHER);
        exchange_vector(state.prec, Eq::n_prec);
      }
    }

    template <int stages>
    void low_order(uint32_t begin,
                   uint32_t end,
                   const rj_state &old_state,
                   const rj_state *const *stage_states,
                   const double *stage_weights,
                   rj_state &new_state)
    {
      StagePointers<stages> st{};
      for (int s = 0; s < stages; ++s) {
        st.U[s] = stage_states[s]->U;
        st.prec[s] = stage_states[s]->prec;
        st.aux[s] = stage_states[s]->aux;
        st.weight[s] = stage_weights[s];
      }
      auto run = [&](auto kernel, auto pointers) {
        launch_rows(kernel,
                    begin,
                    end,
                    view(),
                    prm,
                    begin,
                    end,
                    old_state.U,
                    old_state.prec,
                    old_state.aux,
                    pointers,
                    alpha,
                    dij,
                    (const double *)edge_cache,
                    scalars,
                    new_state.U,
                    r,
                    bounds,
                    pij);
      };
        run(k_low_order<Eq, stages>, st);
    }

    /* hyperbolic_module.template.h:236-1211.
     *
     * With neighbours every sweep whose result is exchanged runs in two launches: rows
     * [0, nx) (the export rows; only they read ghost columns) first, then the exchange on the
     * communication stream concurrently with rows [nx, n_owned) on the main stream. */
    void enqueue_step(const rj_state &old_state,
                      int stages,
                      const rj_state *const *stage_states,
                      const double *stage_weights,
                      rj_state &new_state,
                      TauSource tau_source,
                      CycleScalars *cycle,
                      int substep) override
    {
      RJ_CUDA(cudaSetDevice(od->device));
      if (stages < 0 || stages > 4)
        throw InvalidArgument("rj_hyperbolic_module_step: stages must be in [0, 4]");
      if (old_state.U == new_state.U)
        throw InvalidArgument("rj_hyperbolic_module_step: old and new state must differ");
      const StencilView v = view();
      const uint32_t n_owned = od->n_owned;
      cur = stream;
      const bool ov = overlapping();
      const bool multi = comm && comm->n_ranks > 1;
      /* two row ranges per sweep when overlapping, one otherwise */
      const uint32_t nx = ov ? n_export_rows() : n_owned;

      /* sweep(f): f(begin, end) on the export rows, `exchange` on the communication stream,
       * f on the interior rows meanwhile; the main stream continues after both */
      auto sweep = [&](auto &&f, auto &&exchange) {
        if (ov && overlap_mode == 2) {
          /* the (small) export-row launch itself goes to the high-priority communication stream
           * and runs CONCURRENTLY with the interior rows on the main stream: no under-filled launch
           * in front of the interior sweep (8M DoFs per rank: 80k export rows = half a wave) */
          fork_comm();
          f(0u, nx);
          tag(P_OTHER);
          exchange();
          continue_main();
          f(nx, n_owned);
          join_comm();
        } else if (ov) {
          f(0u, nx);
          fork_comm();
          tag(P_OTHER);
          exchange();
          continue_main();
          f(nx, n_owned);
          join_comm();
        } else {
          f(0u, nx);
          tag(P_OTHER);
          exchange();
        }
      };

      /* single-rank device-resident cycle: the scalar kernels are folded into the sweeps (kernels.cuh:
       * k_dii_tau_final, k_high_order_record); the scalars are clean on entry */
      const bool folded = fold_scalar_kernels && !multi && cycle != nullptr && mp.limiter_iterations != 0;
      tag(P_OTHER);
      if (!folded)
        launch(k_reset_scalars, 1, 1, scalars);

      /* Step 2: d_ij (j > i), alpha_i; then C3 */
      tag(P_DIJ);
      const bool merged_step2 = folded && !profiling && n_owned <= small_mesh_rows;
      if (merged_step2) {
        launch(k_dij_alpha<Eq, use_edge_cache>, size_t(5) * ((n_owned + 127) / 128) * 128, 128, v, prm, od->n_slices,
               old_state.U, old_state.prec, old_state.aux, dij, alpha, edge_cache);
      } else {
        constexpr int W = RJ_DIJ_W; /* warps per slice in the edge-parallel d_ij kernel */
        launch(k_dij<Eq, W>, size_t(od->n_slices) * 32 * W, 32 * W, v, prm, 0u, od->n_slices, old_state.U, old_state.prec, dij);
      }
      sweep(
          [&](uint32_t b, uint32_t e) {
            tag(P_ALPHA);
            if (!merged_step2)
              launch_rows(k_alpha<Eq, use_edge_cache>, b, e, v, prm, b, e, old_state.U, old_state.prec,
                          old_state.aux, alpha, edge_cache);
            if (e == n_owned && (b < e || b == 0)) {
              /* Step 3: boundary d_ij, d_ii, tau_max (need d_ij only) */
              tag(P_BOUNDARY_DIJ);
              launch(k_boundary_dij<Eq>,
                     od->n_pairs,
                     128,
                     od->n_pairs,
                     od->pair_i,
                     od->pair_j,
                     od->pair_e,
                     od->pair_eT,
                     od->pair_cji,
                     od->ld,
                     od->E_owned,
                     prm,
                     old_state.U,
                     old_state.prec,
                     dij);
              tag(P_DII_TAU);
              if (folded)
                launch_rows(k_dii_tau_final, 0, n_owned, v, 0u, n_owned, cfl, dij, scalars, tau_source, cycle);
              else {
                launch_rows(k_dii_tau, 0, n_owned, v, 0u, n_owned, cfl, dij, scalars);
                tag(P_OTHER);
                launch(k_tau_local, 1, 1, scalars, tau_source, (const CycleScalars *)cycle);
              }
            }
          },
          [&] { exchange_vector(alpha, 1); });

      /* C4: all NCCL traffic of this communicator is issued on one stream */
      if (multi) {
        if (ov)
          fork_comm();
        timed_region(P_ALLREDUCE, [&] {
          RJ_NCCL(ncclAllReduce(&scalars->tau_max, &scalars->tau_max, 1, ncclDouble, ncclMin, comm->comm, cur));
        });
        if (ov)
          join_comm();
      }
      if (!folded)
        launch(k_tau_finalize, 1, 1, scalars, tau_source, cycle);

      auto low_order_rows = [&](uint32_t b, uint32_t e) {
        tag(P_LOW_ORDER);
        switch (stages) {
        case 0: low_order<0>(b, e, old_state, stage_states, stage_weights, new_state); break;
        case 1: low_order<1>(b, e, old_state, stage_states, stage_weights, new_state); break;
        case 2: low_order<2>(b, e, old_state, stage_states, stage_weights, new_state); break;
        case 3: low_order<3>(b, e, old_state, stage_states, stage_weights, new_state); break;
        case 4: low_order<4>(b, e, old_state, stage_states, stage_weights, new_state); break;
        }
      };
      auto pij_lij_rows = [&](uint32_t b, uint32_t e) {
        tag(P_PIJ_LIJ);
        if (tile_mask & 1) {
          constexpr int W = RJ_TILE_W;
          const uint32_t s0 = b / 32, s1 = (e + 31) / 32; /* b is slice aligned */
          if (s1 > s0)
            launch_tile(k_pij_lij_tile<Eq, W>, s1 - s0, 32 * W, pij_tile_smem<Eq>(max_row_length), v, prm, s0, s1,
                        max_row_length, new_state.U, r, bounds, scalars, pij, lij);
          return;
        }
        launch_rows(k_pij_lij<Eq>, b, e, v, prm, b, e, new_state.U, r, bounds, scalars, pij, lij);
      };
      const unsigned int n_iterations = mp.limiter_iterations;
      /* Step 4: low-order update, bounds, r_i, P_ij; C6 */
      sweep(low_order_rows, [&] { exchange_vector(r, Eq::k); });

      /* Step 5: P_ij correction and first l_ij; C7 */
      if (n_iterations != 0)
        sweep(pij_lij_rows, [&] { exchange_matrix(lij); });

      /* Steps 6, 7; C8 */
      for (unsigned int pass = 0; pass < n_iterations; ++pass) {
        const bool last_round = (pass + 1 == n_iterations);
        if ((n_iterations == 2) && last_round)
          std::swap(lij, lij_next);
        if (last_round) {
          tag(P_HIGH_ORDER_FINAL);
          if (folded)
            launch_rows(k_high_order_record<Eq>, 0, n_owned, v, prm, 0u, n_owned, bounds, pij, lij, lij_next, new_state.U,
                        scalars, cycle, substep);
          else
            launch_rows(k_high_order<Eq, true>, 0, n_owned, v, prm, 0u, n_owned, bounds, pij, lij, lij_next, new_state.U);
        } else
          sweep(
              [&](uint32_t b, uint32_t e) {
                tag(P_HIGH_ORDER_NEXT);
                if (tile_mask & 2) {
                  constexpr int W = RJ_TILE_W;
                  const uint32_t s0 = b / 32, s1 = (e + 31) / 32;
                  if (s1 > s0)
                    launch_tile(k_high_order_tile<Eq, W>, s1 - s0, 32 * W, high_order_tile_smem<Eq>(max_row_length), v, prm,
                                s0, s1, max_row_length, bounds, pij, lij, lij_next, new_state.U);
                  return;
                }
                  launch_rows(k_high_order<Eq, false>, b, e, v, prm, b, e, bounds, pij, lij, lij_next, new_state.U);
              },
              [&] { exchange_matrix(lij_next); });
      }

      /* C9 and the host side of :571-578,1190-1210.  Inside a device-resident cycle the flags of all
       * substeps are reduced over the ranks ONCE, at the end of the cycle (enqueue_cycle): no
       * substep reads another rank's flag before the host evaluates them. */
      tag(P_OTHER);
      if (multi && !cycle) {
        if (ov)
          fork_comm();
        RJ_NCCL(ncclAllReduce(&scalars->restart_needed, &scalars->restart_needed, 1, ncclInt, ncclMax, comm->comm, cur));
        if (ov)
          join_comm();
      }
      if (cycle && !folded)
        launch(k_cycle_record, 1, 1, (const StepScalars *)scalars, cycle, substep);
    }
  };


  /* ---------------------------------------------------------------------- *
   * TimeIntegrator, source/time_integrator.template.h:207-443.
   *
   * A scheme is a short program over the state slots {0 = the caller's state, 1.. = temp_}:
   * PREPARE(slot, c) = prepare_state_vector(slot, t + c tau); STEP = HyperbolicModule::step with
   * stage slots / weights; SADD = the convex combinations of the SSPRK schemes.  The same
   * program runs in three ways:
   *   host    the reference's control flow: every step() returns tau and the violation flag to
   *           the host (one synchronisation per substep);
   *   device  the whole cycle is enqueued back to back: tau, t + c tau and the per-substep flags
   *           stay in device memory (CycleScalars), the host synchronises once per cycle;
   *   graph   the enqueued cycle is captured into a CUDA graph (kernels, NCCL exchanges, the two
   *           small copies) and replayed; graphs are cached per state-pointer permutation.
   * All three are bit-identical: the device evaluates t + c tau and tau_max / n with the same
   * IEEE operations as the host.
   * ---------------------------------------------------------------------- */
  struct Op {
    enum Kind { PREPARE, STEP, SADD } kind;
    int a = 0, b = 0;           /* PREPARE: slot a; STEP: old a -> new b; SADD: dst a, src b */
    double c = 0., d = 0.;      /* PREPARE: time factor c; SADD: dst = c dst + d src */
    std::vector<int> stages;    /* STEP */
    std::vector<double> weights;
    double divisor = 0.;        /* STEP: > 0 marks the first step of the cycle (tau_max / divisor) */
  };

  struct Scheme {
    std::vector<Op> program;
    int result_slot = 0; /* swapped with slot 0 at the end */
    double tau_multiplier = 1.;
    int n_temp = 0;
    double efficiency = 1.;
  };

  inline Op prepare(int slot, double c) { Op o; o.kind = Op::PREPARE; o.a = slot; o.c = c; return o; }
  inline Op first_step(int from, int to, double divisor) { Op o; o.kind = Op::STEP; o.a = from; o.b = to; o.divisor = divisor; return o; }
  inline Op next_step(int from, int to, std::vector<int> st = {}, std::vector<double> w = {})
  {
    Op o; o.kind = Op::STEP; o.a = from; o.b = to; o.stages = std::move(st); o.weights = std::move(w); return o;
  }
  inline Op sadd(int dst, double c, double d, int src) { Op o; o.kind = Op::SADD; o.a = dst; o.b = src; o.c = c; o.d = d; return o; }

  inline Scheme make_scheme(int scheme)
  {
    Scheme s;
    switch (scheme) { /* time_integrator.template.h:87-137 (storage) and :302-443 (steps) */
    case RJ_SSPRK_22:
      s.program = {prepare(0, 0.), first_step(0, 1, 1.), prepare(1, 1.0), next_step(1, 2), sadd(2, 1.0 / 2.0, 1.0 / 2.0, 0)};
      s.result_slot = 2; s.n_temp = 2; s.tau_multiplier = 1.; s.efficiency = 1.;
      break;
    case RJ_SSPRK_33:
      s.program = {prepare(0, 0.), first_step(0, 1, 1.), prepare(1, 1.0), next_step(1, 2), sadd(2, 1.0 / 4.0, 3.0 / 4.0, 0),
                   prepare(2, 0.5), next_step(2, 1), sadd(1, 2.0 / 3.0, 1.0 / 3.0, 0)};
      s.result_slot = 1; s.n_temp = 2; s.tau_multiplier = 1.; s.efficiency = 1.;
      break;
    case RJ_ERK_11:
      s.program = {prepare(0, 0.), first_step(0, 1, 1.)};
      s.result_slot = 1; s.n_temp = 1; s.tau_multiplier = 1.; s.efficiency = 1.;
      break;
    case RJ_ERK_22:
      s.program = {prepare(0, 0.), first_step(0, 1, 2.), prepare(1, 1.0), next_step(1, 2, {0}, {-1.})};
      s.result_slot = 2; s.n_temp = 2; s.tau_multiplier = 2.; s.efficiency = 2.;
      break;
    case RJ_ERK_33:
      s.program = {prepare(0, 0.), first_step(0, 1, 3.), prepare(1, 1.0), next_step(1, 2, {0}, {-1.}),
                   prepare(2, 2.0), next_step(2, 3, {0, 1}, {0.75, -2.})};
      s.result_slot = 3; s.n_temp = 3; s.tau_multiplier = 3.; s.efficiency = 3.;
      break;
    case RJ_ERK_43:
      s.program = {prepare(0, 0.), first_step(0, 1, 4.), prepare(1, 1.0), next_step(1, 2, {0}, {-1.}),
                   prepare(2, 2.0), next_step(2, 3, {1}, {-1.}), prepare(3, 3.0),
                   next_step(3, 4, {1, 2}, {5. / 3., -10. / 3.})};
      s.result_slot = 4; s.n_temp = 4; s.tau_multiplier = 4.; s.efficiency = 4.;
      break;
    case RJ_ERK_54: { /* time_integrator.template.h:446-510 */
      constexpr double c = 0.2;
      constexpr double a_21 = +0.2;
      constexpr double a_31 = +0.26075582269554909;
      constexpr double a_32 = +0.13924417730445096;
      constexpr double a_41 = -0.25856517872570289;
      constexpr double a_42 = +0.91136274166280729;
      constexpr double a_43 = -0.05279756293710430;
      constexpr double a_51 = +0.21623276431503774;
      constexpr double a_52 = +0.51534223099602405;
      constexpr double a_53 = -0.81662794199265554;
      constexpr double a_54 = +0.88505294668159373;
      constexpr double a_61 = -0.10511678454691901;
      constexpr double a_62 = +0.87880047152100838;
      constexpr double a_63 = -0.58903404061484477;
      constexpr double a_64 = +0.46213380485434047;
      s.program = {prepare(0, 0.), first_step(0, 1, 5.),
                   prepare(1, 1.0), next_step(1, 2, {0}, {(a_31 - a_21) / c}),
                   prepare(2, 2.0), next_step(2, 3, {0, 1}, {(a_41 - a_31) / c, (a_42 - a_32) / c}),
                   prepare(3, 3.0), next_step(3, 4, {0, 1, 2}, {(a_51 - a_41) / c, (a_52 - a_42) / c, (a_53 - a_43) / c}),
                   prepare(4, 4.0),
                   next_step(4, 5, {0, 1, 2, 3}, {(a_61 - a_51) / c, (a_62 - a_52) / c, (a_63 - a_53) / c, (a_64 - a_54) / c})};
      s.result_slot = 5; s.n_temp = 5; s.tau_multiplier = 5.; s.efficiency = 5.;
    } break;
    default: throw InvalidArgument("unknown time stepping scheme");
    }
    return s;
  }

  enum Stepping { STEPPING_HOST = 0, STEPPING_DEVICE = 1, STEPPING_GRAPH = 2 };

  struct TimeIntegrator {
    rj_hyperbolic_module *hm = nullptr;
    int scheme_id = 0;
    Scheme scheme;
    double cfl_min = 0., cfl_max = 0.;
    int recovery = 0;
    int stepping = STEPPING_GRAPH;
    std::vector<rj_state *> temp;

    CycleScalars *d_cycle = nullptr;
    CycleScalars *h_cycle = nullptr; /* pinned: [0] input (t, tau_max), [1] result */

    struct CachedGraph {
      std::vector<const void *> key;
      cudaGraphExec_t exec = nullptr;
      uint64_t launches = 0;
    };
    std::vector<CachedGraph> graphs;
    uint64_t graph_replays = 0, graph_captures = 0;

    rj_state &slot(rj_state &sv, int n) { return n == 0 ? sv : *temp[n - 1]; }

    ~TimeIntegrator()
    {
      if (hm)
        cudaSetDevice(hm->device);
      for (auto &g : graphs)
        if (g.exec)
          cudaGraphExecDestroy(g.exec);
      if (d_cycle)
        cudaFree(d_cycle);
      if (h_cycle)
        cudaFreeHost(h_cycle);
    }
  };
} // namespace rjb

struct rj_time_integrator : rjb::TimeIntegrator {
};

namespace rjb
{
  constexpr double no_tau_max = DBL_MAX;

  /* the reference's control flow: one host round trip per substep */
  int single_step_host(rj_time_integrator &ti, rj_state &sv, double t, double tau_max, double &tau_out)
  {
    auto &hm = *ti.hm;
    double tau = 0.;
    for (const Op &op : ti.scheme.program) {
      switch (op.kind) {
      case Op::PREPARE: hm.prepare_state_vector(ti.slot(sv, op.a), op.c == 0. ? t : t + op.c * tau); break;
      case Op::SADD: {
        const int rc = rj_state_sadd(&ti.slot(sv, op.a), op.c, op.d, &ti.slot(sv, op.b));
        if (rc != RJ_OK)
          return rc;
      } break;
      case Op::STEP: {
        std::vector<const rj_state *> st;
        for (int q : op.stages)
          st.push_back(&ti.slot(sv, q));
        double out = 0.;
        const bool first = op.divisor > 0.;
        const int rc = hm.step(ti.slot(sv, op.a),
                               int(st.size()),
                               st.data(),
                               op.weights.data(),
                               ti.slot(sv, op.b),
                               first ? 0. : tau,
                               first ? tau_max / op.divisor : no_tau_max,
                               &out);
        if (rc != RJ_OK)
          return rc;
        if (first)
          tau = out;
      } break;
      }
    }
    const int rc_swap = rj_state_swap(&sv, &ti.slot(sv, ti.scheme.result_slot));
    if (rc_swap != RJ_OK)
      return rc_swap;
    tau_out = ti.scheme.tau_multiplier * tau;
    return RJ_OK;
  }

  /* enqueue one whole cycle; returns the number of STEP ops */
  int enqueue_cycle(rj_time_integrator &ti, rj_state &sv)
  {
    auto &hm = *ti.hm;
    /* (t, tau_max) of this cycle: first 16 bytes of the pinned input block */
    RJ_CUDA(cudaMemcpyAsync(ti.d_cycle, &ti.h_cycle[0], 2 * sizeof(double), cudaMemcpyHostToDevice, hm.stream));
    int substep = 0;
    for (const Op &op : ti.scheme.program) {
      switch (op.kind) {
      case Op::PREPARE: hm.prepare_state_vector(ti.slot(sv, op.a), 0., ti.d_cycle, op.c); break;
      case Op::SADD:
        if (rj_state_sadd(&ti.slot(sv, op.a), op.c, op.d, &ti.slot(sv, op.b)) != RJ_OK)
          throw CudaError(std::string("TimeIntegrator: sadd failed: ") + g_last_error);
        break;
      case Op::STEP: {
        std::vector<const rj_state *> st;
        for (int q : op.stages)
          st.push_back(&ti.slot(sv, q));
        TauSource src{op.divisor > 0. ? 1 : 2, 0., 0., op.divisor > 0. ? op.divisor : 1.};
        hm.enqueue_step(ti.slot(sv, op.a), int(st.size()), st.data(), op.weights.data(), ti.slot(sv, op.b), src, ti.d_cycle, substep);
        ++substep;
      } break;
      }
    }
    /* one reduction of the restart / numeric-error flags of all substeps over the ranks (C9) */
    if (hm.comm && hm.comm->n_ranks > 1) {
      static_assert(offsetof(CycleScalars, numeric_error) == offsetof(CycleScalars, restart) + max_substeps * sizeof(int),
                    "restart and numeric_error flags are contiguous");
      RJ_NCCL(ncclAllReduce(ti.d_cycle->restart, ti.d_cycle->restart, 2 * max_substeps, ncclInt, ncclMax, hm.comm->comm, hm.stream));
    }
    RJ_CUDA(cudaMemcpyAsync(&ti.h_cycle[1], ti.d_cycle, sizeof(CycleScalars), cudaMemcpyDeviceToHost, hm.stream));
    return substep;
  }

  int single_step_device(rj_time_integrator &ti, rj_state &sv, double t, double tau_max, double &tau_out)
  {
    auto &hm = *ti.hm;
    RJ_CUDA(cudaSetDevice(hm.device));
    ti.h_cycle[0].t = t;
    ti.h_cycle[0].tau_max = tau_max;
    int n_steps = 0;
    for (const Op &op : ti.scheme.program)
      n_steps += op.kind == Op::STEP;

    const bool use_graph = ti.stepping == STEPPING_GRAPH && !hm.profiling;
    if (!use_graph)
      enqueue_cycle(ti, sv);
    else {
      /* a graph is valid for one permutation of the state pointers, one cfl and one device */
      std::vector<const void *> key;
      /* every pointer and switch that is baked into the captured kernel arguments */
      for (const rj_state *s : {static_cast<const rj_state *>(&sv)}) {
        key.push_back(s->U);
        key.push_back(s->prec);
        key.push_back(s->aux);
      }
      for (auto *s : ti.temp) {
        key.push_back(s->U);
        key.push_back(s->prec);
        key.push_back(s->aux);
      }
      key.push_back(reinterpret_cast<const void *>(
          uintptr_t((hm.overlapping() ? unsigned(hm.overlap_mode) : 0u) | (unsigned(hm.tile_mask) << 2) |
                    (hm.fold_scalar_kernels ? 1u << 9 : 0u))));
      uint64_t cfl_bits;
      std::memcpy(&cfl_bits, &hm.cfl, sizeof(cfl_bits));
      key.push_back(reinterpret_cast<const void *>(cfl_bits));
      TimeIntegrator::CachedGraph *hit = nullptr;
      for (auto &g : ti.graphs)
        if (g.key == key)
          hit = &g;
      if (!hit) {
        if (ti.graphs.size() >= 32) { /* pointer permutations are periodic; this is a safety net */
          for (auto &g : ti.graphs)
            cudaGraphExecDestroy(g.exec);
          ti.graphs.clear();
        }
        const uint64_t launches_before = hm.launches;
        cudaGraph_t graph = nullptr;
        RJ_CUDA(cudaStreamBeginCapture(hm.stream, cudaStreamCaptureModeRelaxed));
        try {
          enqueue_cycle(ti, sv);
        } catch (...) {
          cudaStreamEndCapture(hm.stream, &graph);
          if (graph)
            cudaGraphDestroy(graph);
          throw;
        }
        RJ_CUDA(cudaStreamEndCapture(hm.stream, &graph));
        TimeIntegrator::CachedGraph g;
        g.key = key;
        g.launches = hm.launches - launches_before;
        hm.launches = launches_before;
        RJ_CUDA(cudaGraphInstantiate(&g.exec, graph, 0));
        RJ_CUDA(cudaGraphDestroy(graph));
        ti.graphs.push_back(g);
        hit = &ti.graphs.back();
        ti.graph_captures++;
      }
      RJ_CUDA(cudaGraphLaunch(hit->exec, hm.stream));
      hm.launches += hit->launches;
      ti.graph_replays++;
    }
    RJ_CUDA(cudaStreamSynchronize(hm.stream));
    if (hm.profiling)
      hm.collect_profile();

    /* the flags in the order the reference would have met them */
    const CycleScalars &res = ti.h_cycle[1];
    for (int q = 0; q < n_steps; ++q) {
      if (res.numeric_error[q])
        throw NumericError("I'm sorry, Dave. I'm afraid I can't do that.\nWe crashed. "
                           "(tau_max is nan, inf or not positive)");
      if (hm.account_violation(res.restart[q] != 0) != RJ_OK)
        return RJ_RESTART; /* the reference stops here; what the later substeps computed is dropped */
    }
    rj_state_swap(&sv, &ti.slot(sv, ti.scheme.result_slot));
    tau_out = ti.scheme.tau_multiplier * res.tau;
    return RJ_OK;
  }

  /* Each scheme returns RJ_OK / RJ_RESTART and the total step in tau_out. */
  int single_step(rj_time_integrator &ti, rj_state &sv, double t, double tau_max, double &tau_out)
  {
    if (ti.stepping == STEPPING_HOST)
      return single_step_host(ti, sv, t, tau_max, tau_out);
    return single_step_device(ti, sv, t, tau_max, tau_out);
  }
} // namespace rjb


namespace rjb
{
  template <int impl>
  __host__ __device__ inline double pow_by_impl(double x, double y)
  {
    if (impl == RJ_POW_VCL)
      return vcl_pow<false>(x, y);
    if (impl == RJ_POW_VCL_FMA)
      return vcl_pow<true>(x, y);
    return rj_pow(x, y);
  }

  template <int impl>
  __global__ void k_pow_evaluate(size_t n, const double *__restrict__ x, const double *__restrict__ y, double *__restrict__ out)
  {
    const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t < n)
      out[t] = pow_by_impl<impl>(x[t], y[t]);
  }
} // namespace rjb


/* ======================================================================== *
 *                              extern "C"
 * ======================================================================== */

extern "C" {

const char *rj_last_error(void) { return g_last_error.c_str(); }

int rj_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

const char *rj_version(void) { return "ryujin_b200 0.2 (sm_100a)"; }

int rj_pow_implementation(void) { return RJ_POW_IMPL; }


int rj_pow_evaluate(int impl, int on_device, size_t n, const double *x, const double *y, double *out)
{
  return guarded([&] {
    if (impl < 0 || impl > 2 || (n && (!x || !y || !out)))
      throw InvalidArgument("rj_pow_evaluate: invalid argument");
    if (!on_device) {
      for (size_t t = 0; t < n; ++t)
        out[t] = impl == RJ_POW_VCL ? pow_by_impl<RJ_POW_VCL>(x[t], y[t])
                                    : impl == RJ_POW_VCL_FMA ? pow_by_impl<RJ_POW_VCL_FMA>(x[t], y[t])
                                                             : pow_by_impl<RJ_POW_ACCURATE>(x[t], y[t]);
      return RJ_OK;
    }
    if (rj_device_count() == 0)
      throw NoDevice("rj_pow_evaluate: no CUDA device");
    double *d = nullptr;
    RJ_CUDA(cudaMalloc(&d, 3 * n * sizeof(double) + 8));
    try {
      RJ_CUDA(cudaMemcpy(d, x, n * sizeof(double), cudaMemcpyHostToDevice));
      RJ_CUDA(cudaMemcpy(d + n, y, n * sizeof(double), cudaMemcpyHostToDevice));
      const unsigned grid = unsigned((n + 255) / 256);
      if (n) {
        if (impl == RJ_POW_VCL)
          RJ_LAUNCH_KERNEL(k_pow_evaluate<RJ_POW_VCL>, grid, 256, 0, n, d, d + n, d + 2 * n);
        else if (impl == RJ_POW_VCL_FMA)
          RJ_LAUNCH_KERNEL(k_pow_evaluate<RJ_POW_VCL_FMA>, grid, 256, 0, n, d, d + n, d + 2 * n);
        else
          RJ_LAUNCH_KERNEL(k_pow_evaluate<RJ_POW_ACCURATE>, grid, 256, 0, n, d, d + n, d + 2 * n);
      }
      RJ_CUDA(cudaGetLastError());
      RJ_CUDA(cudaMemcpy(out, d + 2 * n, n * sizeof(double), cudaMemcpyDeviceToHost));
    } catch (...) {
      cudaFree(d);
      throw;
    }
    RJ_CUDA(cudaFree(d));
    return RJ_OK;
  });
}

int rj_offline_create(rj_offline **out, const rj_offline_desc *desc, int device)
{
  return guarded([&] {
    if (!out || !desc)
      throw InvalidArgument("rj_offline_create: null argument");
    *out = build_device_offline(*desc, device);
    return RJ_OK;
  });
}

int rj_offline_create_host_layout(rj_offline **out, const rj_offline_desc *desc)
{
  return guarded([&] {
    if (!out || !desc)
      throw InvalidArgument("rj_offline_create_host_layout: null argument");
    *out = build_device_offline(*desc, 0, /*host_only=*/true);
    return RJ_OK;
  });
}

int rj_offline_exchange_sizes(const rj_offline *od, uint64_t sizes[3])
{
  return guarded([&] {
    if (!od)
      throw InvalidArgument("null offline data");
    sizes[0] = od->neighbors.size();
    sizes[1] = od->n_send_rows;
    sizes[2] = od->n_mat_send;
    return RJ_OK;
  });
}

int rj_offline_exchange_get(const rj_offline *od, uint32_t *neighbor_info, uint32_t *send_rows, uint32_t *mat_send_entries)
{
  return guarded([&] {
    if (!od)
      throw InvalidArgument("null offline data");
    if (neighbor_info)
      for (size_t p = 0; p < od->neighbors.size(); ++p) {
        const auto &nb = od->neighbors[p];
        uint32_t *o = neighbor_info + 9 * p;
        o[0] = uint32_t(nb.rank);
        o[1] = nb.ghost_start;
        o[2] = nb.ghost_count;
        o[3] = nb.send_offset;
        o[4] = nb.send_count;
        o[5] = nb.mat_send_offset;
        o[6] = nb.mat_send_count;
        o[7] = nb.mat_recv_offset;
        o[8] = nb.mat_recv_count;
      }
    if (send_rows)
      std::memcpy(send_rows, od->h_send_rows.data(), od->h_send_rows.size() * sizeof(uint32_t));
    if (mat_send_entries)
      std::memcpy(mat_send_entries, od->h_mat_send_entries.data(), od->h_mat_send_entries.size() * sizeof(uint32_t));
    return RJ_OK;
  });
}

void rj_offline_destroy(rj_offline *od) { delete od; }

int rj_offline_layout_sizes(const rj_offline *od, uint64_t sizes[4])
{
  return guarded([&] {
    if (!od)
      throw InvalidArgument("null offline data");
    sizes[0] = od->n_slices;
    sizes[1] = od->E_owned;
    sizes[2] = od->E_total;
    sizes[3] = od->ld;
    return RJ_OK;
  });
}

int rj_offline_layout_get(const rj_offline *od,
                          uint64_t *slice_offsets,
                          uint32_t *row_length,
                          uint32_t *cols,
                          uint32_t *transposed)
{
  return guarded([&] {
    if (!od)
      throw InvalidArgument("null offline data");
    if (!od->host_only)
      RJ_CUDA(cudaSetDevice(od->device));
    if (slice_offsets)
      for (uint32_t g = 0; g <= od->n_slices; ++g)
        slice_offsets[g] = od->h_slice_offsets[g];
    if (row_length)
      for (uint32_t i = 0; i < od->n_owned; ++i)
        row_length[i] = od->h_row_length[i];
    if (od->host_only) {
      if (cols)
        std::memcpy(cols, od->h_cols.data(), size_t(od->E_owned) * sizeof(uint32_t));
      if (transposed)
        std::memcpy(transposed, od->h_transposed.data(), size_t(od->E_owned) * sizeof(uint32_t));
      return RJ_OK;
    }
    if (cols)
      RJ_CUDA(cudaMemcpy(cols, od->cols, size_t(od->E_owned) * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (transposed)
      RJ_CUDA(cudaMemcpy(
          transposed, od->transposed, size_t(od->E_owned) * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return RJ_OK;
  });
}

/* ---- communicator ---- */

int rj_comm_unique_id(uint8_t id[RJ_UNIQUE_ID_BYTES])
{
  return guarded([&] {
    static_assert(sizeof(ncclUniqueId) <= RJ_UNIQUE_ID_BYTES, "unique id size");
    ncclUniqueId uid;
    RJ_NCCL(ncclGetUniqueId(&uid));
    std::memset(id, 0, RJ_UNIQUE_ID_BYTES);
    std::memcpy(id, &uid, sizeof(uid));
    return RJ_OK;
  });
}

int rj_comm_create(rj_comm **out, const uint8_t id[RJ_UNIQUE_ID_BYTES], int rank, int n_ranks, int device)
{
  return guarded([&] {
    if (!out || !id)
      throw InvalidArgument("rj_comm_create: null argument");
    if (rj_device_count() == 0)
      throw NoDevice("rj_comm_create: no CUDA device");
    RJ_CUDA(cudaSetDevice(device));
    auto c = std::unique_ptr<rj_comm>(new rj_comm());
    c->rank = rank;
    c->n_ranks = n_ranks;
    c->device = device;
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof(uid));
    RJ_NCCL(ncclCommInitRank(&c->comm, n_ranks, uid, rank));
    *out = c.release();
    return RJ_OK;
  });
}

void rj_comm_destroy(rj_comm *c)
{
  if (!c)
    return;
  if (c->comm)
    ncclCommDestroy(c->comm);
  delete c;
}

/* ---- parameter defaults ---- */

void rj_equation_desc_default(rj_equation_desc *eq, int equation)
{
  std::memset(eq, 0, sizeof(*eq));
  eq->equation = equation;
  eq->gamma = 7. / 5.;
  eq->reference_density = 1.;
  eq->vacuum_state_relaxation_small = 1.e2;
  eq->vacuum_state_relaxation_large = 1.e4;
  eq->gravity = 9.81;
  eq->manning_friction_coefficient = 0.;
  eq->reference_water_depth = 1.;
  eq->dry_state_relaxation_factor = 2.e-1; /* shallow_water/hyperbolic_system.h:659 */
  eq->dry_state_relaxation_small = 1.e2;
  eq->dry_state_relaxation_large = 1.e4;
  eq->equation_of_state = RJ_EOS_POLYTROPIC_GAS;
  eq->compute_strict_bounds = 1; /* euler_aeos/hyperbolic_system.h:470-474 */
}

void rj_module_params_default(rj_module_params *mp)
{
  mp->indicator_evc_factor = 1.;
  mp->limiter_iterations = 2;
  mp->limiter_newton_tolerance = 1.e-10;
  mp->limiter_newton_max_iterations = 2;
  mp->limiter_relaxation_factor = 1.;
  mp->riemann_newton_tolerance = 1.e-10;
  mp->riemann_newton_max_iterations = 0;
  mp->limit_on_kinetic_energy = 0; /* shallow_water/limiter.h:50-57 */
  mp->limit_on_square_velocity = 1;
}

void rj_initial_values_desc_default(rj_initial_values_desc *iv, int kind)
{
  std::memset(iv, 0, sizeof(*iv));
  iv->kind = kind;
  iv->direction[0] = 1.;
  iv->position[0] = 1.;
  iv->primitive[0] = 1.4;
  iv->primitive[1] = kind == RJ_INITIAL_UNIFORM ? 3. : 0.;
  iv->primitive[2] = 1.;
  iv->primitive_right[0] = 1.4;
  iv->primitive_right[1] = 0.;
  iv->primitive_right[2] = 1.;
  iv->mach_number = 2.;
  iv->beta = 5.;
  switch (kind) {
  case RJ_INITIAL_SW_CIRCULAR_DAM_BREAK: /* initial_state_circular_dam_break.h:30-40 */
    iv->sw_params[0] = 0.5;
    iv->sw_params[1] = 2.5;
    iv->sw_params[2] = 2.5;
    break;
  case RJ_INITIAL_SW_UNIFORM: /* initial_state_uniform.h */
    iv->sw_params[0] = 1.;
    iv->sw_params[1] = 1.;
    break;
  case RJ_INITIAL_SW_RITTER_DAM_BREAK: /* initial_state_ritter_dam_break.h:40-48 */
    iv->sw_params[0] = 0.1;
    iv->sw_params[1] = 0.005;
    break;
  case RJ_INITIAL_SW_SMOOTH_VORTEX: /* initial_state_smooth_vortex.h:38-50 */
    iv->sw_params[0] = 1.;
    iv->sw_params[1] = 2.;
    iv->sw_params[2] = 0.1;
    break;
  default: break;
  }
}

/* ---- module ---- */

int rj_hyperbolic_module_create(rj_hyperbolic_module **out,
                                const rj_offline *od,
                                const rj_equation_desc *eq,
                                const rj_module_params *mp,
                                const rj_initial_values_desc *iv,
                                rj_comm *comm,
                                void *stream)
{
  return guarded([&] {
    if (!out || !od || !eq || !mp || !iv)
      throw InvalidArgument("rj_hyperbolic_module_create: null argument");
    if (mp->limiter_iterations > 2)
      throw InvalidArgument("The number of limiter iterations must be between [0,2]");
    if (od->host_only)
      throw NoDevice("rj_hyperbolic_module_create: offline data is a host-only layout (no CPU fallback)");
    if (!od->neighbors.empty() && !comm)
      throw InvalidArgument("rj_hyperbolic_module_create: offline data has ghost rows but comm is NULL");
    RJ_CUDA(cudaSetDevice(od->device));

    std::unique_ptr<rj_hyperbolic_module> hm;
    auto make = [&](auto tag) {
      using Eq = decltype(tag);
      auto *m = new ModuleImpl<Eq>();
      m->prm = make_params<Eq>(*eq, *mp);
      hm.reset(m);
    };
    if (eq->equation == RJ_EQUATION_EULER) {
      if (od->dim == 1)
        make(Euler<1>());
      else if (od->dim == 2)
        make(Euler<2>());
      else
        make(Euler<3>());
    } else if (eq->equation == RJ_EQUATION_SHALLOW_WATER) {
      /* the reference instantiates shallow water for dim 1 and 2 */
      if (od->dim == 1)
        make(ShallowWater<1>());
      else if (od->dim == 2)
        make(ShallowWater<2>());
      else
        throw InvalidArgument("rj_hyperbolic_module_create: shallow water needs dim 1 or 2");
    } else if (eq->equation == RJ_EQUATION_EULER_AEOS) {
      /* dynamic boundaries trap in the reference (euler_aeos/hyperbolic_system.h:1315) */
      for (int32_t id : od->h_boundary_ids)
        if (id == RJ_BOUNDARY_DYNAMIC)
          throw InvalidArgument("rj_hyperbolic_module_create: euler aeos has no dynamic boundary condition");
      if (od->dim == 1)
        make(EulerAeos<1>());
      else if (od->dim == 2)
        make(EulerAeos<2>());
      else
        make(EulerAeos<3>());
    } else
      throw InvalidArgument("rj_hyperbolic_module_create: equation module not available");

    hm->od = od;
    hm->device = od->device;
    for (uint32_t len : od->h_row_length)
      hm->max_row_length = std::max(hm->max_row_length, len);
    if (const char *e = std::getenv("RYUJIN_B200_TILE"))
      hm->tile_mask = std::atoi(e);
    /* the tiles of the widest slice must fit the 227 kB of shared memory of one block */
    if (high_order_tile_smem<Euler<3>>(hm->max_row_length) > 200 * 1024)
      hm->tile_mask = 0;
    hm->comm = comm;
    hm->eq = *eq;
    hm->mp = *mp;
    if (stream)
      hm->stream = static_cast<cudaStream_t>(stream);
    else {
      RJ_CUDA(cudaStreamCreateWithFlags(&hm->stream, cudaStreamNonBlocking));
      hm->owns_stream = true;
    }
    hm->cur = hm->stream;
    if (comm && comm->n_ranks > 1) {
      /* highest priority: a ready exchange must not queue behind the blocks of the interior sweep */
      int prio_low = 0, prio_high = 0;
      RJ_CUDA(cudaDeviceGetStreamPriorityRange(&prio_low, &prio_high));
      RJ_CUDA(cudaStreamCreateWithPriority(&hm->comm_stream, cudaStreamNonBlocking, prio_high));
      RJ_CUDA(cudaEventCreateWithFlags(&hm->ev_fork, cudaEventDisableTiming));
      RJ_CUDA(cudaEventCreateWithFlags(&hm->ev_join, cudaEventDisableTiming));
    }
    if (const char *env = std::getenv("RYUJIN_B200_SMALL_MESH_ROWS"))
      hm->small_mesh_rows = uint32_t(std::min<long>(std::max<long>(std::atol(env), 0), 0x7fffffffL));
    if (const char *env = std::getenv("RYUJIN_B200_FOLD"))
      hm->fold_scalar_kernels = std::string(env) != "0";
    if (const char *env = std::getenv("RYUJIN_B200_OVERLAP")) {
      hm->overlap = std::string(env) != "0";
      if (hm->overlap)
        hm->overlap_mode = std::atoi(env) == 2 ? 2 : 1;
    }

    /* InitialValues::parse_parameters_callback: normalize direction */
    InitialValuesDevice &d = hm->iv;
    d.kind = iv->kind;
    double nrm = 0.;
    for (int q = 0; q < od->dim; ++q)
      nrm += iv->direction[q] * iv->direction[q];
    nrm = std::sqrt(nrm);
    if (nrm == 0.)
      throw InvalidArgument("Initial direction is set to the zero vector.");
    for (int q = 0; q < 3; ++q) {
      d.direction[q] = q < od->dim ? iv->direction[q] / nrm : 0.;
      d.position[q] = q < od->dim ? iv->position[q] : 0.;
      d.primitive[q] = iv->primitive[q];
      d.primitive_right[q] = iv->primitive_right[q];
    }
    d.mach_number = iv->mach_number;
    d.beta = iv->beta;
    for (int q = 0; q < 4; ++q)
      d.sw[q] = iv->sw_params[q];

    hm->allocate();
    *out = hm.release();
    return RJ_OK;
  });
}

void rj_hyperbolic_module_destroy(rj_hyperbolic_module *hm) { delete hm; }

int rj_hyperbolic_module_set_initial_precomputed(rj_hyperbolic_module *hm, const double *host_values)
{
  return guarded([&] {
    if (!hm || !host_values)
      throw InvalidArgument("null argument");
    if (hm->n_initial_precomputed == 0)
      throw InvalidArgument("this equation has no initial precomputed vector");
    RJ_CUDA(cudaSetDevice(hm->device));
    const int nc = hm->n_initial_precomputed;
    const uint32_t n = hm->od->n_owned;
    RJ_CUDA(cudaMemcpyAsync(
        hm->staging, host_values, size_t(n) * nc * sizeof(double), cudaMemcpyHostToDevice, hm->stream));
    hm->launch(k_aos_to_soa, size_t(n) * nc, 256, n, uint32_t(nc), hm->od->ld, hm->staging, hm->initial_precomputed);
    hm->exchange_vector(hm->initial_precomputed, nc);
    RJ_CUDA(cudaStreamSynchronize(hm->stream));
    return RJ_OK;
  });
}

int rj_hyperbolic_module_prepare_state_vector(rj_hyperbolic_module *hm, rj_state *state, double t)
{
  return guarded([&] {
    if (!hm || !state)
      throw InvalidArgument("null argument");
    hm->prepare_state_vector(*state, t);
    return RJ_OK;
  });
}

int rj_hyperbolic_module_step(rj_hyperbolic_module *hm,
                              const rj_state *old_state,
                              int stages,
                              const rj_state *const *stage_states,
                              const double *stage_weights,
                              rj_state *new_state,
                              double tau,
                              double tau_max,
                              double *tau_out)
{
  return guarded([&] {
    if (!hm || !old_state || !new_state || !tau_out)
      throw InvalidArgument("rj_hyperbolic_module_step: null argument");
    if (stages < 0 || stages > 4)
      throw InvalidArgument("rj_hyperbolic_module_step: stages must be in [0, 4]");
    if (old_state->hm != hm || new_state->hm != hm)
      throw InvalidArgument("rj_hyperbolic_module_step: state vector of another module");
    if (stages > 0 && (!stage_states || !stage_weights))
      throw InvalidArgument("rj_hyperbolic_module_step: null stage arrays with stages > 0");
    for (int q = 0; q < stages; ++q)
      if (!stage_states[q] || stage_states[q]->hm != hm)
        throw InvalidArgument("rj_hyperbolic_module_step: null stage state or stage state of another module");
    return hm->step(*old_state, stages, stage_states, stage_weights, *new_state, tau, tau_max, tau_out);
  });
}

int rj_hyperbolic_module_set_cfl(rj_hyperbolic_module *hm, double cfl)
{
  return guarded([&] {
    if (!hm || !(cfl > 0.))
      throw InvalidArgument("rj_hyperbolic_module_set_cfl: invalid argument");
    hm->cfl = cfl;
    return RJ_OK;
  });
}

double rj_hyperbolic_module_cfl(const rj_hyperbolic_module *hm) { return hm->cfl; }

int rj_hyperbolic_module_set_id_violation_strategy(rj_hyperbolic_module *hm, int strategy)
{
  return guarded([&] {
    if (!hm || (strategy != RJ_STRATEGY_WARN && strategy != RJ_STRATEGY_RAISE_EXCEPTION))
      throw InvalidArgument("invalid id violation strategy");
    hm->strategy = strategy;
    return RJ_OK;
  });
}

int rj_hyperbolic_module_counters(const rj_hyperbolic_module *hm, uint32_t *n_restarts, uint32_t *n_warnings)
{
  return guarded([&] {
    if (!hm)
      throw InvalidArgument("null argument");
    if (n_restarts)
      *n_restarts = hm->n_restarts;
    if (n_warnings)
      *n_warnings = hm->n_warnings;
    return RJ_OK;
  });
}

const double *rj_hyperbolic_module_device_alpha(const rj_hyperbolic_module *hm) { return hm->alpha; }

int rj_hyperbolic_module_synchronize(rj_hyperbolic_module *hm)
{
  return guarded([&] {
    RJ_CUDA(cudaSetDevice(hm->od->device));
    RJ_CUDA(cudaStreamSynchronize(hm->stream));
    return RJ_OK;
  });
}

void *rj_hyperbolic_module_stream(rj_hyperbolic_module *hm) { return hm->stream; }

int rj_hyperbolic_module_sizes(const rj_hyperbolic_module *hm, uint64_t out[6])
{
  return guarded([&] {
    if (!hm)
      throw InvalidArgument("null argument");
    out[0] = hm->k;
    out[1] = hm->n_prec;
    out[2] = hm->n_bounds;
    out[3] = hm->od->n_owned;
    out[4] = hm->od->ld;
    out[5] = hm->od->E_owned;
    return RJ_OK;
  });
}

uint64_t rj_hyperbolic_module_kernel_launches(const rj_hyperbolic_module *hm) { return hm->launches; }

int rj_hyperbolic_module_set_profiling(rj_hyperbolic_module *hm, int on)
{
  return guarded([&] {
    if (!hm)
      throw InvalidArgument("null argument");
    RJ_CUDA(cudaSetDevice(hm->device));
    RJ_CUDA(cudaStreamSynchronize(hm->stream));
    if (hm->profiling)
      hm->collect_profile();
    hm->profiling = on != 0;
    if (on)
      for (int q = 0; q < rj_hyperbolic_module::P_N_SLOTS; ++q) {
        hm->prof_ms[q] = 0.;
        hm->prof_count[q] = 0;
      }
    return RJ_OK;
  });
}

int rj_hyperbolic_module_get_profile(rj_hyperbolic_module *hm, double ms[RJ_N_PROFILE_SLOTS], uint64_t counts[RJ_N_PROFILE_SLOTS])
{
  return guarded([&] {
    if (!hm)
      throw InvalidArgument("null argument");
    static_assert(RJ_N_PROFILE_SLOTS == rj_hyperbolic_module::P_N_SLOTS, "profile slots");
    RJ_CUDA(cudaSetDevice(hm->device));
    RJ_CUDA(cudaStreamSynchronize(hm->stream));
    if (hm->profiling)
      hm->collect_profile();
    for (int q = 0; q < RJ_N_PROFILE_SLOTS; ++q) {
      ms[q] = hm->prof_ms[q];
      counts[q] = hm->prof_count[q];
    }
    return RJ_OK;
  });
}

int rj_hyperbolic_module_get_scratch(rj_hyperbolic_module *hm, int what, double *host_out)
{
  return guarded([&] {
    if (!hm || !host_out)
      throw InvalidArgument("null argument");
    RJ_CUDA(cudaSetDevice(hm->od->device));
    const auto *od = hm->od;
    const uint32_t n = od->n_owned, ld = od->ld;
    auto vector_out = [&](const double *v, int n_comp) {
      hm->launch(k_soa_to_aos, size_t(n) * n_comp, 256, n, uint32_t(n_comp), ld, v, hm->staging);
      RJ_CUDA(cudaMemcpyAsync(
          host_out, hm->staging, size_t(n) * n_comp * sizeof(double), cudaMemcpyDeviceToHost, hm->stream));
      RJ_CUDA(cudaStreamSynchronize(hm->stream));
    };
    auto matrix_out = [&](const double *m, int n_comp) {
      hm->ensure_csr_map();
      double *tmp = nullptr;
      RJ_CUDA(cudaMalloc(&tmp, std::max<size_t>(hm->nnz, 1) * sizeof(double)));
      std::vector<double> h(hm->nnz);
      for (int c = 0; c < n_comp; ++c) {
        hm->launch(k_gather_entries, hm->nnz, 256, hm->nnz, hm->csr_to_sliced, m + size_t(c) * od->E_owned, tmp);
        RJ_CUDA(cudaMemcpyAsync(h.data(), tmp, hm->nnz * sizeof(double), cudaMemcpyDeviceToHost, hm->stream));
        RJ_CUDA(cudaStreamSynchronize(hm->stream));
        for (size_t e = 0; e < hm->nnz; ++e)
          host_out[e * n_comp + c] = h[e];
      }
      cudaFree(tmp);
    };
    switch (what) {
    case 0: vector_out(hm->alpha, 1); break;
    case 1: matrix_out(hm->dij, 1); break;
    case 2: matrix_out(hm->lij, 1); break;
    case 3: matrix_out(hm->lij_next, 1); break;
    case 4: matrix_out(hm->pij, hm->k); break;
    case 5: vector_out(hm->r, hm->k); break;
    case 6: vector_out(hm->bounds, hm->n_bounds); break;
    default: throw InvalidArgument("rj_hyperbolic_module_get_scratch: unknown id");
    }
    return RJ_OK;
  });
}

/* Postprocessor::compute (source/postprocessor.template.h:113-281) for one quantity */
int rj_postprocessor_compute(rj_hyperbolic_module *hm, const rj_state *state, int mode, int first_component, double beta,
                             double *host_out, double *bounds_out)
{
  return guarded([&] {
    if (!hm || !state || !host_out || state->hm != hm)
      throw InvalidArgument("rj_postprocessor_compute: invalid argument");
    const int dim = hm->od->dim;
    if (mode < 0 || mode > 1 || first_component < 0 || first_component + (mode == 1 ? dim : 1) > hm->k)
      throw InvalidArgument("rj_postprocessor_compute: component out of range");
    if (mode == 1 && dim < 2)
      throw InvalidArgument("rj_postprocessor_compute: no vorticity in 1D");
    RJ_CUDA(cudaSetDevice(hm->od->device));
    hm->cur = hm->stream;
    const auto *od = hm->od;
    const uint32_t n = od->n_owned;
    const StencilView v = hm->view();
    PostprocessorScalars *ps = nullptr;
    RJ_CUDA(cudaMalloc(&ps, sizeof(PostprocessorScalars)));
    double *q = hm->staging; /* n_owned doubles */
    try {
      hm->launch(k_postprocess_reset, 1, 1, ps);
      if (dim == 1)
        hm->launch(k_postprocess_quantity<1>, n, 128, v, 0u, n, mode, uint32_t(first_component), state->U, q, ps);
      else if (dim == 2)
        hm->launch(k_postprocess_quantity<2>, n, 128, v, 0u, n, mode, uint32_t(first_component), state->U, q, ps);
      else
        hm->launch(k_postprocess_quantity<3>, n, 128, v, 0u, n, mode, uint32_t(first_component), state->U, q, ps);
      hm->launch(k_postprocess_bounds, 1, 1, ps);
      if (hm->comm && hm->comm->n_ranks > 1) {
        RJ_NCCL(ncclAllReduce(&ps->q_max, &ps->q_max, 1, ncclDouble, ncclMax, hm->comm->comm, hm->stream));
        RJ_NCCL(ncclAllReduce(&ps->q_min, &ps->q_min, 1, ncclDouble, ncclMin, hm->comm->comm, hm->stream));
      }
      hm->launch(k_postprocess_normalize, n, 256, n, beta, (const PostprocessorScalars *)ps, q);
      RJ_CUDA(cudaMemcpyAsync(host_out, q, size_t(n) * sizeof(double), cudaMemcpyDeviceToHost, hm->stream));
      PostprocessorScalars h{};
      RJ_CUDA(cudaMemcpyAsync(&h, ps, sizeof(h), cudaMemcpyDeviceToHost, hm->stream));
      RJ_CUDA(cudaStreamSynchronize(hm->stream));
      if (bounds_out) {
        bounds_out[0] = h.q_max;
        bounds_out[1] = h.q_min;
      }
    } catch (...) {
      cudaFree(ps);
      throw;
    }
    RJ_CUDA(cudaFree(ps));
    return RJ_OK;
  });
}

/* ---- state vectors ---- */

int rj_state_create(rj_state **out, rj_hyperbolic_module *hm)
{
  return guarded([&] {
    if (!out || !hm)
      throw InvalidArgument("null argument");
    RJ_CUDA(cudaSetDevice(hm->od->device));
    auto s = std::unique_ptr<rj_state>(new rj_state());
    s->hm = hm;
    s->device = hm->od->device;
    const size_t ld = hm->od->ld;
    RJ_CUDA(cudaMalloc(&s->U, size_t(hm->k) * ld * sizeof(double)));
    RJ_CUDA(cudaMalloc(&s->prec, size_t(hm->n_prec) * ld * sizeof(double)));
    RJ_CUDA(cudaMalloc(&s->aux, size_t(hm->n_aux) * ld * sizeof(double)));
    RJ_CUDA(cudaMemsetAsync(s->aux, 0, size_t(hm->n_aux) * ld * sizeof(double), hm->stream));
    RJ_CUDA(cudaMemsetAsync(s->U, 0, size_t(hm->k) * ld * sizeof(double), hm->stream));
    RJ_CUDA(cudaMemsetAsync(s->prec, 0, size_t(hm->n_prec) * ld * sizeof(double), hm->stream));
    RJ_CUDA(cudaStreamSynchronize(hm->stream));
    *out = s.release();
    return RJ_OK;
  });
}

void rj_state_destroy(rj_state *s)
{
  if (!s)
    return;
  cudaSetDevice(s->device);
  cudaFree(s->U);
  cudaFree(s->prec);
  cudaFree(s->aux);
  if (auto *x = s->streaming) {
    if (x->h2d)
      cudaStreamSynchronize(x->h2d);
    if (x->d2h)
      cudaStreamSynchronize(x->d2h);
    cudaFree(x->stage_in);
    cudaFree(x->stage_out);
    for (cudaEvent_t e : {x->h2d_done, x->in_consumed, x->out_ready, x->d2h_done})
      if (e)
        cudaEventDestroy(e);
    if (x->h2d)
      cudaStreamDestroy(x->h2d);
    if (x->d2h)
      cudaStreamDestroy(x->d2h);
    delete x;
  }
  delete s;
}

int rj_state_upload(rj_state *s, const double *host_U)
{
  return guarded([&] {
    if (!s || !host_U)
      throw InvalidArgument("null argument");
    auto *hm = s->hm;
    RJ_CUDA(cudaSetDevice(hm->od->device));
    const uint32_t n = hm->od->n_owned, k = hm->k;
    RJ_CUDA(cudaMemcpyAsync(
        hm->staging, host_U, size_t(n) * k * sizeof(double), cudaMemcpyHostToDevice, hm->stream));
    hm->launch(k_aos_to_soa, size_t(n) * k, 256, n, k, hm->od->ld, hm->staging, s->U);
    RJ_CUDA(cudaStreamSynchronize(hm->stream));
    return RJ_OK;
  });
}

int rj_state_download(const rj_state *s, double *host_U)
{
  return guarded([&] {
    if (!s || !host_U)
      throw InvalidArgument("null argument");
    auto *hm = s->hm;
    RJ_CUDA(cudaSetDevice(hm->od->device));
    const uint32_t n = hm->od->n_owned, k = hm->k;
    hm->launch(k_soa_to_aos, size_t(n) * k, 256, n, k, hm->od->ld, s->U, hm->staging);
    RJ_CUDA(cudaMemcpyAsync(
        host_U, hm->staging, size_t(n) * k * sizeof(double), cudaMemcpyDeviceToHost, hm->stream));
    RJ_CUDA(cudaStreamSynchronize(hm->stream));
    return RJ_OK;
  });
}

/* ---- streaming transfers: host <-> device copies that overlap with the sweeps ----
 * The copies run on two private streams (one per direction: both copy engines), the transpositions
 * AoS <-> SoA on the module's stream; events order them, the host never blocks except in
 * rj_state_download_end().  Host buffers must be page locked. */
static rj_state::Streaming &streaming_of(rj_state *s)
{
  if (!s->streaming) {
    auto x = std::unique_ptr<rj_state::Streaming>(new rj_state::Streaming());
    const size_t bytes = std::max<size_t>(size_t(s->hm->od->n_owned) * s->hm->k, 1) * sizeof(double);
    RJ_CUDA(cudaMalloc(&x->stage_in, bytes));
    RJ_CUDA(cudaMalloc(&x->stage_out, bytes));
    RJ_CUDA(cudaStreamCreateWithFlags(&x->h2d, cudaStreamNonBlocking));
    RJ_CUDA(cudaStreamCreateWithFlags(&x->d2h, cudaStreamNonBlocking));
    for (cudaEvent_t *e : {&x->h2d_done, &x->in_consumed, &x->out_ready, &x->d2h_done})
      RJ_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    s->streaming = x.release();
  }
  return *s->streaming;
}

int rj_state_upload_begin(rj_state *s, const double *pinned_host_U)
{
  return guarded([&] {
    if (!s || !pinned_host_U)
      throw InvalidArgument("null argument");
    auto *hm = s->hm;
    RJ_CUDA(cudaSetDevice(hm->od->device));
    auto &x = streaming_of(s);
    if (x.upload_pending)
      throw InvalidArgument("rj_state_upload_begin: the previous upload was not completed with rj_state_upload_end");
    /* the staging buffer is free once the previous transposition out of it has run */
    RJ_CUDA(cudaStreamWaitEvent(x.h2d, x.in_consumed, 0));
    RJ_CUDA(cudaMemcpyAsync(
        x.stage_in, pinned_host_U, size_t(hm->od->n_owned) * hm->k * sizeof(double), cudaMemcpyHostToDevice, x.h2d));
    RJ_CUDA(cudaEventRecord(x.h2d_done, x.h2d));
    x.upload_pending = true;
    return RJ_OK;
  });
}

int rj_state_upload_end(rj_state *s)
{
  return guarded([&] {
    if (!s)
      throw InvalidArgument("null argument");
    auto *hm = s->hm;
    RJ_CUDA(cudaSetDevice(hm->od->device));
    auto &x = streaming_of(s);
    if (!x.upload_pending)
      throw InvalidArgument("rj_state_upload_end without rj_state_upload_begin");
    RJ_CUDA(cudaStreamWaitEvent(hm->stream, x.h2d_done, 0));
    const uint32_t n = hm->od->n_owned, k = hm->k;
    hm->launch(k_aos_to_soa, size_t(n) * k, 256, n, k, hm->od->ld, x.stage_in, s->U);
    RJ_CUDA(cudaEventRecord(x.in_consumed, hm->stream));
    x.upload_pending = false;
    return RJ_OK;
  });
}

int rj_state_download_begin(rj_state *s, double *pinned_host_U)
{
  return guarded([&] {
    if (!s || !pinned_host_U)
      throw InvalidArgument("null argument");
    auto *hm = s->hm;
    RJ_CUDA(cudaSetDevice(hm->od->device));
    auto &x = streaming_of(s);
    if (x.download_pending)
      throw InvalidArgument("rj_state_download_begin: the previous download was not completed with rj_state_download_end");
    const uint32_t n = hm->od->n_owned, k = hm->k;
    hm->launch(k_soa_to_aos, size_t(n) * k, 256, n, k, hm->od->ld, s->U, x.stage_out);
    RJ_CUDA(cudaEventRecord(x.out_ready, hm->stream));
    RJ_CUDA(cudaStreamWaitEvent(x.d2h, x.out_ready, 0));
    RJ_CUDA(cudaMemcpyAsync(pinned_host_U, x.stage_out, size_t(n) * k * sizeof(double), cudaMemcpyDeviceToHost, x.d2h));
    RJ_CUDA(cudaEventRecord(x.d2h_done, x.d2h));
    x.download_pending = true;
    return RJ_OK;
  });
}

int rj_state_download_end(rj_state *s)
{
  return guarded([&] {
    if (!s)
      throw InvalidArgument("null argument");
    auto *hm = s->hm;
    RJ_CUDA(cudaSetDevice(hm->od->device));
    auto &x = streaming_of(s);
    if (!x.download_pending)
      throw InvalidArgument("rj_state_download_end without rj_state_download_begin");
    /* the module's stream is ordered behind the copy as well (the staging buffer is reused by the
     * next rj_state_download_begin; an event recorded on the stream afterwards includes the copy) */
    RJ_CUDA(cudaStreamWaitEvent(hm->stream, x.d2h_done, 0));
    RJ_CUDA(cudaEventSynchronize(x.d2h_done));
    x.download_pending = false;
    return RJ_OK;
  });
}

int rj_state_download_precomputed(const rj_state *s, double *host_prec)
{
  return guarded([&] {
    if (!s || !host_prec)
      throw InvalidArgument("null argument");
    auto *hm = s->hm;
    RJ_CUDA(cudaSetDevice(hm->od->device));
    const uint32_t n = hm->od->n_owned, k = hm->n_prec;
    hm->launch(k_soa_to_aos, size_t(n) * k, 256, n, k, hm->od->ld, s->prec, hm->staging);
    RJ_CUDA(cudaMemcpyAsync(
        host_prec, hm->staging, size_t(n) * k * sizeof(double), cudaMemcpyDeviceToHost, hm->stream));
    RJ_CUDA(cudaStreamSynchronize(hm->stream));
    return RJ_OK;
  });
}

int rj_state_interpolate(rj_state *s, double t)
{
  return guarded([&] {
    if (!s)
      throw InvalidArgument("null argument");
    RJ_CUDA(cudaSetDevice(s->hm->od->device));
    s->hm->interpolate(*s, t);
    return RJ_OK;
  });
}

double *rj_state_device_U(rj_state *s) { return s->U; }
double *rj_state_device_precomputed(rj_state *s) { return s->prec; }

int rj_state_sadd(rj_state *dst, double s, double b, const rj_state *src)
{
  return guarded([&] {
    if (!dst || !src || dst->hm != src->hm)
      throw InvalidArgument("rj_state_sadd: states of different modules");
    auto *hm = dst->hm;
    RJ_CUDA(cudaSetDevice(hm->od->device));
    const size_t n = size_t(hm->k) * hm->od->ld;
    hm->launch(k_sadd, n, 256, n, s, b, dst->U, src->U);
    return RJ_OK;
  });
}

int rj_state_swap(rj_state *a, rj_state *b)
{
  return guarded([&] {
    if (!a || !b || a->hm != b->hm)
      throw InvalidArgument("rj_state_swap: states of different modules");
    std::swap(a->U, b->U);
    std::swap(a->prec, b->prec);
    std::swap(a->aux, b->aux);
    return RJ_OK;
  });
}

/* ---- time integrator ---- */

int rj_time_integrator_create(rj_time_integrator **out,
                              rj_hyperbolic_module *hm,
                              int scheme,
                              double cfl_min,
                              double cfl_max,
                              int cfl_recovery_strategy)
{
  return guarded([&] {
    if (!out || !hm)
      throw InvalidArgument("null argument");
    if (!(cfl_min > 0.))
      throw InvalidArgument("cfl min must be a positive value");
    if (!(cfl_max >= cfl_min))
      throw InvalidArgument("cfl max must be greater than or equal to cfl min");
    auto ti = std::unique_ptr<rj_time_integrator>(new rj_time_integrator());
    ti->hm = hm;
    ti->scheme_id = scheme;
    ti->scheme = make_scheme(scheme);
    ti->cfl_min = cfl_min;
    ti->cfl_max = cfl_max;
    ti->recovery = cfl_recovery_strategy;
    if (const char *env = std::getenv("RYUJIN_B200_STEPPING")) {
      const std::string v(env);
      ti->stepping = v == "host" ? STEPPING_HOST : v == "device" ? STEPPING_DEVICE : v == "graph" ? STEPPING_GRAPH : ti->stepping;
    }
    RJ_CUDA(cudaSetDevice(hm->device));
    RJ_CUDA(cudaMalloc(&ti->d_cycle, sizeof(CycleScalars)));
    RJ_CUDA(cudaMemset(ti->d_cycle, 0, sizeof(CycleScalars)));
    RJ_CUDA(cudaMallocHost(&ti->h_cycle, 2 * sizeof(CycleScalars)));
    std::memset(ti->h_cycle, 0, 2 * sizeof(CycleScalars));
    for (int i = 0; i < ti->scheme.n_temp; ++i) {
      rj_state *s = nullptr;
      const int rc = rj_state_create(&s, hm);
      if (rc != RJ_OK)
        throw CudaError(g_last_error);
      ti->temp.push_back(s);
    }
    hm->cfl = cfl_max;
    *out = ti.release();
    return RJ_OK;
  });
}

void rj_time_integrator_destroy(rj_time_integrator *ti)
{
  if (!ti)
    return;
  for (auto *s : ti->temp)
    rj_state_destroy(s);
  delete ti;
}

/* time_integrator.template.h:207-275 */
int rj_time_integrator_step(rj_time_integrator *ti, rj_state *state, double t, double t_final, double *tau_out)
{
  return guarded([&] {
    if (!ti || !state || !tau_out)
      throw InvalidArgument("null argument");
    auto &hm = *ti->hm;
    /* unbounded: t_final = DBL_MAX or larger (the reference's default argument is
     * numeric_limits::max(), time_integrator.h) or NaN; a negative final time is a final time */
    const double tau_max = (t_final != t_final || t_final >= DBL_MAX) ? DBL_MAX : t_final - t;

    if (ti->recovery == RJ_CFL_RECOVERY_BANG_BANG) {
      hm.strategy = RJ_STRATEGY_RAISE_EXCEPTION;
      hm.cfl = ti->cfl_max;
    }
    double tau = 0.;
    int rc = single_step(*ti, *state, t, tau_max, tau);
    if (rc == RJ_RESTART) {
      if (ti->recovery == RJ_CFL_RECOVERY_NONE)
        throw InvalidArgument("Restart raised with cfl recovery strategy none");
      hm.strategy = RJ_STRATEGY_WARN;
      hm.cfl = ti->cfl_min;
      rc = single_step(*ti, *state, t, tau_max, tau);
    }
    *tau_out = tau;
    return rc;
  });
}

double rj_time_integrator_efficiency(const rj_time_integrator *ti) { return ti->scheme.efficiency; }

int rj_time_integrator_set_stepping(rj_time_integrator *ti, int stepping)
{
  return guarded([&] {
    if (!ti || stepping < STEPPING_HOST || stepping > STEPPING_GRAPH)
      throw InvalidArgument("rj_time_integrator_set_stepping: invalid argument");
    ti->stepping = stepping;
    return RJ_OK;
  });
}

int rj_time_integrator_stepping(const rj_time_integrator *ti) { return ti->stepping; }

int rj_time_integrator_graph_stats(const rj_time_integrator *ti, uint64_t stats[2])
{
  return guarded([&] {
    if (!ti || !stats)
      throw InvalidArgument("null argument");
    stats[0] = ti->graph_captures;
    stats[1] = ti->graph_replays;
    return RJ_OK;
  });
}

/* ---- host-only entry points: mesh generator, layout decoding ---- */

struct rj_host_offline {
  rjb::HostOffline data;
};

int rj_mesh_generate(rj_host_offline **out, const rj_mesh_desc *mesh, int rank, int n_ranks)
{
  return guarded([&] {
    if (!out || !mesh)
      throw InvalidArgument("rj_mesh_generate: null argument");
    auto h = std::unique_ptr<rj_host_offline>(new rj_host_offline());
    generate_cartesian_q1(h->data, *mesh, rank, n_ranks);
    *out = h.release();
    return RJ_OK;
  });
}

void rj_host_offline_destroy(rj_host_offline *h) { delete h; }

int rj_host_offline_desc(const rj_host_offline *h, rj_offline_desc *desc)
{
  return guarded([&] {
    if (!h || !desc)
      throw InvalidArgument("null argument");
    h->data.fill_desc(*desc);
    return RJ_OK;
  });
}

int rj_host_offline_global_indices(const rj_host_offline *h, uint64_t *global)
{
  return guarded([&] {
    if (!h || !global)
      throw InvalidArgument("null argument");
    std::memcpy(global, h->data.global_index.data(), h->data.global_index.size() * sizeof(uint64_t));
    return RJ_OK;
  });
}

uint64_t rj_host_offline_n_global_dofs(const rj_host_offline *h) { return h->data.n_global_dofs; }

int rj_decode_simd_pattern(uint32_t n_rows,
                           uint32_t n_internal,
                           uint32_t simd_length,
                           const uint64_t *row_starts,
                           const uint32_t *column_indices,
                           uint64_t *csr_row_ptr,
                           uint32_t *csr_cols)
{
  return guarded([&] {
    std::vector<uint64_t> rp;
    std::vector<uint32_t> cl;
    decode_simd_pattern(n_rows, n_internal, simd_length, row_starts, column_indices, rp, cl);
    std::memcpy(csr_row_ptr, rp.data(), rp.size() * sizeof(uint64_t));
    std::memcpy(csr_cols, cl.data(), cl.size() * sizeof(uint32_t));
    return RJ_OK;
  });
}

} /* extern "C" */
