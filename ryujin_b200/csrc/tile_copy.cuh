// Bulk tile copies between global and shared memory for the slice-tiled sweeps: the TMA engine of
// sm_90+/sm_100a in its 1D form (PTX cp.async.bulk, SASS UBLKCP) completing on an mbarrier.
//
// Why 1D bulk copies fit this path: in the warp-sliced ELL layout (device_offline.cu) ALL entries
// of a slice -- 32 rows x width columns -- of one matrix component are one contiguous, 256-byte
// aligned run of width * 256 bytes (c_ij, m_ij, d_ij, l_ij, P_ij, the column indices).  A block
// that owns a slice therefore stages each matrix component it needs with ONE bulk copy issued by
// one thread: no per-thread address arithmetic, no registers holding data in flight, and the
// copy engine keeps a whole tile (6.9 kB per component for Q1 in 3D) in flight per instruction,
// which is what the latency-bound limiter sweeps need.  Neighbour state
// windows (U_j of the slice's columns) are contiguous runs whenever the DoF ordering keeps the
// neighbours of consecutive rows consecutive; device_offline.cu detects the runs per slice.
//
// The reference's counterpart is the vectorised row gather / SIMD transpose of
// SparseMatrixSIMD::get_tensor (source/sparse_matrix_simd.h:389-448) and
// MultiComponentVector::get_tensor(js) (source/multicomponent_vector.h:291-324).
//
// tests/host_emulation compiles this header with g++: there a bulk copy is a memcpy at issue time
// and the barrier a counter the coroutine scheduler spins on, with the same phase semantics.
#pragma once

#include <cstdint>
#include <cstring>

namespace rjb
{
  namespace tile
  {
#if defined(RJ_HOST_EMULATION)
    void emu_yield(); /* host_emulation.cc */

    struct alignas(8) Barrier {
      uint32_t phase;
      int32_t pending_arrivals;
      int64_t pending_tx;
      int32_t count;
    };

    inline void init(Barrier *bar, int count)
    {
      bar->phase = 0;
      bar->count = count;
      bar->pending_arrivals = count;
      bar->pending_tx = 0;
    }

    inline void emu_complete(Barrier *bar)
    {
      if (bar->pending_arrivals == 0 && bar->pending_tx == 0) {
        bar->phase ^= 1u;
        bar->pending_arrivals = bar->count;
      }
    }

    inline void fence_init() {}

    /* arrive and announce `bytes` of copies that will complete on this barrier */
    inline void arrive_expect(Barrier *bar, uint32_t bytes)
    {
      bar->pending_tx += bytes;
      --bar->pending_arrivals;
      emu_complete(bar);
    }

    inline void load(void *dst_shared, const void *src_global, uint32_t bytes, Barrier *bar)
    {
      std::memcpy(dst_shared, src_global, bytes);
      bar->pending_tx -= bytes;
      emu_complete(bar);
    }

    inline void wait(Barrier *bar, uint32_t parity)
    {
      while (bar->phase == parity)
        emu_yield();
    }

    inline void fence_async() {}
    inline void store(void *dst_global, const void *src_shared, uint32_t bytes) { std::memcpy(dst_global, src_shared, bytes); }
    inline void store_commit() {}
    inline void store_wait_read() {}
    inline void store_wait_all() {}
#else
    struct alignas(8) Barrier {
      unsigned long long state;
    };

    __device__ __forceinline__ uint32_t shared_address(const void *p)
    {
      return static_cast<uint32_t>(__cvta_generic_to_shared(p));
    }

    __device__ __forceinline__ void init(Barrier *bar, int count)
    {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(shared_address(bar)), "r"(count) : "memory");
    }

    /* makes the initialised barrier visible to the async proxy (the copy engine) */
    __device__ __forceinline__ void fence_init()
    {
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    __device__ __forceinline__ void arrive_expect(Barrier *bar, uint32_t bytes)
    {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(shared_address(bar)), "r"(bytes)
                   : "memory");
    }

    /* global -> shared, 16-byte aligned on both sides, bytes % 16 == 0; completes on `bar` */
    __device__ __forceinline__ void load(void *dst_shared, const void *src_global, uint32_t bytes, Barrier *bar)
    {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       shared_address(dst_shared)),
                   "l"(src_global),
                   "r"(bytes),
                   "r"(shared_address(bar))
                   : "memory");
    }

    __device__ __forceinline__ void wait(Barrier *bar, uint32_t parity)
    {
      const uint32_t address = shared_address(bar);
      uint32_t done;
      do {
        asm volatile("{\n"
                     ".reg .pred p;\n"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                     "selp.u32 %0, 1, 0, p;\n"
                     "}\n"
                     : "=r"(done)
                     : "r"(address), "r"(parity)
                     : "memory");
      } while (!done);
    }

    /* orders generic-proxy writes to shared memory before a bulk store reads them */
    __device__ __forceinline__ void fence_async()
    {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }

    /* shared -> global; completion tracked by the issuing thread's bulk groups */
    __device__ __forceinline__ void store(void *dst_global, const void *src_shared, uint32_t bytes)
    {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_global),
                   "r"(shared_address(src_shared)),
                   "r"(bytes)
                   : "memory");
    }

    __device__ __forceinline__ void store_commit()
    {
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }

    /* the shared memory source of all committed stores may be overwritten / released */
    __device__ __forceinline__ void store_wait_read()
    {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }

    __device__ __forceinline__ void store_wait_all()
    {
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
#endif
  } // namespace tile
} // namespace rjb
