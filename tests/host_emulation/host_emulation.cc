// TEST INFRASTRUCTURE ONLY -- block execution engine of the host emulation (see cuda_runtime.h).
// Threads of a block are ucontext coroutines, scheduled round-robin on one OS thread.
#include "cuda_runtime.h"

#include <stdexcept>

rj_emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

namespace
{
  constexpr size_t stack_bytes = 256 * 1024;

  struct Thread {
    ucontext_t context;
    char *stack = nullptr;
    bool done = false;
  };

  struct Warp {
    unsigned int arrived = 0, done = 0, lanes = 0;
    unsigned long generation = 0;
    unsigned long long buffer[2][32];
  };

  ucontext_t scheduler_context;
  std::vector<Thread> threads;
  Warp warps[32];
  rj_emu_thunk *kernel_call = nullptr;
  unsigned int current = 0, block_threads = 0, block_done = 0, block_arrived = 0;
  unsigned long block_generation = 0;

  void yield() { swapcontext(&threads[current].context, &scheduler_context); }

  void entry()
  {
    kernel_call->run();
    threads[current].done = true;
    /* uc_link returns to the scheduler */
  }

  void run_block()
  {
    if (threads.size() < block_threads)
      threads.resize(block_threads);
    block_done = block_arrived = 0;
    for (unsigned int w = 0; w < (block_threads + 31) / 32; ++w) {
      warps[w].arrived = warps[w].done = 0;
      warps[w].lanes = std::min(32u, block_threads - 32 * w);
    }
    for (unsigned int t = 0; t < block_threads; ++t) {
      Thread &th = threads[t];
      if (!th.stack)
        th.stack = static_cast<char *>(std::malloc(stack_bytes));
      th.done = false;
      getcontext(&th.context);
      th.context.uc_stack.ss_sp = th.stack;
      th.context.uc_stack.ss_size = stack_bytes;
      th.context.uc_link = &scheduler_context;
      makecontext(&th.context, entry, 0);
    }
    std::vector<unsigned char> counted(block_threads, 0);
    while (block_done < block_threads) {
      for (unsigned int t = 0; t < block_threads; ++t) {
        if (counted[t])
          continue;
        current = t;
        threadIdx.x = t;
        swapcontext(&scheduler_context, &threads[t].context);
        if (threads[t].done) {
          counted[t] = 1;
          ++block_done;
          ++warps[t / 32].done;
        }
      }
    }
  }
} // namespace

void rj_emu_syncthreads()
{
  const unsigned long my_generation = block_generation;
  ++block_arrived;
  while (block_generation == my_generation) {
    if (block_arrived + block_done == block_threads) { /* exited threads count as arrived */
      block_arrived = 0;
      ++block_generation;
      break;
    }
    yield();
  }
}

unsigned long long rj_emu_shfl_xor(unsigned long long value, int lane_mask)
{
  Warp &w = warps[current / 32];
  const unsigned int lane = current % 32;
  const unsigned long my_generation = w.generation;
  const unsigned int parity = my_generation & 1u;
  w.buffer[parity][lane] = value;
  ++w.arrived;
  while (w.generation == my_generation) {
    if (w.arrived + w.done == w.lanes) {
      w.arrived = 0;
      ++w.generation;
      break;
    }
    yield();
  }
  const unsigned int source = lane ^ static_cast<unsigned int>(lane_mask);
  return source < w.lanes ? w.buffer[parity][source] : value;
}

void rj_emu_run_grid(unsigned int grid, unsigned int block, rj_emu_thunk &call)
{
  if (block == 0 || block > 1024)
    throw std::runtime_error("host emulation: invalid block size");
  kernel_call = &call;
  block_threads = block;
  blockDim.x = block;
  gridDim.x = grid;
  for (unsigned int b = 0; b < grid; ++b) {
    blockIdx.x = b;
    run_block();
  }
}
