// TEST INFRASTRUCTURE ONLY -- block execution engine of the host emulation (see cuda_runtime.h).
// Threads of a block are ucontext coroutines, scheduled round-robin on one OS thread.
#include "cuda_runtime.h"

#include <algorithm>
#include <stdexcept>
#include <string>

thread_local rj_emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
thread_local rj_emu_graph *rj_emu_capture = nullptr;

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

  thread_local ucontext_t scheduler_context;
  thread_local std::vector<Thread> threads;
  thread_local Warp warps[32];
  thread_local rj_emu_thunk *kernel_call = nullptr;
  thread_local unsigned int current = 0, block_threads = 0, block_done = 0, block_arrived = 0;
  thread_local unsigned long block_generation = 0;
  thread_local bool reverse_threads = false;
  thread_local std::vector<unsigned char> dynamic_smem; /* of the running block */

  void yield() { swapcontext(&threads[current].context, &scheduler_context); }

  void entry()
  {
    kernel_call->run();
    threads[current].done = true;
    /* uc_link returns to the scheduler */
  }

  void run_block()
  {
    /* uninitialised shared memory is poisoned like fresh device memory */
    std::fill(dynamic_smem.begin(), dynamic_smem.end(), (unsigned char)0xff);
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
      for (unsigned int n = 0; n < block_threads; ++n) {
        const unsigned int t = reverse_threads ? block_threads - 1 - n : n;
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

unsigned char *rj_emu_dynamic_smem()
{
  /* 128-byte aligned like the device's dynamic shared memory window */
  unsigned char *p = dynamic_smem.data();
  return p + (128 - reinterpret_cast<uintptr_t>(p) % 128) % 128;
}

namespace rjb
{
  namespace tile
  {
    void emu_yield() { yield(); }
  } // namespace tile
} // namespace rjb

void rj_emu_run_grid(unsigned int grid, unsigned int block, rj_emu_thunk &call, size_t smem_bytes)
{
  dynamic_smem.assign(smem_bytes + 128, (unsigned char)0xff);
  if (block == 0 || block > 1024)
    throw std::runtime_error("host emulation: invalid block size");
  kernel_call = &call;
  block_threads = block;
  blockDim.x = block;
  gridDim.x = grid;
  /* Order-dependence check (a poor man's race detector): RJ_EMU_ORDER=reverse runs the blocks of a
   * launch from last to first and starts the threads of a block from the highest index;
   * RJ_EMU_ORDER=shuffle visits the blocks in a pseudo-random order.  A kernel in which one thread
   * consumes what another thread of the same launch produces gives different results then. */
  static const std::string order = std::getenv("RJ_EMU_ORDER") ? std::getenv("RJ_EMU_ORDER") : "";
  reverse_threads = (order == "reverse");
  if (order == "shuffle") {
    std::vector<unsigned int> perm(grid);
    for (unsigned int b = 0; b < grid; ++b)
      perm[b] = b;
    unsigned long long state = 0x9e3779b97f4a7c15ull ^ grid;
    for (unsigned int b = grid; b > 1; --b) {
      state = state * 6364136223846793005ull + 1442695040888963407ull;
      std::swap(perm[b - 1], perm[(state >> 33) % b]);
    }
    for (unsigned int b : perm) {
      blockIdx.x = b;
      run_block();
    }
    return;
  }
  for (unsigned int n = 0; n < grid; ++n) {
    blockIdx.x = reverse_threads ? grid - 1 - n : n;
    run_block();
  }
}


/* ------------------------------------------------------------------------------------------ *
 * In-process NCCL emulation (nccl.h): ranks are host threads.
 * ------------------------------------------------------------------------------------------ */
#include "nccl.h"

#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>

namespace
{
  struct World {
    std::mutex mutex;
    std::condition_variable changed;
    int n_ranks = 0;
    std::map<std::pair<int, int>, std::deque<std::vector<char>>> mailbox; /* (source, destination) */
    /* all-reduce rendezvous */
    int arrived = 0;
    unsigned long generation = 0;
    std::vector<char> accumulator, result;
  };

  std::mutex registry_mutex;
  std::map<std::string, std::shared_ptr<World>> registry;
  unsigned long next_id = 1;

  size_t type_size(ncclDataType_t t) { return t == ncclDouble ? 8 : 4; }

  template <typename T> void reduce_into(T *acc, const T *in, size_t n, ncclRedOp_t op)
  {
    for (size_t q = 0; q < n; ++q)
      acc[q] = op == ncclMin ? std::min(acc[q], in[q]) : std::max(acc[q], in[q]);
  }
} // namespace

struct rj_emu_comm {
  std::shared_ptr<World> world;
  int rank = 0;
};

const char *ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : "invalid use of the NCCL emulation"; }

ncclResult_t ncclGetUniqueId(ncclUniqueId *id)
{
  std::lock_guard<std::mutex> lock(registry_mutex);
  std::memset(id->internal, 0, sizeof(id->internal));
  const std::string text = "rj-emu-" + std::to_string(next_id++);
  std::memcpy(id->internal, text.data(), text.size());
  return ncclSuccess;
}

ncclResult_t ncclCommInitRank(ncclComm_t *comm, int n_ranks, ncclUniqueId id, int rank)
{
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks)
    return ncclInvalidUsage;
  const std::string key(id.internal, strnlen(id.internal, sizeof(id.internal)));
  std::lock_guard<std::mutex> lock(registry_mutex);
  auto &w = registry[key];
  if (!w) {
    w = std::make_shared<World>();
    w->n_ranks = n_ranks;
  }
  if (w->n_ranks != n_ranks)
    return ncclInvalidUsage;
  *comm = new rj_emu_comm{w, rank};
  return ncclSuccess;
}

ncclResult_t ncclCommDestroy(ncclComm_t comm)
{
  delete comm;
  return ncclSuccess;
}

ncclResult_t ncclGroupStart() { return ncclSuccess; }
ncclResult_t ncclGroupEnd() { return ncclSuccess; }

namespace
{
  ncclResult_t send_now(const void *buffer, size_t count, ncclDataType_t type, int peer, ncclComm_t comm);
  ncclResult_t recv_now(void *buffer, size_t count, ncclDataType_t type, int peer, ncclComm_t comm);
  ncclResult_t allreduce_now(const void *send, void *recv, size_t count, ncclDataType_t type, ncclRedOp_t op, ncclComm_t comm);
  void must_succeed(ncclResult_t r)
  {
    if (r != ncclSuccess)
      throw std::runtime_error("host emulation: NCCL call failed while a captured graph was replayed");
  }
} // namespace

/* inside a stream capture the calls are recorded and run when the graph is launched */
ncclResult_t ncclSend(const void *buffer, size_t count, ncclDataType_t type, int peer, ncclComm_t comm, cudaStream_t)
{
  if (rj_emu_capture) {
    rj_emu_capture->nodes.emplace_back([=]() { must_succeed(send_now(buffer, count, type, peer, comm)); });
    return ncclSuccess;
  }
  return send_now(buffer, count, type, peer, comm);
}
ncclResult_t ncclRecv(void *buffer, size_t count, ncclDataType_t type, int peer, ncclComm_t comm, cudaStream_t)
{
  if (rj_emu_capture) {
    rj_emu_capture->nodes.emplace_back([=]() { must_succeed(recv_now(buffer, count, type, peer, comm)); });
    return ncclSuccess;
  }
  return recv_now(buffer, count, type, peer, comm);
}
ncclResult_t ncclAllReduce(const void *send, void *recv, size_t count, ncclDataType_t type, ncclRedOp_t op, ncclComm_t comm, cudaStream_t)
{
  if (rj_emu_capture) {
    rj_emu_capture->nodes.emplace_back([=]() { must_succeed(allreduce_now(send, recv, count, type, op, comm)); });
    return ncclSuccess;
  }
  return allreduce_now(send, recv, count, type, op, comm);
}

namespace
{
ncclResult_t send_now(const void *buffer, size_t count, ncclDataType_t type, int peer, ncclComm_t comm)
{
  World &w = *comm->world;
  std::vector<char> message(count * type_size(type));
  std::memcpy(message.data(), buffer, message.size());
  {
    std::lock_guard<std::mutex> lock(w.mutex);
    w.mailbox[{comm->rank, peer}].push_back(std::move(message));
  }
  w.changed.notify_all();
  return ncclSuccess;
}

ncclResult_t recv_now(void *buffer, size_t count, ncclDataType_t type, int peer, ncclComm_t comm)
{
  World &w = *comm->world;
  std::unique_lock<std::mutex> lock(w.mutex);
  auto &queue = w.mailbox[{peer, comm->rank}];
  w.changed.wait(lock, [&] { return !queue.empty(); });
  std::vector<char> message = std::move(queue.front());
  queue.pop_front();
  if (message.size() != count * type_size(type))
    return ncclInvalidUsage; /* sender and receiver disagree on the message size */
  std::memcpy(buffer, message.data(), message.size());
  return ncclSuccess;
}

ncclResult_t allreduce_now(const void *send, void *recv, size_t count, ncclDataType_t type, ncclRedOp_t op, ncclComm_t comm)
{
  World &w = *comm->world;
  const size_t bytes = count * type_size(type);
  std::unique_lock<std::mutex> lock(w.mutex);
  const unsigned long my_generation = w.generation;
  if (w.arrived == 0) {
    w.accumulator.assign(static_cast<const char *>(send), static_cast<const char *>(send) + bytes);
  } else {
    if (w.accumulator.size() != bytes)
      return ncclInvalidUsage;
    if (type == ncclDouble)
      reduce_into(reinterpret_cast<double *>(w.accumulator.data()), static_cast<const double *>(send), count, op);
    else
      reduce_into(reinterpret_cast<int *>(w.accumulator.data()), static_cast<const int *>(send), count, op);
  }
  if (++w.arrived == w.n_ranks) {
    w.result = w.accumulator;
    w.arrived = 0;
    ++w.generation;
    w.changed.notify_all();
  } else
    w.changed.wait(lock, [&] { return w.generation != my_generation; });
  std::memcpy(recv, w.result.data(), bytes);
  return ncclSuccess;
}
} // namespace
