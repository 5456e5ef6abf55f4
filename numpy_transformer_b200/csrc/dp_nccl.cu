// Data-parallel gradient all-reduce over NCCL (NVLink 5 / NVSwitch): one communicator per
// process, collectives on a dedicated comm stream ordered against the compute stream by events
// (capturable inside the step graph).  The reference has no distributed path; SURVEY.md §8(e).
#include "common.cuh"
#include <nccl.h>

namespace nt {
static ncclComm_t g_comm = nullptr;
static cudaStream_t g_comm_stream = nullptr;
static cudaEvent_t g_ev_ready = nullptr, g_ev_done = nullptr;
static bool g_pending = false;
#define NT_NCCL(expr)                                                                 \
  do {                                                                                \
    ncclResult_t _r = (expr);                                                         \
    if (_r != ncclSuccess) {                                                          \
      nt::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, ncclGetErrorString(_r)); \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

// ---------------------------------------------------------------- small-message all-reduce over NVLink peer memory
// The README ViT's gradient arena is 3.4 MB: an NCCL ring / tree call on it is latency-bound (~0.1-0.25 ms, a third of the
// 0.5 ms step, and the part that follows the fused backward cannot overlap anything).  For such sizes every rank maps every
// peer's gradient arena (CUDA IPC) and ONE kernel does a two-shot reduction in place: rank r sums slice r of all the arenas
// through NVLink loads and stores the sum back into slice r of all of them.  2 * 7/8 * n * 4 bytes cross NVLink per GPU
// (2.8 MB for 1.6 MB of gradients: a few microseconds at 700 GB/s); two flag barriers in peer memory bracket it.
constexpr int P2P_MAX = 8;
struct P2PChan {            // flags of one channel (= one stream); lives in every rank's IPC-shared flag page
  unsigned start[P2P_MAX];  // start[p] = epoch: rank p's gradients of this call are final
  unsigned end[P2P_MAX];    // end[p]   = epoch: rank p has stored its slice into MY arena
  unsigned epoch;           // local: last completed call
  unsigned done;            // local: CTAs of the running call that have finished
  unsigned pad[14];
};
struct P2PArgs {
  float* buf[P2P_MAX];      // every rank's gradient arena (own one included), as mapped into THIS process
  P2PChan* flags[P2P_MAX];  // every rank's flag page
  int rank, world;
};
static P2PArgs g_p2p{};
static bool g_p2p_ready = false;
static void* g_p2p_flags_local = nullptr;
static void* g_p2p_opened[2 * P2P_MAX] = {};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// bounded: a rank that never arrives (crashed peer, mismatched call sequence) traps this kernel instead of hanging the GPU
__device__ __forceinline__ void spin_until(const unsigned* p, unsigned epoch) {
  for (unsigned it = 0; it < (1u << 27); ++it) {
    if ((int)(ld_acquire_sys(p) - epoch) >= 0) return;
    if (it > 64) __nanosleep(32);
  }
  __trap();
}

__global__ void __launch_bounds__(512) p2p_allreduce_kernel(P2PArgs a, int chan, size_t off, size_t n) {
  pdl_prologue();
  P2PChan* mine = a.flags[a.rank] + chan;
  const unsigned epoch = mine->epoch + 1;          // every CTA reads it before the last one to finish advances it
  const int tid = threadIdx.x;
  if (blockIdx.x == 0 && tid < a.world) {          // stream order: this rank's gradients are final when the kernel starts
    __threadfence_system();
    st_release_sys(&(a.flags[tid] + chan)->start[a.rank], epoch);
  }
  if (tid < a.world) spin_until(&mine->start[tid], epoch);
  __syncthreads();
  // slice of this rank, in float4 units
  const size_t n4 = n / 4, per = (n4 + a.world - 1) / a.world;
  const size_t lo = (size_t)a.rank * per, hi = lo + per < n4 ? lo + per : n4;
  for (size_t i = lo + (size_t)blockIdx.x * blockDim.x + tid; i < hi; i += (size_t)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p = 0; p < P2P_MAX; p++)
      if (p < a.world) {
        const float4 v = reinterpret_cast<const float4*>(a.buf[p] + off)[i];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
#pragma unroll
    for (int p = 0; p < P2P_MAX; p++)
      if (p < a.world) reinterpret_cast<float4*>(a.buf[p] + off)[i] = acc;
  }
  __threadfence_system();                          // this thread's peer stores are out
  __shared__ unsigned s_last;
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&mine->done, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (s_last) {                                    // last CTA of this rank to finish: every store of the rank is out
    // one thread per peer: signal and wait in parallel (a loop of release stores by one thread costs a fence round trip each)
    if (tid < a.world) {
      __threadfence_system();
      st_release_sys(&(a.flags[tid] + chan)->end[a.rank], epoch);
      spin_until(&mine->end[tid], epoch);          // rank `tid` has stored its slice into MY arena
    }
    __syncthreads();
    if (tid == 0) {
      mine->done = 0;
      mine->epoch = epoch;
      __threadfence();
    }
  }
}
}  // namespace nt
using namespace nt;

extern "C" {

int nt_p2p_flags_alloc(void** dptr) {
  NT_ENTRY();
  NT_CHECK(cudaMalloc(dptr, 4096));                // a whole allocation of its own: CUDA IPC exports allocations
  NT_CHECK(cudaMemset(*dptr, 0, 4096));
  g_p2p_flags_local = *dptr;
  return 0;
}
int nt_p2p_ipc_handle(void* dptr, void* handle64) {
  NT_ENTRY();
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
  cudaIpcMemHandle_t h;
  NT_CHECK(cudaIpcGetMemHandle(&h, dptr));
  memcpy(handle64, &h, sizeof(h));
  return 0;
}
/* grad_handles / flag_handles: world x 64 bytes, rank-ordered (this rank's own entries are ignored). */
int nt_p2p_setup(int rank, int world, float* my_grads, const void* grad_handles, void* my_flags, const void* flag_handles) {
  NT_ENTRY();
  NT_REQUIRE(world >= 2 && world <= P2P_MAX && rank >= 0 && rank < world, "nt_p2p_setup: bad rank / world %d / %d", rank, world);
  NT_REQUIRE(!g_p2p_ready, "nt_p2p_setup: already set up");
  g_p2p = P2PArgs{};
  g_p2p.rank = rank; g_p2p.world = world;
  for (int p = 0; p < world; p++) {
    if (p == rank) { g_p2p.buf[p] = my_grads; g_p2p.flags[p] = (P2PChan*)my_flags; continue; }
    cudaIpcMemHandle_t hg, hf;
    memcpy(&hg, (const char*)grad_handles + 64 * p, 64);
    memcpy(&hf, (const char*)flag_handles + 64 * p, 64);
    void *pg = nullptr, *pf = nullptr;
    NT_CHECK(cudaIpcOpenMemHandle(&pg, hg, cudaIpcMemLazyEnablePeerAccess));
    NT_CHECK(cudaIpcOpenMemHandle(&pf, hf, cudaIpcMemLazyEnablePeerAccess));
    g_p2p.buf[p] = (float*)pg; g_p2p.flags[p] = (P2PChan*)pf;
    g_p2p_opened[2 * p] = pg; g_p2p_opened[2 * p + 1] = pf;
  }
  g_p2p_ready = true;
  return 0;
}
/* In-place sum over all ranks of grads[off, off + n) (float units; off and n multiples of 4).  channel 0: on the compute
 * stream; channel 1: on the comm stream after everything enqueued on the compute stream so far (joined by nt_dp_join). */
int nt_p2p_allreduce_sum(size_t off, size_t n, int channel) {
  NT_ENTRY();
  NT_REQUIRE(g_p2p_ready, "nt_p2p_allreduce_sum: nt_p2p_setup has not been called");
  NT_FUSE_FLUSH();
  NT_REQUIRE(off % 4 == 0 && n % 4 == 0 && (channel == 0 || channel == 1), "nt_p2p_allreduce_sum: off / n must be multiples of 4");
  if (n == 0) return 0;
  const size_t per4 = (n / 4 + g_p2p.world - 1) / g_p2p.world;
  static int threads = 0, use_pdl = -1;
  // 256 threads per CTA spread the NVLink loads / stores of a 1.7 MB exchange over twice as many SMs as 512 (8 GPUs, README ViT
  // step: 0.5346 -> 0.5321 ms; 128 threads: the same)
  if (!threads) { const char* e = getenv("NTB200_P2P_THREADS"); threads = e ? atoi(e) : 256; if (threads < 32 || threads > 512) threads = 256; }
  if (use_pdl < 0) { const char* e = getenv("NTB200_P2P_PDL"); use_pdl = (e && e[0] == '0') ? 0 : 1; }
  int grid = (int)((per4 + threads - 1) / threads);
  grid = grid < 1 ? 1 : (grid > 128 ? 128 : grid);
  cudaStream_t st = g_stream;
  if (channel == 1) {
    NT_REQUIRE(g_comm_stream, "nt_p2p_allreduce_sum: channel 1 needs nt_dp_init (comm stream)");
    NT_CHECK(cudaEventRecord(g_ev_ready, g_stream));
    NT_CHECK(cudaStreamWaitEvent(g_comm_stream, g_ev_ready, 0));
    st = g_comm_stream;
    g_pending = true;
  }
  if (channel == 0 && use_pdl) {
    // on the compute stream the kernel may be scheduled under the tail of its predecessor (it starts with griddepcontrol.wait)
    NT_CHECK(launch_k(p2p_allreduce_kernel, dim3(grid), dim3(threads), 0, g_p2p, channel, off, n));
  } else {
    p2p_allreduce_kernel<<<grid, threads, 0, st>>>(g_p2p, channel, off, n);
  }
  NT_LAUNCH_OK();
  return 0;
}
int nt_p2p_shutdown(void) {
  if (!g_p2p_ready) return 0;
  if (g_stream) cudaStreamSynchronize(g_stream);
  if (g_comm_stream) cudaStreamSynchronize(g_comm_stream);
  for (int i = 0; i < 2 * P2P_MAX; i++)
    if (g_p2p_opened[i]) { cudaIpcCloseMemHandle(g_p2p_opened[i]); g_p2p_opened[i] = nullptr; }
  if (g_p2p_flags_local) { cudaFree(g_p2p_flags_local); g_p2p_flags_local = nullptr; }
  g_p2p_ready = false;
  return 0;
}

int nt_dp_unique_id(void* uid128) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  NT_NCCL(ncclGetUniqueId(&id));
  memcpy(uid128, &id, sizeof(id));
  return 0;
}

int nt_dp_init(const void* uid128, int rank, int world) {
  NT_ENTRY();
  NT_REQUIRE(!g_comm, "nt_dp_init: communicator already initialised");
  ncclUniqueId id;
  memcpy(&id, uid128, sizeof(id));
  NT_NCCL(ncclCommInitRank(&g_comm, world, id, rank));
  NT_CHECK(cudaStreamCreateWithFlags(&g_comm_stream, cudaStreamNonBlocking));
  NT_CHECK(cudaEventCreateWithFlags(&g_ev_ready, cudaEventDisableTiming));
  NT_CHECK(cudaEventCreateWithFlags(&g_ev_done, cudaEventDisableTiming));
  return 0;
}

int nt_dp_allreduce_sum(float* buf, size_t n) {
  NT_ENTRY();
  NT_REQUIRE(g_comm, "nt_dp_allreduce_sum: nt_dp_init has not been called");
  NT_FUSE_FLUSH();
  // comm stream waits for everything enqueued on the compute stream so far (the bucket's grads)
  NT_CHECK(cudaEventRecord(g_ev_ready, g_stream));
  NT_CHECK(cudaStreamWaitEvent(g_comm_stream, g_ev_ready, 0));
  NT_NCCL(ncclAllReduce(buf, buf, n, ncclFloat, ncclSum, g_comm, g_comm_stream));
  g_pending = true;
  return 0;
}

int nt_dp_broadcast(void* buf, size_t n, int root) {
  NT_ENTRY();
  NT_REQUIRE(g_comm, "nt_dp_broadcast: nt_dp_init has not been called");
  NT_FUSE_FLUSH();
  NT_NCCL(ncclBroadcast(buf, buf, n, ncclFloat, root, g_comm, g_stream));
  return 0;
}

int nt_dp_join(void) {
  NT_ENTRY();
  if (!g_pending) return 0;
  NT_CHECK(cudaEventRecord(g_ev_done, g_comm_stream));
  NT_CHECK(cudaStreamWaitEvent(g_stream, g_ev_done, 0));
  g_pending = false;
  return 0;
}

int nt_dp_shutdown(void) {
  if (g_comm) {
    // all collectives have completed once both streams drain; abort tears the communicator down
    // without the cross-rank handshake of ncclCommDestroy (which blocks while captured graphs that
    // contain NCCL kernels are still alive in another rank)
    if (g_stream) cudaStreamSynchronize(g_stream);
    if (g_comm_stream) cudaStreamSynchronize(g_comm_stream);
    ncclCommAbort(g_comm);
    g_comm = nullptr;
  }
  if (g_comm_stream) { cudaStreamDestroy(g_comm_stream); g_comm_stream = nullptr; }
  if (g_ev_ready) { cudaEventDestroy(g_ev_ready); g_ev_ready = nullptr; }
  if (g_ev_done) { cudaEventDestroy(g_ev_done); g_ev_done = nullptr; }
  return 0;
}

}  // extern "C"
