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
}  // namespace nt
using namespace nt;

extern "C" {

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
