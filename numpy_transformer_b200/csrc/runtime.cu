// Runtime plumbing of libntb200.so: device/stream, pooled allocator, copies, events, graphs.
#include "common.cuh"
#include <cstdarg>
#include <cstdlib>
#include <map>
#include <vector>
#include <mutex>
#include <unordered_map>

namespace nt {
cudaStream_t g_stream = nullptr;
int g_sm_count = 148;
int g_gemm_mode = 1;   // default: tcgen05 3xTF32 (fp32-class accuracy)
long long g_launches = 0;
bool g_capturing = false;
bool g_use_pdl = true;
int* g_dev_error = nullptr;
static int g_device = -1;
static thread_local char g_err[1024] = "";
static char g_err_global[1024] = "";
static std::mutex g_mu;
static std::map<size_t, std::vector<void*>> g_free;      // size class -> free blocks
static std::unordered_map<void*, size_t> g_size;         // live + pooled blocks
static size_t g_reserved = 0;
static void* g_flush_buf = nullptr;
static cudaStream_t g_main_stream = nullptr, g_side_stream = nullptr;
static cudaEvent_t g_ev_fork = nullptr, g_ev_join = nullptr;
static bool g_side_active = false, g_side_dirty = false;
static bool g_side_enabled = true;
static const size_t kFlushBytes = 256ull << 20;
struct GraphRec { cudaGraphExec_t exec; cudaGraph_t graph; int kernels; };

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  strncpy(g_err_global, g_err, sizeof(g_err_global) - 1);
}
bool ensure_init() {
  if (g_stream) return true;
  set_error("ntb200: nt_init() has not been called (no CUDA device bound)");
  return false;
}
static size_t size_class(size_t b) {
  if (b < 512) return 512;
  if (b < (1u << 20)) return (b + 511) & ~size_t(511);
  return (b + (1u << 16) - 1) & ~size_t((1u << 16) - 1);
}
}  // namespace nt
using namespace nt;

extern "C" {

const char* nt_last_error(void) { return g_err_global; }

int nt_device_count(int* n) {
  cudaError_t e = cudaGetDeviceCount(n);
  if (e != cudaSuccess) { *n = 0; set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

int nt_init(int device) {
  if (g_stream) {
    NT_REQUIRE(device == g_device, "nt_init: already bound to device %d", g_device);
    return 0;
  }
  int n = 0;
  NT_CHECK(cudaGetDeviceCount(&n));
  NT_REQUIRE(n > 0, "nt_init: no CUDA device visible; this backend has no CPU fallback");
  NT_REQUIRE(device >= 0 && device < n, "nt_init: device %d out of range (%d visible)", device, n);
  NT_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  NT_CHECK(cudaGetDeviceProperties(&prop, device));
  NT_REQUIRE(prop.major == 10, "nt_init: device is sm_%d%d; this library is built for sm_100a only",
             prop.major, prop.minor);
  g_sm_count = prop.multiProcessorCount;
  if (const char* e = getenv("NTB200_PDL")) g_use_pdl = (e[0] != '0');
  NT_CHECK(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
  g_main_stream = g_stream;
  NT_CHECK(cudaStreamCreateWithFlags(&g_side_stream, cudaStreamNonBlocking));
  NT_CHECK(cudaEventCreateWithFlags(&g_ev_fork, cudaEventDisableTiming));
  NT_CHECK(cudaEventCreateWithFlags(&g_ev_join, cudaEventDisableTiming));
  NT_CHECK(cudaHostAlloc((void**)&g_dev_error, sizeof(int), cudaHostAllocMapped));
  *g_dev_error = 0;
  g_device = device;
  return 0;
}

// after a synchronisation: did a kernel flag a bad index since the last check?
static int check_dev_error(const char* where) {
  if (!g_dev_error || *(volatile int*)g_dev_error == 0) return 0;
  const int code = *(volatile int*)g_dev_error;
  *(volatile int*)g_dev_error = 0;
  set_error("%s: %s", where,
            code == NT_DEVERR_TOKEN ? "index out of range in an embedding lookup (token id >= num_embeddings; the reference raises "
                                      "IndexError, net/Embedding.py:46)"
                                    : "label out of range in cross_entropy (label >= number of classes)");
  return 4;
}

int nt_sm_count(int* n) { NT_ENTRY(); *n = g_sm_count; return 0; }

int nt_shutdown(void) {
  if (!g_stream) return 0;
  cudaStreamSynchronize(g_stream);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& kv : g_free) for (void* p : kv.second) cudaFree(p);
    g_free.clear();
    g_size.clear();
    g_reserved = 0;
  }
  if (g_flush_buf) { cudaFree(g_flush_buf); g_flush_buf = nullptr; }
  cudaStreamDestroy(g_stream);
  g_stream = nullptr;
  g_device = -1;
  return 0;
}

int nt_malloc(void** dptr, size_t bytes) {
  NT_ENTRY();
  size_t cls = size_class(bytes);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_free.find(cls);
    if (it != g_free.end() && !it->second.empty()) {
      *dptr = it->second.back();
      it->second.pop_back();
      return 0;
    }
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, cls);
  if (e != cudaSuccess) {
    set_error("nt_malloc(%zu): %s", bytes, cudaGetErrorString(e));
    return 1;
  }
  std::lock_guard<std::mutex> lk(g_mu);
  g_size[p] = cls;
  g_reserved += cls;
  *dptr = p;
  return 0;
}

int nt_free(void* dptr) {
  if (!dptr) return 0;
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_size.find(dptr);
  if (it == g_size.end()) { set_error("nt_free: unknown pointer %p", dptr); return 2; }
  // stream-ordered reuse: every consumer runs on g_stream, so the block can be handed out again
  g_free[it->second].push_back(dptr);
  return 0;
}

int nt_pool_bytes(size_t* reserved) { *reserved = g_reserved; return 0; }

int nt_host_alloc(void** hptr, size_t bytes) { NT_ENTRY(); NT_CHECK(cudaMallocHost(hptr, bytes)); return 0; }
int nt_host_free(void* hptr) { NT_CHECK(cudaFreeHost(hptr)); return 0; }

int nt_h2d(void* dst, const void* src, size_t bytes) {
  NT_ENTRY();
  NT_FUSE_FLUSH();
  NT_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_stream));
  return 0;
}
int nt_d2h_async(void* dst, const void* src, size_t bytes) {
  NT_ENTRY();
  NT_FUSE_FLUSH();
  NT_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_stream));
  return 0;
}
int nt_d2h(void* dst, const void* src, size_t bytes) {
  NT_ENTRY();
  NT_FUSE_FLUSH();
  NT_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_stream));
  NT_CHECK(cudaStreamSynchronize(g_stream));
  return check_dev_error("nt_d2h");
}
int nt_d2d(void* dst, const void* src, size_t bytes) {
  NT_ENTRY();
  NT_FUSE_FLUSH();
  NT_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, g_stream));
  return 0;
}
int nt_memset(void* dst, int value, size_t bytes) {
  NT_ENTRY();
  NT_FUSE_FLUSH();
  NT_CHECK(cudaMemsetAsync(dst, value, bytes, g_stream));
  return 0;
}
int nt_sync(void) { NT_ENTRY(); NT_FUSE_FLUSH(); NT_CHECK(cudaStreamSynchronize(g_stream)); return check_dev_error("nt_sync"); }
int nt_stream_handle(void** s) { NT_ENTRY(); *s = (void*)g_stream; return 0; }

int nt_event_create(void** ev) {
  NT_ENTRY();
  cudaEvent_t e;
  NT_CHECK(cudaEventCreate(&e));
  *ev = (void*)e;
  return 0;
}
int nt_event_record_external(void* ev) {
  // inside stream capture this becomes an event-record NODE whose timestamp can be read after replay
  NT_ENTRY();
  NT_FUSE_FLUSH();
  NT_CHECK(cudaEventRecordWithFlags((cudaEvent_t)ev, g_stream, cudaEventRecordExternal));
  return 0;
}
int nt_event_record(void* ev) { NT_ENTRY(); NT_FUSE_FLUSH(); NT_CHECK(cudaEventRecord((cudaEvent_t)ev, g_stream)); return 0; }
int nt_event_elapsed_ms(void* e0, void* e1, float* ms) {
  NT_CHECK(cudaEventSynchronize((cudaEvent_t)e1));
  NT_CHECK(cudaEventElapsedTime(ms, (cudaEvent_t)e0, (cudaEvent_t)e1));
  return 0;
}
int nt_event_synchronize(void* ev) { NT_CHECK(cudaEventSynchronize((cudaEvent_t)ev)); return 0; }
int nt_event_destroy(void* ev) { NT_CHECK(cudaEventDestroy((cudaEvent_t)ev)); return 0; }

int nt_graph_begin(void) {
  NT_ENTRY();
  NT_REQUIRE(!g_capturing, "nt_graph_begin: already capturing");
  NT_CHECK(cudaStreamBeginCapture(g_stream, cudaStreamCaptureModeRelaxed));
  g_capturing = true;
  return 0;
}
int nt_graph_end(void** graph_exec, int* n_kernel_nodes) {
  NT_ENTRY();
  NT_REQUIRE(g_capturing, "nt_graph_end: not capturing");
  if (nt::g_fusing) nt::fuse_flush();            // a capture that ends inside a fuse region still holds everything recorded
  // whatever happens below, the capture is over when this call returns: a failed join (or a kernel launch that
  // invalidated the capture) must not leave g_capturing set / the stream in capture mode, or every later
  // nt_graph_begin would fail
  const int join_rc = nt_side_join();
  g_capturing = false;
  g_stream = g_main_stream;
  g_side_active = g_side_dirty = false;
  cudaGraph_t graph = nullptr;
  const cudaError_t end_rc = cudaStreamEndCapture(g_main_stream, &graph);
  if (join_rc) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    return 1;                          // nt_side_join recorded the error text
  }
  NT_CHECK(end_rc);
  NT_REQUIRE(graph != nullptr, "nt_graph_end: the capture was invalidated");
  size_t n = 0;
  NT_CHECK(cudaGraphGetNodes(graph, nullptr, &n));
  std::vector<cudaGraphNode_t> nodes(n);
  if (n) NT_CHECK(cudaGraphGetNodes(graph, nodes.data(), &n));
  int kernels = 0;
  for (size_t i = 0; i < n; i++) {
    cudaGraphNodeType t;
    NT_CHECK(cudaGraphNodeGetType(nodes[i], &t));
    if (t == cudaGraphNodeTypeKernel) kernels++;
  }
  cudaGraphExec_t exec = nullptr;
  NT_CHECK(cudaGraphInstantiate(&exec, graph, 0));
  GraphRec* r = new GraphRec{exec, graph, kernels};
  *graph_exec = (void*)r;
  if (n_kernel_nodes) *n_kernel_nodes = kernels;
  return 0;
}
int nt_graph_launch(void* graph_exec) {
  NT_ENTRY();
  NT_FUSE_FLUSH();
  GraphRec* r = (GraphRec*)graph_exec;
  NT_CHECK(cudaGraphLaunch(r->exec, g_stream));
  g_launches += r->kernels;
  return 0;
}
int nt_graph_destroy(void* graph_exec) {
  GraphRec* r = (GraphRec*)graph_exec;
  if (!r) return 0;
  cudaGraphExecDestroy(r->exec);
  cudaGraphDestroy(r->graph);
  delete r;
  return 0;
}

// Side branch: while a graph is being captured, work enqueued between nt_side_begin/nt_side_end goes to a
// second stream forked from the current point of the main stream (the weight-gradient GEMMs, which nothing
// on the backward critical path consumes).  nt_side_join makes the main stream wait for all of it.  Outside
// capture the calls are no-ops: the pooled allocator's reuse rule assumes a single stream, and only a
// captured step pins every buffer for the lifetime of the graph.
int nt_set_side_branch(int enabled) { g_side_enabled = (enabled != 0); return 0; }
int nt_side_begin(void) {
  NT_ENTRY();
  if (nt::g_fusing) return 0;                     // a phase program keeps program order: no side stream while recording
  if (!g_capturing || !g_side_enabled || g_side_active) return 0;
  NT_CHECK(cudaEventRecord(g_ev_fork, g_main_stream));
  NT_CHECK(cudaStreamWaitEvent(g_side_stream, g_ev_fork, 0));
  g_stream = g_side_stream;
  g_side_active = true;
  g_side_dirty = true;
  return 0;
}
int nt_side_end(void) {
  NT_ENTRY();
  if (!g_side_active) return 0;
  g_stream = g_main_stream;
  g_side_active = false;
  return 0;
}
int nt_side_join(void) {
  NT_ENTRY();
  if (g_side_active) nt_side_end();
  if (!g_side_dirty) return 0;
  NT_CHECK(cudaEventRecord(g_ev_join, g_side_stream));
  NT_CHECK(cudaStreamWaitEvent(g_main_stream, g_ev_join, 0));
  g_side_dirty = false;
  return 0;
}

int nt_launch_count(long long* n) { *n = g_launches; return 0; }

int nt_set_gemm_mode(int mode) {
  NT_REQUIRE(mode >= 0 && mode <= 2, "nt_set_gemm_mode: mode %d not in {0,1,2}", mode);
  g_gemm_mode = mode;
  return 0;
}
int nt_get_gemm_mode(int* mode) { *mode = g_gemm_mode; return 0; }

int nt_flush_l2(void) {
  NT_ENTRY();
  NT_FUSE_FLUSH();
  if (!g_flush_buf) NT_CHECK(cudaMalloc(&g_flush_buf, kFlushBytes));
  NT_CHECK(cudaMemsetAsync(g_flush_buf, 0, kFlushBytes, g_stream));
  return 0;
}

}  // extern "C"
