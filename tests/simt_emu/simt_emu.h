// A minimal CPU SIMT emulator — TEST INFRASTRUCTURE, not product code.
//
// It runs the device code of numpy_transformer_b200/csrc/phase_kernel.cuh unchanged on the host so that the phase-program
// kernel (one thread-block cluster, CTA barriers, warp shuffles, cluster barriers) can be checked against the oracle in
// the `-m "not gpu"` suite.  Every CUDA thread is a fibre (ucontext) with its own stack; barriers and shuffles yield to a
// round-robin scheduler; the scheduling order can be reversed, which exposes most missing barriers (a consumer that runs
// before its producer reads stale data in one of the two orders).  Shared memory is one buffer per CTA; "global memory"
// is ordinary host memory (NumPy arrays in the tests).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <ucontext.h>
#include <sys/mman.h>

#define NT_EMU 1
#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__

struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }

namespace emu {

struct Barrier {
  int count = 0, expected = 0;
  uint64_t gen = 0;
};

struct Fiber {
  ucontext_t ctx;
  void* stack = nullptr;
  int tid = 0, cta = 0;
  bool done = false;
  Barrier* wait_bar = nullptr;
  uint64_t wait_gen = 0;
};

struct Cta {
  Barrier bar;
  std::vector<Barrier> warp_bar;
  std::vector<uint32_t> warp_buf;      // [nwarps][32]
  std::vector<unsigned char> smem;
};

struct Launch {
  int nctas = 0, nthreads = 0;
  std::vector<Fiber> fibers;
  std::vector<Cta> ctas;
  Barrier cluster_bar;
  ucontext_t sched;
  Fiber* cur = nullptr;
  void (*entry)(const void*) = nullptr;
  const void* param = nullptr;
  long switches = 0;
};

extern Launch* g_launch;

static inline Fiber* self() { return g_launch->cur; }
static inline void yield_to_scheduler() {
  Fiber* f = g_launch->cur;
  swapcontext(&f->ctx, &g_launch->sched);
}
static inline void barrier_wait(Barrier& b) {
  Fiber* f = self();
  const uint64_t g = b.gen;
  if (++b.count == b.expected) {
    b.count = 0;
    b.gen++;
    return;
  }
  f->wait_bar = &b;
  f->wait_gen = g;
  yield_to_scheduler();
}

static const size_t kStack = 256 * 1024;

static void trampoline() {
  Launch* L = g_launch;
  Fiber* f = L->cur;
  L->entry(L->param);
  f->done = true;
  swapcontext(&f->ctx, &L->sched);
}

// run entry(param) as nctas x nthreads fibres; reverse != 0 walks the fibres from the last to the first
static inline int run(void (*entry)(const void*), const void* param, int nctas, int nthreads, size_t smem_bytes, int reverse) {
  Launch L;
  g_launch = &L;
  L.nctas = nctas;
  L.nthreads = nthreads;
  L.entry = entry;
  L.param = param;
  L.ctas.resize(nctas);
  const int nwarps = (nthreads + 31) / 32;
  for (auto& c : L.ctas) {
    c.bar.expected = nthreads;
    c.warp_bar.resize(nwarps);
    for (int w = 0; w < nwarps; w++) c.warp_bar[w].expected = (w == nwarps - 1 && nthreads % 32) ? nthreads % 32 : 32;
    c.warp_buf.assign((size_t)nwarps * 32, 0u);
    c.smem.assign(smem_bytes + 64, 0);
  }
  L.cluster_bar.expected = nctas * nthreads;
  const int n = nctas * nthreads;
  L.fibers.resize(n);
  for (int i = 0; i < n; i++) {
    Fiber& f = L.fibers[i];
    f.cta = i / nthreads;
    f.tid = i % nthreads;
    f.stack = mmap(nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (f.stack == MAP_FAILED) { fprintf(stderr, "simt_emu: mmap failed\n"); return 1; }
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = kStack;
    f.ctx.uc_link = nullptr;
    makecontext(&f.ctx, trampoline, 0);
  }
  int rc = 0;
  for (;;) {
    int alive = 0, progressed = 0;
    for (int s = 0; s < n; s++) {
      Fiber& f = L.fibers[reverse ? n - 1 - s : s];
      if (f.done) continue;
      alive++;
      if (f.wait_bar) {
        if (f.wait_bar->gen == f.wait_gen) continue;       // still blocked
        f.wait_bar = nullptr;
      }
      L.cur = &f;
      L.switches++;
      swapcontext(&L.sched, &f.ctx);
      progressed++;
    }
    if (!alive) break;
    if (!progressed) { fprintf(stderr, "simt_emu: deadlock (%d fibres blocked on barriers not everybody reaches)\n", alive); rc = 2; break; }
  }
  for (auto& f : L.fibers) munmap(f.stack, kStack);
  g_launch = nullptr;
  return rc;
}

}  // namespace emu

// ---- what the device code sees (the CUDA build defines the same names over the real intrinsics)
static inline int ph_tid() { return emu::self()->tid; }
static inline int ph_rank() { return emu::self()->cta; }
static inline int ph_nrank() { return emu::g_launch->nctas; }
static inline void ph_syncthreads() { emu::barrier_wait(emu::g_launch->ctas[emu::self()->cta].bar); }
static inline void ph_cluster_sync() { emu::barrier_wait(emu::g_launch->cluster_bar); }
static inline float* ph_smem() {
  unsigned char* p = emu::g_launch->ctas[emu::self()->cta].smem.data();
  return reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(p) + 15) & ~uintptr_t(15));
}
static inline void ph_prologue() {}
static inline uint32_t ph_shfl_xor_bits(uint32_t v, int off) {
  emu::Fiber* f = emu::self();
  emu::Cta& c = emu::g_launch->ctas[f->cta];
  const int warp = f->tid >> 5, lane = f->tid & 31;
  c.warp_buf[warp * 32 + lane] = v;
  emu::barrier_wait(c.warp_bar[warp]);
  const uint32_t r = c.warp_buf[warp * 32 + (lane ^ off)];
  emu::barrier_wait(c.warp_bar[warp]);
  return r;
}
static inline float ph_shfl_xor(float v, int off) {
  uint32_t b;
  memcpy(&b, &v, 4);
  b = ph_shfl_xor_bits(b, off);
  memcpy(&v, &b, 4);
  return v;
}
static inline int ph_shfl_xor(int v, int off) { return (int)ph_shfl_xor_bits((uint32_t)v, off); }
