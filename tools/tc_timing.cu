// Development aid: phase timestamps of gemm_tc_kernel (globaltimer) for a few shapes.
// build: nvcc -DNT_TC_TIMING ... (see tools/build_tc_timing.sh)
#include "../numpy_transformer_b200/csrc/gemm_tc.cu"
#include <vector>
#include <algorithm>
using namespace nt;

static void run(const char* name, int M, int K, int N, int which, int passes, int act, bool extras) {
  float *x, *w, *y, *b, *pre, *res;
  size_t big = (size_t)std::max(M, K) * std::max(std::max(K, N), M) + 16;
  cudaMalloc(&x, big * 4); cudaMalloc(&w, big * 4); cudaMalloc(&y, big * 4); cudaMalloc(&b, big * 4);
  cudaMalloc(&pre, big * 4); cudaMalloc(&res, big * 4);
  cudaMemset(x, 0, big * 4); cudaMemset(w, 0, big * 4); cudaMemset(y, 0, big * 4); cudaMemset(b, 0, big * 4);
  cudaMemset(pre, 0, big * 4); cudaMemset(res, 0, big * 4);
  GemmDesc d{};
  if (which == 0) {        // fwd NN
    d.A = x; d.B = w; d.C = y; d.M = M; d.N = N; d.K = K; d.sAm = K; d.sAk = 1; d.sBk = N; d.sBn = 1; d.ldc = N;
    d.ep.bias = b; d.ep.act = act; if (extras) { d.ep.pre_out = pre; d.ep.addend = res; }
  } else if (which == 1) { // dX NT
    d.A = x; d.B = w; d.C = y; d.M = M; d.N = K; d.K = N; d.sAm = N; d.sAk = 1; d.sBk = 1; d.sBn = N; d.ldc = K;
    if (extras) { d.ep.dact = 2; d.ep.pre_in = pre; d.ep.addend = res; }
  } else {                 // dW TN
    d.A = x; d.B = w; d.C = y; d.M = K; d.N = N; d.K = M; d.sAm = 1; d.sAk = K; d.sBk = N; d.sBn = 1; d.ldc = N;
    d.ep.atomic = 1; d.splitk = 2;
  }
  bool handled = false;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int it = 0; it < 5; it++) {
    cudaEventRecord(e0, g_stream);
    gemm_tc(d, passes, &handled);
    cudaEventRecord(e1, g_stream);
    cudaStreamSynchronize(g_stream);
    float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
  }
  cudaError_t err = cudaGetLastError();
  std::vector<unsigned long long> t(16384 * 8);
  cudaMemcpyFromSymbol(t.data(), g_tc_times, sizeof(unsigned long long) * 16384 * 8);
  // averages relative to the earliest CTA start
  unsigned long long base = ~0ull;
  int ncta = 0;
  for (int c = 0; c < 16384; c++) if (t[c * 8]) { base = std::min(base, t[c * 8]); ncta++; }
  double avg[8] = {0};
  unsigned long long last = 0;
  for (int c = 0; c < 16384; c++) if (t[c * 8]) { for (int e = 0; e < 8; e++) avg[e] += (double)(t[c * 8 + e] - (e == 0 ? base : t[c * 8])) / ncta; last = std::max(last, t[c*8+7]); }
  printf("%-28s passes=%d handled=%d err=%d event=%.1fus ctas=%d span=%.1fus | start+%.1f setup %.1f full0 %.1f conv0_done %.1f conv_done %.1f mma_issued %.1f tmem_full %.1f epi_done %.1f (us after CTA start)\n",
         name, passes, (int)handled, (int)err, best * 1e3, ncta, (last - base) / 1e3, avg[0] / 1e3, avg[1] / 1e3, avg[2] / 1e3, avg[3] / 1e3, avg[4] / 1e3, avg[5] / 1e3, avg[6] / 1e3, avg[7] / 1e3);
  cudaMemset(x, 0, 4);
  void* sym; cudaGetSymbolAddress(&sym, g_tc_times); cudaMemset(sym, 0, sizeof(unsigned long long) * 16384 * 8);
  cudaFree(x); cudaFree(w); cudaFree(y); cudaFree(b); cudaFree(pre); cudaFree(res);
}

int main() {
  if (nt_init(0)) { printf("init failed: %s\n", nt_last_error()); return 1; }
  for (int passes : {1, 3}) {
    run("fwd 4900x16x96 bias", 4900, 16, 96, 0, passes, 0, false);
    run("fwd 4900x96x288 bias", 4900, 96, 288, 0, passes, 0, false);
    run("fwd 4900x96x192 gelu+pre+res", 4900, 96, 192, 0, passes, 2, true);
    run("fwd 100x4704x96", 100, 4704, 96, 0, passes, 0, false);
    run("dX 4900x96x288", 4900, 96, 288, 1, passes, 0, false);
    run("dX 4900x96x192 dgelu+add", 4900, 96, 192, 1, passes, 0, true);
    run("dW 4900x96x288", 4900, 96, 288, 2, passes, 0, false);
    run("fwd 50176x384x1152", 50176, 384, 1152, 0, passes, 0, false);
  }
  return 0;
}
