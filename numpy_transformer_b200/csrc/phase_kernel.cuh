// Device code of the phase-program kernel (see phase_prog.h).  One thread-block cluster runs a whole list of per-op
// phases: every CTA takes a slice of each phase (output columns of a Linear, weight rows of its backward, (sample, head)
// items of the attention, rows of a LayerNorm ...), results go through global memory (L2), and a hardware cluster barrier
// stands where a kernel boundary stood in the per-op path.  Exact fp32 FMAs throughout (these shapes have <= 16 rows: a
// 128-row tensor-core tile would be >= 87 % padding), so the results agree with the per-op kernels to fp32 rounding.
//
// The file is written against a handful of ph_* primitives so that the SAME source also compiles as host C++ under the
// CPU SIMT emulator of tests/simt_emu (fibres for threads, real barriers and shuffles), which is how the `not gpu` test
// suite checks it against the fp64 restatement of the reference.  Rules that keep both builds honest: no early returns (every thread reaches every
// barrier), warp shuffles only under warp-uniform control flow, dynamic shared memory only.
#pragma once
#include <cmath>
#include "phase_prog.h"

#ifndef NT_EMU
__device__ __forceinline__ int ph_tid() { return (int)threadIdx.x; }
__device__ __forceinline__ int ph_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return (int)r;
}
__device__ __forceinline__ int ph_nrank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return (int)r;
}
__device__ __forceinline__ void ph_syncthreads() { __syncthreads(); }
// release / acquire at cluster scope: every global (and shared) write of every thread of the cluster before the barrier is
// visible to every thread after it; the acquire side also invalidates the L1 lines of the waiting SM
__device__ __forceinline__ void ph_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
extern __shared__ __align__(16) float ph_smem_[];
__device__ __forceinline__ float* ph_smem() { return ph_smem_; }
__device__ __forceinline__ float ph_shfl_xor(float v, int off) { return __shfl_xor_sync(0xffffffffu, v, off); }
__device__ __forceinline__ int ph_shfl_xor(int v, int off) { return __shfl_xor_sync(0xffffffffu, v, off); }
__device__ __forceinline__ void ph_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ long long ph_clock() { return clock64(); }
#else
static inline long long ph_clock() { return 0; }
#endif

namespace ntph {

constexpr int NT = NT_PH_THREADS, NW = NT / 32;
#define kInf INFINITY

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += ph_shfl_xor(v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, ph_shfl_xor(v, o));
  return v;
}
// (value, index) maximum with the smallest index on ties (np.argmax)
__device__ __forceinline__ void warp_argmax(float& v, int& idx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = ph_shfl_xor(v, o);
    const int oi = ph_shfl_xor(idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
}
__device__ __forceinline__ float gelu(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float x) {
  const float s = x * 0.70710678118654752f;
  return 0.5f + 0.5f * erff(s) + x * 0.3989422804014327f * expf(-s * s);
}
template <typename T>
__device__ __forceinline__ T* arg(const NtPhase& ph, int k) { return reinterpret_cast<T*>(static_cast<uintptr_t>(ph.p[k])); }
__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
__device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }

// ---------------------------------------------------------------------------------------------------------------- Linear
// y[M][N] = act(LN?(x)[M][K] w[K][N] + b) + residual.  The CTA owns a contiguous run of column quads; x (transposed, zero
// padded to MT rows) sits in shared memory; a lane is (column quad, k-group): QP = next power of two of the quads in
// flight, 32 / QP k-groups per warp, NW warps -> the weight panel of the slice is streamed once with float4 loads, U of
// them in flight per thread; k-groups fold by shuffles inside the warp, across warps through shared memory.
template <int MT>
__device__ void linear_fwd(const NtPhase& ph, float* sm, int rank, int C) {
  constexpr int XS = MT >= 4 ? MT + 4 : MT;             // row pitch of the transposed panel (bank spread for float4 reads)
  constexpr int RC = MT < 4 ? MT : 4;                   // rows per cross-warp reduction round
  constexpr int U = MT <= 2 ? 8 : (MT <= 8 ? 4 : 2);    // weight loads in flight per thread
  const int tid = ph_tid(), warp = tid >> 5, lane = tid & 31;
  const int M = ph.i[0], K = ph.i[1], N = ph.i[2], act = ph.i[3];
  const float* x = arg<const float>(ph, 0);
  const float* w = arg<const float>(ph, 1);
  const float* b = arg<const float>(ph, 2);
  float* y = arg<float>(ph, 3);
  float* pre = arg<float>(ph, 4);
  const float* res = arg<const float>(ph, 5);
  const float* ln_g = arg<const float>(ph, 6);
  const float* ln_b = arg<const float>(ph, 7);
  float* xs = sm;                                                         // [K][XS]
  float4* red = reinterpret_cast<float4*>(sm + ((K * XS + 3) & ~3));      // [NW][RC][32]
  for (int e = tid; e < MT * K; e += NT) {
    const int m = e / K, k = e - m * K;
    xs[k * XS + m] = m < M ? x[e] : 0.f;
  }
  ph_syncthreads();
  if (ln_g) {                      // net/layernorm.py:100-114 on the staged rows: biased variance, eps 1e-5
    for (int m = warp; m < M; m += NW) {
      float s = 0.f;
      for (int k = lane; k < K; k += 32) s += xs[k * XS + m];
      const float mean = warp_sum(s) / K;
      float q = 0.f;
      for (int k = lane; k < K; k += 32) { const float d = xs[k * XS + m] - mean; q += d * d; }
      const float rstd = 1.0f / sqrtf(warp_sum(q) / K + 1e-5f);
      for (int k = lane; k < K; k += 32) xs[k * XS + m] = (xs[k * XS + m] - mean) * rstd * ln_g[k] + ln_b[k];
    }
    ph_syncthreads();
  }
  const int NQ = (N + 3) >> 2, per = (NQ + C - 1) / C;
  const int q_begin = rank * per, q_end = imin(NQ, q_begin + per);
  const bool vec = ((N & 3) == 0) && aligned16(w);
  for (int qb = q_begin; qb < q_end; qb += 32) {
    const int nqc = imin(32, q_end - qb);
    int QP = 1;
    while (QP < nqc) QP <<= 1;
    const int ql = lane & (QP - 1), kgl = lane / QP, KGW = 32 / QP, kg = warp * KGW + kgl, KG = NW * KGW;
    const bool active = ql < nqc;
    const int n0 = (qb + ql) * 4;
    float acc[MT][4];
#pragma unroll
    for (int r = 0; r < MT; r++) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
    if (active) {
      for (int k0 = kg; k0 < K; k0 += KG * U) {
        float wv[U][4];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int k = k0 + u * KG;
          wv[u][0] = wv[u][1] = wv[u][2] = wv[u][3] = 0.f;
          if (k < K) {
            const float* wr = w + (long long)k * N + n0;
            if (vec) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(wr));
              wv[u][0] = t.x; wv[u][1] = t.y; wv[u][2] = t.z; wv[u][3] = t.w;
            } else {
#pragma unroll
              for (int c = 0; c < 4; c++) if (n0 + c < N) wv[u][c] = __ldg(wr + c);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int k = k0 + u * KG;
          if (k < K) {
#pragma unroll
            for (int r = 0; r < MT; r++) {
              const float xv = xs[k * XS + r];
#pragma unroll
              for (int c = 0; c < 4; c++) acc[r][c] = fmaf(xv, wv[u][c], acc[r][c]);
            }
          }
        }
      }
    }
    for (int off = 16; off >= QP; off >>= 1) {
#pragma unroll
      for (int r = 0; r < MT; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[r][c] += ph_shfl_xor(acc[r][c], off);
    }
#pragma unroll
    for (int mc = 0; mc < MT; mc += RC) {
      if (kgl == 0 && active) {
#pragma unroll
        for (int r = 0; r < RC; r++)
          red[(warp * RC + r) * 32 + ql] = make_float4(acc[mc + r][0], acc[mc + r][1], acc[mc + r][2], acc[mc + r][3]);
      }
      ph_syncthreads();
      for (int e = tid; e < RC * nqc; e += NT) {
        const int r = e / nqc, q = e - r * nqc, m = mc + r;
        if (m < M) {
          float s[4] = {0.f, 0.f, 0.f, 0.f};
          for (int wp = 0; wp < NW; wp++) {
            const float4 t = red[(wp * RC + r) * 32 + q];
            s[0] += t.x; s[1] += t.y; s[2] += t.z; s[3] += t.w;
          }
#pragma unroll
          for (int c = 0; c < 4; c++) {
            const int n = (qb + q) * 4 + c;
            if (n < N) {
              float v = s[c];
              if (b) v += b[n];
              const long long o = (long long)m * N + n;
              if (pre) pre[o] = v;
              if (act == 1) v = fmaxf(v, 0.f);
              else if (act == 2) v = gelu(v);
              if (res) v += res[o];
              y[o] = v;
            }
          }
        }
      }
      ph_syncthreads();
    }
  }
}

// dx[M][K] = (dy[M][N] w[K][N]^T) o act'(pre) + addend: one warp per weight row (read once, coalesced), dy in shared memory
template <int MT>
__device__ void linear_bwd_input(const NtPhase& ph, float* sm, int rank, int C) {
  const int tid = ph_tid(), warp = tid >> 5, lane = tid & 31;
  const int M = ph.i[0], K = ph.i[1], N = ph.i[2], act = ph.i[3];
  const float* dy = arg<const float>(ph, 0);
  const float* w = arg<const float>(ph, 1);
  float* dx = arg<float>(ph, 2);
  const float* pre = arg<const float>(ph, 3);
  const float* addend = arg<const float>(ph, 4);
  const int NP = (N + 3) & ~3;
  float* sdy = sm;                                     // [MT][NP]
  for (int e = tid; e < MT * NP; e += NT) {
    const int m = e / NP, n = e - m * NP;
    sdy[e] = (m < M && n < N) ? dy[(long long)m * N + n] : 0.f;
  }
  ph_syncthreads();
  const bool vec = ((N & 3) == 0) && aligned16(w);
  for (int k = rank * NW + warp; k < K; k += C * NW) {
    const float* wr = w + (long long)k * N;
    float acc[MT];
#pragma unroll
    for (int r = 0; r < MT; r++) acc[r] = 0.f;
    for (int n = 4 * lane; n < NP; n += 128) {
      float t[4] = {0.f, 0.f, 0.f, 0.f};
      if (vec) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(wr + n));
        t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
      } else {
#pragma unroll
        for (int c = 0; c < 4; c++) if (n + c < N) t[c] = __ldg(wr + n + c);
      }
#pragma unroll
      for (int r = 0; r < MT; r++) {
        const float4 d = *reinterpret_cast<const float4*>(sdy + r * NP + n);
        acc[r] = fmaf(t[0], d.x, fmaf(t[1], d.y, fmaf(t[2], d.z, fmaf(t[3], d.w, acc[r]))));
      }
    }
#pragma unroll
    for (int r = 0; r < MT; r++) acc[r] = warp_sum(acc[r]);
#pragma unroll
    for (int r = 0; r < MT; r++) {
      if (lane == r && r < M) {
        const long long o = (long long)r * K + k;
        float v = acc[r];
        if (act == 1) v = pre[o] > 0.f ? v : 0.f;
        else if (act == 2) v *= gelu_grad(pre[o]);
        if (addend) v += addend[o];
        dx[o] = v;
      }
    }
  }
}

// dw[K][N] += x[M][K]^T dy[M][N], db[N] += colsum(dy): one warp per weight row, every element has one owner (plain
// read-modify-write keeps the reference's accumulate-until-setzero semantics, fullconnect.py:118-122)
template <int MT>
__device__ void linear_bwd_params(const NtPhase& ph, float* sm, int rank, int C) {
  const int tid = ph_tid(), warp = tid >> 5, lane = tid & 31;
  const int M = ph.i[0], K = ph.i[1], N = ph.i[2];
  const float* x = arg<const float>(ph, 0);
  const float* dy = arg<const float>(ph, 1);
  float* dw = arg<float>(ph, 2);
  float* db = arg<float>(ph, 3);
  const int NP = (N + 3) & ~3;
  float* sdy = sm;                                     // [MT][NP]
  float* sx = sm + MT * NP;                            // [K][MT]
  for (int e = tid; e < MT * NP; e += NT) {
    const int m = e / NP, n = e - m * NP;
    sdy[e] = (m < M && n < N) ? dy[(long long)m * N + n] : 0.f;
  }
  for (int e = tid; e < MT * K; e += NT) {
    const int m = e / K, k = e - m * K;
    sx[k * MT + m] = m < M ? x[e] : 0.f;
  }
  ph_syncthreads();
  const bool vec = ((N & 3) == 0) && aligned16(dw);
  for (int k = rank * NW + warp; k < K; k += C * NW) {
    float xv[MT];
#pragma unroll
    for (int r = 0; r < MT; r++) xv[r] = sx[k * MT + r];
    float* dr = dw + (long long)k * N;
    for (int n = 4 * lane; n < NP; n += 128) {
      float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int r = 0; r < MT; r++) {
        const float4 d = *reinterpret_cast<const float4*>(sdy + r * NP + n);
        a[0] = fmaf(xv[r], d.x, a[0]); a[1] = fmaf(xv[r], d.y, a[1]); a[2] = fmaf(xv[r], d.z, a[2]); a[3] = fmaf(xv[r], d.w, a[3]);
      }
      if (vec) {
        float4 t = *reinterpret_cast<float4*>(dr + n);
        t.x += a[0]; t.y += a[1]; t.z += a[2]; t.w += a[3];
        *reinterpret_cast<float4*>(dr + n) = t;
      } else {
#pragma unroll
        for (int c = 0; c < 4; c++) if (n + c < N) dr[n + c] += a[c];
      }
    }
  }
  if (db) {
    for (int n = rank * NT + tid; n < N; n += C * NT) {
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < MT; r++) s += sdy[r * NP + n];
      db[n] += s;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------- LayerNorm
__device__ void ln_fwd(const NtPhase& ph, int rank, int C) {
  const int tid = ph_tid(), warp = tid >> 5, lane = tid & 31;
  const int M = ph.i[0], D = ph.i[1];
  const float eps = ph.f[0];
  const float* x = arg<const float>(ph, 0);
  const float* gamma = arg<const float>(ph, 1);
  const float* beta = arg<const float>(ph, 2);
  float* y = arg<float>(ph, 3);
  float* mean_out = arg<float>(ph, 4);
  float* rstd_out = arg<float>(ph, 5);
  for (int m = rank * NW + warp; m < M; m += C * NW) {
    const float* xr = x + (long long)m * D;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s += xr[c];
    const float mean = warp_sum(s) / D;
    float v = 0.f;
    for (int c = lane; c < D; c += 32) { const float t = xr[c] - mean; v += t * t; }
    const float rstd = 1.0f / sqrtf(warp_sum(v) / D + eps);             // biased variance, np.var
    if (lane == 0) {
      if (mean_out) mean_out[m] = mean;
      if (rstd_out) rstd_out[m] = rstd;
    }
    float* yr = y + (long long)m * D;
    for (int c = lane; c < D; c += 32) yr[c] = (xr[c] - mean) * rstd * gamma[c] + beta[c];
  }
}

// dx = rstd (g dy - mean(g dy) - xhat mean(g dy xhat)) (+ addend); dgamma += sum_r xhat dy; dbeta += sum_r dy
__device__ void ln_bwd(const NtPhase& ph, int rank, int C) {
  const int tid = ph_tid(), warp = tid >> 5, lane = tid & 31;
  const int M = ph.i[0], D = ph.i[1];
  const float* dy = arg<const float>(ph, 0);
  const float* x = arg<const float>(ph, 1);
  const float* mean = arg<const float>(ph, 2);
  const float* rstd = arg<const float>(ph, 3);
  const float* gamma = arg<const float>(ph, 4);
  float* dx = arg<float>(ph, 5);
  float* dgamma = arg<float>(ph, 6);
  float* dbeta = arg<float>(ph, 7);
  const float* addend = arg<const float>(ph, 8);
  if (dx) {
    for (int m = rank * NW + warp; m < M; m += C * NW) {
      const long long off = (long long)m * D;
      const float mu = mean[m], rs = rstd[m];
      float s1 = 0.f, s2 = 0.f;
      for (int c = lane; c < D; c += 32) {
        const float gdy = gamma[c] * dy[off + c];
        const float xh = (x[off + c] - mu) * rs;
        s1 += gdy;
        s2 += gdy * xh;
      }
      s1 = warp_sum(s1) / D;
      s2 = warp_sum(s2) / D;
      for (int c = lane; c < D; c += 32) {
        const float gdy = gamma[c] * dy[off + c];
        const float xh = (x[off + c] - mu) * rs;
        float o = rs * (gdy - s1 - xh * s2);
        if (addend) o += addend[off + c];
        dx[off + c] = o;
      }
    }
  }
  if (dgamma || dbeta) {
    for (int c = rank * NT + tid; c < D; c += C * NT) {
      float sg = 0.f, sb = 0.f;
      for (int r = 0; r < M; r++) {
        const float d = dy[(long long)r * D + c];
        sg += (x[(long long)r * D + c] - mean[r]) * rstd[r] * d;
        sb += d;
      }
      if (dgamma) dgamma[c] += sg;
      if (dbeta) dbeta[c] += sb;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------- attention
// N <= 32 tokens, dh % 4 == 0: one CTA per (sample, head) item, everything of the item in shared memory.
// out = softmax(q k^T / sqrt(dh) [+ causal]) v (+ residual); P [B][H][N][N] is the reference's atg__.
__device__ void attn_fwd(const NtPhase& ph, float* sm, int rank, int C) {
  const int tid = ph_tid();
  const int B = ph.i[0], N = ph.i[1], H = ph.i[2], dh = ph.i[3], mask_kind = ph.i[4];
  const float* qkv = arg<const float>(ph, 0);
  float* out = arg<float>(ph, 1);
  float* P = arg<float>(ph, 2);
  const float* residual = arg<const float>(ph, 3);
  const int D = H * dh, DP = dh + 4, dh4 = dh >> 2, NP1 = N + 1;
  float* sq = sm;                  // [N][DP]
  float* sk = sq + N * DP;
  float* sv = sk + N * DP;
  float* sp = sv + N * DP;         // [N][N + 1]
  const float scale = 1.0f / sqrtf((float)dh);
  for (int item = rank; item < B * H; item += C) {
    const int b = item / H, h = item - b * H;
    const float* base = qkv + (long long)b * N * 3 * D + h * dh;
    for (int e = tid; e < N * dh4; e += NT) {
      const int i = e / dh4, d4 = (e - i * dh4) * 4;
      const float* r = base + (long long)i * 3 * D + d4;
      *reinterpret_cast<float4*>(sq + i * DP + d4) = *reinterpret_cast<const float4*>(r);
      *reinterpret_cast<float4*>(sk + i * DP + d4) = *reinterpret_cast<const float4*>(r + D);
      *reinterpret_cast<float4*>(sv + i * DP + d4) = *reinterpret_cast<const float4*>(r + 2 * D);
    }
    ph_syncthreads();
    for (int e = tid; e < N * N; e += NT) {
      const int i = e / N, j = e - i * N;
      float s = 0.f;
      for (int d = 0; d < dh; d += 4) {
        const float4 a = *reinterpret_cast<const float4*>(sq + i * DP + d);
        const float4 k4 = *reinterpret_cast<const float4*>(sk + j * DP + d);
        s = fmaf(a.x, k4.x, s); s = fmaf(a.y, k4.y, s); s = fmaf(a.z, k4.z, s); s = fmaf(a.w, k4.w, s);
      }
      s *= scale;                                                     // attdecoderblock.py:57
      if (mask_kind == 1 && j > i) s = -kInf;                         // causal: 0 / -inf lower-triangular mask
      sp[i * NP1 + j] = s;
    }
    ph_syncthreads();
    if (tid < N) {
      float* row = sp + tid * NP1;
      float mx = -kInf;
      for (int j = 0; j < N; j++) mx = fmaxf(mx, row[j]);
      float s = 0.f;
      for (int j = 0; j < N; j++) { const float e_ = expf(row[j] - mx); row[j] = e_; s += e_; }
      const float inv = 1.0f / s;
      float* prow = P ? P + ((long long)item * N + tid) * N : nullptr;
      for (int j = 0; j < N; j++) {
        const float pv = row[j] * inv;
        row[j] = pv;
        if (prow) prow[j] = pv;
      }
    }
    ph_syncthreads();
    for (int e = tid; e < N * dh4; e += NT) {
      const int i = e / dh4, d4 = (e - i * dh4) * 4;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int j = 0; j < N; j++) {
        const float pv = sp[i * NP1 + j];
        const float4 v4 = *reinterpret_cast<const float4*>(sv + j * DP + d4);
        a0 = fmaf(pv, v4.x, a0); a1 = fmaf(pv, v4.y, a1); a2 = fmaf(pv, v4.z, a2); a3 = fmaf(pv, v4.w, a3);
      }
      const long long o = ((long long)b * N + i) * D + h * dh + d4;
      if (residual) {
        const float4 rr = *reinterpret_cast<const float4*>(residual + o);
        a0 += rr.x; a1 += rr.y; a2 += rr.z; a3 += rr.w;
      }
      *reinterpret_cast<float4*>(out + o) = make_float4(a0, a1, a2, a3);
    }
    ph_syncthreads();
  }
}

// dV = P^T dO, dP = dO V^T, dS = P o (dP - rowsum(dP o P)), dQ = dS K / sqrt(dh), dK = dS^T Q / sqrt(dh)
__device__ void attn_bwd(const NtPhase& ph, float* sm, int rank, int C) {
  const int tid = ph_tid();
  const int B = ph.i[0], N = ph.i[1], H = ph.i[2], dh = ph.i[3];
  const float* dout = arg<const float>(ph, 0);
  const float* qkv = arg<const float>(ph, 1);
  const float* P = arg<const float>(ph, 2);
  float* dqkv = arg<float>(ph, 3);
  const int D = H * dh, DP = dh + 4, dh4 = dh >> 2, NP1 = N + 1;
  float* sq = sm;
  float* sk = sq + N * DP;
  float* sv = sk + N * DP;
  float* sdo = sv + N * DP;
  float* sp = sdo + N * DP;        // [N][N + 1]
  float* sds = sp + N * NP1;       // [N][N + 1]
  const float scale = 1.0f / sqrtf((float)dh);
  for (int item = rank; item < B * H; item += C) {
    const int b = item / H, h = item - b * H;
    const float* base = qkv + (long long)b * N * 3 * D + h * dh;
    for (int e = tid; e < N * dh4; e += NT) {
      const int i = e / dh4, d4 = (e - i * dh4) * 4;
      const float* r = base + (long long)i * 3 * D + d4;
      *reinterpret_cast<float4*>(sq + i * DP + d4) = *reinterpret_cast<const float4*>(r);
      *reinterpret_cast<float4*>(sk + i * DP + d4) = *reinterpret_cast<const float4*>(r + D);
      *reinterpret_cast<float4*>(sv + i * DP + d4) = *reinterpret_cast<const float4*>(r + 2 * D);
      *reinterpret_cast<float4*>(sdo + i * DP + d4) =
          *reinterpret_cast<const float4*>(dout + ((long long)b * N + i) * D + h * dh + d4);
    }
    for (int e = tid; e < N * N; e += NT) {
      const int i = e / N, j = e - i * N;
      sp[i * NP1 + j] = P[(long long)item * N * N + e];
    }
    ph_syncthreads();
    for (int e = tid; e < N * N; e += NT) {                            // dP = dO V^T (attention.py:74)
      const int i = e / N, j = e - i * N;
      float s = 0.f;
      for (int d = 0; d < dh; d += 4) {
        const float4 g = *reinterpret_cast<const float4*>(sdo + i * DP + d);
        const float4 v4 = *reinterpret_cast<const float4*>(sv + j * DP + d);
        s = fmaf(g.x, v4.x, s); s = fmaf(g.y, v4.y, s); s = fmaf(g.z, v4.z, s); s = fmaf(g.w, v4.w, s);
      }
      sds[i * NP1 + j] = s;
    }
    ph_syncthreads();
    if (tid < N) {                                                     // closed form of the per-row Jacobian (activation.py:86-91)
      float s = 0.f;
      for (int j = 0; j < N; j++) s += sds[tid * NP1 + j] * sp[tid * NP1 + j];
      for (int j = 0; j < N; j++) sds[tid * NP1 + j] = sp[tid * NP1 + j] * (sds[tid * NP1 + j] - s);
    }
    ph_syncthreads();
    float* obase = dqkv + (long long)b * N * 3 * D + h * dh;
    for (int e = tid; e < 3 * N * dh4; e += NT) {
      const int which = e / (N * dh4), f = e - which * (N * dh4);
      const int r = f / dh4, d4 = (f - r * dh4) * 4;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int t = 0; t < N; t++) {
        float cf;
        const float* src;
        if (which == 0) { cf = sds[r * NP1 + t]; src = sk + t * DP + d4; }           // dQ[r] = sum_j dS[r][j] k[j]
        else if (which == 1) { cf = sds[t * NP1 + r]; src = sq + t * DP + d4; }      // dK[r] = sum_i dS[i][r] q[i]
        else { cf = sp[t * NP1 + r]; src = sdo + t * DP + d4; }                      // dV[r] = sum_i P[i][r] dO[i]
        const float4 v4 = *reinterpret_cast<const float4*>(src);
        a0 = fmaf(cf, v4.x, a0); a1 = fmaf(cf, v4.y, a1); a2 = fmaf(cf, v4.z, a2); a3 = fmaf(cf, v4.w, a3);
      }
      const float sc = which == 2 ? 1.0f : scale;
      *reinterpret_cast<float4*>(obase + (long long)r * 3 * D + which * D + d4) = make_float4(a0 * sc, a1 * sc, a2 * sc, a3 * sc);
    }
    ph_syncthreads();
  }
}

// one new query per (sample, head) over the cached keys / values; `append`: the item's k / v slice is stored into cache
// row len-1 first (the rows are re-read by the same CTA after the barrier)
__device__ void attn_decode(const NtPhase& ph, float* sm, int rank, int C) {
  const int tid = ph_tid(), warp = tid >> 5, lane = tid & 31;
  const int B = ph.i[0], Tmax = ph.i[1], H = ph.i[2], dh = ph.i[3], append = ph.i[4];
  const float* qkv_new = arg<const float>(ph, 0);
  float* kc = arg<float>(ph, 1);
  float* vc = arg<float>(ph, 2);
  float* out = arg<float>(ph, 3);
  const int32_t* d_pos = arg<const int32_t>(ph, 4);
  const int len = imin(*d_pos + 1, Tmax);
  const int D = H * dh;
  float* sq = sm;                      // [dh]
  float* so = sq + dh;                 // [NW][dh]
  float* sr = so + NW * dh;            // [2 NW]
  float* sp = sr + 2 * NW;             // [Tmax]
  const float scale = 1.0f / sqrtf((float)dh);
  const bool vec = ((dh & 3) == 0) && ((D & 3) == 0) && aligned16(kc);
  for (int item = rank; item < B * H; item += C) {
    const int b = item / H, h = item - b * H;
    for (int d = tid; d < dh; d += NT) {
      const long long qo = (long long)b * 3 * D + h * dh + d;
      sq[d] = qkv_new[qo];
      if (append) {
        const long long co = ((long long)b * Tmax + (len - 1)) * D + h * dh + d;
        kc[co] = qkv_new[qo + D];
        vc[co] = qkv_new[qo + 2 * D];
      }
    }
    ph_syncthreads();
    const float* kb = kc + (long long)b * Tmax * D + h * dh;
    const float* vb = vc + (long long)b * Tmax * D + h * dh;
    float mx = -kInf;
    for (int j = tid; j < len; j += NT) {
      const float* kr = kb + (long long)j * D;
      float s = 0.f;
      if (vec) {
        for (int d = 0; d < dh; d += 4) {
          const float4 k4 = *reinterpret_cast<const float4*>(kr + d);
          s = fmaf(sq[d], k4.x, s); s = fmaf(sq[d + 1], k4.y, s); s = fmaf(sq[d + 2], k4.z, s); s = fmaf(sq[d + 3], k4.w, s);
        }
      } else {
        for (int d = 0; d < dh; d++) s = fmaf(sq[d], kr[d], s);
      }
      s *= scale;
      sp[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    if (lane == 0) sr[warp] = mx;
    ph_syncthreads();
    mx = sr[0];
    for (int wp = 1; wp < NW; wp++) mx = fmaxf(mx, sr[wp]);
    float sum = 0.f;
    for (int j = tid; j < len; j += NT) { const float e_ = expf(sp[j] - mx); sp[j] = e_; sum += e_; }
    sum = warp_sum(sum);
    if (lane == 0) sr[NW + warp] = sum;
    ph_syncthreads();
    float tot = 0.f;
    for (int wp = 0; wp < NW; wp++) tot += sr[NW + wp];
    const float inv = 1.0f / tot;
    for (int d = lane; d < dh; d += 32) {
      float acc = 0.f;
      for (int j = warp; j < len; j += NW) acc = fmaf(sp[j], vb[(long long)j * D + d], acc);
      so[warp * dh + d] = acc;
    }
    ph_syncthreads();
    for (int d = tid; d < dh; d += NT) {
      float s = 0.f;
      for (int wp = 0; wp < NW; wp++) s += so[wp * dh + d];
      out[(long long)b * D + h * dh + d] = s * inv;
    }
    ph_syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------- loss
// net/loss.py:46-51 per row (one warp each): softmax, -log p[label], gradient (p - y) / n, arg-max of the returned
// probabilities with np.argmax's first-index tie-break
__device__ void ce_rows(const NtPhase& ph, int rank, int C) {
  const int tid = ph_tid(), warp = tid >> 5, lane = tid & 31;
  const int rows = ph.i[0], cols = ph.i[1];
  const float inv_n = ph.f[0];
  const float* logits = arg<const float>(ph, 0);
  const int32_t* labels = arg<const int32_t>(ph, 1);
  const float* onehot = arg<const float>(ph, 2);
  float* row_loss = arg<float>(ph, 3);
  float* grad = arg<float>(ph, 4);
  float* prob = arg<float>(ph, 5);
  int32_t* argmax_out = arg<int32_t>(ph, 6);
  int* err = arg<int>(ph, 7);
  for (int row = rank * NW + warp; row < rows; row += C * NW) {
    const float* lr = logits + (long long)row * cols;
    float mx = -kInf;
    for (int j = lane; j < cols; j += 32) mx = fmaxf(mx, lr[j]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int j = lane; j < cols; j += 32) s += expf(lr[j] - mx);
    s = warp_sum(s);
    const float inv = 1.0f / s;
    const int lab = labels ? labels[row] : -1;
    if (labels && (lab < 0 || lab >= cols) && lane == 0 && err) *err = 2;        // NT_DEVERR_LABEL: flagged, loss 0
    float lossacc = 0.f, pm = -kInf;
    int pi = 0x7fffffff;
    for (int j = lane; j < cols; j += 32) {
      const float p = expf(lr[j] - mx) * inv;
      const float yv = onehot ? onehot[(long long)row * cols + j] : (j == lab ? 1.f : 0.f);
      if (prob) prob[(long long)row * cols + j] = p;
      if (grad) grad[(long long)row * cols + j] = (p - yv) * inv_n;
      if (yv != 0.f) lossacc -= yv * logf(p + 1e-10f);
      if (p > pm) { pm = p; pi = j; }
    }
    lossacc = warp_sum(lossacc);
    warp_argmax(pm, pi);
    if (lane == 0) {
      row_loss[row] = lossacc;
      if (argmax_out) argmax_out[row] = pi;
    }
  }
}

__device__ void ce_reduce(const NtPhase& ph, int rank) {
  if (rank == 0 && ph_tid() == 0) {
    const float* row_loss = arg<const float>(ph, 0);
    float s = 0.f;
    for (int i = 0; i < ph.i[0]; i++) s += row_loss[i];
    *arg<float>(ph, 1) = s * ph.f[0];
  }
}

// ---------------------------------------------------------------------------------------------------------------- embeddings
__device__ void embed_fwd(const NtPhase& ph, int rank, int C) {
  const int tid = ph_tid(), warp = tid >> 5, lane = tid & 31;
  const int BT = ph.i[0], T = ph.i[1], D = ph.i[2], vocab = ph.i[3];
  const int32_t* idx = arg<const int32_t>(ph, 0);
  const float* etok = arg<const float>(ph, 1);
  const float* epos = arg<const float>(ph, 2);
  float* out = arg<float>(ph, 3);
  int* err = arg<int>(ph, 4);
  for (int row = rank * NW + warp; row < BT; row += C * NW) {
    int tok = idx[row];
    if (tok < 0) tok += vocab;                        // NumPy wraps negative indices (Embedding.py:46)
    const bool bad = tok < 0 || tok >= vocab;         // ... and raises IndexError beyond the table: flag it, read nothing
    if (bad && lane == 0 && err) *err = 1;            // NT_DEVERR_TOKEN
    const int t = row % T;
    for (int c = lane; c < D; c += 32) {
      float v = bad ? 0.f : etok[(long long)tok * D + c];
      if (epos) v += epos[(long long)t * D + c];
      out[(long long)row * D + c] = v;
    }
  }
}

// detok[tok[row]] += dy[row], depos[t] += sum_b dy[b, t]: a thread owns one feature column of both tables, rows in order
// (repeated tokens need no atomics and the sum order is fixed)
__device__ void embed_bwd(const NtPhase& ph, int rank, int C) {
  const int tid = ph_tid();
  const int B = ph.i[0], T = ph.i[1], D = ph.i[2], vocab = ph.i[3];
  const int32_t* idx = arg<const int32_t>(ph, 0);
  const float* dy = arg<const float>(ph, 1);
  float* detok = arg<float>(ph, 2);
  float* depos = arg<float>(ph, 3);
  int* err = arg<int>(ph, 4);
  for (int c = rank * NT + tid; c < D; c += C * NT) {
    if (detok) {
      for (int row = 0; row < B * T; row++) {
        int tok = idx[row];
        if (tok < 0) tok += vocab;
        if (tok < 0 || tok >= vocab) {                // never scatter outside the gradient table
          if (c == 0 && err) *err = 1;
        } else {
          detok[(long long)tok * D + c] += dy[(long long)row * D + c];
        }
      }
    }
    if (depos) {
      for (int t = 0; t < T; t++) {
        float s = 0.f;
        for (int b = 0; b < B; b++) s += dy[((long long)b * T + t) * D + c];
        depos[(long long)t * D + c] += s;
      }
    }
  }
}

__device__ void decode_embed(const NtPhase& ph, int rank, int C) {
  const int tid = ph_tid(), warp = tid >> 5, lane = tid & 31;
  const int B = ph.i[0], D = ph.i[1], Tmax = ph.i[2], vocab = ph.i[3];
  const int32_t* idx = arg<const int32_t>(ph, 0);
  const float* etok = arg<const float>(ph, 1);
  const float* epos = arg<const float>(ph, 2);
  float* out = arg<float>(ph, 3);
  const int32_t* d_pos = arg<const int32_t>(ph, 4);
  int* err = arg<int>(ph, 5);
  const int t = imin(*d_pos, Tmax - 1);
  for (int row = rank * NW + warp; row < B; row += C * NW) {
    int tok = idx[row];
    if (tok < 0) tok += vocab;
    const bool bad = tok < 0 || tok >= vocab;
    if (bad && lane == 0 && err) *err = 1;
    for (int c = lane; c < D; c += 32)
      out[(long long)row * D + c] = (bad ? 0.f : etok[(long long)tok * D + c]) + epos[(long long)t * D + c];
  }
}

// greedy choice: tok[b] = argmax(logits[b, :cols]), also recorded at hist[b, *d_pos + 1]
__device__ void argmax_next(const NtPhase& ph, float* sm, int rank, int C) {
  const int tid = ph_tid(), warp = tid >> 5, lane = tid & 31;
  const int ld = ph.i[0], cols = ph.i[1], B = ph.i[2], hist_ld = ph.i[3];
  const float* logits = arg<const float>(ph, 0);
  int32_t* tok = arg<int32_t>(ph, 1);
  int32_t* hist = arg<int32_t>(ph, 2);
  const int32_t* d_pos = arg<const int32_t>(ph, 3);
  float* sv = sm;                                  // [NW]
  int* si = reinterpret_cast<int*>(sm + NW);       // [NW]
  for (int row = rank; row < B; row += C) {
    float v = -kInf;
    int idx = 0x7fffffff;
    for (int c = tid; c < cols; c += NT) {
      const float xv = logits[(long long)row * ld + c];
      if (xv > v) { v = xv; idx = c; }             // ascending c per thread: the first maximum is kept
    }
    warp_argmax(v, idx);
    if (lane == 0) { sv[warp] = v; si[warp] = idx; }
    ph_syncthreads();
    if (tid == 0) {
      for (int wp = 1; wp < NW; wp++)
        if (sv[wp] > v || (sv[wp] == v && si[wp] < idx)) { v = sv[wp]; idx = si[wp]; }
      tok[row] = idx;
      const int p = *d_pos + 1;
      if (hist && p < hist_ld) hist[(long long)row * hist_ld + p] = idx;
    }
    ph_syncthreads();
  }
}

#define NTPH_ROWS_DISPATCH(fn, M, ...)                 \
  do {                                                 \
    if ((M) <= 1) fn<1>(__VA_ARGS__);                  \
    else if ((M) <= 2) fn<2>(__VA_ARGS__);             \
    else if ((M) <= 4) fn<4>(__VA_ARGS__);             \
    else if ((M) <= 8) fn<8>(__VA_ARGS__);             \
    else fn<16>(__VA_ARGS__);                          \
  } while (0)

// the interpreter: phases in program order, a cluster barrier after the ones something later depends on
__device__ void run_program(const NtPhaseProgram& prog) {
  float* sm = ph_smem();
  const int rank = ph_rank(), C = ph_nrank();
  long long* stamps = reinterpret_cast<long long*>(static_cast<uintptr_t>(prog.stamps));
  for (int i = 0; i < prog.n; i++) {
    const NtPhase& ph = prog.ph[i];
    const long long t0 = stamps ? ph_clock() : 0;
    switch (ph.op) {
      case NT_PH_LINEAR_FWD: NTPH_ROWS_DISPATCH(linear_fwd, ph.i[0], ph, sm, rank, C); break;
      case NT_PH_LINEAR_BWD_INPUT: NTPH_ROWS_DISPATCH(linear_bwd_input, ph.i[0], ph, sm, rank, C); break;
      case NT_PH_LINEAR_BWD_PARAMS: NTPH_ROWS_DISPATCH(linear_bwd_params, ph.i[0], ph, sm, rank, C); break;
      case NT_PH_LN_FWD: ln_fwd(ph, rank, C); break;
      case NT_PH_LN_BWD: ln_bwd(ph, rank, C); break;
      case NT_PH_ATTN_FWD: attn_fwd(ph, sm, rank, C); break;
      case NT_PH_ATTN_BWD: attn_bwd(ph, sm, rank, C); break;
      case NT_PH_CE_ROWS: ce_rows(ph, rank, C); break;
      case NT_PH_CE_REDUCE: ce_reduce(ph, rank); break;
      case NT_PH_EMBED_FWD: embed_fwd(ph, rank, C); break;
      case NT_PH_EMBED_BWD: embed_bwd(ph, rank, C); break;
      case NT_PH_DECODE_EMBED: decode_embed(ph, rank, C); break;
      case NT_PH_ATTN_DECODE: attn_decode(ph, sm, rank, C); break;
      case NT_PH_ARGMAX_NEXT: argmax_next(ph, sm, rank, C); break;
      case NT_PH_INCREMENT:
        if (rank == 0 && ph_tid() == 0) *arg<int32_t>(ph, 0) += 1;
        break;
      default: break;
    }
    long long t1 = 0;
    if (stamps) {                        // debug timing: the CTA's body ends when its last thread is through
      ph_syncthreads();
      t1 = ph_clock();
    }
    if (ph.flags & NT_PHF_SYNC_AFTER) ph_cluster_sync();
    else ph_syncthreads();               // the next phase reuses the shared-memory staging area
    if (stamps && ph_tid() == 0) {
      long long* s = stamps + ((long long)i * C + rank) * 3;
      s[0] = t0; s[1] = t1; s[2] = ph_clock();
    }
  }
}

}  // namespace ntph
