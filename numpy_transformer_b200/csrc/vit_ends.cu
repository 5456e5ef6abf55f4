// The two ends of the ViT step for the launch-bound configuration (transformer_of_image.py:56-77, README sizes):
//   stem  PatchEmbed_flatten.forward/backward (PatchEmbed.py:22-49) + Position_learnable.forward (Position_add.py:20-23)
//   head  final layer_norm (net/layernorm.py) -> flatten -> classify_layer (classify.py:20-47: FC0 -> ReLU -> FC1)
//         -> cross_entropy_loss (net/loss.py:46-51) -> the backward of all of them
// On the per-op path these are ~17 dependent launches of a few microseconds each around the two fused block-stack
// kernels (vit_block.cu); here they are five: stem forward, head FC0 (+ final LayerNorm), head middle (ReLU, FC1,
// loss, argmax, their gradients), head backward (dFC0 + final LayerNorm backward) and stem backward.  The shapes
// are tiny (45 MFLOP in the head, 7.5 MFLOP in the stem), so plain fp32 FFMA with register tiles is enough and
// the results are exact fp32.
#include "common.cuh"

namespace nt {
namespace ve {

// ---------------------------------------------------------------- stem
// one CTA per image: patches -> lin = patches Wpe + b -> LayerNorm -> + pos
__global__ void __launch_bounds__(256) stem_fwd_kernel(NtVitEnds e, int C, int Hi, int Wi, int P, int D) {
  pdl_prologue();
  extern __shared__ __align__(16) float sm[];
  const int hl = Hi / P, wl = Wi / P, KP = C * hl * wl, N = P * P;
  float* sPat = sm;                    // [N][KP]
  float* sW = sPat + N * KP;           // [KP][D]
  float* sLin = sW + KP * D;           // [N][D]
  float* sPos = sLin + N * D;          // [N][D]  position table (zeros if none)
  float* sGB = sPos + N * D;           // [2][D]  gamma, beta
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* img = e.images + (size_t)b * C * Hi * Wi;
  // every global operand is fetched up front by independent loads (the kernel is pure latency otherwise)
  for (int i = tid; i < N * D; i += 256) sPos[i] = e.pos ? __ldg(e.pos + i) : 0.f;
  for (int i = tid; i < D; i += 256) { sGB[i] = __ldg(e.pe_g + i); sGB[D + i] = __ldg(e.pe_b + i); }
  for (int i = tid; i < N * KP; i += 256) {        // PatchEmbed.py:27-36: patches row-major, (c,h,w) inside a patch
    int r = i;
    const int w = r % wl; r /= wl;
    const int h = r % hl; r /= hl;
    const int c = r % C; r /= C;
    const int pj = r % P, pi = r / P;
    sPat[i] = img[((size_t)c * Hi + pi * hl + h) * Wi + pj * wl + w];
  }
  for (int i = tid; i < KP * D; i += 256) sW[i] = __ldg(e.wpe + i);
  __syncthreads();
  float* lin_g = e.pe_lin + (size_t)b * N * D;
  for (int o = tid; o < N * D; o += 256) {
    const int n = o / D, d = o - n * D;
    float acc = __ldg(e.bpe + d);
    for (int k = 0; k < KP; k++) acc = fmaf(sPat[n * KP + k], sW[k * D + d], acc);
    sLin[o] = acc;
    lin_g[o] = acc;                                  // kept for the LayerNorm backward
  }
  __syncthreads();
  float* x0 = e.x0 + (size_t)b * N * D;
  for (int n = warp; n < N; n += 8) {                // net/layernorm.py:100-114 per token
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s += sLin[n * D + d];
    const float mean = warp_sum(s) / D;
    float q = 0.f;
    for (int d = lane; d < D; d += 32) { const float t = sLin[n * D + d] - mean; q += t * t; }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / D + 1e-5f);
    if (lane == 0) { e.pe_stat[((size_t)b * N + n) * 2] = mean; e.pe_stat[((size_t)b * N + n) * 2 + 1] = rstd; }
    for (int d = lane; d < D; d += 32)                // + Position_add.py:20-23
      x0[n * D + d] = (sLin[n * D + d] - mean) * rstd * sGB[d] + sGB[D + d] + sPos[n * D + d];
  }
}

// one CTA per image: LayerNorm backward of the stem, then dWpe += patches^T dlin, dbpe += colsum(dlin)
__global__ void __launch_bounds__(256) stem_bwd_kernel(NtVitEnds e, const float* __restrict__ dx0, int C, int Hi, int Wi, int P,
                                                       int D) {
  pdl_prologue();
  extern __shared__ __align__(16) float sm[];
  const int hl = Hi / P, wl = Wi / P, KP = C * hl * wl, N = P * P;
  float* sPat = sm;                    // [N][KP]
  float* sDl = sPat + N * KP;          // [N][D]   dy, then d(lin)
  float* sRed = sDl + N * D;           // [8][2][D] per-warp dgamma / dbeta partials
  float* sLn = sRed + 16 * D;          // [N][D]   lin
  float* sG = sLn + N * D;             // [D] gamma
  float* sSt = sG + D;                 // [N][2] mean, rstd
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* img = e.images + (size_t)b * C * Hi * Wi;
  for (int i = tid; i < N * D; i += 256) {
    sDl[i] = dx0[(size_t)b * N * D + i];
    sLn[i] = e.pe_lin[(size_t)b * N * D + i];
  }
  for (int i = tid; i < D; i += 256) sG[i] = __ldg(e.pe_g + i);
  for (int i = tid; i < 2 * N; i += 256) sSt[i] = e.pe_stat[(size_t)b * N * 2 + i];
  for (int i = tid; i < N * KP; i += 256) {
    int r = i;
    const int w = r % wl; r /= wl;
    const int h = r % hl; r /= hl;
    const int c = r % C; r /= C;
    const int pj = r % P, pi = r / P;
    sPat[i] = img[((size_t)c * Hi + pi * hl + h) * Wi + pj * wl + w];
  }
  for (int i = tid; i < 16 * D; i += 256) sRed[i] = 0.f;
  __syncthreads();
  for (int n = warp; n < N; n += 8) {                // net/layernorm.py:126-133
    const float mean = sSt[2 * n], rstd = sSt[2 * n + 1];
    float s1 = 0.f, s2 = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float g = sG[d] * sDl[n * D + d], xh = (sLn[n * D + d] - mean) * rstd;
      s1 += g; s2 += g * xh;
    }
    s1 = warp_sum(s1) / D; s2 = warp_sum(s2) / D;
    for (int d = lane; d < D; d += 32) {
      const float dyv = sDl[n * D + d], xh = (sLn[n * D + d] - mean) * rstd;
      sDl[n * D + d] = rstd * (sG[d] * dyv - s1 - xh * s2);
      sRed[(warp * 2) * D + d] += xh * dyv;
      sRed[(warp * 2 + 1) * D + d] += dyv;
    }
  }
  __syncthreads();
  for (int d = tid; d < 2 * D; d += 256) {
    float v = 0.f;
    for (int w = 0; w < 8; w++) v += sRed[(w * 2 + d / D) * D + d % D];
    atomicAdd((d < D ? e.d_pe_g : e.d_pe_b) + d % D, v);
  }
  for (int o = tid; o < KP * D; o += 256) {          // net/fullconnect.py:118-121
    const int k = o / D, d = o - k * D;
    float acc = 0.f;
    for (int n = 0; n < N; n++) acc = fmaf(sPat[n * KP + k], sDl[n * D + d], acc);
    atomicAdd(e.d_wpe + o, acc);
  }
  for (int d = tid; d < D; d += 256) {
    float acc = 0.f;
    for (int n = 0; n < N; n++) acc += sDl[n * D + d];
    atomicAdd(e.d_bpe + d, acc);
  }
}

// ---------------------------------------------------------------- head
// 4 x 4 register tile GEMM on shared operands:  out(r, c) = sum_k At[k][r] * Bm[k][c]  (both k-major, pitches lda / ldb)
template <class F>
__device__ __forceinline__ void tile_gemm(const float* At, int lda, const float* Bm, int ldb, int R, int Cc, int K, int tid,
                                          int nthreads, F emit) {
  const int r4 = (R + 3) >> 2, c4 = (Cc + 3) >> 2;
  for (int tile = tid; tile < r4 * c4; tile += nthreads) {
    const int r0 = (tile / c4) * 4, c0 = (tile % c4) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    for (int k = 0; k < K; k++) {
      const float4 a = *reinterpret_cast<const float4*>(At + k * lda + r0);
      const float4 bb = *reinterpret_cast<const float4*>(Bm + k * ldb + c0);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (r0 + i < R && c0 + j < Cc) emit(r0 + i, c0 + j, acc[i][j]);
  }
}

__device__ __forceinline__ int pad4(int n) { return (n + 3) & ~3; }

// z = LN(xl) rows (b, t) of one token t for all images, TRANSPOSED into shared memory: zT[d][b] (pitch Bp).
// `stage` ([B][D] floats) receives the rows first, by independent coalesced loads; D % 4 == 0.
__device__ __forceinline__ void token_ln(const NtVitEnds& e, float* zT, float* stage, int Bp, int B, int N, int D, int t,
                                         bool write_stats, bool use_saved, int tid, int nthreads) {
  const int d4 = D >> 2;
  for (int i = tid; i < B * d4; i += nthreads) {
    const int b = i / d4, c = (i - b * d4) * 4;
    *reinterpret_cast<float4*>(stage + b * D + c) = *reinterpret_cast<const float4*>(e.xl + ((size_t)b * N + t) * D + c);
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
  for (int b = warp; b < Bp; b += nwarps) {
    if (b < B) {
      const float* x = stage + b * D;
      float mean, rstd;
      if (use_saved) {
        mean = e.tl_stat[((size_t)b * N + t) * 2]; rstd = e.tl_stat[((size_t)b * N + t) * 2 + 1];
      } else {
        float s = 0.f;
        for (int d = lane; d < D; d += 32) s += x[d];
        mean = warp_sum(s) / D;
        float q = 0.f;
        for (int d = lane; d < D; d += 32) { const float v = x[d] - mean; q += v * v; }
        rstd = 1.0f / sqrtf(warp_sum(q) / D + 1e-5f);
        if (write_stats && lane == 0) { e.tl_stat[((size_t)b * N + t) * 2] = mean; e.tl_stat[((size_t)b * N + t) * 2 + 1] = rstd; }
      }
      for (int d = lane; d < D; d += 32) zT[d * Bp + b] = (x[d] - mean) * rstd;     // xhat; the affine is applied by the caller
    } else {
      for (int d = lane; d < D; d += 32) zT[d * Bp + b] = 0.f;
    }
  }
}

// K3: one CTA per token t: h0[b][j] += sum_d z[b][t][d] W0[t*D + d][j]      (classify.py:22 on the flattened LN output)
__global__ void __launch_bounds__(256) head_fc0_kernel(NtVitEnds e, int B, int N, int D, int D0) {
  pdl_prologue();
  extern __shared__ __align__(16) float sm[];
  const int Bp = pad4(B), D0p = pad4(D0);
  float* zT = sm;                      // [D][Bp]
  float* sW = zT + D * Bp;             // [D][D0p]
  float* stage = sW + D * D0p;         // [B][D]
  const int t = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < D * D0p; i += 256) {           // weight slice in flight before the LayerNorm needs its rows
    const int d = i / D0p, j = i - d * D0p;
    sW[i] = j < D0 ? __ldg(e.w0 + ((size_t)t * D + d) * D0 + j) : 0.f;
  }
  token_ln(e, zT, stage, Bp, B, N, D, t, true, false, tid, 256);
  __syncthreads();
  for (int i = tid; i < D * Bp; i += 256) {            // affine of the final LayerNorm
    const int d = i / Bp;
    zT[i] = (i - d * Bp) < B ? zT[i] * __ldg(e.tl_g + d) + __ldg(e.tl_b + d) : 0.f;
  }
  __syncthreads();
  tile_gemm(zT, Bp, sW, D0p, B, D0, D, tid, 256, [&](int b, int j, float v) { atomicAdd(e.h0 + (size_t)b * D0 + j, v); });
}

// K4: one CTA: ReLU, FC1, softmax / loss / argmax, and the gradients back to d(h0)
__global__ void __launch_bounds__(256) head_mid_kernel(NtVitEnds e, int B, int D0, int C1, float inv_n) {
  pdl_prologue();
  extern __shared__ __align__(16) float sm[];
  float* sH = sm;                      // [B][D0]   relu(h0 + b0)
  float* sW1 = sH + B * D0;            // [D0][C1]
  float* sL = sW1 + D0 * C1;           // [B][C1]   logits -> dlogits
  float* sRow = sL + B * C1;           // [B] per-row loss
  __shared__ float sv[8];
  const int tid = threadIdx.x;
  for (int i = tid; i < D0 * C1; i += 256) sW1[i] = __ldg(e.w1 + i);
  for (int i = tid; i < B * D0; i += 256) {            // loads only: they all go out together
    const float v = e.h0[i] + __ldg(e.b0 + i % D0);
    sH[i] = v > 0.f ? v : 0.f;                         // net/activation.py:11-14
  }
  __syncthreads();
  for (int i = tid; i < B * D0; i += 256) e.h0[i] = 0.f;   // accumulator ready for the next step
  for (int o = tid; o < B * C1; o += 256) {            // classify.py:27
    const int b = o / C1, c = o - b * C1;
    float acc = __ldg(e.b1 + c);
    for (int j = 0; j < D0; j++) acc = fmaf(sH[b * D0 + j], sW1[j * C1 + c], acc);
    sL[o] = acc;
    e.logits[o] = acc;
  }
  __syncthreads();
  for (int b = tid; b < B; b += 256) {                 // net/loss.py:46-51 + argmax of the returned softmax
    float mx = -INFINITY;
    for (int c = 0; c < C1; c++) mx = fmaxf(mx, sL[b * C1 + c]);
    float s = 0.f;
    for (int c = 0; c < C1; c++) s += expf(sL[b * C1 + c] - mx);
    const float inv = 1.0f / s;
    const int lab = e.labels[b];
    float pm = -INFINITY, loss = 0.f;
    int pi = 0;
    for (int c = 0; c < C1; c++) {
      const float p = expf(sL[b * C1 + c] - mx) * inv;
      e.probs[b * C1 + c] = p;
      if (p > pm) { pm = p; pi = c; }                  // first index wins ties, like np.argmax
      const float y = c == lab ? 1.f : 0.f;
      if (c == lab) loss = -logf(p + 1e-10f);
      sL[b * C1 + c] = (p - y) * inv_n;                // dlogits
    }
    e.argmax[b] = pi;
    sRow[b] = loss;
  }
  __syncthreads();
  {                                                    // deterministic block reduction of the row losses
    float s = 0.f;
    for (int b = tid; b < B; b += 256) s += sRow[b];
    s = warp_sum(s);
    if ((tid & 31) == 0) sv[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
      float tot = 0.f;
      for (int w = 0; w < 8; w++) tot += sv[w];
      *e.loss = tot * inv_n;
    }
  }
  for (int o = tid; o < D0 * C1; o += 256) {           // dW1 += h^T dlogits   (fullconnect.py:118-121)
    const int j = o / C1, c = o - j * C1;
    float acc = 0.f;
    for (int b = 0; b < B; b++) acc = fmaf(sH[b * D0 + j], sL[b * C1 + c], acc);
    e.d_w1[o] += acc;
  }
  for (int c = tid; c < C1; c += 256) {
    float acc = 0.f;
    for (int b = 0; b < B; b++) acc += sL[b * C1 + c];
    e.d_b1[c] += acc;
  }
  for (int o = tid; o < B * D0; o += 256) {            // d(h0) = (dlogits W1^T) * relu'   (classify.py:40-43)
    const int b = o / D0, j = o - b * D0;
    float acc = 0.f;
    for (int c = 0; c < C1; c++) acc = fmaf(sL[b * C1 + c], sW1[j * C1 + c], acc);
    const float v = sH[o] > 0.f ? acc : 0.f;
    e.dh0[o] = v;
    sH[o] = v;                                         // each element is read and overwritten by its own thread
  }
  __syncthreads();
  for (int j = tid; j < D0; j += 256) {                // db0 += colsum(dh0)
    float acc = 0.f;
    for (int b = 0; b < B; b++) acc += sH[b * D0 + j];
    e.d_b0[j] += acc;
  }
}

// K5: one CTA per token t: dz = dh0 W0_t^T, dW0_t += z^T dh0, then the final LayerNorm backward of rows (b, t)
__global__ void __launch_bounds__(256) head_bwd_kernel(NtVitEnds e, int B, int N, int D, int D0) {
  pdl_prologue();
  extern __shared__ __align__(16) float sm[];
  const int Bp = pad4(B), Dp = pad4(D), D0p = pad4(D0);
  float* zT = sm;                      // [D][Bp]     xhat, then z
  float* sWt = zT + D * Bp;            // [D0][Dp]    W0_t transposed: sWt[j][d]
  float* sDhT = sWt + D0 * Dp;         // [D0][Bp]    dh0 transposed
  float* sDh = sDhT + D0 * Bp;         // [B][D0p]    dh0
  float* sDz = sDh + B * D0p;          // [B][Dp]
  float* sRed = sDz + B * Dp;          // [8][2][D]
  const int t = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < D0 * Dp; i += 256) {
    const int j = i / Dp, d = i - j * Dp;
    sWt[i] = d < D ? __ldg(e.w0 + ((size_t)t * D + d) * D0 + j) : 0.f;
  }
  for (int i = tid; i < D0 * Bp; i += 256) {
    const int j = i / Bp, b = i - j * Bp;
    sDhT[i] = b < B ? e.dh0[(size_t)b * D0 + j] : 0.f;
  }
  for (int i = tid; i < B * D0p; i += 256) {
    const int b = i / D0p, j = i - b * D0p;
    sDh[i] = j < D0 ? e.dh0[(size_t)b * D0 + j] : 0.f;
  }
  for (int i = tid; i < 16 * D; i += 256) sRed[i] = 0.f;
  token_ln(e, zT, sDz, Bp, B, N, D, t, false, true, tid, 256);      // rows staged in sDz (free until the first GEMM)
  __syncthreads();
  // dz[b][d] = sum_j dh0[b][j] W0[t*D+d][j]          (fullconnect.py:114)
  tile_gemm(sDhT, Bp, sWt, Dp, B, D, D0, tid, 256, [&](int b, int d, float v) { sDz[b * Dp + d] = v; });
  __syncthreads();
  // final LayerNorm backward on rows (b, t) with xhat still in zT   (net/layernorm.py:126-133)
  float* dxl = e.dxl;
  for (int b = warp; b < B; b += 8) {
    const float rstd = e.tl_stat[((size_t)b * N + t) * 2 + 1];
    float s1 = 0.f, s2 = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float g = __ldg(e.tl_g + d) * sDz[b * Dp + d], xh = zT[d * Bp + b];
      s1 += g; s2 += g * xh;
    }
    s1 = warp_sum(s1) / D; s2 = warp_sum(s2) / D;
    for (int d = lane; d < D; d += 32) {
      const float dzv = sDz[b * Dp + d], xh = zT[d * Bp + b];
      dxl[((size_t)b * N + t) * D + d] = rstd * (__ldg(e.tl_g + d) * dzv - s1 - xh * s2);
      sRed[(warp * 2) * D + d] += xh * dzv;
      sRed[(warp * 2 + 1) * D + d] += dzv;
    }
  }
  __syncthreads();
  for (int d = tid; d < 2 * D; d += 256) {
    float v = 0.f;
    for (int w = 0; w < 8; w++) v += sRed[(w * 2 + d / D) * D + d % D];
    atomicAdd((d < D ? e.d_tl_g : e.d_tl_b) + d % D, v);
  }
  for (int i = tid; i < D * Bp; i += 256) {            // xhat -> z = xhat * gamma + beta (operand of dW0)
    const int d = i / Bp;
    zT[i] = (i - d * Bp) < B ? zT[i] * __ldg(e.tl_g + d) + __ldg(e.tl_b + d) : 0.f;
  }
  __syncthreads();
  // dW0[t*D + d][j] += sum_b z[b][d] dh0[b][j]        (fullconnect.py:118-121); each element has exactly one owner
  float* dw = e.d_w0 + (size_t)t * D * D0;
  // operands k-major over b: A = z as [b][d] is zT transposed -> use sDz's space?  zT is [d][b]; tile_gemm wants
  // At[k][r] with k = b: build it from zT on the fly through sDz (dz is consumed)
  for (int i = tid; i < B * Dp; i += 256) {
    const int b = i / Dp, d = i - b * Dp;
    sDz[i] = d < D ? zT[d * Bp + b] : 0.f;
  }
  __syncthreads();
  tile_gemm(sDz, Dp, sDh, D0p, D, D0, B, tid, 256, [&](int d, int j, float v) { dw[(size_t)d * D0 + j] += v; });
}

static size_t stem_fwd_smem(int N, int KP, int D) { return sizeof(float) * ((size_t)N * KP + (size_t)KP * D + 2 * (size_t)N * D + 2 * (size_t)D); }
static size_t stem_bwd_smem(int N, int KP, int D) { return sizeof(float) * ((size_t)N * KP + 2 * (size_t)N * D + 17 * (size_t)D + 2 * (size_t)N); }
static int p4(int n) { return (n + 3) & ~3; }
static size_t fc0_smem(int B, int D, int D0) { return sizeof(float) * ((size_t)D * p4(B) + (size_t)D * p4(D0) + (size_t)B * D); }
static size_t mid_smem(int B, int D0, int C1) { return sizeof(float) * ((size_t)B * D0 + (size_t)D0 * C1 + (size_t)B * C1 + B); }
static size_t bwd_smem(int B, int D, int D0) {
  return sizeof(float) * ((size_t)D * p4(B) + (size_t)D0 * p4(D) + (size_t)D0 * p4(B) + (size_t)B * p4(D0) + (size_t)B * p4(D) + 16 * (size_t)D);
}
static const size_t SMEM_MAX = 227 * 1024;

template <typename K>
static int opt_in(K kernel, size_t smem) {
  if (smem > 48 * 1024) NT_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
  return 0;
}

}  // namespace ve
}  // namespace nt
using namespace nt;

extern "C" {

int nt_vit_ends_supported(int B, int C, int Hi, int Wi, int n_patch, int D, int D0, int C1, int* ok) {
  *ok = 0;
  if (B < 1 || n_patch < 1 || Hi % n_patch || Wi % n_patch) return 0;
  const int N = n_patch * n_patch, KP = C * (Hi / n_patch) * (Wi / n_patch);
  if (ve::stem_fwd_smem(N, KP, D) > ve::SMEM_MAX || ve::stem_bwd_smem(N, KP, D) > ve::SMEM_MAX) return 0;
  if (ve::fc0_smem(B, D, D0) > ve::SMEM_MAX || ve::mid_smem(B, D0, C1) > ve::SMEM_MAX || ve::bwd_smem(B, D, D0) > ve::SMEM_MAX) return 0;
  if (D % 4 || D0 % 4) return 0;
  *ok = 1;
  return 0;
}

int nt_vit_stem_fwd(const NtVitEnds* e, int B, int C, int Hi, int Wi, int n_patch, int D) {
  NT_ENTRY();
  NT_REQUIRE(B >= 0 && n_patch > 0 && Hi % n_patch == 0 && Wi % n_patch == 0, "nt_vit_stem_fwd: bad shape");
  if (B == 0) return 0;
  const int N = n_patch * n_patch, KP = C * (Hi / n_patch) * (Wi / n_patch);
  const size_t smem = ve::stem_fwd_smem(N, KP, D);
  NT_REQUIRE(smem <= ve::SMEM_MAX, "nt_vit_stem_fwd: shape needs %zu bytes of shared memory", smem);
  if (ve::opt_in(ve::stem_fwd_kernel, smem)) return 1;
  NT_CHECK(launch_k(ve::stem_fwd_kernel, dim3(B), dim3(256), smem, *e, C, Hi, Wi, n_patch, D));
  NT_LAUNCH_OK();
  return 0;
}

int nt_vit_stem_bwd(const NtVitEnds* e, const float* dx0, int B, int C, int Hi, int Wi, int n_patch, int D) {
  NT_ENTRY();
  NT_REQUIRE(B >= 0 && n_patch > 0 && Hi % n_patch == 0 && Wi % n_patch == 0, "nt_vit_stem_bwd: bad shape");
  if (B == 0) return 0;
  const int N = n_patch * n_patch, KP = C * (Hi / n_patch) * (Wi / n_patch);
  const size_t smem = ve::stem_bwd_smem(N, KP, D);
  NT_REQUIRE(smem <= ve::SMEM_MAX, "nt_vit_stem_bwd: shape needs %zu bytes of shared memory", smem);
  if (ve::opt_in(ve::stem_bwd_kernel, smem)) return 1;
  NT_CHECK(launch_k(ve::stem_bwd_kernel, dim3(B), dim3(256), smem, *e, dx0, C, Hi, Wi, n_patch, D));
  NT_LAUNCH_OK();
  return 0;
}

int nt_vit_head(const NtVitEnds* e, int B, int N, int D, int D0, int C1, float n_norm) {
  NT_ENTRY();
  NT_REQUIRE(B > 0 && N > 0 && D > 0 && D0 > 0 && C1 > 0 && n_norm > 0.f && D % 4 == 0 && D0 % 4 == 0, "nt_vit_head: bad shape");
  const size_t s0 = ve::fc0_smem(B, D, D0), s1 = ve::mid_smem(B, D0, C1), s2 = ve::bwd_smem(B, D, D0);
  NT_REQUIRE(s0 <= ve::SMEM_MAX && s1 <= ve::SMEM_MAX && s2 <= ve::SMEM_MAX, "nt_vit_head: batch %d / widths %d, %d do not fit shared memory", B, D, D0);
  if (ve::opt_in(ve::head_fc0_kernel, s0) || ve::opt_in(ve::head_mid_kernel, s1) || ve::opt_in(ve::head_bwd_kernel, s2)) return 1;
  NT_CHECK(launch_k(ve::head_fc0_kernel, dim3(N), dim3(256), s0, *e, B, N, D, D0));
  NT_LAUNCH_OK();
  NT_CHECK(launch_k(ve::head_mid_kernel, dim3(1), dim3(256), s1, *e, B, D0, C1, 1.0f / n_norm));
  NT_LAUNCH_OK();
  NT_CHECK(launch_k(ve::head_bwd_kernel, dim3(N), dim3(256), s2, *e, B, N, D, D0));
  NT_LAUNCH_OK();
  return 0;
}

}
