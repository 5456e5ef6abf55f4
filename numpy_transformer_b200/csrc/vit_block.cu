// Fused ViT encoder blocks (attention.py:20-89) for the launch-bound shapes (N <= 64 tokens, D in {32,64,96}).
//
// One CTA per image walks ALL blocks of the stack: the residual stream of an image never depends on another
// image, so the forward of a whole stack (and its backward) is a single launch instead of ~9 launches per
// block.  Inside a CTA (8 warps):
//   * every contraction is mma.sync.m16n8k16 BF16 with fp32 accumulation, error-compensated: each fp32 operand
//     is split x = hi + lo (two bf16) and a product is lo*hi + hi*lo + hi*hi (fp32-class accuracy, ~1e-5
//     relative; tools/mma_red_bench.cu measures the legacy tensor path at 0.92 mma/ns/SM for both TF32 k8 and
//     BF16 k16, so the BF16 split does the same work in half the instructions of a TF32 split);
//   * activations that feed an MMA live in shared memory already split (hi/lo bf16, row pitch = 16 mod 128 bytes
//     so every ldmatrix phase is conflict-free); fragments come from ldmatrix(.trans), never scalar LDS;
//   * weights are pre-split once per step into mma-B-fragment order (vit_wsplit_kernel) so a lane fetches the
//     hi and lo fragments of one (k16, n8) tile with ONE coalesced 16-byte load, straight into registers;
//   * fp32 tensors that feed row-wise math (residual stream, LayerNorm inputs) go through global memory (L2/L1
//     resident, written and re-read by the same CTA) — they are the tensors the backward needs anyway;
//   * weight gradients are contracted over the image's tokens in the CTA and accumulated with red.global.add.v2
//     (the reference's `+=` into *_delta until setzero).
// Tokens are padded to 64 rows (4 m16 tiles); padded rows are kept exactly zero in every shared operand.
#include "mma_bf16x3.cuh"

namespace nt {
namespace vb {

constexpr int MAXB = 16;      // blocks per launch (kernel parameter space)

struct Args {
  NtVitBlock blk[MAXB];
  int nblocks;
  int N;
  const float* x;       // [B,N,D] input of block 0
  const float* dout;    // [B,N,D] gradient w.r.t. the last block's output      (backward)
  float* dx;            // [B,N,D] gradient w.r.t. x                              (backward)
  float* dh;            // [B,N,D] scratch: gradient w.r.t. h                     (backward)
  float* tmp;           // [B,N,D] scratch: LayerNorm-backward input              (backward)
};

// ---------------------------------------------------------------- async bulk copies (TMA, 1-D) + mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 24); ++it) {     // bounded: a protocol bug traps instead of hanging the GPU
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global (bulk async-group of the issuing thread)
__device__ __forceinline__ void bulk_store(void* dst, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// GELU (exact-erf form, net/activation.py:139-148) and its derivative from one shared exponential.
// erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7): erf(s) = 1 - (a1 t + .. + a5 t^5) exp(-s^2), t = 1/(1 + p|s|).
__device__ __forceinline__ void gelu_both(float x, float& y, float& dy) {
  const float s = x * 0.70710678118654752f;
  const float e = __expf(-s * s);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, fabsf(s), 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float er = copysignf(fmaf(-poly * t, e, 1.0f), s);
  const float cdf = fmaf(0.5f, er, 0.5f);
  y = x * cdf;
  dy = fmaf(x * 0.3989422804014327f, e, cdf);
}

// interleaved butterfly reductions of R independent values (ILP across the rows a warp owns)
template <int R>
__device__ __forceinline__ void warp_sum_multi(float (&v)[R]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int i = 0; i < R; i++) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
}

// LayerNorm of rows of a global [N][D] fp32 tensor into a split shared operand (net/layernorm.py:100-114);
// rows >= N are zero-filled.  A warp owns rows warp, warp+8, ...; all of them are loaded and reduced together.
// mean / rstd go to shared memory when requested (the backward formula needs them).
template <int D>
__device__ __forceinline__ void ln_rows(const float* src, int src_ld, const float* __restrict__ gamma,
                                        const float* __restrict__ beta, const SArr& dst, int N, float* s_mean,
                                        float* s_rstd, int warp, int lane) {
  constexpr int E = D / 32, RPW = ROWS / 8;
  float gm[E], bt[E], v[RPW][E], s[RPW], q[RPW];
#pragma unroll
  for (int e = 0; e < E; e++) { gm[e] = __ldg(gamma + lane + 32 * e); bt[e] = __ldg(beta + lane + 32 * e); }
#pragma unroll
  for (int i = 0; i < RPW; i++) {
    const int r = warp + 8 * i;
    s[i] = 0.f;
#pragma unroll
    for (int e = 0; e < E; e++) { v[i][e] = r < N ? src[r * src_ld + lane + 32 * e] : 0.f; s[i] += v[i][e]; }
  }
  warp_sum_multi<RPW>(s);
#pragma unroll
  for (int i = 0; i < RPW; i++) {
    s[i] *= (1.0f / D);
    q[i] = 0.f;
#pragma unroll
    for (int e = 0; e < E; e++) { v[i][e] -= s[i]; q[i] += v[i][e] * v[i][e]; }
  }
  warp_sum_multi<RPW>(q);
#pragma unroll
  for (int i = 0; i < RPW; i++) {
    const int r = warp + 8 * i;
    const float rstd = rsqrtf(q[i] * (1.0f / D) + 1e-5f);
    if (s_mean && lane == 0) { s_mean[r] = s[i]; s_rstd[r] = rstd; }
#pragma unroll
    for (int e = 0; e < E; e++) st1(dst, r, lane + 32 * e, r < N ? v[i][e] * rstd * gm[e] + bt[e] : 0.f);
  }
}

// dx = addend + rstd*(g*dy - mean(g*dy) - xhat*mean(g*dy*xhat))  (net/layernorm.py:126-128) per row; the result
// goes to global fp32 (`out`) and, split, to `dst`.  dgamma/dbeta (and optionally the column sums of `addend`,
// i.e. the bias gradient of the layer whose output gradient `addend` is) are accumulated per lane over the
// warp's rows and flushed with one atomic per column per warp.
template <int D>
__device__ __forceinline__ void ln_bwd_rows(const float* dy, int dy_ld, const float* xin, const float* __restrict__ gamma,
                                            const float* addend, float* out, const SArr& dst, int N,
                                            const float* s_mean, const float* s_rstd, float* s_red, int warp, int lane) {
  constexpr int E = D / 32, RPW = ROWS / 8;
  float gm[E], ag[E], ab[E], aa[E];
#pragma unroll
  for (int e = 0; e < E; e++) { gm[e] = __ldg(gamma + lane + 32 * e); ag[e] = ab[e] = aa[e] = 0.f; }
  float g[RPW][E], xh[RPW][E], va[RPW][E], s1[RPW], s2[RPW], rs[RPW];
#pragma unroll
  for (int i = 0; i < RPW; i++) {
    const int r = warp + 8 * i;
#pragma unroll
    for (int e = 0; e < E; e++) {
      const int o = r * D + lane + 32 * e;
      g[i][e] = r < N ? dy[r * dy_ld + lane + 32 * e] : 0.f;
      xh[i][e] = r < N ? xin[o] : 0.f;
      va[i][e] = r < N ? addend[o] : 0.f;
    }
  }
#pragma unroll
  for (int i = 0; i < RPW; i++) {
    const int r = warp + 8 * i;
    const float mu = s_mean[r];
    rs[i] = s_rstd[r];
    s1[i] = s2[i] = 0.f;
#pragma unroll
    for (int e = 0; e < E; e++) {
      const float d = g[i][e];
      xh[i][e] = r < N ? (xh[i][e] - mu) * rs[i] : 0.f;
      ag[e] += xh[i][e] * d;
      ab[e] += d;
      aa[e] += va[i][e];
      g[i][e] = gm[e] * d;
      s1[i] += g[i][e];
      s2[i] += g[i][e] * xh[i][e];
    }
  }
  warp_sum_multi<RPW>(s1);
  warp_sum_multi<RPW>(s2);
#pragma unroll
  for (int i = 0; i < RPW; i++) {
    const int r = warp + 8 * i;
    const float m1 = s1[i] * (1.0f / D), m2 = s2[i] * (1.0f / D);
#pragma unroll
    for (int e = 0; e < E; e++) {
      const float o = va[i][e] + rs[i] * (g[i][e] - m1 - xh[i][e] * m2);
      if (r < N) out[r * D + lane + 32 * e] = o;
      st1(dst, r, lane + 32 * e, r < N ? o : 0.f);
    }
  }
  // per-warp column partials -> shared memory [8 warps][3][D]; ln_bwd_flush() adds them up after the next barrier
#pragma unroll
  for (int e = 0; e < E; e++) {
    float* p = s_red + warp * 3 * D + lane + 32 * e;
    p[0] = ag[e];
    p[D] = ab[e];
    p[2 * D] = aa[e];
  }
}

// one atomic per column per CTA: dgamma += sum xhat*dy, dbeta += sum dy, and (optionally) the bias gradient of the
// layer whose output gradient was the addend (net/layernorm.py:130-133, net/fullconnect.py:122)
template <int D>
__device__ __forceinline__ void ln_bwd_flush(const float* s_red, float* dgamma, float* dbeta, float* dbias, int tid) {
  for (int c = tid; c < 3 * D; c += 256) {
    float* dst = c < D ? dgamma + c : (c < 2 * D ? dbeta + (c - D) : (dbias ? dbias + (c - 2 * D) : nullptr));
    if (!dst) continue;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) v += s_red[w * 3 * D + c];
    atomicAdd(dst, v);
  }
}

// column sums of the two fragment rows held by the 8 `g` lanes -> lanes with g == 0 hold the totals
__device__ __forceinline__ float colsum_g(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}

template <int D, int H>
struct Cfg {
  static constexpr int DH = D / H;
  static constexpr int LD = D + 8;          // pitch of [64][D] operands: 2*LD bytes = 16 (mod 128) or 80 (mod 128)
  static constexpr int LD2 = 2 * D + 8;
  static constexpr int ARR = ROWS * LD * 2 * 2;        // bytes of one [64][D] hi/lo pair
  static constexpr int ARR2 = ROWS * LD2 * 2 * 2;
  static constexpr int PARR = ROWS * LDP * 2 * 2;      // one [64][64] hi/lo pair
  static constexpr int PW = 16 * LDP * 2 * 2;          // one warp's [16][64] hi/lo pair
  // weight-fragment scratch, in uint4 units: forward order (qkv, fc0, fc1) then backward order
  static constexpr int F_QKV = 0;
  static constexpr int F_FC0 = F_QKV + 3 * D * D / 4;
  static constexpr int F_FC1 = F_FC0 + 2 * D * D / 4;
  static constexpr int B_QKV = F_FC1 + 2 * D * D / 4;
  static constexpr int B_FC0 = B_QKV + 3 * D * D / 4;
  static constexpr int B_FC1 = B_FC0 + 2 * D * D / 4;
  static constexpr int F_TOTAL = B_FC1 + 2 * D * D / 4;     // = 3.5 D^2 uint4 = 8 bytes per weight
  // per-image operand save area written by the forward, bulk-loaded by the backward (bytes):
  // split LN(x) | split q,k,v | split LN1(h) | split GELU(pre2) | mean, rstd of LN and LN1
  static constexpr int SV_U = 0;
  static constexpr int SV_QKV = SV_U + ARR;
  static constexpr int SV_U1 = SV_QKV + 3 * ARR;
  static constexpr int SV_G = SV_U1 + ARR;
  static constexpr int SV_STAT = SV_G + ARR2;
  static constexpr int SV_BYTES = SV_STAT + 4 * ROWS * (int)sizeof(float);
  static constexpr int XLD = D + 8;                     // pitch (floats) of the fp32 residual / LN-input buffers
  static constexpr int XARR = ROWS * XLD * 4;
  static constexpr size_t FWD_SMEM = 4 * (size_t)ARR + 4 * ROWS * sizeof(float) + XARR;
  static constexpr size_t BWD_SMEM = 4 * (size_t)ARR + 4 * (size_t)PARR + 4 * ROWS * sizeof(float) + 16 + XARR + 8 * 3 * D * sizeof(float);
  static_assert(D % 32 == 0 && DH % 8 == 0 && DH * H == D, "fused ViT block: D % 32 == 0, dh % 8 == 0");
  static_assert(ARR2 <= 2 * ARR && ARR2 <= 4 * PARR && ARR <= 4 * PARR, "aliasing plan");
  static_assert(ARR % 16 == 0 && ARR2 % 16 == 0 && SV_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");
};

// ---------------------------------------------------------------- weight pre-split
// W [K][Nw] row-major fp32 -> B fragments.  fwd order: B(k, n) = W[k][n]; bwd order: B(k, n) = W[n][k].
__device__ __forceinline__ void wsplit_entry(const float* __restrict__ W, int ldw, bool bwd, int KS, int NT, int e,
                                             uint4* out) {
  const int lane = e & 31, tile = e >> 5, nt = tile % NT, ks = tile / NT;
  const int g = lane >> 2, t = lane & 3;
  float v[4];
  if (!bwd) {
    const int n = 8 * nt + g, k0 = 16 * ks + 2 * t;
    v[0] = W[(size_t)k0 * ldw + n]; v[1] = W[(size_t)(k0 + 1) * ldw + n];
    v[2] = W[(size_t)(k0 + 8) * ldw + n]; v[3] = W[(size_t)(k0 + 9) * ldw + n];
  } else {
    const float* r = W + (size_t)(8 * nt + g) * ldw + 16 * ks + 2 * t;
    v[0] = r[0]; v[1] = r[1]; v[2] = r[8]; v[3] = r[9];
  }
  uint4 o;
  split2(v[0], v[1], o.x, o.z);
  split2(v[2], v[3], o.y, o.w);
  out[e] = o;
  (void)KS;
}

template <int D, int H>
__global__ void __launch_bounds__(256) vit_wsplit_kernel(Args a) {
  pdl_prologue();
  using C = Cfg<D, H>;
  const NtVitBlock& P = a.blk[blockIdx.y];
  uint4* wf = reinterpret_cast<uint4*>(P.wfrag);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < C::F_TOTAL; e += gridDim.x * blockDim.x) {
    if (e < C::F_FC0) wsplit_entry(P.wqkv, 3 * D, false, D / 16, 3 * D / 8, e - C::F_QKV, wf + C::F_QKV);
    else if (e < C::F_FC1) wsplit_entry(P.w0, 2 * D, false, D / 16, 2 * D / 8, e - C::F_FC0, wf + C::F_FC0);
    else if (e < C::B_QKV) wsplit_entry(P.w1, D, false, 2 * D / 16, D / 8, e - C::F_FC1, wf + C::F_FC1);
    else if (e < C::B_FC0) wsplit_entry(P.wqkv, 3 * D, true, 3 * D / 16, D / 8, e - C::B_QKV, wf + C::B_QKV);
    else if (e < C::B_FC1) wsplit_entry(P.w0, 2 * D, true, 2 * D / 16, D / 8, e - C::B_FC0, wf + C::B_FC0);
    else wsplit_entry(P.w1, D, true, D / 16, 2 * D / 8, e - C::B_FC1, wf + C::B_FC1);
  }
}

// ---------------------------------------------------------------- forward
template <int D, int H>
__global__ void __launch_bounds__(256, 1) vit_fwd_kernel(Args a) {
  pdl_prologue();
  using C = Cfg<D, H>;
  constexpr int DH = C::DH;
  extern __shared__ __align__(128) char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;
  const int N = a.N, b = blockIdx.x;
  const SArr sU = make_arr(smem, C::LD);
  const SArr sQ = make_arr(smem + C::ARR, C::LD);
  const SArr sK = make_arr(smem + 2 * C::ARR, C::LD);
  const SArr sV = make_arr(smem + 3 * C::ARR, C::LD);
  const SArr sG = make_arr(smem + C::ARR, C::LD2);                       // aliases sQ|sK after attention
  float* s_stat = reinterpret_cast<float*>(smem + 4 * C::ARR);   // mean, rstd (LN) | mean, rstd (LN1)
  // token contractions always run over the 64 padded rows / keys (pad rows are exact zeros, padded keys are masked):
  // compile-time trip counts let the compiler software-pipeline the fragment loads
  constexpr int NPAIR = ROWS / 16, KT = ROWS / 16;
  const float scale = rsqrtf((float)DH);
  const size_t img = (size_t)b * N * D;
  // fp32 residual stream of this image: x -> h -> out in place (every element is updated by the thread that read it)
  float* sX = reinterpret_cast<float*>(smem + 4 * C::ARR + 4 * ROWS * sizeof(float));
  for (int e = tid; e < ROWS * (D / 4); e += 256) {
    const int r = e / (D / 4), c4 = (e - r * (D / 4)) * 4;
    *reinterpret_cast<float4*>(sX + r * C::XLD + c4) =
        r < N ? __ldg(reinterpret_cast<const float4*>(a.x + img + r * D + c4)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();

  for (int bi = 0; bi < a.nblocks; bi++) {
    const NtVitBlock& P = a.blk[bi];
    const uint4* wf = reinterpret_cast<const uint4*>(P.wfrag);
    char* sv = reinterpret_cast<char*>(P.save) + (size_t)b * C::SV_BYTES;
    uint4 bq0[3 * D / 32];
    wf_first(bq0, wf + C::F_QKV + (wn * (3 * D / 32)) * 32 + lane);
    // ---- u = LN(x)                                                      attention.py:22
    ln_rows<D>(sX, C::XLD, P.ln_g, P.ln_b, sU, N, s_stat, s_stat + ROWS, warp, lane);
    fence_async_smem();
    __syncthreads();
    if (tid == 0) bulk_store(sv + C::SV_U, sU.ptr, C::ARR);               // operand of dWqkv in the backward
    // ---- qkv = u Wqkv + b                                               attention.py:23-24
    {
      constexpr int NTW = 3 * D / 32;
      float acc[2][NTW][4];
      zero_acc(acc);
      const uint32_t a_lane = sU.addr + (wm * 32 + (lane & 15)) * sU.ldb + ((lane >> 4) * 8) * 2;
      gemm_wf<2, NTW, D / 16, 1>(acc, [&](int ks) { return a_lane + ks * 32; }, sU.lo, 16 * sU.ldb,
                                 wf + C::F_QKV + (wn * NTW) * 32 + lane, (3 * D / 8) * 32, bq0);
      float* qg = P.qkv + (size_t)b * N * 3 * D;
      float2 bb[NTW];
#pragma unroll
      for (int nt = 0; nt < NTW; nt++) bb[nt] = __ldg(reinterpret_cast<const float2*>(P.bqkv + (wn * NTW + nt) * 8 + 2 * t));
#pragma unroll
      for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < NTW; nt++) {
          const int col = (wn * NTW + nt) * 8 + 2 * t;
          const int which = col / D, cc = col - which * D;
          const SArr& dst = which == 0 ? sQ : (which == 1 ? sK : sV);
#pragma unroll
          for (int hf = 0; hf < 2; hf++) {
            const int row = wm * 32 + mt * 16 + g + hf * 8;
            float v0 = acc[mt][nt][2 * hf] + bb[nt].x, v1 = acc[mt][nt][2 * hf + 1] + bb[nt].y;
            if (row < N) *reinterpret_cast<float2*>(qg + (size_t)row * 3 * D + col) = make_float2(v0, v1);
            else v0 = v1 = 0.f;
            st2(dst, row, cc, v0, v1);
          }
        }
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) bulk_store(sv + C::SV_QKV, sQ.ptr, 3 * C::ARR);         // q, k, v operands of the attention backward
    // ---- attention, one (head, 16-query tile) per warp pass              attention.py:31-45
    for (int item = warp; item < 4 * H; item += 8) {
      const int hd = item >> 2, i0 = (item & 3) * 16;
      if (i0 >= N) continue;
      float s[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; nt++) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      gemm_rowk_nk<DH>(s, sQ, i0, sK, hd * DH, NPAIR, lane);
      const int r0 = i0 + g, r1 = r0 + 8;
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; nt++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const int col = nt * 8 + 2 * t + (c & 1);
          const float v = col < N ? s[nt][c] * scale : -INFINITY;          // attention.py:37; padded keys never win
          s[nt][c] = v;
          if (c < 2) mx0 = fmaxf(mx0, v); else mx1 = fmaxf(mx1, v);
        }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; nt++) {
        s[nt][0] = __expf(s[nt][0] - mx0); s[nt][1] = __expf(s[nt][1] - mx0);
        s[nt][2] = __expf(s[nt][2] - mx1); s[nt][3] = __expf(s[nt][3] - mx1);
        s0 += s[nt][0] + s[nt][1];
        s1 += s[nt][2] + s[nt][3];
      }
      s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
      const float inv0 = r0 < N ? __fdividef(1.0f, s0) : 0.f, inv1 = r1 < N ? __fdividef(1.0f, s1) : 0.f;
      float* pg = P.P + ((size_t)b * H + hd) * N * N;
#pragma unroll
      for (int nt = 0; nt < 8; nt++) {
        const int col = nt * 8 + 2 * t;
        s[nt][0] *= inv0; s[nt][1] *= inv0; s[nt][2] *= inv1; s[nt][3] *= inv1;
        if (r0 < N) { if (col < N) pg[r0 * N + col] = s[nt][0]; if (col + 1 < N) pg[r0 * N + col + 1] = s[nt][1]; }   // atg__
        if (r1 < N) { if (col < N) pg[r1 * N + col] = s[nt][2]; if (col + 1 < N) pg[r1 * N + col + 1] = s[nt][3]; }
      }
      float o[DH / 8][4];
#pragma unroll
      for (int nd = 0; nd < DH / 8; nd++) o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f;
      gemm_regA_kn<DH>(o, s, sV, hd * DH, KT, lane);     // P feeds the MMA straight from its accumulator fragments
      float* hg = P.h + img;
      float2 xr0[DH / 8], xr1[DH / 8];
#pragma unroll
      for (int nd = 0; nd < DH / 8; nd++) {
        const int col = hd * DH + nd * 8 + 2 * t;
        xr0[nd] = *reinterpret_cast<const float2*>(sX + r0 * C::XLD + col);
        xr1[nd] = *reinterpret_cast<const float2*>(sX + r1 * C::XLD + col);
      }
#pragma unroll
      for (int nd = 0; nd < DH / 8; nd++) {
        const int col = hd * DH + nd * 8 + 2 * t;
        const float2 h0 = make_float2(xr0[nd].x + o[nd][0], xr0[nd].y + o[nd][1]);                 // attention.py:46
        const float2 h1 = make_float2(xr1[nd].x + o[nd][2], xr1[nd].y + o[nd][3]);
        if (r0 < N) { *reinterpret_cast<float2*>(sX + r0 * C::XLD + col) = h0; *reinterpret_cast<float2*>(hg + r0 * D + col) = h0; }
        if (r1 < N) { *reinterpret_cast<float2*>(sX + r1 * C::XLD + col) = h1; *reinterpret_cast<float2*>(hg + r1 * D + col) = h1; }
      }
      __syncwarp();
    }
    if (tid == 0) bulk_store_wait_read();        // the bulk stores have finished READING sU / sQ|sK|sV
    uint4 bf0[2 * D / 32];
    wf_first(bf0, wf + C::F_FC0 + (wn * (2 * D / 32)) * 32 + lane);
    __syncthreads();
    // ---- u1 = LN1(h)                                                     attention.py:48
    ln_rows<D>(sX, C::XLD, P.ln1_g, P.ln1_b, sU, N, s_stat + 2 * ROWS, s_stat + 3 * ROWS, warp, lane);
    fence_async_smem();
    __syncthreads();
    if (tid == 0) bulk_store(sv + C::SV_U1, sU.ptr, C::ARR);              // operand of dW0 in the backward
    // ---- pre2 = u1 W0 + b0 ; G = GELU(pre2)                              attention.py:49-50
    uint4 bg0[D / 32];
    {
      constexpr int NTW = 2 * D / 32;
      float acc[2][NTW][4];
      zero_acc(acc);
      const uint32_t a_lane = sU.addr + (wm * 32 + (lane & 15)) * sU.ldb + ((lane >> 4) * 8) * 2;
      gemm_wf<2, NTW, D / 16, 2>(acc, [&](int ks) { return a_lane + ks * 32; }, sU.lo, 16 * sU.ldb,
                                 wf + C::F_FC0 + (wn * NTW) * 32 + lane, (2 * D / 8) * 32, bf0);
      wf_first(bg0, wf + C::F_FC1 + (wn * (D / 32)) * 32 + lane);
      float* dg = P.dgelu + (size_t)b * N * 2 * D;
      float2 bb[NTW];
#pragma unroll
      for (int nt = 0; nt < NTW; nt++) bb[nt] = __ldg(reinterpret_cast<const float2*>(P.b0 + (wn * NTW + nt) * 8 + 2 * t));
#pragma unroll
      for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < NTW; nt++) {
          const int col = (wn * NTW + nt) * 8 + 2 * t;
#pragma unroll
          for (int hf = 0; hf < 2; hf++) {
            const int row = wm * 32 + mt * 16 + g + hf * 8;
            float y0 = 0.f, y1 = 0.f;
            if (row < N) {
              float d0, d1;
              gelu_both(acc[mt][nt][2 * hf] + bb[nt].x, y0, d0);
              gelu_both(acc[mt][nt][2 * hf + 1] + bb[nt].y, y1, d1);
              *reinterpret_cast<float2*>(dg + (size_t)row * 2 * D + col) = make_float2(d0, d1);   // GELU'(pre2) for the backward
            }
            st2(sG, row, col, y0, y1);
          }
        }
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) bulk_store(sv + C::SV_G, sG.ptr, C::ARR2);              // operand of dW1 in the backward
    if (tid < 4 * ROWS / 4)                                                // LN / LN1 statistics for the backward formula
      reinterpret_cast<float4*>(sv + C::SV_STAT)[tid] = reinterpret_cast<const float4*>(s_stat)[tid];
    // ---- out = h + G W1 + b1                                              attention.py:51-53
    {
      constexpr int NTW = D / 32;
      float acc[2][NTW][4];
      zero_acc(acc);
      const uint32_t a_lane = sG.addr + (wm * 32 + (lane & 15)) * sG.ldb + ((lane >> 4) * 8) * 2;
      gemm_wf<2, NTW, 2 * D / 16, 3>(acc, [&](int ks) { return a_lane + ks * 32; }, sG.lo, 16 * sG.ldb,
                                     wf + C::F_FC1 + (wn * NTW) * 32 + lane, (D / 8) * 32, bg0);
      float* og = P.out + img;
      float2 bb[NTW];
#pragma unroll
      for (int nt = 0; nt < NTW; nt++) bb[nt] = __ldg(reinterpret_cast<const float2*>(P.b1 + (wn * NTW + nt) * 8 + 2 * t));
#pragma unroll
      for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < NTW; nt++)
#pragma unroll
          for (int hf = 0; hf < 2; hf++) {
            const int row = wm * 32 + mt * 16 + g + hf * 8, col = (wn * NTW + nt) * 8 + 2 * t;
            if (row < N) {
              float2* xs = reinterpret_cast<float2*>(sX + row * C::XLD + col);
              const float2 hr = *xs;
              const float2 ov = make_float2(acc[mt][nt][2 * hf] + bb[nt].x + hr.x, acc[mt][nt][2 * hf + 1] + bb[nt].y + hr.y);
              *xs = ov;
              *reinterpret_cast<float2*>(og + row * D + col) = ov;
            }
          }
    }
    if (tid == 0) bulk_store_wait_read();        // sU / sG may be overwritten by the next block
    __syncthreads();
  }
  if (tid == 0) bulk_store_wait_all();
}

// ---------------------------------------------------------------- backward
template <int D, int H>
__global__ void __launch_bounds__(256, 1) vit_bwd_kernel(Args a) {
  pdl_prologue();
  using C = Cfg<D, H>;
  constexpr int DH = C::DH, NDT = DH / 8;
  extern __shared__ __align__(128) char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;
  const int N = a.N, b = blockIdx.x;
  // regions: R_D | R_Q | R_K | R_V | R_PS(4 score planes) | stats | mbarrier
  char* rD = smem;
  char* rQ = smem + C::ARR;
  char* rK = smem + 2 * C::ARR;
  char* rV = smem + 3 * C::ARR;
  char* rPS = smem + 4 * C::ARR;
  float* s_stat = reinterpret_cast<float*>(smem + 4 * C::ARR + 4 * C::PARR);   // mean, rstd (LN) | mean, rstd (LN1)
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_stat + 4 * ROWS);
  const SArr sD = make_arr(rD, C::LD);
  const SArr sQ = make_arr(rQ, C::LD), sK = make_arr(rK, C::LD), sV = make_arr(rV, C::LD);
  const SArr sU1 = make_arr(rQ, C::LD);          // MLP phase: LN1(h)
  const SArr sDP = make_arr(rK, C::LD2);         // MLP phase: d(pre2)   (spans R_K|R_V)
  const SArr sG = make_arr(rPS, C::LD2);         // MLP phase: GELU(pre2)
  const SArr sU = make_arr(rPS, C::LD);          // QKV phase: LN(x)
  constexpr int KT = ROWS / 16, NPAIR = ROWS / 16;   // token contractions cover the 64 padded rows (pads are exact zeros)
  const float scale = rsqrtf((float)DH);
  const size_t img = (size_t)b * N * D;
  float* dh = a.dh + img;
  float* sT = reinterpret_cast<float*>(smem + 4 * C::ARR + 4 * C::PARR + 4 * ROWS * sizeof(float) + 16);   // fp32 [64][XLD]
  float* s_red = sT + ROWS * C::XLD;                                                                      // [8][3][D]
  float* dxo = a.dx + img;
  const float* dcur = a.dout + img;
  uint32_t parity = 0;

  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // incoming gradient -> split operand
  for (int r = warp; r < ROWS; r += 8)
#pragma unroll
    for (int e = 0; e < D / 32; e++) st1(sD, r, lane + 32 * e, r < N ? dcur[r * D + lane + 32 * e] : 0.f);
  __syncthreads();

  for (int bi = a.nblocks - 1; bi >= 0; bi--) {
    const NtVitBlock& P = a.blk[bi];
    const uint4* wf = reinterpret_cast<const uint4*>(P.wfrag);
    const float* xin = (bi ? a.blk[bi - 1].out : a.x) + img;
    const float* hg = P.h + img;
    const char* sv = reinterpret_cast<const char*>(P.save) + (size_t)b * C::SV_BYTES;
    uint4 b10[2 * D / 32];
    wf_first(b10, wf + C::B_FC1 + (wn * (2 * D / 32)) * 32 + lane);
    // ---- S0: operands the forward saved: u1 = LN1(h) -> R_Q, G = GELU(pre2) -> R_PS (async), LN statistics
    if (tid == 0) {
      mbar_expect_tx(bar, C::ARR + C::ARR2);
      bulk_load(rQ, sv + C::SV_U1, C::ARR, bar);
      bulk_load(rPS, sv + C::SV_G, C::ARR2, bar);
    }
    if (tid < 4 * ROWS / 4) reinterpret_cast<float4*>(s_stat)[tid] = reinterpret_cast<const float4*>(sv + C::SV_STAT)[tid];
    // ---- S1: dG = d W1^T ; dpre = dG * gelu'(pre2)                          attention.py:58-59
    {
      constexpr int NTW = 2 * D / 32;
      float acc[2][NTW][4];
      zero_acc(acc);
      const uint32_t a_lane = sD.addr + (wm * 32 + (lane & 15)) * sD.ldb + ((lane >> 4) * 8) * 2;
      gemm_wf<2, NTW, D / 16, 2>(acc, [&](int ks) { return a_lane + ks * 32; }, sD.lo, 16 * sD.ldb,
                                 wf + C::B_FC1 + (wn * NTW) * 32 + lane, (2 * D / 8) * 32, b10);
      const float* dg = P.dgelu + (size_t)b * N * 2 * D;
      float2 gp[2][NTW][2];
#pragma unroll
      for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < NTW; nt++)
#pragma unroll
          for (int hf = 0; hf < 2; hf++) {
            const int row = wm * 32 + mt * 16 + g + hf * 8, col = (wn * NTW + nt) * 8 + 2 * t;
            gp[mt][nt][hf] = row < N ? *reinterpret_cast<const float2*>(dg + (size_t)row * 2 * D + col) : make_float2(0.f, 0.f);
          }
#pragma unroll
      for (int nt = 0; nt < NTW; nt++) {
        const int col = (wn * NTW + nt) * 8 + 2 * t;
        float c0 = 0.f, c1 = 0.f;
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
          for (int hf = 0; hf < 2; hf++) {
            const int row = wm * 32 + mt * 16 + g + hf * 8;
            const float d0 = acc[mt][nt][2 * hf] * gp[mt][nt][hf].x, d1 = acc[mt][nt][2 * hf + 1] * gp[mt][nt][hf].y;
            st2(sDP, row, col, d0, d1);
            c0 += d0; c1 += d1;
          }
        c0 = colsum_g(c0); c1 = colsum_g(c1);                                 // db0 += colsum(dpre)  fullconnect.py:122
        if (g == 0) { atomicAdd(P.d_b0 + col, c0); atomicAdd(P.d_b0 + col + 1, c1); }
      }
    }
    uint4 b00[D / 32];
    wf_first(b00, wf + C::B_FC0 + (wn * (D / 32)) * 32 + lane);
    mbar_wait(bar, parity); parity ^= 1;
    __syncthreads();
    // ---- S2: dW1 += G^T d                                                  fullconnect.py:118-121
    {
      constexpr int MT = D / 32, NTW = D / 16;
      float acc[MT][NTW][4];
      zero_acc(acc);
      const int m0 = (warp >> 1) * (MT * 16), n0 = (warp & 1) * (NTW * 8);
      gemm_tt<MT, NTW>(acc, sG, m0, sD, n0, KT, lane);
#pragma unroll
      for (int mt = 0; mt < MT; mt++)
#pragma unroll
        for (int nt = 0; nt < NTW; nt++) {
          float* w = P.d_w1 + (size_t)(m0 + mt * 16 + g) * D + n0 + nt * 8 + 2 * t;
          red2(w, acc[mt][nt][0], acc[mt][nt][1]);
          red2(w + 8 * D, acc[mt][nt][2], acc[mt][nt][3]);
        }
    }
    // ---- S3: dU1 = dpre W0^T -> tmp                                         attention.py:60
    {
      constexpr int NTW = D / 32;
      float acc[2][NTW][4];
      zero_acc(acc);
      const uint32_t a_lane = sDP.addr + (wm * 32 + (lane & 15)) * sDP.ldb + ((lane >> 4) * 8) * 2;
      gemm_wf<2, NTW, 2 * D / 16, 3>(acc, [&](int ks) { return a_lane + ks * 32; }, sDP.lo, 16 * sDP.ldb,
                                     wf + C::B_FC0 + (wn * NTW) * 32 + lane, (D / 8) * 32, b00);
#pragma unroll
      for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < NTW; nt++)
#pragma unroll
          for (int hf = 0; hf < 2; hf++) {
            const int row = wm * 32 + mt * 16 + g + hf * 8, col = (wn * NTW + nt) * 8 + 2 * t;
            *reinterpret_cast<float2*>(sT + row * C::XLD + col) = make_float2(acc[mt][nt][2 * hf], acc[mt][nt][2 * hf + 1]);
          }
    }
    // ---- S4: dW0 += u1^T dpre
    {
      constexpr int MT = D / 32, NTW = D / 16;
      float acc[MT][NTW][4];
      zero_acc(acc);
      const int m0 = wm * (MT * 16), n0 = wn * (NTW * 8);
      gemm_tt<MT, NTW>(acc, sU1, m0, sDP, n0, KT, lane);
#pragma unroll
      for (int mt = 0; mt < MT; mt++)
#pragma unroll
        for (int nt = 0; nt < NTW; nt++) {
          float* w = P.d_w0 + (size_t)(m0 + mt * 16 + g) * 2 * D + n0 + nt * 8 + 2 * t;
          red2(w, acc[mt][nt][0], acc[mt][nt][1]);
          red2(w + 8 * 2 * D, acc[mt][nt][2], acc[mt][nt][3]);
        }
    }
    fence_async_smem();
    __syncthreads();
    // ---- q, k, v (split, as the forward left them) -> R_Q|R_K|R_V, overlapped with the LayerNorm backward
    if (tid == 0) {
      mbar_expect_tx(bar, 3 * C::ARR);
      bulk_load(rQ, sv + C::SV_QKV, 3 * C::ARR, bar);
    }
    // ---- S5: dh = d + LN1-backward(dU1) ; db1 += colsum(d)                  attention.py:61-62
    ln_bwd_rows<D>(sT, C::XLD, hg, P.ln1_g, dcur, dh, sD, N, s_stat + 2 * ROWS, s_stat + 3 * ROWS, s_red, warp, lane);
    mbar_wait(bar, parity); parity ^= 1;
    __syncthreads();
    ln_bwd_flush<D>(s_red, P.d_ln1_g, P.d_ln1_b, P.d_b1, tid);
    // ---- S6: attention backward, two heads per round
    const int rounds = (H + 1) / 2;
    for (int rnd = 0; rnd < rounds; rnd++) {
      const int hl = warp >> 2, hd = 2 * rnd + hl, i0 = (warp & 3) * 16;
      const bool head_ok = hd < H;
      const bool valid = head_ok && i0 < N;
      const SArr sP = make_arr(rPS + (2 * hl) * C::PARR, LDP);
      const SArr sdS = make_arr(rPS + (2 * hl + 1) * C::PARR, LDP);
      const int r0 = i0 + g, r1 = r0 + 8;
      float dq[NDT][4], dk[NDT][4], dv[NDT][4];
#pragma unroll
      for (int nd = 0; nd < NDT; nd++)
#pragma unroll
        for (int c = 0; c < 4; c++) dq[nd][c] = dk[nd][c] = dv[nd][c] = 0.f;
      // phase 1: dP = dO V^T, dS = P o (dP - rowsum(dP o P)), dQ = dS K / sqrt(dh)      attention.py:74-78
      if (valid) {
        const float* pg = P.P + ((size_t)b * H + hd) * N * N;
        float p[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
          const int col = nt * 8 + 2 * t;
          p[nt][0] = (r0 < N && col < N) ? pg[r0 * N + col] : 0.f;
          p[nt][1] = (r0 < N && col + 1 < N) ? pg[r0 * N + col + 1] : 0.f;
          p[nt][2] = (r1 < N && col < N) ? pg[r1 * N + col] : 0.f;
          p[nt][3] = (r1 < N && col + 1 < N) ? pg[r1 * N + col + 1] : 0.f;
        }
        float s[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; nt++) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        gemm_rowk_nk<DH>(s, sD, i0, sV, hd * DH, NPAIR, lane);
        float t0 = 0.f, t1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
          t0 += s[nt][0] * p[nt][0] + s[nt][1] * p[nt][1];
          t1 += s[nt][2] * p[nt][2] + s[nt][3] * p[nt][3];
        }
        t0 += __shfl_xor_sync(0xffffffffu, t0, 1); t0 += __shfl_xor_sync(0xffffffffu, t0, 2);
        t1 += __shfl_xor_sync(0xffffffffu, t1, 1); t1 += __shfl_xor_sync(0xffffffffu, t1, 2);
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
          const int col = nt * 8 + 2 * t;
          st2(sP, r0, col, p[nt][0], p[nt][1]);
          st2(sP, r1, col, p[nt][2], p[nt][3]);
          s[nt][0] = p[nt][0] * (s[nt][0] - t0); s[nt][1] = p[nt][1] * (s[nt][1] - t0);
          s[nt][2] = p[nt][2] * (s[nt][2] - t1); s[nt][3] = p[nt][3] * (s[nt][3] - t1);
          st2(sdS, r0, col, s[nt][0], s[nt][1]);          // phase 2 needs dS transposed: through shared memory
          st2(sdS, r1, col, s[nt][2], s[nt][3]);
        }
        gemm_regA_kn<DH>(dq, s, sK, hd * DH, KT, lane);   // dQ takes dS straight from the accumulator fragments
      } else if (head_ok) {
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
          const int col = nt * 8 + 2 * t;
          st2(sP, r0, col, 0.f, 0.f); st2(sP, r1, col, 0.f, 0.f);
          st2(sdS, r0, col, 0.f, 0.f); st2(sdS, r1, col, 0.f, 0.f);
        }
      }
      __syncthreads();
      // phase 2: this warp's 16 KEYS: dK = dS^T Q / sqrt(dh), dV = P^T dO                attention.py:76,79
      if (valid) {
        gemm_colk_kn<DH>(dk, sdS, i0, sQ, hd * DH, KT, lane);
        gemm_colk_kn<DH>(dv, sP, i0, sD, hd * DH, KT, lane);
      }
      if (rnd == rounds - 1) fence_async_smem();
      __syncthreads();
      if (rnd == rounds - 1 && tid == 0) {       // score planes are dead: fetch u = LN(x) (operand of dWqkv) into R_PS
        mbar_expect_tx(bar, C::ARR);
        bulk_load(rPS, sv + C::SV_U, C::ARR, bar);
      }
      // q/k/v of these heads are dead: overwrite them with dq/dk/dv (attention.py:81-83) + bias gradient
      if (head_ok) {
#pragma unroll
        for (int nd = 0; nd < NDT; nd++) {
          const int col = hd * DH + nd * 8 + 2 * t;
          st2(sQ, r0, col, dq[nd][0] * scale, dq[nd][1] * scale);
          st2(sQ, r1, col, dq[nd][2] * scale, dq[nd][3] * scale);
          st2(sK, r0, col, dk[nd][0] * scale, dk[nd][1] * scale);
          st2(sK, r1, col, dk[nd][2] * scale, dk[nd][3] * scale);
          st2(sV, r0, col, dv[nd][0], dv[nd][1]);
          st2(sV, r1, col, dv[nd][2], dv[nd][3]);
          float c[6] = {(dq[nd][0] + dq[nd][2]) * scale, (dq[nd][1] + dq[nd][3]) * scale,
                        (dk[nd][0] + dk[nd][2]) * scale, (dk[nd][1] + dk[nd][3]) * scale,
                        dv[nd][0] + dv[nd][2], dv[nd][1] + dv[nd][3]};
#pragma unroll
          for (int i = 0; i < 6; i++) c[i] = colsum_g(c[i]);
          if (g == 0 && valid) {
            atomicAdd(P.d_bqkv + col, c[0]); atomicAdd(P.d_bqkv + col + 1, c[1]);
            atomicAdd(P.d_bqkv + D + col, c[2]); atomicAdd(P.d_bqkv + D + col + 1, c[3]);
            atomicAdd(P.d_bqkv + 2 * D + col, c[4]); atomicAdd(P.d_bqkv + 2 * D + col + 1, c[5]);
          }
        }
      }
    }
    uint4 bu0[D / 32];
    wf_first(bu0, wf + C::B_QKV + (wn * (D / 32)) * 32 + lane);
    mbar_wait(bar, parity); parity ^= 1;
    __syncthreads();
    // ---- S8: dU = dqkv Wqkv^T -> tmp ; dWqkv += u^T dqkv                   attention.py:85-86
    {
      constexpr int NTW = D / 32;
      float acc[2][NTW][4];
      zero_acc(acc);
      const uint32_t a_lane = sQ.addr + (wm * 32 + (lane & 15)) * sQ.ldb + ((lane >> 4) * 8) * 2;
      // k runs over the 3D columns of dqkv = [dq | dk | dv], three arrays C::ARR bytes apart
      gemm_wf<2, NTW, 3 * D / 16, 3>(acc, [&](int ks) { const int seg = ks / (D / 16); return a_lane + seg * C::ARR + (ks - seg * (D / 16)) * 32; },
                                     sQ.lo, 16 * sQ.ldb, wf + C::B_QKV + (wn * NTW) * 32 + lane, (D / 8) * 32, bu0);
#pragma unroll
      for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < NTW; nt++)
#pragma unroll
          for (int hf = 0; hf < 2; hf++) {
            const int row = wm * 32 + mt * 16 + g + hf * 8, col = (wn * NTW + nt) * 8 + 2 * t;
            *reinterpret_cast<float2*>(sT + row * C::XLD + col) = make_float2(acc[mt][nt][2 * hf], acc[mt][nt][2 * hf + 1]);
          }
    }
#pragma unroll 1
    for (int seg = 0; seg < 3; seg++) {
      constexpr int MT = D / 32, NTW = D / 32;
      float acc[MT][NTW][4];
      zero_acc(acc);
      const int m0 = wm * (MT * 16), n0 = wn * (NTW * 8);
      const SArr& sB = seg == 0 ? sQ : (seg == 1 ? sK : sV);
      gemm_tt<MT, NTW>(acc, sU, m0, sB, n0, KT, lane);
#pragma unroll
      for (int mt = 0; mt < MT; mt++)
#pragma unroll
        for (int nt = 0; nt < NTW; nt++) {
          float* w = P.d_wqkv + (size_t)(m0 + mt * 16 + g) * 3 * D + seg * D + n0 + nt * 8 + 2 * t;
          red2(w, acc[mt][nt][0], acc[mt][nt][1]);
          red2(w + 8 * 3 * D, acc[mt][nt][2], acc[mt][nt][3]);
        }
    }
    __syncthreads();
    // ---- S9: dx = dh + LN-backward(dU)                                     attention.py:87-88
    ln_bwd_rows<D>(sT, C::XLD, xin, P.ln_g, dh, dxo, sD, N, s_stat, s_stat + ROWS, s_red, warp, lane);
    fence_async_smem();
    __syncthreads();
    ln_bwd_flush<D>(s_red, P.d_ln_g, P.d_ln_b, nullptr, tid);
    dcur = dxo;
  }
}

template <int D, int H>
static int ensure_attrs() {
  using C = Cfg<D, H>;
  static bool attr = false;
  if (!attr) {
    NT_CHECK(cudaFuncSetAttribute(vit_fwd_kernel<D, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::FWD_SMEM));
    NT_CHECK(cudaFuncSetAttribute(vit_bwd_kernel<D, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::BWD_SMEM));
    attr = true;
  }
  return 0;
}
template <int D, int H>
static int launch_split(const Args& a, int) {
  using C = Cfg<D, H>;
  NT_CHECK(launch_k(vit_wsplit_kernel<D, H>, dim3(ceil_div(C::F_TOTAL, 256 * 4), a.nblocks), dim3(256), 0, a));
  NT_LAUNCH_OK();
  return 0;
}
template <int D, int H>
static int launch_fwd(const Args& a, int B) {
  using C = Cfg<D, H>;
  if (int rc = ensure_attrs<D, H>()) return rc;
  NT_CHECK(launch_k(vit_fwd_kernel<D, H>, dim3(B), dim3(256), C::FWD_SMEM, a));
  NT_LAUNCH_OK();
  return 0;
}
template <int D, int H>
static int launch_bwd(const Args& a, int B) {
  using C = Cfg<D, H>;
  if (int rc = ensure_attrs<D, H>()) return rc;
  NT_CHECK(launch_k(vit_bwd_kernel<D, H>, dim3(B), dim3(256), C::BWD_SMEM, a));
  NT_LAUNCH_OK();
  return 0;
}

static bool supported(int N, int D, int H) {
  if (N < 1 || N > ROWS) return false;
  return (D == 96 && H == 4) || (D == 64 && H == 4) || (D == 32 && H == 2) || (D == 64 && H == 2) || (D == 96 && H == 6);
}

#define VB_DISPATCH(fn, ...)                                               \
  do {                                                                     \
    if (D == 96 && H == 4) return fn<96, 4>(__VA_ARGS__);                  \
    if (D == 96 && H == 6) return fn<96, 6>(__VA_ARGS__);                  \
    if (D == 64 && H == 4) return fn<64, 4>(__VA_ARGS__);                  \
    if (D == 64 && H == 2) return fn<64, 2>(__VA_ARGS__);                  \
    if (D == 32 && H == 2) return fn<32, 2>(__VA_ARGS__);                  \
  } while (0)

static int run_split(const Args& a, int D, int H) { VB_DISPATCH(launch_split, a, 0); return 2; }
static int run_fwd(const Args& a, int B, int D, int H) { VB_DISPATCH(launch_fwd, a, B); return 2; }
static int run_bwd(const Args& a, int B, int D, int H) { VB_DISPATCH(launch_bwd, a, B); return 2; }

}  // namespace vb
}  // namespace nt
using namespace nt;

extern "C" {

int nt_vit_blocks_supported(int N, int D, int H, int* ok) {
  *ok = vb::supported(N, D, H) ? 1 : 0;
  return 0;
}

int nt_vit_blocks_wfrag_bytes(int D, size_t* bytes) {
  *bytes = (size_t)56 * D * D;          // 8 bytes per weight element of the block's three Linear layers (7 D^2)
  return 0;
}

int nt_vit_blocks_save_bytes(int D, size_t* bytes_per_image) {
  const size_t arr = (size_t)vb::ROWS * (D + 8) * 4, arr2 = (size_t)vb::ROWS * (2 * D + 8) * 4;
  *bytes_per_image = 5 * arr + arr2 + 4 * vb::ROWS * sizeof(float);
  return 0;
}

int nt_vit_blocks_prepare(const NtVitBlock* blocks, int nblocks, int D, int H) {
  NT_ENTRY();
  NT_REQUIRE(vb::supported(1, D, H), "nt_vit_blocks_prepare: unsupported shape D=%d H=%d", D, H);
  NT_REQUIRE(nblocks >= 1 && nblocks <= vb::MAXB, "nt_vit_blocks_prepare: 1..%d blocks per call, got %d", vb::MAXB, nblocks);
  vb::Args a{};
  for (int i = 0; i < nblocks; i++) a.blk[i] = blocks[i];
  a.nblocks = nblocks;
  return vb::run_split(a, D, H);
}

int nt_vit_blocks_fwd(const NtVitBlock* blocks, int nblocks, const float* x, int B, int N, int D, int H, int prepared) {
  NT_ENTRY();
  NT_REQUIRE(vb::supported(N, D, H), "nt_vit_blocks_fwd: unsupported shape N=%d D=%d H=%d", N, D, H);
  NT_REQUIRE(nblocks >= 1 && nblocks <= vb::MAXB, "nt_vit_blocks_fwd: 1..%d blocks per call, got %d", vb::MAXB, nblocks);
  if (B == 0) return 0;
  vb::Args a{};
  for (int i = 0; i < nblocks; i++) a.blk[i] = blocks[i];
  a.nblocks = nblocks; a.N = N; a.x = x;
  if (!prepared) {
    if (int rc = vb::run_split(a, D, H)) return rc;
  }
  return vb::run_fwd(a, B, D, H);
}

int nt_vit_blocks_bwd(const NtVitBlock* blocks, int nblocks, const float* x, const float* dout, float* dx,
                      float* scratch, int B, int N, int D, int H) {
  NT_ENTRY();
  NT_REQUIRE(vb::supported(N, D, H), "nt_vit_blocks_bwd: unsupported shape N=%d D=%d H=%d", N, D, H);
  NT_REQUIRE(nblocks >= 1 && nblocks <= vb::MAXB, "nt_vit_blocks_bwd: 1..%d blocks per call, got %d", vb::MAXB, nblocks);
  if (B == 0) return 0;
  vb::Args a{};
  for (int i = 0; i < nblocks; i++) a.blk[i] = blocks[i];
  a.nblocks = nblocks; a.N = N; a.x = x; a.dout = dout; a.dx = dx;
  a.dh = scratch; a.tmp = scratch + (size_t)B * N * D;
  return vb::run_bwd(a, B, D, H);
}

}
