// Attention core on the 5th-generation tensor cores for dh = 64, N <= 64 tokens (the scaled ViT: 6 heads x 64, 49 tokens):
// attention.py:31-46 (forward) and :63-84 (backward) for ALL heads of an image in one CTA, persistent over the images.
//
// Same construction as the fused ViT blocks (vit_block_tc.cu): q, k, v (and dO, P, dS) live in shared memory as bf16
// hi / lo planes in the 128-byte-swizzled layout — a [token][64] plane is one 8 KB chunk that serves as a K-major operand
// (Q K^T, dO V^T, dS K) and, the same bytes, as an MN-major one (P V, P^T dO, dS^T Q); every product is three tcgen05.mma
// passes (lo*hi + hi*lo + hi*hi, fp32 accumulator in TMEM) issued by one thread; accumulator row i is TMEM lane i, so a
// THREAD owns a token row and softmax / dS are row-local.  Two heads are in flight (two "slots"); a row is shared by two
// threads (32 columns each), which exchange their partial row maxima / sums through shared memory.
//
// An alternative to the mma.sync kernels of attention_bf16.cu at these shapes (one CTA per (image, head)); parity-tested
// through the same C ABI (tests/test_gpu_attention_tc.py), opt-in — see enabled() below for the measurements.
#include "common.cuh"
#include "umma.cuh"
#include <cstdlib>

namespace nt {
namespace atc {
using namespace um;

constexpr int DH = 64, ROWS = 64, SLOTS = 2;
constexpr uint32_t CH = ROWS * 128;              // one [64][64] bf16 plane
constexpr int THREADS = 448;                     // 14 warps: token warps {0,1,4,5,8,9,12,13}, MMA issuer 3
constexpr int TOK_THREADS = 256;

// shared memory: per slot the planes (hi, lo) of q, k, v [, dO, P] then the exchange area and the barriers
constexpr int FWD_PLANES = 4, BWD_PLANES = 5;    // q k v P | q k v dO P
constexpr uint32_t fwd_slot_bytes = FWD_PLANES * 2 * CH, bwd_slot_bytes = BWD_PLANES * 2 * CH;
constexpr uint32_t XCH_BYTES = SLOTS * 64 * 2 * 2 * 4;      // [slot][row][half][2] floats
// + one chunk of padding behind the planes: an M = 128 A operand reads 16 KB from its start (rows 64..127 are never used,
// but the addresses must stay inside the allocation)
constexpr uint32_t FWD_SMEM = SLOTS * fwd_slot_bytes + CH + XCH_BYTES + 32 * 8 + 16;
constexpr uint32_t BWD_SMEM = SLOTS * bwd_slot_bytes + CH + XCH_BYTES + 32 * 8 + 16;
static_assert(BWD_SMEM <= 232448, "shared memory budget");

struct Args {
  const float* qkv;      // [B,N,3*H*64]
  float* out;            // [B,N,H*64]            forward
  float* P;              // [B,H,N,N] or null     forward
  const float* residual; // [B,N,H*64] or null    forward
  uint32_t* o_hi;        // bf16 pairs of `out` (row-major) or null
  uint32_t* o_lo;
  const float* dout;     // [B,N,H*64]            backward
  float* dqkv;           // [B,N,3*H*64]          backward
  uint32_t* g_hi;        // bf16 pairs of dqkv or null
  uint32_t* g_lo;
  int B, N, H, causal;
};

enum { B_QKV = 0 /*2*/, B_S = 2, B_P = 4, B_O = 6, B_DP = 8, B_DV = 10, B_DS = 12, B_DQK = 14, B_COUNT = 16 };

__device__ __forceinline__ void split_store32(uint8_t* hi, uint8_t* lo, int r, int c0, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 4; j++) {
    float w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = v[8 * j + i];
    store8_split(hi, lo, r, c0 + 8 * j, CH, w);
  }
}
__device__ __forceinline__ void load32(const float* p, bool live, float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const float4 t = live ? __ldg(reinterpret_cast<const float4*>(p) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
}
// 32 fp32 values of row `row` -> fp32 global (+ optional bf16 hi / lo pair planes), all row-major with pitch `ld`
__device__ __forceinline__ void store32_out(float* dst, uint32_t* hi, uint32_t* lo, long long idx, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 8; j++)
    *reinterpret_cast<float4*>(dst + idx + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  if (hi) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      uint4 h, l;
      split2(v[8 * j], v[8 * j + 1], h.x, l.x);
      split2(v[8 * j + 2], v[8 * j + 3], h.y, l.y);
      split2(v[8 * j + 4], v[8 * j + 5], h.z, l.z);
      split2(v[8 * j + 6], v[8 * j + 7], h.w, l.w);
      *reinterpret_cast<uint4*>(hi + (idx >> 1) + 4 * j) = h;
      *reinterpret_cast<uint4*>(lo + (idx >> 1) + 4 * j) = l;
    }
  }
}
// three-pass product D (+)= A B with both operands given as (hi, lo) start addresses and per-K-step advances
__device__ __forceinline__ void mma3(uint32_t tacc, uint32_t a_hi, uint32_t a_lo, bool a_mn, uint32_t b_hi, uint32_t b_lo, bool b_mn,
                                     uint32_t idesc) {
#pragma unroll
  for (int ks = 0; ks < 4; ks++) {
    const uint32_t ao = ks * (a_mn ? MNMAJ_STEP : KMAJ_STEP), bo = ks * (b_mn ? MNMAJ_STEP : KMAJ_STEP);
    const uint64_t ah = a_mn ? desc_mnmajor(a_hi + ao, CH) : desc_kmajor(a_hi + ao);
    const uint64_t al = a_mn ? desc_mnmajor(a_lo + ao, CH) : desc_kmajor(a_lo + ao);
    const uint64_t bh = b_mn ? desc_mnmajor(b_hi + bo, CH) : desc_kmajor(b_hi + bo);
    const uint64_t bl = b_mn ? desc_mnmajor(b_lo + bo, CH) : desc_kmajor(b_lo + bo);
    mma_bf16(tacc, al, bh, idesc, ks ? 1u : 0u);
    mma_bf16(tacc, ah, bl, idesc, 1u);
    mma_bf16(tacc, ah, bh, idesc, 1u);
  }
}

// softmax of one half row: t[32] holds the raw scores of keys 32 half .. 32 half + 31; on return the probabilities.
// The two threads of a row exchange their partial maximum / sum through `xch` ([row][half][2]); `bar_id` is the named
// barrier of the slot (128 threads).
__device__ __forceinline__ void softmax_half(uint32_t (&t)[32], int r, int half, int N, bool causal, float scale, bool live,
                                             float* xch, int bar_id) {
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < 32; j++) {
    const int col = 32 * half + j;
    const bool ok = col < N && !(causal && col > r);
    const float s = ok ? __uint_as_float(t[j]) * scale : -INFINITY;       // scale, then mask (attention.py:37-39)
    t[j] = __float_as_uint(s);
    mx = fmaxf(mx, s);
  }
  xch[(r * 2 + half) * 2] = mx;
  named_bar_sync(bar_id, 128);
  mx = fmaxf(mx, xch[(r * 2 + (half ^ 1)) * 2]);
  if (!(mx > -INFINITY)) mx = 0.f;
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 32; j++) {
    const float e = __expf(__uint_as_float(t[j]) - mx);
    t[j] = __float_as_uint(e);
    sum += e;
  }
  xch[(r * 2 + half) * 2 + 1] = sum;
  named_bar_sync(bar_id, 128);
  sum += xch[(r * 2 + (half ^ 1)) * 2 + 1];
  const float inv = (live && sum > 0.f) ? __fdividef(1.0f, sum) : 0.f;
#pragma unroll
  for (int j = 0; j < 32; j++) t[j] = __float_as_uint(__uint_as_float(t[j]) * inv);
}

template <bool BWD>
__global__ void __launch_bounds__(THREADS, 1) attn_tc_kernel(Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023u) __trap();
  constexpr uint32_t slot_bytes = BWD ? bwd_slot_bytes : fwd_slot_bytes;
  constexpr int NPL = BWD ? BWD_PLANES : FWD_PLANES;
  float* xch_all = reinterpret_cast<float*>(smem + SLOTS * slot_bytes + CH);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SLOTS * slot_bytes + CH + XCH_BYTES);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 32);

  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = a.N, H = a.H, D = H * DH;
  const bool is_tok = (warp & 3) < 2 && warp < 14;
  const int q = warp & 3, g = warp >> 2, slot = g >> 1, half = g & 1, r = 32 * q + lane;
  const int rounds_per_img = (H + SLOTS - 1) / SLOTS;

  if (tid == 0) {
    for (int i = 0; i < B_COUNT; i++) mbar_init(&bar[i], 1);
    for (int s = 0; s < SLOTS; s++) { mbar_init(&bar[B_QKV + s], 4); mbar_init(&bar[B_P + s], 4); mbar_init(&bar[B_DS + s], 4); }
    mbar_fence_init();
  }
  for (uint32_t i = tid * 16; i < SLOTS * slot_bytes + CH; i += THREADS * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  if (warp == 3) tmem_alloc<512>(tslot);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tslot;
  pdl_wait();
  // plane p (0 q, 1 k, 2 v, [3 dO,] last P) of slot s: hi at plane_addr, lo CH behind
  auto plane = [&](int s, int p) -> uint8_t* { return smem + s * slot_bytes + (uint32_t)p * 2 * CH; };
  constexpr int PL_Q = 0, PL_K = 1, PL_V = 2, PL_DO = 3, PL_P = NPL - 1;
  constexpr uint32_t TCOLS = BWD ? 192 : 128;     // TMEM columns per slot

  if (warp == 3) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t id64 = idesc_bf16(128, 64, 0, 0), id_pv = idesc_bf16(128, 64, 0, 1), id_tt = idesc_bf16(128, 64, 1, 1);
      uint32_t cnt[SLOTS] = {0, 0};              // rounds in which the slot was active (= barrier phases used)
      auto P_ = [&](int s, int p) -> uint32_t { return smem_u32(plane(s, p)); };
      auto wait = [&](int b, uint32_t ph) { mbar_wait(&bar[b], ph); tc_fence_after(); };
      for (int img = blockIdx.x; img < a.B; img += gridDim.x)
        for (int rd = 0; rd < rounds_per_img; rd++) {
          for (int s = 0; s < SLOTS; s++) {
            if (rd * SLOTS + s >= H) continue;
            const uint32_t ph = cnt[s] & 1, tc = tbase + TCOLS * s;
            wait(B_QKV + s, ph);
            mma3(tc, P_(s, PL_Q), P_(s, PL_Q) + CH, false, P_(s, PL_K), P_(s, PL_K) + CH, false, id64);          // S = Q K^T
            mma_commit(&bar[B_S + s]);
            if (BWD) {
              mma3(tc + 64, P_(s, PL_DO), P_(s, PL_DO) + CH, false, P_(s, PL_V), P_(s, PL_V) + CH, false, id64);   // dP = dO V^T
              mma_commit(&bar[B_DP + s]);
            }
          }
          for (int s = 0; s < SLOTS; s++) {
            if (rd * SLOTS + s >= H) continue;
            const uint32_t ph = cnt[s] & 1, tc = tbase + TCOLS * s;
            wait(B_P + s, ph);
            if (!BWD) {
              mma3(tc + 64, P_(s, PL_P), P_(s, PL_P) + CH, false, P_(s, PL_V), P_(s, PL_V) + CH, true, id_pv);     // O = P V
              mma_commit(&bar[B_O + s]);
            } else {
              mma3(tc + 128, P_(s, PL_P), P_(s, PL_P) + CH, true, P_(s, PL_DO), P_(s, PL_DO) + CH, true, id_tt);   // dV = P^T dO
              mma_commit(&bar[B_DV + s]);
            }
          }
          if (BWD)
            for (int s = 0; s < SLOTS; s++) {
              if (rd * SLOTS + s >= H) continue;
              const uint32_t ph = cnt[s] & 1, tc = tbase + TCOLS * s;
              wait(B_DS + s, ph);
              mma3(tc, P_(s, PL_P), P_(s, PL_P) + CH, false, P_(s, PL_K), P_(s, PL_K) + CH, true, id_pv);          // dQ = dS K
              mma3(tc + 64, P_(s, PL_P), P_(s, PL_P) + CH, true, P_(s, PL_Q), P_(s, PL_Q) + CH, true, id_tt);      // dK = dS^T Q
              mma_commit(&bar[B_DQK + s]);
            }
          for (int s = 0; s < SLOTS; s++)
            if (rd * SLOTS + s < H) cnt[s]++;
        }
    }
    __syncwarp();
  } else if (is_tok) {
    // ===================== token warps: thread = (token row r, head slot, 32-column half) =====================
    const bool live = r < N;
    const float scale = rsqrtf((float)DH);
    const uint32_t lane_addr = tbase + ((uint32_t)(32 * q) << 16) + TCOLS * slot;
    float* xch = xch_all + slot * 64 * 2 * 2;
    const int bar_id = 1 + slot;
    uint32_t cnt = 0;
    auto arrive = [&](int b) {
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[b]);
    };
    auto wait = [&](int b, uint32_t ph) { mbar_wait(&bar[b], ph); tc_fence_after(); };
    for (int img = blockIdx.x; img < a.B; img += gridDim.x)
      for (int rd = 0; rd < rounds_per_img; rd++) {
        const int h = rd * SLOTS + slot;
        if (h >= H) continue;
        const uint32_t ph = cnt & 1;
        cnt++;
        const long long row = (long long)img * N + r;
        // ---- operands: this thread's 32 columns of q, k, v (and dO) of head h -> planes
        {
          float v[32];
          const float* base = a.qkv + row * 3 * D + h * DH + 32 * half;
#pragma unroll 1
          for (int part = 0; part < 3; part++) {
            load32(base + part * D, live, v);
            split_store32(plane(slot, part), plane(slot, part) + CH, r, 32 * half, v);
          }
          if (BWD) {
            load32(a.dout + row * D + h * DH + 32 * half, live, v);
            split_store32(plane(slot, PL_DO), plane(slot, PL_DO) + CH, r, 32 * half, v);
          }
        }
        arrive(B_QKV + slot);
        // ---- P = softmax(S / sqrt(dh) (+ causal mask)), this thread's 32 keys        attention.py:37-41
        uint32_t p[32];
        wait(B_S + slot, ph);
        tmem_ld32(lane_addr + 32 * half, p);
        tmem_wait_ld();
        softmax_half(p, r, half, N, a.causal != 0, scale, live, xch, bar_id);
        {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] = __uint_as_float(p[j]);
          split_store32(plane(slot, PL_P), plane(slot, PL_P) + CH, r, 32 * half, v);
          if (!BWD && a.P && live) {                                  // the reference's atg__ (attention.py:41)
            float* pg = a.P + (((long long)img * H + h) * N + r) * N + 32 * half;
#pragma unroll
            for (int j = 0; j < 32; j++)
              if (32 * half + j < N) pg[j] = v[j];
          }
        }
        arrive(B_P + slot);
        if (!BWD) {
          // ---- out = P V (+ residual)                                                attention.py:42-46
          wait(B_O + slot, ph);
          uint32_t o[32];
          tmem_ld32(lane_addr + 64 + 32 * half, o);
          tmem_wait_ld();
          if (live) {
            const long long oi = row * D + h * DH + 32 * half;
            float v[32];
            if (a.residual) load32(a.residual + oi, true, v);
#pragma unroll
            for (int j = 0; j < 32; j++) v[j] = __uint_as_float(o[j]) + (a.residual ? v[j] : 0.f);
            store32_out(a.out, a.o_hi, a.o_lo, oi, v);
          }
        } else {
          // ---- dS = P o (dP - rowsum(dP o P)) / sqrt(dh)                              attention.py:74-75
          wait(B_DP + slot, ph);
          uint32_t dp[32];
          tmem_ld32(lane_addr + 64 + 32 * half, dp);
          tmem_wait_ld();
          float t = 0.f;
#pragma unroll
          for (int j = 0; j < 32; j++) t += __uint_as_float(dp[j]) * __uint_as_float(p[j]);
          xch[(r * 2 + half) * 2] = t;
          named_bar_sync(bar_id, 128);
          t += xch[(r * 2 + (half ^ 1)) * 2];
          named_bar_sync(bar_id, 128);                              // (the buffer is rewritten by the next round's softmax)
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] = live ? __uint_as_float(p[j]) * (__uint_as_float(dp[j]) - t) * scale : 0.f;
          wait(B_DV + slot, ph);                                     // dV has read P: the plane takes dS
          split_store32(plane(slot, PL_P), plane(slot, PL_P) + CH, r, 32 * half, v);
          arrive(B_DS + slot);
          // ---- dq, dk, dv of head h -> dqkv (attention.py:76-83)
          wait(B_DQK + slot, ph);
#pragma unroll 1
          for (int part = 0; part < 3; part++) {
            uint32_t o[32];
            tmem_ld32(lane_addr + (part == 0 ? 0 : (part == 1 ? 64 : 128)) + 32 * half, o);
            tmem_wait_ld();
            if (live) {
#pragma unroll
              for (int j = 0; j < 32; j++) v[j] = __uint_as_float(o[j]);
              store32_out(a.dqkv, a.g_hi, a.g_lo, row * 3 * D + part * D + h * DH + 32 * half, v);
            }
          }
        }
      }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc<512>(tbase);
  }
}

// OPT-IN (NTB200_ATT_TC=1).  Measured on B200 at the scaled-ViT shape (B = 1024, N = 49, 6 heads; tools/att_time.py): forward
// 257 us, backward 406 us against 152 / 272 us for the mma.sync kernels of attention_bf16.cu.  One image per CTA with eight
// token warps walks load -> S -> softmax -> P V -> store as a serial chain per head (12 us per round of two heads), while the
// mma.sync kernels keep four CTAs per SM in flight and hide the same latencies; closing that needs the next round's q / k / v
// staged during the current one.
static bool enabled() {
  const char* e = getenv("NTB200_ATT_TC");
  return e && e[0] == '1';
}
}  // namespace atc

bool attention_tc_ok(int N, int H, int dh, int mask_kind, bool want_scores) {
  return atc::enabled() && dh == 64 && N >= 16 && N <= 64 && H >= 1 && mask_kind != NT_MASK_DENSE && !want_scores && g_gemm_mode == 1;
}

static int attn_tc_launch(const atc::Args& a, bool bwd) {
  static bool cfg = false;
  if (!cfg) {
    NT_CHECK(cudaFuncSetAttribute(atc::attn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)atc::FWD_SMEM));
    NT_CHECK(cudaFuncSetAttribute(atc::attn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)atc::BWD_SMEM));
    cfg = true;
  }
  const int grid = a.B < g_sm_count ? a.B : g_sm_count;
  if (bwd) NT_CHECK(launch_k(atc::attn_tc_kernel<true>, dim3(grid), dim3(atc::THREADS), atc::BWD_SMEM, a));
  else NT_CHECK(launch_k(atc::attn_tc_kernel<false>, dim3(grid), dim3(atc::THREADS), atc::FWD_SMEM, a));
  NT_LAUNCH_OK();
  return 0;
}

int attention_fwd_tc(const float* qkv, int mask_kind, float* out, float* P, int B, int N, int H, const float* residual, void* o_hi,
                     void* o_lo) {
  atc::Args a{};
  a.qkv = qkv; a.out = out; a.P = P; a.residual = residual; a.o_hi = (uint32_t*)o_hi; a.o_lo = (uint32_t*)o_lo;
  a.B = B; a.N = N; a.H = H; a.causal = mask_kind == NT_MASK_CAUSAL ? 1 : 0;
  return attn_tc_launch(a, false);
}

int attention_bwd_tc(const float* dout, const float* qkv, float* dqkv, int B, int N, int H, int causal, void* g_hi, void* g_lo) {
  atc::Args a{};
  a.qkv = qkv; a.dout = dout; a.dqkv = dqkv; a.g_hi = (uint32_t*)g_hi; a.g_lo = (uint32_t*)g_lo;
  a.B = B; a.N = N; a.H = H; a.causal = causal;
  return attn_tc_launch(a, true);
}

}  // namespace nt
