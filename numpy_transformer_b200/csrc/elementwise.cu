// Bandwidth kernels: activations, adds, patchify, embedding gather/scatter, optimiser updates.
#include "common.cuh"

namespace nt {

enum { OP_GELU_F, OP_GELU_B, OP_RELU_F, OP_RELU_B, OP_ADD, OP_SCALE, OP_SIGMOID_F, OP_SIGMOID_B, OP_TANH_F, OP_TANH_B,
       OP_SILU_F, OP_SILU_B };

template <int OP>
__global__ void __launch_bounds__(256) ew_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                 float* __restrict__ out, size_t n, float alpha) {
  pdl_prologue();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n4 = n / 4;
  const bool al = ((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)out)) & 15) == 0;
  auto f = [&](float x, float y) -> float {
    if (OP == OP_GELU_F) return gelu_f(x);
    if (OP == OP_GELU_B) return x * gelu_grad_f(y);          // a = dy, b = x
    if (OP == OP_RELU_F) return x < 0.f ? 0.f : x;
    if (OP == OP_RELU_B) return y > 0.f ? x : 0.f;           // a = dy, b = pre
    if (OP == OP_ADD) return x + y;
    if (OP == OP_SIGMOID_F) return 1.0f / (1.0f + expf(-x));                 // net/activation.py:27-30
    if (OP == OP_SIGMOID_B) return y * (1.0f - y) * x;                        // :32-33   a = dy, b = y
    if (OP == OP_TANH_F) return tanhf(x);                                     // :43-47
    if (OP == OP_TANH_B) return (1.0f - y * y) * x;                           // :49-50   a = dy, b = y
    if (OP == OP_SILU_F) return x / (1.0f + expf(-x));                        // :59-63
    if (OP == OP_SILU_B) {                                                    // :65-67   a = dy, b = x
      const float d = 1.0f / (1.0f + expf(-y));                               // (div + out * e^-x * div) = d (1 + x (1 - d))
      return d * (1.0f + y * (1.0f - d)) * x;
    }
    return x * alpha;
  };
  if (al) {
    for (; i < n4; i += stride) {
      float4 x = reinterpret_cast<const float4*>(a)[i];
      float4 y = b ? reinterpret_cast<const float4*>(b)[i] : make_float4(0, 0, 0, 0);
      float4 o = make_float4(f(x.x, y.x), f(x.y, y.y), f(x.z, y.z), f(x.w, y.w));
      reinterpret_cast<float4*>(out)[i] = o;
    }
    for (size_t j = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
      out[j] = f(a[j], b ? b[j] : 0.f);
  } else {
    for (; i < n; i += stride) out[i] = f(a[i], b ? b[i] : 0.f);
  }
}

template <int OP>
static int launch_ew(const float* a, const float* b, float* out, size_t n, float alpha = 1.f) {
  if (n == 0) return 0;
  long long blocks = (long long)((n / 4 + 255) / 256);
  if (blocks < 1) blocks = 1;
  long long cap = (long long)g_sm_count * 8;
  if (blocks > cap) blocks = cap;
  NT_CHECK(nt::launch_k(ew_kernel<OP>, dim3((int)blocks), dim3(256), 0, a, b, out, n, alpha));
  NT_LAUNCH_OK();
  return 0;
}

__global__ void add_rows_kernel(const float* __restrict__ x, const float* __restrict__ table, float* __restrict__ out,
                                long long total, int period, int D) {
  pdl_prologue();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long pd = (long long)period * D;
  for (; i < total; i += stride) out[i] = x[i] + table[i % pd];
}

// (B,C,Hi,Wi) -> (B, P*P, C*hl*wl), (c,h,w) order inside a patch  (PatchEmbed.py:27-36)
__global__ void patchify_kernel(const float* __restrict__ img, float* __restrict__ out, int B, int C, int Hi, int Wi,
                                int P, int hl, int wl) {
  pdl_prologue();
  const long long total = (long long)B * P * P * C * hl * wl;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    long long r = i;
    int w = r % wl; r /= wl;
    int h = r % hl; r /= hl;
    int c = r % C; r /= C;
    int pj = r % P; r /= P;
    int pi = r % P; r /= P;
    int b = (int)r;
    out[i] = img[(((long long)b * C + c) * Hi + pi * hl + h) * Wi + pj * wl + w];
  }
}

// ---- batch construction on the device
// images[b] = lut[src[perm[first + b]]]: the shuffled, /255-normalised MNIST slice of transformer_of_image.py:99-100,
// 128-137 without touching the host; lut[v] = float32(v / 255.0) is computed on the host so that the values are bit-identical
// to the reference's fp64 division followed by the fp32 upload cast.  labels likewise.
__global__ void gather_images_u8_kernel(const unsigned char* __restrict__ src, const int32_t* __restrict__ perm, int first,
                                        const float* __restrict__ lut, float* __restrict__ out, int B, int pixels) {
  pdl_prologue();
  const long long total = (long long)B * pixels;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / pixels), p = (int)(i - (long long)b * pixels);
    out[i] = lut[src[(long long)perm[first + b] * pixels + p]];
  }
}

__global__ void gather_i32_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ perm, int first,
                                  int32_t* __restrict__ out, int B) {
  pdl_prologue();
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) out[b] = src[perm[first + b]];
}

// inputs[b, t] = corpus[start[b] + t], labels[b, t] = corpus[start[b] + t + 1]   (gpt/gpt_train_potry3000.py:106-133)
__global__ void gather_windows_kernel(const int32_t* __restrict__ corpus, const int32_t* __restrict__ start,
                                      int32_t* __restrict__ inputs, int32_t* __restrict__ labels, int B, int T) {
  pdl_prologue();
  const long long total = (long long)B * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / T), t = (int)(i - (long long)b * T);
    const long long o = (long long)start[b] + t;
    inputs[i] = corpus[o];
    labels[i] = corpus[o + 1];
  }
}

// ---- convolution as im2col + Linear (net/Convolution.py:66-84, 104-131, 163-225)
struct ConvGeom { int N, C, H, W, KH, KW, SH, SW, PH, PW, OH, OW; };

// col[(n, oh, ow)][(c, kh, kw)] = x[n, c, oh*SH + kh - PH, ow*SW + kw - PW] (zero outside the image)
__global__ void im2col_kernel(const float* __restrict__ x, float* __restrict__ col, ConvGeom g) {
  pdl_prologue();
  const long long K = (long long)g.C * g.KH * g.KW, total = (long long)g.N * g.OH * g.OW * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int kw = r % g.KW; r /= g.KW;
    const int kh = r % g.KH; r /= g.KH;
    const int c = r % g.C; r /= g.C;
    const int ow = r % g.OW; r /= g.OW;
    const int oh = r % g.OH; r /= g.OH;
    const int n = (int)r;
    const int h = oh * g.SH + kh - g.PH, w = ow * g.SW + kw - g.PW;
    col[i] = (h >= 0 && h < g.H && w >= 0 && w < g.W) ? x[(((long long)n * g.C + c) * g.H + h) * g.W + w] : 0.f;
  }
}

// dx[n, c, h, w] = sum of dcol over every window position that covers the pixel: a gather, so no atomics and a
// deterministic summation order (the reference's rotated-kernel full convolution, Convolution.py:196-223)
__global__ void col2im_kernel(const float* __restrict__ dcol, float* __restrict__ dx, ConvGeom g) {
  pdl_prologue();
  const long long K = (long long)g.C * g.KH * g.KW, total = (long long)g.N * g.C * g.H * g.W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int w = r % g.W; r /= g.W;
    const int h = r % g.H; r /= g.H;
    const int c = r % g.C; r /= g.C;
    const int n = (int)r;
    float acc = 0.f;
    for (int kh = 0; kh < g.KH; kh++) {
      const int hn = h + g.PH - kh;
      if (hn < 0 || hn % g.SH) continue;
      const int oh = hn / g.SH;
      if (oh >= g.OH) continue;
      for (int kw = 0; kw < g.KW; kw++) {
        const int wn = w + g.PW - kw;
        if (wn < 0 || wn % g.SW) continue;
        const int ow = wn / g.SW;
        if (ow >= g.OW) continue;
        acc += dcol[(((long long)n * g.OH + oh) * g.OW + ow) * K + ((long long)c * g.KH + kh) * g.KW + kw];
      }
    }
    dx[i] = acc;
  }
}

// y[b][c][r] = x[b][r][c]: NHWC <-> NCHW and the (OC, C*KH*KW) <-> (C*KH*KW, OC) views of the kernel tensor
__global__ void transpose_last2_kernel(const float* __restrict__ x, float* __restrict__ y, int rows, int cols) {
  pdl_prologue();
  __shared__ float tile[32][33];
  const long long base = (long long)blockIdx.z * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = x[base + (long long)r * cols + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) y[base + (long long)c * rows + r] = tile[threadIdx.x][j];
  }
}

__global__ void embedding_fwd_kernel(const int32_t* __restrict__ idx, const float* __restrict__ etok,
                                     const float* __restrict__ epos, float* __restrict__ out, int BT, int T, int D,
                                     int vocab, int* err) {
  pdl_prologue();
  const int row = blockIdx.x;
  if (row >= BT) return;
  int tok = idx[row];
  if (tok < 0) tok += vocab;                    // NumPy wraps negative indices (Embedding.py:46: self.params[inputs])
  const bool bad = tok < 0 || tok >= vocab;     // ... and raises IndexError beyond the table: flag it, read nothing
  if (bad && threadIdx.x == 0) *err = NT_DEVERR_TOKEN;
  const int t = row % T;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float v = bad ? 0.f : etok[(long long)tok * D + c];
    if (epos) v += epos[(long long)t * D + c];
    out[(long long)row * D + c] = v;
  }
}

// one decode step: out[b] = etok[idx[b]] + epos[*d_pos]  (PatchEmbed.py:132-138 for the single new position)
__global__ void decode_embed_kernel(const int32_t* __restrict__ idx, const float* __restrict__ etok,
                                    const float* __restrict__ epos, float* __restrict__ out, int D, int Tmax,
                                    const int32_t* __restrict__ d_pos, int vocab, int* err) {
  pdl_prologue();
  const int row = blockIdx.x, t = min(*d_pos, Tmax - 1);
  int tok = idx[row];
  if (tok < 0) tok += vocab;
  const bool bad = tok < 0 || tok >= vocab;
  if (bad && threadIdx.x == 0) *err = NT_DEVERR_TOKEN;
  for (int c = threadIdx.x; c < D; c += blockDim.x)
    out[(long long)row * D + c] = (bad ? 0.f : etok[(long long)tok * D + c]) + epos[(long long)t * D + c];
}

// greedy choice on device: tok[b] = argmax(logits[b, :cols]) (first index on ties, like np.argmax), also recorded at
// hist[b, *d_pos + 1] so a whole continuation is read back with one copy
__global__ void __launch_bounds__(256) argmax_next_kernel(const float* __restrict__ logits, int ld, int cols,
                                                          int32_t* __restrict__ tok, int32_t* __restrict__ hist, int hist_ld,
                                                          const int32_t* __restrict__ d_pos) {
  pdl_prologue();
  __shared__ float sv[8];
  __shared__ int si[8];
  const int row = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float v = -INFINITY;
  int idx = 0x7fffffff;
  for (int c = tid; c < cols; c += 256) {
    const float x = logits[(long long)row * ld + c];
    if (x > v) { v = x; idx = c; }                    // ascending c per thread: the first maximum is kept
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  if (lane == 0) { sv[warp] = v; si[warp] = idx; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; w++)
      if (sv[w] > v || (sv[w] == v && si[w] < idx)) { v = sv[w]; idx = si[w]; }
    tok[row] = idx;
    const int p = *d_pos + 1;
    if (hist && p < hist_ld) hist[(long long)row * hist_ld + p] = idx;
  }
}

__global__ void embedding_bwd_tok_kernel(const int32_t* __restrict__ idx, const float* __restrict__ dy,
                                         float* __restrict__ detok, int BT, int D, int vocab, int* err) {
  pdl_prologue();
  const int row = blockIdx.x;
  if (row >= BT) return;
  int tok = idx[row];
  if (tok < 0) tok += vocab;
  if (tok < 0 || tok >= vocab) {                // never scatter outside the gradient table
    if (threadIdx.x == 0) *err = NT_DEVERR_TOKEN;
    return;
  }
  for (int c = threadIdx.x; c < D; c += blockDim.x) atomicAdd(detok + (long long)tok * D + c, dy[(long long)row * D + c]);
}

__global__ void embedding_bwd_pos_kernel(const float* __restrict__ dy, float* __restrict__ depos, int B, int T, int D) {
  pdl_prologue();
  const long long total = (long long)T * D;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float s = 0.f;
  for (int b = 0; b < B; b++) s += dy[(long long)b * total + i];
  atomicAdd(depos + i, s);     // += : accumulate until setzero (PatchEmbed.py:142-143)
}

__global__ void __launch_bounds__(256) sgd_kernel(float* __restrict__ p, float* __restrict__ g, size_t n, float lr,
                                                  const float* __restrict__ lr_dev, int zero_grad) {
  pdl_prologue();
  if (lr_dev) lr = *lr_dev;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    p[i] -= g[i] * lr;
    if (zero_grad) g[i] = 0.f;
  }
}

// net/fullconnect.py:137-149: m̂ = m/sqrt(1-b1^t), v̂ = v/sqrt(1-b2^t), p -= m̂*lr/(sqrt(v̂)+1e-8)
__global__ void __launch_bounds__(256) adam_star_kernel(float* __restrict__ p, float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, size_t n,
                                                        float lr, const float* __restrict__ lr_dev, int t,
                                                        const int32_t* __restrict__ t_dev, int zero_grad) {
  pdl_prologue();
  if (lr_dev) lr = *lr_dev;
  if (t_dev) t = *t_dev;
  const float b1 = 0.9f, omb1 = 0.1f, b2 = 0.999f, omb2 = 0.001f, eps = 1e-8f;
  // bias corrections once per thread in fp64: 1-0.999^t loses ~1e-5 relative in fp32 for small t
  const float c1 = (float)(1.0 / sqrt(1.0 - pow(0.9, (double)t)));
  const float c2 = (float)(1.0 / sqrt(1.0 - pow(0.999, (double)t)));
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i];
    const float mi = b1 * m[i] + omb1 * gi;
    const float vi = b2 * v[i] + omb2 * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= (mi * c1) * lr / (sqrtf(vi * c2) + eps);
    if (zero_grad) g[i] = 0.f;
  }
}

__global__ void increment_kernel(int32_t* c) {
  pdl_prologue(); *c += 1; }

static int grid_for(size_t n) {
  long long blocks = (long long)((n + 255) / 256);
  long long cap = (long long)g_sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace nt
using namespace nt;

extern "C" {

int nt_gelu_fwd(const float* x, float* y, size_t n) { NT_ENTRY(); return launch_ew<OP_GELU_F>(x, nullptr, y, n); }
int nt_gelu_bwd(const float* dy, const float* x, float* dx, size_t n) { NT_ENTRY(); return launch_ew<OP_GELU_B>(dy, x, dx, n); }
int nt_relu_fwd(const float* x, float* y, size_t n) { NT_ENTRY(); return launch_ew<OP_RELU_F>(x, nullptr, y, n); }
int nt_relu_bwd(const float* dy, const float* pre, float* dx, size_t n) { NT_ENTRY(); return launch_ew<OP_RELU_B>(dy, pre, dx, n); }
int nt_add(const float* a, const float* b, float* out, size_t n) { NT_ENTRY(); return launch_ew<OP_ADD>(a, b, out, n); }
int nt_sigmoid_fwd(const float* x, float* y, size_t n) { NT_ENTRY(); return launch_ew<OP_SIGMOID_F>(x, nullptr, y, n); }
int nt_sigmoid_bwd(const float* dy, const float* y, float* dx, size_t n) { NT_ENTRY(); return launch_ew<OP_SIGMOID_B>(dy, y, dx, n); }
int nt_tanh_fwd(const float* x, float* y, size_t n) { NT_ENTRY(); return launch_ew<OP_TANH_F>(x, nullptr, y, n); }
int nt_tanh_bwd(const float* dy, const float* y, float* dx, size_t n) { NT_ENTRY(); return launch_ew<OP_TANH_B>(dy, y, dx, n); }
int nt_silu_fwd(const float* x, float* y, size_t n) { NT_ENTRY(); return launch_ew<OP_SILU_F>(x, nullptr, y, n); }
int nt_silu_bwd(const float* dy, const float* x, float* dx, size_t n) { NT_ENTRY(); return launch_ew<OP_SILU_B>(dy, x, dx, n); }
int nt_scale(float* x, float alpha, size_t n) { NT_ENTRY(); return launch_ew<OP_SCALE>(x, nullptr, x, n, alpha); }

int nt_add_rows(const float* x, const float* table, float* out, int rows, int period, int D) {
  NT_ENTRY();
  NT_REQUIRE(period > 0 && D > 0, "nt_add_rows: bad period/D");
  long long total = (long long)rows * D;
  if (total == 0) return 0;
  NT_CHECK(nt::launch_k(add_rows_kernel, dim3(grid_for(total)), dim3(256), 0, x, table, out, total, period, D));
  NT_LAUNCH_OK();
  return 0;
}

int nt_patchify(const float* images, float* out, int B, int C, int Hi, int Wi, int n_patch) {
  NT_ENTRY();
  NT_REQUIRE(n_patch > 0 && Hi >= n_patch && Wi >= n_patch, "nt_patchify: bad geometry");
  int hl = Hi / n_patch, wl = Wi / n_patch;
  long long total = (long long)B * n_patch * n_patch * C * hl * wl;
  if (total == 0) return 0;
  NT_CHECK(nt::launch_k(patchify_kernel, dim3(grid_for(total)), dim3(256), 0, images, out, B, C, Hi, Wi, n_patch, hl, wl));
  NT_LAUNCH_OK();
  return 0;
}

int nt_gather_images_u8(const unsigned char* src, const int32_t* perm, int first, const float* lut256, float* out, int B,
                        int pixels) {
  NT_ENTRY();
  NT_REQUIRE(src && perm && lut256 && out && first >= 0 && B >= 0 && pixels > 0, "nt_gather_images_u8: bad arguments");
  if (B == 0) return 0;
  NT_CHECK(nt::launch_k(gather_images_u8_kernel, dim3(grid_for((size_t)B * pixels)), dim3(256), 0, src, perm, first, lut256, out, B, pixels));
  NT_LAUNCH_OK();
  return 0;
}

int nt_gather_i32(const int32_t* src, const int32_t* perm, int first, int32_t* out, int B) {
  NT_ENTRY();
  NT_REQUIRE(src && perm && out && first >= 0 && B >= 0, "nt_gather_i32: bad arguments");
  if (B == 0) return 0;
  NT_CHECK(nt::launch_k(gather_i32_kernel, dim3(ceil_div(B, 256)), dim3(256), 0, src, perm, first, out, B));
  NT_LAUNCH_OK();
  return 0;
}

int nt_gather_windows(const int32_t* corpus, const int32_t* start, int32_t* inputs, int32_t* labels, int B, int T) {
  NT_ENTRY();
  NT_REQUIRE(corpus && start && inputs && labels && B >= 0 && T > 0, "nt_gather_windows: bad arguments");
  if (B == 0) return 0;
  NT_CHECK(nt::launch_k(gather_windows_kernel, dim3(grid_for((size_t)B * T)), dim3(256), 0, corpus, start, inputs, labels, B, T));
  NT_LAUNCH_OK();
  return 0;
}

static bool conv_geom(ConvGeom& g, int N, int C, int H, int W, int KH, int KW, int SH, int SW, int PH, int PW) {
  if (N < 0 || C <= 0 || H <= 0 || W <= 0 || KH <= 0 || KW <= 0 || SH <= 0 || SW <= 0 || PH < 0 || PW < 0) return false;
  if (H + 2 * PH < KH || W + 2 * PW < KW) return false;
  g = ConvGeom{N, C, H, W, KH, KW, SH, SW, PH, PW, (H + 2 * PH - KH) / SH + 1, (W + 2 * PW - KW) / SW + 1};
  return true;
}

int nt_im2col(const float* x, float* col, int N, int C, int H, int W, int KH, int KW, int SH, int SW, int PH, int PW) {
  NT_ENTRY();
  ConvGeom g;
  NT_REQUIRE(conv_geom(g, N, C, H, W, KH, KW, SH, SW, PH, PW), "nt_im2col: bad geometry");
  const long long total = (long long)N * g.OH * g.OW * C * KH * KW;
  if (total == 0) return 0;
  NT_CHECK(nt::launch_k(im2col_kernel, dim3(grid_for(total)), dim3(256), 0, x, col, g));
  NT_LAUNCH_OK();
  return 0;
}

int nt_col2im(const float* dcol, float* dx, int N, int C, int H, int W, int KH, int KW, int SH, int SW, int PH, int PW) {
  NT_ENTRY();
  ConvGeom g;
  NT_REQUIRE(conv_geom(g, N, C, H, W, KH, KW, SH, SW, PH, PW), "nt_col2im: bad geometry");
  const long long total = (long long)N * C * H * W;
  if (total == 0) return 0;
  NT_CHECK(nt::launch_k(col2im_kernel, dim3(grid_for(total)), dim3(256), 0, dcol, dx, g));
  NT_LAUNCH_OK();
  return 0;
}

int nt_transpose_last2(const float* x, float* y, int batch, int rows, int cols) {
  NT_ENTRY();
  NT_REQUIRE(batch >= 0 && rows >= 0 && cols >= 0 && batch <= 65535, "nt_transpose_last2: bad shape");
  if ((long long)batch * rows * cols == 0) return 0;
  NT_CHECK(nt::launch_k(transpose_last2_kernel, dim3(ceil_div(cols, 32), ceil_div(rows, 32), batch), dim3(32, 8), 0, x, y, rows, cols));
  NT_LAUNCH_OK();
  return 0;
}

int nt_embedding_fwd(const int32_t* idx, const float* etok, const float* epos, float* out, int B, int T, int D, int vocab) {
  NT_ENTRY();
  NT_REQUIRE(vocab > 0, "nt_embedding_fwd: num_embeddings must be positive");
  if (B * T == 0) return 0;
  NT_FUSE_TRY(nt::fuse_embed_fwd(idx, etok, epos, out, B, T, D, vocab));
  NT_CHECK(nt::launch_k(embedding_fwd_kernel, dim3(B * T), dim3(D >= 256 ? 256 : 128), 0, idx, etok, epos, out, B * T, T, D, vocab,
                        g_dev_error));
  NT_LAUNCH_OK();
  return 0;
}

int nt_embedding_bwd(const int32_t* idx, const float* dy, float* detok, float* depos, int B, int T, int D, int vocab) {
  NT_ENTRY();
  if (B * T == 0) return 0;
  NT_FUSE_TRY(nt::fuse_embed_bwd(idx, dy, detok, depos, B, T, D, vocab));
  if (detok) {
    NT_REQUIRE(vocab > 0, "nt_embedding_bwd: num_embeddings must be positive");
    NT_CHECK(nt::launch_k(embedding_bwd_tok_kernel, dim3(B * T), dim3(D >= 256 ? 256 : 128), 0, idx, dy, detok, B * T, D, vocab,
                          g_dev_error));
    NT_LAUNCH_OK();
  }
  if (depos) {
    NT_CHECK(nt::launch_k(embedding_bwd_pos_kernel, dim3(ceil_div((long long)T * D, 256)), dim3(256), 0, dy, depos, B, T, D));
    NT_LAUNCH_OK();
  }
  return 0;
}

int nt_sgd(float* p, float* g, size_t n, float lr, const float* lr_dev, int zero_grad) {
  NT_ENTRY();
  if (n == 0) return 0;
  NT_CHECK(nt::launch_k(sgd_kernel, dim3(grid_for(n)), dim3(256), 0, p, g, n, lr, lr_dev, zero_grad));
  NT_LAUNCH_OK();
  return 0;
}

int nt_adam_star(float* p, float* g, float* m, float* v, size_t n, float lr, const float* lr_dev, int t,
                 const int32_t* t_dev, int zero_grad) {
  NT_ENTRY();
  if (n == 0) return 0;
  NT_CHECK(nt::launch_k(adam_star_kernel, dim3(grid_for(n)), dim3(256), 0, p, g, m, v, n, lr, lr_dev, t, t_dev, zero_grad));
  NT_LAUNCH_OK();
  return 0;
}

int nt_decode_embed(const int32_t* idx, const float* etok, const float* epos, float* out, int B, int D, int Tmax,
                    const int32_t* d_pos, int vocab) {
  NT_ENTRY();
  NT_REQUIRE(idx && etok && epos && out && d_pos && D > 0 && Tmax > 0 && vocab > 0, "nt_decode_embed: bad arguments");
  if (B == 0) return 0;
  NT_FUSE_TRY(nt::fuse_decode_embed(idx, etok, epos, out, B, D, Tmax, d_pos, vocab));
  NT_CHECK(nt::launch_k(decode_embed_kernel, dim3(B), dim3(D >= 256 ? 256 : 128), 0, idx, etok, epos, out, D, Tmax, d_pos, vocab,
                        g_dev_error));
  NT_LAUNCH_OK();
  return 0;
}

int nt_argmax_next(const float* logits, int ld, int cols, int B, int32_t* tok, int32_t* hist, int hist_ld, const int32_t* d_pos) {
  NT_ENTRY();
  NT_REQUIRE(logits && tok && d_pos && cols > 0 && ld >= cols, "nt_argmax_next: bad arguments");
  if (B == 0) return 0;
  NT_FUSE_TRY(nt::fuse_argmax_next(logits, ld, cols, B, tok, hist, hist_ld, d_pos));
  NT_CHECK(nt::launch_k(argmax_next_kernel, dim3(B), dim3(256), 0, logits, ld, cols, tok, hist, hist_ld, d_pos));
  NT_LAUNCH_OK();
  return 0;
}

int nt_increment(int32_t* counter) {
  NT_ENTRY();
  NT_FUSE_TRY(nt::fuse_increment(counter));
  NT_CHECK(nt::launch_k(increment_kernel, dim3(1), dim3(1), 0, counter));
  NT_LAUNCH_OK();
  return 0;
}

}  // extern "C"
