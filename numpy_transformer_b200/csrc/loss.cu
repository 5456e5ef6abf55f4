// Cross-entropy (net/loss.py:46-51) fused with argmax of the returned softmax
// (transformer_of_image.py:151, gpt_train_potry3000.py:290): one CTA per row.
#include "common.cuh"
#include <cmath>

namespace nt {

__device__ __forceinline__ void block_argmax(float& v, int& idx, float* sv, int* si) {
  // (value, index) max with first-index tie-break, matching np.argmax
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) { sv[warp] = v; si[warp] = idx; }
  __syncthreads();
  if (warp == 0) {
    v = lane < nw ? sv[lane] : -INFINITY;
    idx = lane < nw ? si[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, v, o);
      int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    if (lane == 0) { sv[0] = v; si[0] = idx; }
  }
  __syncthreads();
  v = sv[0];
  idx = si[0];
  __syncthreads();
}

__device__ __forceinline__ float block_sum(float v, float* sv) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) sv[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < nw ? sv[lane] : 0.f;
    v = warp_sum(v);
    if (lane == 0) sv[0] = v;
  }
  __syncthreads();
  v = sv[0];
  __syncthreads();
  return v;
}

__global__ void __launch_bounds__(128) ce_row_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels,
                                                     const float* __restrict__ onehot, float inv_n,
                                                     float* __restrict__ row_loss, float* __restrict__ grad,
                                                     float* __restrict__ prob, int32_t* __restrict__ argmax_out,
                                                     int cols, int* err) {
  pdl_prologue();
  __shared__ float sv[4];
  __shared__ int si[4];
  const long long row = blockIdx.x;
  const float* lr = logits + row * cols;
  float mx = -INFINITY;
  int mi = 0x7fffffff;
  for (int j = threadIdx.x; j < cols; j += blockDim.x) {
    float v = lr[j];
    if (v > mx) { mx = v; mi = j; }
  }
  block_argmax(mx, mi, sv, si);
  float s = 0.f;
  for (int j = threadIdx.x; j < cols; j += blockDim.x) s += expf(lr[j] - mx);
  s = block_sum(s, sv);
  const float inv = 1.0f / s;
  const int lab = labels ? labels[row] : -1;
  if (labels && (lab < 0 || lab >= cols) && threadIdx.x == 0) *err = NT_DEVERR_LABEL;   // no class to credit: flagged, loss 0
  float lossacc = 0.f;
  float pm = -INFINITY;
  int pi = 0x7fffffff;
  for (int j = threadIdx.x; j < cols; j += blockDim.x) {
    const float p = expf(lr[j] - mx) * inv;
    const float y = onehot ? onehot[row * cols + j] : (j == lab ? 1.f : 0.f);
    if (prob) prob[row * cols + j] = p;
    if (grad) grad[row * cols + j] = (p - y) * inv_n;
    if (y != 0.f) lossacc -= y * logf(p + 1e-10f);
    if (p > pm) { pm = p; pi = j; }            // strided scan keeps the smallest j on ties per thread
  }
  lossacc = block_sum(lossacc, sv);
  block_argmax(pm, pi, sv, si);                // argmax over the probabilities that are returned
  if (threadIdx.x == 0) {
    row_loss[row] = lossacc;
    if (argmax_out) argmax_out[row] = pi;
  }
}

// deterministic single-block reduction of the per-row losses
__global__ void __launch_bounds__(256) ce_reduce_kernel(const float* __restrict__ row_loss, float* __restrict__ loss,
                                                        int rows, float inv_n) {
  pdl_prologue();
  __shared__ float sv[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) s += row_loss[i];
  s = block_sum(s, sv);
  if (threadIdx.x == 0) *loss = s * inv_n;
}

// Element-wise losses of net/loss.py:53-57 (BCE on logits) and :68-71 (MSE): gradient (and sigmoid) per element, the
// scalar as per-CTA partial sums folded with one atomic each into a zeroed accumulator.
template <int KIND>      // 0: BCE with logits, 1: MSE
__global__ void __launch_bounds__(256) elem_loss_kernel(const float* __restrict__ pred, const float* __restrict__ label,
                                                        float* __restrict__ loss, float* __restrict__ grad,
                                                        float* __restrict__ prob, size_t n, float inv_n) {
  pdl_prologue();
  __shared__ float sv[8];
  float acc = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float p = pred[i], y = label[i];
    if (KIND == 0) {
      const float s = 1.0f / (1.0f + expf(-p));
      acc -= y * logf(s + 1e-10f) + (1.0f - y) * logf((1.0f - s) + 1e-10f);
      if (grad) grad[i] = (s - y) * inv_n;
      if (prob) prob[i] = s;
    } else {
      const float d = p - y;
      acc += d * d;
      if (grad) grad[i] = 2.0f * d * inv_n;
    }
  }
  acc = block_sum(acc, sv);
  if (threadIdx.x == 0) atomicAdd(loss, acc * inv_n);
}

template <int KIND>
static int elem_loss(const float* pred, const float* label, float* loss, float* grad, float* prob, size_t n) {
  NT_FUSE_FLUSH();
  NT_CHECK(cudaMemsetAsync(loss, 0, sizeof(float), g_stream));
  if (n == 0) return 0;
  size_t blocks = (n + 256 * 8 - 1) / (256 * 8);
  if (blocks > (size_t)g_sm_count * 4) blocks = (size_t)g_sm_count * 4;
  NT_CHECK(nt::launch_k(elem_loss_kernel<KIND>, dim3((unsigned)blocks), dim3(256), 0, pred, label, loss, grad, prob, n, 1.0f / (float)n));
  NT_LAUNCH_OK();
  return 0;
}

}  // namespace nt
using namespace nt;

extern "C" int nt_bce_logit_loss(const float* logits, const float* label, float* loss, float* grad, float* sigmoid, size_t n) {
  NT_ENTRY();
  NT_REQUIRE(logits && label && loss, "nt_bce_logit_loss: logits, label and loss are required");
  return elem_loss<0>(logits, label, loss, grad, sigmoid, n);
}

extern "C" int nt_mse_loss(const float* predict, const float* label, float* loss, float* grad, size_t n) {
  NT_ENTRY();
  NT_REQUIRE(predict && label && loss, "nt_mse_loss: predict, label and loss are required");
  return elem_loss<1>(predict, label, loss, grad, nullptr, n);
}

extern "C" int nt_cross_entropy(const float* logits, const int32_t* labels, const float* onehot, float n_norm,
                                float* loss, float* row_loss, float* grad, float* prob, int32_t* argmax, int rows,
                                int cols) {
  NT_ENTRY();
  NT_REQUIRE(rows > 0 && cols > 0, "nt_cross_entropy: bad shape rows=%d cols=%d", rows, cols);
  NT_REQUIRE((labels != nullptr) != (onehot != nullptr), "nt_cross_entropy: pass exactly one of labels / onehot");
  NT_REQUIRE(n_norm > 0.f && row_loss && loss, "nt_cross_entropy: n_norm, row_loss and loss are required");
  NT_FUSE_TRY(nt::fuse_cross_entropy(logits, labels, onehot, n_norm, loss, row_loss, grad, prob, argmax, rows, cols));
  const float inv_n = 1.0f / n_norm;
  NT_CHECK(nt::launch_k(ce_row_kernel, dim3(rows), dim3(128), 0, logits, labels, onehot, inv_n, row_loss, grad, prob, argmax, cols,
                        g_dev_error));
  NT_LAUNCH_OK();
  NT_CHECK(nt::launch_k(ce_reduce_kernel, dim3(1), dim3(256), 0, row_loss, loss, rows, inv_n));
  NT_LAUNCH_OK();
  return 0;
}
