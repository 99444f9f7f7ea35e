// ffn_train.cuh -- one optimisation step of the 39-64-32-16-3 FFN on the device.
//
// Reference: learning/ffn_trainer.py:104-120,139-148 -- Keras-1 Sequential(Dense64 relu relu Dense32 relu
// Dense16 relu Dense3 softmax), loss categorical_crossentropy, optimizer 'adadelta' (Keras-1 defaults lr 1.0,
// rho 0.95, epsilon 1e-8), model.train_on_batch(batch_x, batch_y).  Restated semantics:
//   p = softmax(logits);  loss = -mean_i sum_c y_ic log(clip(p_ic, 1e-7, 1 - 1e-7))
//   g = dloss/dtheta (clip inactive => dlogits = (p - y) / B)
//   a <- rho a + (1 - rho) g^2;  u = g sqrt(d + eps) / sqrt(a + eps);  theta <- theta - lr u;  d <- rho d + (1 - rho) u^2
//
// Two kernels per step, both deterministic (no floating-point atomics):
//   ffn_train_grad_kernel   one CTA = 128 batch rows.  Forward and backward per row (thread = row, weights broadcast
//                           from shared memory, activations / deltas parked in shared memory with odd pitches), then
//                           the weight gradients of the tile as K x N dot products over the 128 rows (thread = output
//                           entries) -> per-CTA partial gradient + partial loss in global memory.
//   ffn_train_update_kernel thread = parameter: sums the partials in CTA order, applies Adadelta in place.
#pragma once

#include <cuda_runtime.h>

#include "vad_core.cuh"

namespace vadb {

constexpr int kTrainRows = 128;                        // batch rows per CTA
constexpr int kNParams = kNFeat * kH1 + kH1 + kH1 * kH2 + kH2 + kH2 * kH3 + kH3 + kH3 * kNCls + kNCls;  // 5219
// parameter vector layout == FfnParams: W1 b1 W2 b2 W3 b3 W4 b4
constexpr int kOffW1 = 0, kOffB1 = kOffW1 + kNFeat * kH1, kOffW2 = kOffB1 + kH1, kOffB2 = kOffW2 + kH1 * kH2,
              kOffW3 = kOffB2 + kH2, kOffB3 = kOffW3 + kH2 * kH3, kOffW4 = kOffB3 + kH3, kOffB4 = kOffW4 + kH3 * kNCls;
static_assert(kOffB4 + kNCls == kNParams, "layout");
// shared-memory pitches (floats), all odd: thread r touching [r * pitch + k] is bank-conflict free
constexpr int kPX = 39, kPH1 = 65, kPH2 = 33, kPH3 = 17, kPD4 = 5;
constexpr int kTrainSmemFloats = kNParams + 1 + kTrainRows * (kPX + 2 * kPH1 + 2 * kPH2 + 2 * kPH3 + kPD4) + 8;
constexpr int kTrainSmemBytes = kTrainSmemFloats * 4;

// dW[k][n] = sum_r a[r][k] d[r][n] (and db[n] = sum_r d[r][n]) of one layer -> partial gradient of this CTA
template <int K, int N, int PA, int PD>
__device__ __forceinline__ void train_layer_grad(const float* a, const float* d, int rows, float* gW, float* gb) {
  for (int o = threadIdx.x; o < K * N; o += kTrainRows) {
    const int k = o / N, n = o - k * N;
    float s0 = 0.0f, s1 = 0.0f;
    int r = 0;
    for (; r + 1 < rows; r += 2) {
      s0 = fmaf(a[r * PA + k], d[r * PD + n], s0);
      s1 = fmaf(a[(r + 1) * PA + k], d[(r + 1) * PD + n], s1);
    }
    if (r < rows) s0 = fmaf(a[r * PA + k], d[r * PD + n], s0);
    gW[o] = s0 + s1;
  }
  for (int n = threadIdx.x; n < N; n += kTrainRows) {
    float s = 0.0f;
    for (int r = 0; r < rows; ++r) s += d[r * PD + n];
    gb[n] = s;
  }
}

__global__ void __launch_bounds__(kTrainRows) ffn_train_grad_kernel(const float* params, const float* x,
                                                                    const uint8_t* y, long long n_rows, float inv_b,
                                                                    float* partial /*[grid][kNParams + 1]*/) {
  extern __shared__ __align__(16) float sm[];
  float* w = sm;                                   // parameters
  float* sx = w + kNParams + 1;                    // [rows][39]
  float* sh1 = sx + kTrainRows * kPX;              // activations (post-ReLU)
  float* sh2 = sh1 + kTrainRows * kPH1;
  float* sh3 = sh2 + kTrainRows * kPH2;
  float* sd1 = sh3 + kTrainRows * kPH3;            // deltas (dloss / dpre-activation)
  float* sd2 = sd1 + kTrainRows * kPH1;
  float* sd3 = sd2 + kTrainRows * kPH2;
  float* sd4 = sd3 + kTrainRows * kPH3;
  float* sloss = sd4 + kTrainRows * kPD4;
  const int tid = threadIdx.x;
  for (int i = tid; i < kNParams; i += kTrainRows) w[i] = params[i];
  const long long row0 = static_cast<long long>(blockIdx.x) * kTrainRows;
  const int rows = static_cast<int>(min(static_cast<long long>(kTrainRows), n_rows - row0));
  for (int i = tid; i < rows * kNFeat; i += kTrainRows) {            // coalesced tile load
    const int r = i / kNFeat, k = i - r * kNFeat;
    sx[r * kPX + k] = x[(row0 + r) * kNFeat + k];
  }
  __syncthreads();
  float loss = 0.0f;
  if (tid < rows) {
    const int r = tid;
    // ---- forward (learning/ffn_trainer.py:106-116) ----
    for (int o = 0; o < kH1; ++o) {
      float s = w[kOffB1 + o];
#pragma unroll 13
      for (int i = 0; i < kNFeat; ++i) s = fmaf(sx[r * kPX + i], w[kOffW1 + i * kH1 + o], s);
      sh1[r * kPH1 + o] = fmaxf(s, 0.0f);
    }
    for (int o = 0; o < kH2; ++o) {
      float s = w[kOffB2 + o];
#pragma unroll 16
      for (int i = 0; i < kH1; ++i) s = fmaf(sh1[r * kPH1 + i], w[kOffW2 + i * kH2 + o], s);
      sh2[r * kPH2 + o] = fmaxf(s, 0.0f);
    }
    for (int o = 0; o < kH3; ++o) {
      float s = w[kOffB3 + o];
#pragma unroll 16
      for (int i = 0; i < kH2; ++i) s = fmaf(sh2[r * kPH2 + i], w[kOffW3 + i * kH3 + o], s);
      sh3[r * kPH3 + o] = fmaxf(s, 0.0f);
    }
    float lg[kNCls];
#pragma unroll
    for (int o = 0; o < kNCls; ++o) {
      float s = w[kOffB4 + o];
#pragma unroll
      for (int i = 0; i < kH3; ++i) s = fmaf(sh3[r * kPH3 + i], w[kOffW4 + i * kNCls + o], s);
      lg[o] = s;
    }
    // ---- softmax + categorical cross-entropy ----
    const float mx = fmaxf(lg[0], fmaxf(lg[1], lg[2]));
    const float e0 = expf(lg[0] - mx), e1 = expf(lg[1] - mx), e2 = expf(lg[2] - mx);
    const float inv = 1.0f / (e0 + e1 + e2);
    const float p[kNCls] = {e0 * inv, e1 * inv, e2 * inv};
    const int cls = y[row0 + r];
    const float pc = fminf(fmaxf(cls == 0 ? p[0] : cls == 1 ? p[1] : p[2], 1e-7f), 1.0f - 1e-7f);
    loss = -logf(pc);
    // ---- backward ----
    float d4[kNCls];
#pragma unroll
    for (int o = 0; o < kNCls; ++o) {
      d4[o] = (p[o] - (o == cls ? 1.0f : 0.0f)) * inv_b;
      sd4[r * kPD4 + o] = d4[o];
    }
    for (int i = 0; i < kH3; ++i) {
      float s = 0.0f;
#pragma unroll
      for (int o = 0; o < kNCls; ++o) s = fmaf(d4[o], w[kOffW4 + i * kNCls + o], s);
      sd3[r * kPH3 + i] = sh3[r * kPH3 + i] > 0.0f ? s : 0.0f;
    }
    for (int i = 0; i < kH2; ++i) {
      float s = 0.0f;
#pragma unroll
      for (int o = 0; o < kH3; ++o) s = fmaf(sd3[r * kPH3 + o], w[kOffW3 + i * kH3 + o], s);
      sd2[r * kPH2 + i] = sh2[r * kPH2 + i] > 0.0f ? s : 0.0f;
    }
    for (int i = 0; i < kH1; ++i) {
      float s = 0.0f;
#pragma unroll 16
      for (int o = 0; o < kH2; ++o) s = fmaf(sd2[r * kPH2 + o], w[kOffW2 + i * kH2 + o], s);
      sd1[r * kPH1 + i] = sh1[r * kPH1 + i] > 0.0f ? s : 0.0f;
    }
  }
  // block loss (fixed order: warp shuffles, then the four warp sums)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) loss += __shfl_down_sync(0xffffffffu, loss, o);
  if ((tid & 31) == 0) sloss[tid >> 5] = loss;
  __syncthreads();
  float* g = partial + static_cast<long long>(blockIdx.x) * (kNParams + 1);
  train_layer_grad<kNFeat, kH1, kPX, kPH1>(sx, sd1, rows, g + kOffW1, g + kOffB1);
  train_layer_grad<kH1, kH2, kPH1, kPH2>(sh1, sd2, rows, g + kOffW2, g + kOffB2);
  train_layer_grad<kH2, kH3, kPH2, kPH3>(sh2, sd3, rows, g + kOffW3, g + kOffB3);
  train_layer_grad<kH3, kNCls, kPH3, kPD4>(sh3, sd4, rows, g + kOffW4, g + kOffB4);
  if (tid == 0) g[kNParams] = ((sloss[0] + sloss[1]) + (sloss[2] + sloss[3])) * inv_b;
}

// state: [0] parameters, [1] accumulated squared gradients a, [2] accumulated squared updates d (kNParams each)
__global__ void ffn_train_update_kernel(float* params, float* acc_g, float* acc_u, const float* partial, int n_part,
                                        float lr, float rho, float eps, float* loss_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > kNParams) return;
  float g = 0.0f;
  for (int c = 0; c < n_part; ++c) g += partial[static_cast<long long>(c) * (kNParams + 1) + i];
  if (i == kNParams) {
    if (loss_out) *loss_out = g;
    return;
  }
  const float a = rho * acc_g[i] + (1.0f - rho) * g * g;
  const float u = g * sqrtf(acc_u[i] + eps) / sqrtf(a + eps);
  params[i] -= lr * u;
  acc_g[i] = a;
  acc_u[i] = rho * acc_u[i] + (1.0f - rho) * u * u;
}

}  // namespace vadb
