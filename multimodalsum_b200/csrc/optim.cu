// Fused gradient-norm + clip + AdamW over the flat parameter arenas (HBM-bound, 30 B per parameter).
//   clip:  torch.nn.utils.clip_grad_norm_(params, max_norm)            src/multimodal_train.py:361-362
//   step:  transformers 3.0.2 AdamW.step                               src/transformer/optimization.py:208-267
//          (bias-corrected, eps added to sqrt(v), decoupled decay applied AFTER the Adam update with lr*wd)
// Per-64-element flags select which arena blocks are updated / decayed (param groups; reference quirk Q1 leaves the
// no-decay group empty so those tensors are never touched).  The bf16 compute copy is written in the same pass.
#include "common.cuh"
#include "../../include/mmsum_b200.h"

namespace mmsum {

__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ partial) {
  __shared__ float red[8];
  float s = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 4 <= n) { const float4 v = *reinterpret_cast<const float4*>(g + i); s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w; }
    else for (long long k = i; k < n; ++k) s += g[k] * g[k];
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = red[threadIdx.x];
    v += __shfl_xor_sync(0xffu, v, 4); v += __shfl_xor_sync(0xffu, v, 2); v += __shfl_xor_sync(0xffu, v, 1);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
  }
}
__global__ void __launch_bounds__(1024) sumsq_final_kernel(const float* __restrict__ partial, int n, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 1024) s += partial[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) { float v = red[threadIdx.x]; v = warp_sum(v); if (threadIdx.x == 0) out[0] = v; }
}

// flags[i / 64]: bit0 = update this block, bit1 = apply weight decay
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ w, bf16* __restrict__ w16, const float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v, const uint8_t* __restrict__ flags,
                                                    long long n, float lr, float beta1, float beta2, float eps, float wd,
                                                    float step_size, const float* __restrict__ sumsq, float max_norm) {
  float clip = 1.f;
  if (sumsq != nullptr && max_norm > 0.f) {
    const float c = max_norm / (sqrtf(sumsq[0]) + 1e-6f);
    clip = c < 1.f ? c : 1.f;
  }
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    const uint8_t f = flags[i >> 6];
    float4 wv = *reinterpret_cast<const float4*>(w + i);
    if (f & 1) {
      const float4 gv = *reinterpret_cast<const float4*>(g + i);
      float4 mv = *reinterpret_cast<const float4*>(m + i);
      float4 vv = *reinterpret_cast<const float4*>(v + i);
      float* wp = &wv.x; const float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float gk = gp[k] * clip;
        mp[k] = beta1 * mp[k] + (1.f - beta1) * gk;
        vp[k] = beta2 * vp[k] + (1.f - beta2) * gk * gk;
        wp[k] -= step_size * mp[k] / (sqrtf(vp[k]) + eps);
        if (f & 2) wp[k] -= lr * wd * wp[k];
      }
      *reinterpret_cast<float4*>(w + i) = wv;
      *reinterpret_cast<float4*>(m + i) = mv;
      *reinterpret_cast<float4*>(v + i) = vv;
    }
    uint2 o; o.x = pack_bf16(wv.x, wv.y); o.y = pack_bf16(wv.z, wv.w);
    *reinterpret_cast<uint2*>(w16 + i) = o;
  }
}

}  // namespace mmsum
using namespace mmsum;

extern "C" int mmsum_grad_sumsq(const float* g, int64_t n, float* partial, int32_t n_partial, float* out, void* stream) {
  if (!g || !partial || !out || n <= 0 || n_partial <= 0) return MMSUM_ERR_INVALID;
  int blocks = n_partial < 148 * 8 ? n_partial : 148 * 8;
  grad_sumsq_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(g, n, partial);
  MMSUM_CHECK_LAUNCH();
  sumsq_final_kernel<<<1, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(partial, blocks, out);
  MMSUM_CHECK_LAUNCH();
  return 0;
}

extern "C" int mmsum_adamw_step(float* w32, void* w16, const float* g, float* m, float* v, const uint8_t* flags, int64_t n,
                                float lr, float beta1, float beta2, float eps, float weight_decay, float step_size,
                                const float* sumsq, float max_norm, void* stream) {
  if (!w32 || !w16 || !g || !m || !v || !flags || n <= 0 || (n % 64)) return MMSUM_ERR_INVALID;
  adamw_kernel<<<148 * 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(w32, reinterpret_cast<bf16*>(w16), g, m, v, flags, n, lr,
                                                                            beta1, beta2, eps, weight_decay, step_size, sumsq, max_norm);
  MMSUM_CHECK_LAUNCH();
  return 0;
}
