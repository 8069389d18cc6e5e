// Multi-tensor AdamW: ONE launch updates every parameter of the model.
//
// Replaces torch.optim.AdamW as built by the reference's optimizer factory (open_clip_train/optim.py:453-454,
// decoupled weight decay, per-group lr / weight_decay) with the arithmetic of torch's fused kernel
// (ATen fused_adam_utils.cuh, ADAMW mode, no amsgrad / maximize / grad scaler), so the parameter trajectory is the
// reference's: under `--precision bf16` parameters, gradients and both moments are bf16 tensors, every update is
// computed in fp32 and rounded once per tensor element (round-to-nearest-even), fp32 parameters stay fp32 throughout.
//
// The whole tensor table (<= 512 tensors: pointers, sizes, per-tensor lr / weight decay) travels in the kernel
// parameter block (CUDA >= 12.1: up to 32 KB), so there is no table upload and no per-tensor launch; a block finds its
// tensor by binary search over the prefix sum of chunk counts.  HBM traffic = the algorithmic minimum: read p, g, m,
// v once, write p, m, v once (14 B per bf16 parameter, 28 B per fp32 parameter).
#include "common.cuh"

namespace clipn {

constexpr int kAdamMaxTensors = CLIPN_ADAMW_MAX_TENSORS;
constexpr int kAdamChunk = 8192;  // elements per block: 256 threads x 4 iterations x 8 elements

struct AdamTable {
  void* p[kAdamMaxTensors];
  const void* g[kAdamMaxTensors];
  void* m[kAdamMaxTensors];
  void* v[kAdamMaxTensors];
  int chunk_prefix[kAdamMaxTensors + 1];  // first chunk of tensor i; [n] = total chunks
  int numel[kAdamMaxTensors];
  float lr[kAdamMaxTensors];
  float wd[kAdamMaxTensors];
  unsigned char is_bf16[kAdamMaxTensors];
  int n;
  float beta1, beta2, eps, bc1, bc2_sqrt;
};
static_assert(sizeof(AdamTable) <= 32000, "AdamW table must fit the 32 KB kernel parameter block");

__device__ __forceinline__ void adamw_math(float& p, float g, float& m, float& v, float lr, float wd, float w1, float beta2,
                                           float w2, float eps, float step_size, float inv_bc2_sqrt) {
  p -= lr * wd * p;                    // decoupled weight decay
  m = m + w1 * (g - m);                // lerp(m, g, 1 - beta1)
  v = beta2 * v + w2 * g * g;
  const float denom = sqrtf(v) * inv_bc2_sqrt + eps;
  p -= step_size * m / denom;
}

__global__ void __launch_bounds__(256) adamw_multi_kernel(const __grid_constant__ AdamTable t) {
  // which tensor does this chunk belong to? (upper bound over chunk_prefix)
  const int chunk = blockIdx.x;
  int lo = 0, hi = t.n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (t.chunk_prefix[mid] <= chunk) lo = mid; else hi = mid;
  }
  const int ti = lo;
  const int64_t base = static_cast<int64_t>(chunk - t.chunk_prefix[ti]) * kAdamChunk;
  const int n = t.numel[ti];
  const float lr = t.lr[ti], wd = t.wd[ti];
  const float w1 = 1.f - t.beta1, w2 = 1.f - t.beta2;
  const float step_size = lr / t.bc1, inv_bc2 = 1.f / t.bc2_sqrt;
  if (t.is_bf16[ti]) {
    __nv_bfloat16* P = reinterpret_cast<__nv_bfloat16*>(t.p[ti]);
    const __nv_bfloat16* G = reinterpret_cast<const __nv_bfloat16*>(t.g[ti]);
    __nv_bfloat16* M = reinterpret_cast<__nv_bfloat16*>(t.m[ti]);
    __nv_bfloat16* V = reinterpret_cast<__nv_bfloat16*>(t.v[ti]);
    const bool vec = ((reinterpret_cast<uintptr_t>(P) | reinterpret_cast<uintptr_t>(G) | reinterpret_cast<uintptr_t>(M) |
                       reinterpret_cast<uintptr_t>(V)) & 15) == 0;
#pragma unroll 1
    for (int it = 0; it < kAdamChunk / (256 * 8); ++it) {
      const int64_t i0 = base + (static_cast<int64_t>(it) * 256 + threadIdx.x) * 8;
      if (i0 >= n) break;
      if (vec && i0 + 8 <= n) {
        float p[8], g[8], m[8], v[8];
        unpack_bf16x8(*reinterpret_cast<const uint4*>(P + i0), p);
        unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(G + i0)), g);
        unpack_bf16x8(*reinterpret_cast<const uint4*>(M + i0), m);
        unpack_bf16x8(*reinterpret_cast<const uint4*>(V + i0), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) adamw_math(p[j], g[j], m[j], v[j], lr, wd, w1, t.beta2, w2, t.eps, step_size, inv_bc2);
        *reinterpret_cast<uint4*>(P + i0) = pack_bf16x8(p);
        *reinterpret_cast<uint4*>(M + i0) = pack_bf16x8(m);
        *reinterpret_cast<uint4*>(V + i0) = pack_bf16x8(v);
      } else {
        for (int64_t i = i0; i < i0 + 8 && i < n; ++i) {
          float p = __bfloat162float(P[i]), g = __bfloat162float(G[i]), m = __bfloat162float(M[i]), v = __bfloat162float(V[i]);
          adamw_math(p, g, m, v, lr, wd, w1, t.beta2, w2, t.eps, step_size, inv_bc2);
          P[i] = __float2bfloat16_rn(p); M[i] = __float2bfloat16_rn(m); V[i] = __float2bfloat16_rn(v);
        }
      }
    }
  } else {
    float* P = reinterpret_cast<float*>(t.p[ti]);
    const float* G = reinterpret_cast<const float*>(t.g[ti]);
    float* M = reinterpret_cast<float*>(t.m[ti]);
    float* V = reinterpret_cast<float*>(t.v[ti]);
    const bool vec = ((reinterpret_cast<uintptr_t>(P) | reinterpret_cast<uintptr_t>(G) | reinterpret_cast<uintptr_t>(M) |
                       reinterpret_cast<uintptr_t>(V)) & 15) == 0;
#pragma unroll 1
    for (int it = 0; it < kAdamChunk / (256 * 4); ++it) {
      const int64_t i0 = base + (static_cast<int64_t>(it) * 256 + threadIdx.x) * 4;
      if (i0 >= n) break;
      if (vec && i0 + 4 <= n) {
        float4 p = *reinterpret_cast<const float4*>(P + i0), g = __ldg(reinterpret_cast<const float4*>(G + i0));
        float4 m = *reinterpret_cast<const float4*>(M + i0), v = *reinterpret_cast<const float4*>(V + i0);
        adamw_math(p.x, g.x, m.x, v.x, lr, wd, w1, t.beta2, w2, t.eps, step_size, inv_bc2);
        adamw_math(p.y, g.y, m.y, v.y, lr, wd, w1, t.beta2, w2, t.eps, step_size, inv_bc2);
        adamw_math(p.z, g.z, m.z, v.z, lr, wd, w1, t.beta2, w2, t.eps, step_size, inv_bc2);
        adamw_math(p.w, g.w, m.w, v.w, lr, wd, w1, t.beta2, w2, t.eps, step_size, inv_bc2);
        *reinterpret_cast<float4*>(P + i0) = p;
        *reinterpret_cast<float4*>(M + i0) = m;
        *reinterpret_cast<float4*>(V + i0) = v;
      } else {
        for (int64_t i = i0; i < i0 + 4 && i < n; ++i) {
          float p = P[i], g = G[i], m = M[i], v = V[i];
          adamw_math(p, g, m, v, lr, wd, w1, t.beta2, w2, t.eps, step_size, inv_bc2);
          P[i] = p; M[i] = m; V[i] = v;
        }
      }
    }
  }
}

}  // namespace clipn

using namespace clipn;

extern "C" int clipn_adamw_multi(const clipn_adamw_tensor* tensors, int32_t n, float beta1, float beta2, float eps,
                                 float bias_correction1, float bias_correction2_sqrt, clipn_stream_t stream) {
  CLIPN_REQUIRE(tensors != nullptr && n > 0, "adamw: empty tensor list");
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int first = 0; first < n; first += kAdamMaxTensors) {  // > 512 tensors: one launch per 512
    AdamTable t;
    const int cnt = (n - first) < kAdamMaxTensors ? (n - first) : kAdamMaxTensors;
    int chunks = 0;
    for (int i = 0; i < cnt; ++i) {
      const clipn_adamw_tensor& a = tensors[first + i];
      CLIPN_REQUIRE(a.param && a.grad && a.exp_avg && a.exp_avg_sq, "adamw: null tensor pointer");
      CLIPN_REQUIRE(a.numel > 0 && a.numel < (1ll << 31), "adamw: tensor size out of range");
      t.p[i] = a.param; t.g[i] = a.grad; t.m[i] = a.exp_avg; t.v[i] = a.exp_avg_sq;
      t.numel[i] = static_cast<int>(a.numel);
      t.lr[i] = a.lr; t.wd[i] = a.weight_decay;
      t.is_bf16[i] = a.is_bf16 ? 1 : 0;
      t.chunk_prefix[i] = chunks;
      chunks += static_cast<int>((a.numel + kAdamChunk - 1) / kAdamChunk);
    }
    t.chunk_prefix[cnt] = chunks;
    t.n = cnt;
    t.beta1 = beta1; t.beta2 = beta2; t.eps = eps; t.bc1 = bias_correction1; t.bc2_sqrt = bias_correction2_sqrt;
    adamw_multi_kernel<<<chunks, 256, 0, st>>>(t);
    CLIPN_CHECK_CUDA(cudaGetLastError());
  }
  return CLIPN_OK;
}
