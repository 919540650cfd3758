// Fused parameter update (LAMB / AdamW + EMA) over all parameters of the model in one or two launches.
// HBM-bound streaming kernels: one CTA per <= 8192-element chunk of one parameter tensor, 16-byte accesses when the
// chunk is 16-byte aligned.  Every reduction (global gradient norm, per-tensor norms) is DETERMINISTIC: CTAs write
// partial sums, a small kernel adds them in a fixed order -- data-parallel replicas that hold bit-identical gradients
// after the all-reduce therefore stay bit-identical (float atomics would let them drift by an ulp per step).
// Algorithmic bytes per parameter element: AdamW 16 read + 12 write (+8 with EMA); LAMB stage 1 16 + 12, stage 2
// 8 + 4 (+8 with EMA).  See include/octic_b200.h for the arithmetic and the reference call sites.
#include "octic_capi_internal.h"

namespace octic {

constexpr int OPT_THREADS = 256;

struct OptArgs {
  float max_grad_norm, beta1, beta2, beta3, eps, inv_bc1, inv_bc2, lr, ema_mom;
};

__device__ __forceinline__ float block_sum(float x, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = x;
  __syncthreads();
  float t = (threadIdx.x < OPT_THREADS / 32) ? sh[threadIdx.x] : 0.f;
  if (w == 0) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  __syncthreads();
  return t;   // valid in thread 0
}

__global__ void __launch_bounds__(OPT_THREADS) optim_sqnorm_kernel(const float* __restrict__ x, long n,
                                                                   float* __restrict__ partials) {
  __shared__ float sh[OPT_THREADS / 32];
  float acc = 0.f;
  const long stride = static_cast<long>(gridDim.x) * OPT_THREADS;
  const long tid = static_cast<long>(blockIdx.x) * OPT_THREADS + threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const long n4 = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (long i = tid; i < n4; i += stride) {
      const float4 a = x4[i];
      acc += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    }
    for (long i = (n4 << 2) + tid; i < n; i += stride) acc += x[i] * x[i];
  } else {
    for (long i = tid; i < n; i += stride) acc += x[i] * x[i];
  }
  const float t = block_sum(acc, sh);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

// out[j] = sum of `count` partials in a fixed order (thread-strided, then the shuffle tree); one CTA per output.
// mode 0: one output from partials[0..count);  mode 1: per tensor, two interleaved partials per chunk.
__global__ void __launch_bounds__(OPT_THREADS)
optim_reduce_kernel(const float* __restrict__ partials, int count, const octic_optim_seg* __restrict__ segs,
                    float* __restrict__ out) {
  __shared__ float sh[OPT_THREADS / 32];
  if (segs == nullptr) {
    float acc = 0.f;
    for (int i = threadIdx.x; i < count; i += OPT_THREADS) acc += partials[i];
    const float t = block_sum(acc, sh);
    if (threadIdx.x == 0) out[0] = t;
    return;
  }
  const octic_optim_seg sg = segs[blockIdx.x];
  float a = 0.f, b = 0.f;
  for (int i = threadIdx.x; i < sg.num_chunks; i += OPT_THREADS) {
    a += partials[2 * (sg.first_chunk + i)];
    b += partials[2 * (sg.first_chunk + i) + 1];
  }
  const float ta = block_sum(a, sh);
  const float tb = block_sum(b, sh);
  if (threadIdx.x == 0) {
    out[2 * blockIdx.x] = ta;
    out[2 * blockIdx.x + 1] = tb;
  }
}

// one element of stage 1; returns the update u and advances the moments
__device__ __forceinline__ float adam_update(float p, float g, float& m, float& v, float wd, float inv_clip,
                                             const OptArgs& a) {
  const float sg = g * inv_clip;
  m = a.beta1 * m + a.beta3 * sg;
  v = a.beta2 * v + (1.f - a.beta2) * sg * sg;
  const float denom = sqrtf(v * a.inv_bc2) + a.eps;
  return (m * a.inv_bc1) / denom + wd * p;
}

template <bool APPLY>
__global__ void __launch_bounds__(OPT_THREADS)
optim_stage1_kernel(const octic_optim_chunk* __restrict__ chunks, const octic_optim_seg* __restrict__ segs,
                    float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, float* __restrict__ chunk_norms,
                    const float* __restrict__ gnorm_sq, OptArgs a) {
  __shared__ float sh[OPT_THREADS / 32];
  const octic_optim_chunk c = chunks[blockIdx.x];
  const octic_optim_seg sg = segs[c.seg];
  float inv_clip = 1.f;
  if (gnorm_sq != nullptr && a.max_grad_norm > 0.f) {
    const float gn = sqrtf(*gnorm_sq);
    inv_clip = gn > a.max_grad_norm ? a.max_grad_norm / gn : 1.f;
  }
  float* __restrict__ pp = c.p;
  float* __restrict__ ep = c.ema;
  float* __restrict__ gp = g + c.off;
  float* __restrict__ mp = m + c.off;
  float* __restrict__ vp = v + c.off;
  const float step = a.lr * sg.lr_scale;
  float psq = 0.f, usq = 0.f;
  const bool vec = ((reinterpret_cast<uintptr_t>(pp) | reinterpret_cast<uintptr_t>(gp) | reinterpret_cast<uintptr_t>(mp) |
                     reinterpret_cast<uintptr_t>(vp) | reinterpret_cast<uintptr_t>(ep)) & 15) == 0;
  const int len4 = vec ? (c.len & ~3) : 0;
  for (int i = threadIdx.x * 4; i < len4; i += OPT_THREADS * 4) {
    float4 P = *reinterpret_cast<const float4*>(pp + i);
    float4 G = *reinterpret_cast<const float4*>(gp + i);
    float4 M = *reinterpret_cast<const float4*>(mp + i);
    float4 V = *reinterpret_cast<const float4*>(vp + i);
    float4 U;
    U.x = adam_update(P.x, G.x, M.x, V.x, sg.weight_decay, inv_clip, a);
    U.y = adam_update(P.y, G.y, M.y, V.y, sg.weight_decay, inv_clip, a);
    U.z = adam_update(P.z, G.z, M.z, V.z, sg.weight_decay, inv_clip, a);
    U.w = adam_update(P.w, G.w, M.w, V.w, sg.weight_decay, inv_clip, a);
    *reinterpret_cast<float4*>(mp + i) = M;
    *reinterpret_cast<float4*>(vp + i) = V;
    if (APPLY) {
      P.x -= step * U.x; P.y -= step * U.y; P.z -= step * U.z; P.w -= step * U.w;
      *reinterpret_cast<float4*>(pp + i) = P;
      if (ep != nullptr) {
        float4 E = *reinterpret_cast<const float4*>(ep + i);
        E.x = a.ema_mom * E.x + (1.f - a.ema_mom) * P.x; E.y = a.ema_mom * E.y + (1.f - a.ema_mom) * P.y;
        E.z = a.ema_mom * E.z + (1.f - a.ema_mom) * P.z; E.w = a.ema_mom * E.w + (1.f - a.ema_mom) * P.w;
        *reinterpret_cast<float4*>(ep + i) = E;
      }
    } else {
      psq += P.x * P.x + P.y * P.y + P.z * P.z + P.w * P.w;
      usq += U.x * U.x + U.y * U.y + U.z * U.z + U.w * U.w;
      *reinterpret_cast<float4*>(gp + i) = U;
    }
  }
  for (int i = len4 + threadIdx.x; i < c.len; i += OPT_THREADS) {
    float P = pp[i], M = mp[i], V = vp[i];
    const float U = adam_update(P, gp[i], M, V, sg.weight_decay, inv_clip, a);
    mp[i] = M;
    vp[i] = V;
    if (APPLY) {
      P -= step * U;
      pp[i] = P;
      if (ep != nullptr) ep[i] = a.ema_mom * ep[i] + (1.f - a.ema_mom) * P;
    } else {
      psq += P * P;
      usq += U * U;
      gp[i] = U;
    }
  }
  if (!APPLY) {
    const float ps = block_sum(psq, sh);
    const float us = block_sum(usq, sh);
    if (threadIdx.x == 0) {
      chunk_norms[2 * blockIdx.x] = ps;
      chunk_norms[2 * blockIdx.x + 1] = us;
    }
  }
}

__global__ void __launch_bounds__(OPT_THREADS)
optim_lamb_stage2_kernel(const octic_optim_chunk* __restrict__ chunks, const octic_optim_seg* __restrict__ segs,
                         const float* __restrict__ u, const float* __restrict__ seg_norms, float lr, int use_nvlamb,
                         float ema_mom) {
  const octic_optim_chunk c = chunks[blockIdx.x];
  const octic_optim_seg sg = segs[c.seg];
  float ratio = lr * sg.lr_scale;
  if (use_nvlamb || sg.weight_decay != 0.f) {
    const float pn = sqrtf(seg_norms[2 * c.seg]), un = sqrtf(seg_norms[2 * c.seg + 1]);
    if (pn != 0.f && un != 0.f) ratio *= pn / un;
  }
  float* __restrict__ pp = c.p;
  float* __restrict__ ep = c.ema;
  const float* __restrict__ up = u + c.off;
  const bool vec = ((reinterpret_cast<uintptr_t>(pp) | reinterpret_cast<uintptr_t>(up) | reinterpret_cast<uintptr_t>(ep)) & 15) == 0;
  const int len4 = vec ? (c.len & ~3) : 0;
  for (int i = threadIdx.x * 4; i < len4; i += OPT_THREADS * 4) {
    float4 P = *reinterpret_cast<const float4*>(pp + i);
    const float4 U = *reinterpret_cast<const float4*>(up + i);
    P.x -= ratio * U.x; P.y -= ratio * U.y; P.z -= ratio * U.z; P.w -= ratio * U.w;
    *reinterpret_cast<float4*>(pp + i) = P;
    if (ep != nullptr) {
      float4 E = *reinterpret_cast<const float4*>(ep + i);
      E.x = ema_mom * E.x + (1.f - ema_mom) * P.x; E.y = ema_mom * E.y + (1.f - ema_mom) * P.y;
      E.z = ema_mom * E.z + (1.f - ema_mom) * P.z; E.w = ema_mom * E.w + (1.f - ema_mom) * P.w;
      *reinterpret_cast<float4*>(ep + i) = E;
    }
  }
  for (int i = len4 + threadIdx.x; i < c.len; i += OPT_THREADS) {
    const float P = pp[i] - ratio * up[i];
    pp[i] = P;
    if (ep != nullptr) ep[i] = ema_mom * ep[i] + (1.f - ema_mom) * P;
  }
}

}  // namespace octic

using namespace octic;

extern "C" {

int octic_optim_sqnorm(const float* x, long n, float* partials, float* out, void* stream) {
  if (x == nullptr || out == nullptr || partials == nullptr || n < 0) return OCTIC_ERR_ARG;
  long blocks = (n / 4 + OPT_THREADS - 1) / OPT_THREADS;
  if (blocks > OCTIC_OPTIM_SQNORM_PARTIALS) blocks = OCTIC_OPTIM_SQNORM_PARTIALS;
  if (blocks < 1) blocks = 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  optim_sqnorm_kernel<<<static_cast<unsigned>(blocks), OPT_THREADS, 0, s>>>(x, n, partials);
  optim_reduce_kernel<<<1, OPT_THREADS, 0, s>>>(partials, static_cast<int>(blocks), nullptr, out);
  return cudaGetLastError() == cudaSuccess ? OCTIC_OK : OCTIC_ERR_CUDA;
}

int octic_optim_stage1(const octic_optim_chunk* chunks, int nchunks, const octic_optim_seg* segs, float* g, float* m,
                       float* v, float* chunk_norms, const float* gnorm_sq, float max_grad_norm, float beta1,
                       float beta2, float beta3, float eps, float bc1, float bc2, float lr, int apply,
                       float ema_momentum, void* stream) {
  if (chunks == nullptr || segs == nullptr || g == nullptr || m == nullptr || v == nullptr || nchunks < 0)
    return OCTIC_ERR_ARG;
  if (!apply && chunk_norms == nullptr) return OCTIC_ERR_ARG;
  if (!(bc1 > 0.f) || !(bc2 > 0.f)) return OCTIC_ERR_ARG;
  if (nchunks == 0) return OCTIC_OK;
  OptArgs a{max_grad_norm, beta1, beta2, beta3, eps, 1.f / bc1, 1.f / bc2, lr, ema_momentum};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (apply)
    optim_stage1_kernel<true><<<nchunks, OPT_THREADS, 0, s>>>(chunks, segs, g, m, v, chunk_norms, gnorm_sq, a);
  else
    optim_stage1_kernel<false><<<nchunks, OPT_THREADS, 0, s>>>(chunks, segs, g, m, v, chunk_norms, gnorm_sq, a);
  return cudaGetLastError() == cudaSuccess ? OCTIC_OK : OCTIC_ERR_CUDA;
}

int octic_optim_lamb_stage2(const octic_optim_chunk* chunks, int nchunks, const octic_optim_seg* segs, int nsegs,
                            const float* u, const float* chunk_norms, float* seg_norms, float lr, int use_nvlamb,
                            float ema_momentum, void* stream) {
  if (chunks == nullptr || segs == nullptr || u == nullptr || chunk_norms == nullptr || seg_norms == nullptr ||
      nchunks < 0 || nsegs < 0)
    return OCTIC_ERR_ARG;
  if (nchunks == 0 || nsegs == 0) return OCTIC_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  optim_reduce_kernel<<<nsegs, OPT_THREADS, 0, s>>>(chunk_norms, 0, segs, seg_norms);
  optim_lamb_stage2_kernel<<<nchunks, OPT_THREADS, 0, s>>>(chunks, segs, u, seg_norms, lr, use_nvlamb, ema_momentum);
  return cudaGetLastError() == cudaSuccess ? OCTIC_OK : OCTIC_ERR_CUDA;
}

}  // extern "C"
