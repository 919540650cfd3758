// HBM-bound kernels of the octic block: D8 GELU, LayerNormD8 / LayerNorm, layer-scale backward, column sums,
// power-spectrum invariant, hybrid bridge, im2col and casts.  All of them stream packed [T, D] rows with
// 16-byte accesses; per-token reductions are warp-shuffle, column (parameter-gradient) reductions are
// register accumulators + one red.add per column per CTA.
#include "octic_capi_internal.h"
#include <stdlib.h>
#include "gelu_math.cuh"

namespace octic {

constexpr float kSqrt2Over4 = 0.35355339059327376f;

// erf via Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, i.e. below fp32 GELU's own rounding for |x| < 4):
// one MUFU.RCP + one MUFU.EX2 + 6 FMA instead of erff's ~30 instructions (the D8 GELU needs 8 erf per channel)
__device__ __forceinline__ float erf_as(float x) {
  const float z = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));   // MUFU.RCP
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  return copysignf(1.0f - poly * t * __expf(-z * z), x);
}
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erf_as(x * kInvSqrt2)); }
// d/dx gelu(x) = Phi(x) + x * phi(x)      (reference octic_vits/d8_gelu.py:16-26)
// The erf approximation needs exp(-(x/sqrt2)^2) = exp(-x^2/2), the same exponential as the pdf: one MUFU.EX2 and
// one MUFU.RCP per element in total (the kernels using this are otherwise co-bound by the MUFU pipe).
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float z = fabsf(x) * kInvSqrt2;
  const float e = __expf(-0.5f * x * x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));   // MUFU.RCP
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erfv = copysignf(1.0f - poly * t * e, x);
  return fmaf(x, kInvSqrt2Pi * e, 0.5f * (1.0f + erfv));
}

// Isotypic -> regular (inverse D8 Fourier transform), reference octic_vits/d8_utils.py:276-303.
__device__ __forceinline__ void i2r(const float (&x)[8], float (&y)[8]) {
  const float a = x[0] + x[1], b = x[0] - x[1], c = x[2] + x[3], d = x[2] - x[3];
  const float e = x[4] + x[5], f = x[4] - x[5], g = x[6] + x[7], h = x[6] - x[7];
  const float apc = a + c, amc = a - c, bpd = b + d, bmd = b - d;
  const float eph = e + h, emh = e - h, fpg = f + g, fmg = f - g;
  y[0] = kSqrt2Over4 * (apc + eph);
  y[1] = kSqrt2Over4 * (amc + fmg);
  y[2] = kSqrt2Over4 * (apc - eph);
  y[3] = kSqrt2Over4 * (amc - fmg);
  y[4] = kSqrt2Over4 * (bpd - fpg);
  y[5] = kSqrt2Over4 * (bmd - emh);
  y[6] = kSqrt2Over4 * (bpd + fpg);
  y[7] = kSqrt2Over4 * (bmd + emh);
}
// Regular -> isotypic (forward D8 Fourier transform), reference octic_vits/d8_utils.py:317-344.
__device__ __forceinline__ void r2i(const float (&x)[8], float (&y)[8]) {
  const float a = x[0] + x[1], b = x[0] - x[1], c = x[2] + x[3], d = x[2] - x[3];
  const float e = x[4] + x[5], f = x[4] - x[5], g = x[6] + x[7], h = x[6] - x[7];
  const float apc = a + c, cma = c - a, bpd = b + d, bmd = b - d;
  const float epg = e + g, gme = g - e, fph = f + h, fmh = f - h;
  y[0] = kSqrt2Over4 * (apc + epg);
  y[1] = kSqrt2Over4 * (apc - epg);
  y[2] = kSqrt2Over4 * (bpd + fph);
  y[3] = kSqrt2Over4 * (bpd - fph);
  y[4] = kSqrt2Over4 * (gme - cma);
  y[5] = kSqrt2Over4 * (bmd + fmh);
  y[6] = kSqrt2Over4 * (bmd - fmh);
  y[7] = kSqrt2Over4 * (gme + cma);
}

// same butterflies as i2r / r2i above on channel pairs; `s` is the output scale (sqrt2/4, or sqrt2/8 when the
// 0.5 of the GELU is folded in)
__device__ __forceinline__ void i2r_2(const f2 (&x)[8], f2 (&y)[8], f2 s) {
  const f2 a = add2(x[0], x[1]), b = sub2(x[0], x[1]), c = add2(x[2], x[3]), d = sub2(x[2], x[3]);
  const f2 e = add2(x[4], x[5]), f = sub2(x[4], x[5]), g = add2(x[6], x[7]), h = sub2(x[6], x[7]);
  const f2 apc = add2(a, c), amc = sub2(a, c), bpd = add2(b, d), bmd = sub2(b, d);
  const f2 eph = add2(e, h), emh = sub2(e, h), fpg = add2(f, g), fmg = sub2(f, g);
  y[0] = mul2(s, add2(apc, eph));
  y[1] = mul2(s, add2(amc, fmg));
  y[2] = mul2(s, sub2(apc, eph));
  y[3] = mul2(s, sub2(amc, fmg));
  y[4] = mul2(s, sub2(bpd, fpg));
  y[5] = mul2(s, sub2(bmd, emh));
  y[6] = mul2(s, add2(bpd, fpg));
  y[7] = mul2(s, add2(bmd, emh));
}
__device__ __forceinline__ void r2i_2(const f2 (&x)[8], f2 (&y)[8], f2 s) {
  const f2 a = add2(x[0], x[1]), b = sub2(x[0], x[1]), c = add2(x[2], x[3]), d = sub2(x[2], x[3]);
  const f2 e = add2(x[4], x[5]), f = sub2(x[4], x[5]), g = add2(x[6], x[7]), h = sub2(x[6], x[7]);
  const f2 apc = add2(a, c), cma = sub2(c, a), bpd = add2(b, d), bmd = sub2(b, d);
  const f2 epg = add2(e, g), gme = sub2(g, e), fph = add2(f, h), fmh = sub2(f, h);
  y[0] = mul2(s, add2(apc, epg));
  y[1] = mul2(s, sub2(apc, epg));
  y[2] = mul2(s, add2(bpd, fph));
  y[3] = mul2(s, sub2(bpd, fph));
  y[4] = mul2(s, sub2(gme, cma));
  y[5] = mul2(s, add2(bmd, fmh));
  y[6] = mul2(s, sub2(bmd, fmh));
  y[7] = mul2(s, add2(gme, cma));
}
// ----------------------------------------------- vector I/O -----------------------------------------------
template <typename T, int V> struct Vec;
template <> struct Vec<float, 4> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec<float, 1> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[1]) { v[0] = *p; }
  static __device__ __forceinline__ void store(float* p, const float (&v)[1]) { *p = v[0]; }
};
template <> struct Vec<__nv_bfloat16, 8> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
      const float2 f = __bfloat1622float2(b);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&b);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <> struct Vec<__nv_bfloat16, 4> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x));
    const float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
    v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y;
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[4]) {
    const __nv_bfloat162 b0 = __floats2bfloat162_rn(v[0], v[1]);
    const __nv_bfloat162 b1 = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<const uint32_t*>(&b0);
    t.y = *reinterpret_cast<const uint32_t*>(&b1);
    *reinterpret_cast<uint2*>(p) = t;
  }
};
template <> struct Vec<__nv_bfloat16, 1> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[1]) { v[0] = __bfloat162float(*p); }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[1]) { *p = __float2bfloat16(v[0]); }
};

// column offset (in units of C) of 8-tuple component k inside the packed row
// (x4 = E[0,:C], x5 = E[1,:C], x6 = E[0,C:], x7 = E[1,C:]; reference octic_vits/d8_utils.py:370-385)
__device__ __constant__ int kCompOff[8] = {0, 1, 2, 3, 4, 6, 5, 7};

// ------------------------------------------------- D8 GELU -------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(256) gelu_d8_fwd_kernel(const T* __restrict__ x, long ldx, T* __restrict__ y, long ldy,
                                                          long T_rows, int C) {
  const int nv = C / V;
  const long total = T_rows * nv;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const long t = idx / nv;
    const int c = static_cast<int>(idx - t * nv) * V;
    float in[8][V];
#pragma unroll
    for (int k = 0; k < 8; ++k) Vec<T, V>::load(x + t * ldx + kCompOff[k] * C + c, in[k]);
    float out[8][V];
    if constexpr ((V & 1) == 0) {
#pragma unroll
      for (int v = 0; v < V; v += 2) {
        f2 a[8], r[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = mk2(in[k][v], in[k][v + 1]);
        i2r_2(a, r, sp2(kSqrt2Over4));
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = gelu_x2_2(r[k]);
        r2i_2(r, a, sp2(0.5f * kSqrt2Over4));
#pragma unroll
        for (int k = 0; k < 8; ++k) un2(a[k], out[k][v], out[k][v + 1]);
      }
    } else {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        float a[8], r[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = in[k][v];
        i2r(a, r);
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = gelu_f(r[k]);
        r2i(r, a);
#pragma unroll
        for (int k = 0; k < 8; ++k) out[k][v] = a[k];
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) Vec<T, V>::store(y + t * ldy + kCompOff[k] * C + c, out[k]);
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(256) gelu_d8_bwd_kernel(const T* __restrict__ g, long ldg, const T* __restrict__ x,
                                                          long ldx, T* __restrict__ gin, long ldgin, long T_rows, int C) {
  const int nv = C / V;
  const long total = T_rows * nv;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const long t = idx / nv;
    const int c = static_cast<int>(idx - t * nv) * V;
    float xin[8][V], gg[8][V];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      Vec<T, V>::load(x + t * ldx + kCompOff[k] * C + c, xin[k]);
      Vec<T, V>::load(g + t * ldg + kCompOff[k] * C + c, gg[k]);
    }
    float out[8][V];
    // (the packed-pair form that speeds up the forward kernel by 16 % made this one slower: 232 -> 281 us, measured)
    {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        float a[8], u[8], b[8], gu[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { a[k] = xin[k][v]; b[k] = gg[k][v]; }
        i2r(a, u);
        i2r(b, gu);   // R2I = I2R^T, so the cotangent is pulled back with I2R (octic_vits/d8_gelu.py:283-321)
#pragma unroll
        for (int k = 0; k < 8; ++k) u[k] = gelu_grad_f(u[k]) * gu[k];
        r2i(u, a);
#pragma unroll
        for (int k = 0; k < 8; ++k) out[k][v] = a[k];
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) Vec<T, V>::store(gin + t * ldgin + kCompOff[k] * C + c, out[k]);
  }
}

// plain GELU backward (dense half): gin = g * gelu'(x), bf16
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const __nv_bfloat16* __restrict__ g,
                                                       const __nv_bfloat16* __restrict__ x,
                                                       __nv_bfloat16* __restrict__ gin, long n8) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    float a[8], b[8], o[8];
    Vec<__nv_bfloat16, 8>::load(g + i * 8, a);
    Vec<__nv_bfloat16, 8>::load(x + i * 8, b);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = a[k] * gelu_grad_f(b[k]);
    Vec<__nv_bfloat16, 8>::store(gin + i * 8, o);
  }
}

// ------------------------------------------------ symmetric parameter expansion ------------------------------------------------
// The D8-symmetric lifting filters (d8_layers.py:329-373) and the unfolded positional embedding (d8_utils.py:388-451) are
// LINEAR maps of the stored half-size parameters in which every output element is a signed combination of at most K
// stored elements (K = 2 forward, 8 for the transposed map of backward).  The index / coefficient tables are built once
// on the host from the reference formula itself (pushing a one-hot basis through it); these two kernels apply them:
//   row map:  out[r, o]  (+)= sum_k coef[o, k] * in[r, idx[o, k]]     (filters: r = (c_out, c_in), o = filter tap)
//   pos map:  out[o, c]  (+)= sum_k coef[o, k] * in[idx[o, k], c]     (positional embedding: o = position, c = channel)
// One launch each replaces ~20 eager flip / rot90 / cat / mul / add launches per parameter (and as many in backward).
__global__ void __launch_bounds__(256) sparse_rowmap_kernel(const float* __restrict__ in, long ld_in, float* __restrict__ out,
                                                            long ld_out, long rows, int n_out, int K,
                                                            const int* __restrict__ idx, const float* __restrict__ coef,
                                                            int accumulate) {
  const long total = rows * n_out;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / n_out;
    const int o = static_cast<int>(i - r * n_out);
    const float* src = in + r * ld_in;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc = fmaf(__ldg(coef + o * K + k), src[__ldg(idx + o * K + k)], acc);
    float* dst = out + r * ld_out + o;
    *dst = accumulate ? *dst + acc : acc;
  }
}
__global__ void __launch_bounds__(256) sparse_posmap_kernel(const float* __restrict__ in, long ld_in, float* __restrict__ out,
                                                            long ld_out, int n_out, int cols, int K,
                                                            const int* __restrict__ idx, const float* __restrict__ coef,
                                                            int accumulate) {
  const long total = static_cast<long>(n_out) * cols;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int o = static_cast<int>(i / cols), c = static_cast<int>(i - static_cast<long>(o) * cols);
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc = fmaf(__ldg(coef + o * K + k), in[static_cast<long>(__ldg(idx + o * K + k)) * ld_in + c], acc);
    float* dst = out + static_cast<long>(o) * ld_out + c;
    *dst = accumulate ? *dst + acc : acc;
  }
}

// ------------------------------------------------ column sums ------------------------------------------------
// out[c] += sum_t x[t, c];  block = 32 x 8 threads, each thread owns 2 columns and strides rows.
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, long ldx, long T_rows,
                                                          int n_cols, float* __restrict__ out, int rows_per_block) {
  __shared__ float red[8][64];
  const int c = (blockIdx.x * 32 + threadIdx.x) * 2;
  const long r0 = static_cast<long>(blockIdx.y) * rows_per_block;
  const long r1 = min(T_rows, r0 + rows_per_block);
  float s0 = 0.f, s1 = 0.f;
  if (c < n_cols) {
    for (long r = r0 + threadIdx.y; r < r1; r += 8) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(x + r * ldx + c));
      s0 += f.x; s1 += f.y;
    }
  }
  red[threadIdx.y][threadIdx.x * 2] = s0;
  red[threadIdx.y][threadIdx.x * 2 + 1] = s1;
  __syncthreads();
  if (threadIdx.y == 0 && c < n_cols) {
#pragma unroll
    for (int i = 1; i < 8; ++i) { s0 += red[i][threadIdx.x * 2]; s1 += red[i][threadIdx.x * 2 + 1]; }
    atomicAdd(out + c, s0);
    if (c + 1 < n_cols) atomicAdd(out + c + 1, s1);
  }
}

// ------------------------------------------------ LayerNorm(D8) ------------------------------------------------
// One warp per token; the whole row lives in registers (NCH float4 chunks per lane, D <= 128 * NCH).
// D8 = true : six mean/variance groups {A1,A2,B1,B2,E0,E1}, one shared std (d8_layers.py:166-186)
// D8 = false: ordinary LayerNorm.
template <bool D8>
__device__ __forceinline__ int seg_of(int col, int C) {
  if (!D8) return 0;
  return col < 4 * C ? col / C : 4 + (col - 4 * C) / (2 * C);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename TY, bool D8, int NCH>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, long ldx,
                                                            const float* __restrict__ alpha,
                                                            const float* __restrict__ beta, float eps,
                                                            TY* __restrict__ y, long ldy, float* __restrict__ stats,
                                                            long T_rows, int D) {
  constexpr int NSEG = D8 ? 6 : 1;
  const int lane = threadIdx.x & 31;
  const int C = D / 8;
  const int nchunks = D / 4;
  const long warp_global = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const long nwarps = (static_cast<long>(gridDim.x) * blockDim.x) >> 5;
  // per-lane chunk geometry is token independent: the segment id of each of the lane's chunks is packed 3 bits apiece
  // into one register (7 = chunk past the row), so the row loads below issue back to back with no index math between
  unsigned long long segpack = 0;
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    const int ch = lane + 32 * j;
    segpack |= static_cast<unsigned long long>(ch < nchunks ? seg_of<D8>(ch * 4, C) : 7) << (3 * j);
  }
  static_assert(NCH <= 21, "segpack holds 21 chunks");
  for (long t = warp_global; t < T_rows; t += nwarps) {
    float v[NCH][4];
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int ch = min(lane + 32 * j, nchunks - 1);      // clamped: the load is unconditional, the value is masked
      Vec<float, 4>::load(x + t * ldx + ch * 4, v[j]);
    }
    int seg[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) seg[j] = static_cast<int>((segpack >> (3 * j)) & 7ull);
    float sum[NSEG];
#pragma unroll
    for (int s = 0; s < NSEG; ++s) sum[s] = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const float s4 = (v[j][0] + v[j][1]) + (v[j][2] + v[j][3]);
#pragma unroll
      for (int s = 0; s < NSEG; ++s) sum[s] += (seg[j] == s) ? s4 : 0.f;
    }
    float mean[NSEG], var[NSEG];
#pragma unroll
    for (int s = 0; s < NSEG; ++s) {
      const float n = D8 ? (s < 4 ? C : 2 * C) : D;
      mean[s] = warp_sum(sum[s]) / n;
      var[s] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      float mu = 0.f;
#pragma unroll
      for (int s = 0; s < NSEG; ++s) mu = (seg[j] == s) ? mean[s] : mu;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) { v[j][i] -= mu; q += v[j][i] * v[j][i]; }
#pragma unroll
      for (int s = 0; s < NSEG; ++s) var[s] += (seg[j] == s) ? q : 0.f;
    }
    float S = 0.f;
#pragma unroll
    for (int s = 0; s < NSEG; ++s) {
      const float n = D8 ? (s < 4 ? C : 2 * C) : D;
      const float w = D8 ? (s < 4 ? 1.0f : 0.5f) : 1.0f;
      S += w * (warp_sum(var[s]) / n);
    }
    // D8: std = (sqrt2/4) * sqrt(S + eps)  ->  rstd = sqrt(8 / (S + eps));   plain: rstd = 1/sqrt(var + eps)
    const float rstd = D8 ? sqrtf(8.0f / (S + eps)) : rsqrtf(S + eps);
    if (stats != nullptr && lane == 0) {
      if (D8) {
#pragma unroll
        for (int s = 0; s < NSEG; ++s) stats[t * 8 + s] = mean[s];
        stats[t * 8 + 6] = rstd;
      } else {
        stats[t * 2] = mean[0];
        stats[t * 2 + 1] = rstd;
      }
    }
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      if (seg[j] != 7) {
        const int col = (lane + 32 * j) * 4;
        float a[4], o[4];
        Vec<float, 4>::load(alpha + col, a);
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = v[j][i] * rstd * a[i];
        if (beta != nullptr && (!D8 || col < C)) {
          float b[4];
          Vec<float, 4>::load(beta + col, b);
#pragma unroll
          for (int i = 0; i < 4; ++i) o[i] += b[i];
        }
        Vec<TY, 4>::store(y + t * ldy + col, o);
      }
    }
  }
}

// LayerNormD8 forward, lane-contiguous fast path (C % 16 == 0: ViT-S/L/H widths 48, 128, 160).  One warp per token;
// lane l owns the CPL = C/16 consecutive float4 chunks [l*CPL, (l+1)*CPL), which all lie inside ONE irrep segment
// (lanes 0-3 A1, 4-7 A2, 8-11 B1, 12-15 B2, 16-23 E row 0, 24-31 E row 1).  The six means / variances of
// d8_layers.py:166-186 are therefore plain group reductions (two or three xor-shuffles) with no per-chunk segment
// selection: ~5x fewer instructions and half the registers of the generic kernel.
template <typename TY, int CPL>
__global__ void __launch_bounds__(256) layernorm_d8_fwd_lc_kernel(const float* __restrict__ x, long ldx,
                                                                  const float* __restrict__ alpha,
                                                                  const float* __restrict__ beta, float eps,
                                                                  TY* __restrict__ y, long ldy, float* __restrict__ stats,
                                                                  long T_rows, int C) {
  const int lane = threadIdx.x & 31;
  const long warp_global = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const long nwarps = (static_cast<long>(gridDim.x) * blockDim.x) >> 5;
  const int col0 = lane * CPL * 4;
  const bool is_e = lane >= 16;
  const float inv_n = 1.0f / static_cast<float>(is_e ? 2 * C : C);
  for (long t = warp_global; t < T_rows; t += nwarps) {
    float v[CPL][4];
#pragma unroll
    for (int j = 0; j < CPL; ++j) Vec<float, 4>::load(x + t * ldx + col0 + 4 * j, v[j]);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < CPL; ++j) { s0 += v[j][0] + v[j][1]; s1 += v[j][2] + v[j][3]; }
    float sm = s0 + s1;
    sm += __shfl_xor_sync(0xffffffffu, sm, 1);
    sm += __shfl_xor_sync(0xffffffffu, sm, 2);
    {
      const float o4 = __shfl_xor_sync(0xffffffffu, sm, 4);
      if (is_e) sm += o4;
    }
    const float mean = sm * inv_n;
    float q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[j][i] -= mean;
      q0 += v[j][0] * v[j][0] + v[j][1] * v[j][1];
      q1 += v[j][2] * v[j][2] + v[j][3] * v[j][3];
    }
    float q = q0 + q1;
    q += __shfl_xor_sync(0xffffffffu, q, 1);
    q += __shfl_xor_sync(0xffffffffu, q, 2);
    {
      const float o4 = __shfl_xor_sync(0xffffffffu, q, 4);
      if (is_e) q += o4;
    }
    const float var = q * inv_n;
    // S = var_A1 + var_A2 + var_B1 + var_B2 + (var_E0 + var_E1) / 2
    const float S = (__shfl_sync(0xffffffffu, var, 0) + __shfl_sync(0xffffffffu, var, 4)) +
                    (__shfl_sync(0xffffffffu, var, 8) + __shfl_sync(0xffffffffu, var, 12)) +
                    0.5f * (__shfl_sync(0xffffffffu, var, 16) + __shfl_sync(0xffffffffu, var, 24));
    const float rstd = sqrtf(8.0f / (S + eps));       // std = (sqrt2/4) * sqrt(S + eps)
    if (stats != nullptr) {
      if (lane < 16 && (lane & 3) == 0) stats[t * 8 + (lane >> 2)] = mean;
      if (lane == 16 || lane == 24) stats[t * 8 + 4 + ((lane - 16) >> 3)] = mean;
      if (lane == 0) stats[t * 8 + 6] = rstd;
    }
    float o[CPL][4];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      float a[4];
      Vec<float, 4>::load(alpha + col0 + 4 * j, a);
#pragma unroll
      for (int i = 0; i < 4; ++i) o[j][i] = v[j][i] * rstd * a[i];
      if (beta != nullptr && lane < 4) {              // lanes 0-3 hold exactly the A1 columns [0, C)
        float bb[4];
        Vec<float, 4>::load(beta + col0 + 4 * j, bb);
#pragma unroll
        for (int i = 0; i < 4; ++i) o[j][i] += bb[i];
      }
    }
    TY* yr = y + t * ldy + col0;
    if constexpr (sizeof(TY) == 2 && (CPL % 2) == 0) {
#pragma unroll
      for (int j = 0; j < CPL; j += 2) {               // 16-byte stores: two chunks = 8 bf16
        float w8[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) { w8[i] = o[j][i]; w8[4 + i] = o[j + 1 < CPL ? j + 1 : j][i]; }
        Vec<TY, 8>::store(yr + 4 * j, w8);
      }
    } else {
#pragma unroll
      for (int j = 0; j < CPL; ++j) Vec<TY, 4>::store(yr + 4 * j, o[j]);
    }
  }
}

// LayerNormD8 forward, lane-INTERLEAVED variant for C % 128 == 0 ... i.e. CPL = D / 128 in {8, 10} (ViT-L / ViT-H): lane l
// owns the float4 chunks l, l + 32, ..., so every load / store instruction of the warp covers 512 (256) contiguous bytes.
// The lane-contiguous kernel above needs few instructions but each of its 16-byte accesses touches 32 different 128-byte
// lines: ncu shows it bound by L1 wavefronts (480 per token), 3.4 TB/s at 22 % issue utilisation
// (profiles/r02_ncu_summary.md).  Here a 1-D irrep segment spans 4 * CPL >= 32 chunks, so the 32 chunks of one j
// (chunk = lane + 32 j) lie in at most TWO segments, split at a compile-time lane threshold: the six sums are six
// registers with static indices, followed by six full-warp reductions.
template <int CPL>
struct LnIl {
  static constexpr int S1 = 4 * CPL;                                   // chunks per 1-D irrep segment (E rows: 2 * S1)
  static __host__ __device__ constexpr int seg(int c) { return c < 4 * S1 ? c / S1 : 4 + (c - 4 * S1) / (2 * S1); }
  static __host__ __device__ constexpr int seg_begin(int s) { return s < 4 ? s * S1 : 4 * S1 + (s - 4) * 2 * S1; }
};
template <typename TY, int CPL>
__global__ void __launch_bounds__(256, 3) layernorm_d8_fwd_il_kernel(const float* __restrict__ x, long ldx,
                                                                  const float* __restrict__ alpha,
                                                                  const float* __restrict__ beta, float eps,
                                                                  TY* __restrict__ y, long ldy, float* __restrict__ stats,
                                                                  long T_rows, int C) {
  using G = LnIl<CPL>;
  static_assert(G::S1 >= 32, "at most two segments per 32 consecutive chunks");
  const int lane = threadIdx.x & 31;
  const long warp_global = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const long nwarps = (static_cast<long>(gridDim.x) * blockDim.x) >> 5;
  const float inv_n1 = 1.0f / static_cast<float>(C), inv_n2 = 1.0f / static_cast<float>(2 * C);
  for (long t = warp_global; t < T_rows; t += nwarps) {
    float v[CPL][4];
    const float* xr = x + t * ldx + lane * 4;
#pragma unroll
    for (int j = 0; j < CPL; ++j) Vec<float, 4>::load(xr + 128 * j, v[j]);
    float sum[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      constexpr int dummy = 0; (void)dummy;
      const int slo = G::seg(32 * j), shi = G::seg(32 * j + 31);
      const float s4 = (v[j][0] + v[j][1]) + (v[j][2] + v[j][3]);
      if (slo == shi) {
        sum[slo] += s4;
      } else {
        const bool lo = lane < G::seg_begin(shi) - 32 * j;
        sum[slo] += lo ? s4 : 0.f;
        sum[shi] += lo ? 0.f : s4;
      }
    }
    float mean[6];
#pragma unroll
    for (int s_ = 0; s_ < 6; ++s_) mean[s_] = warp_sum(sum[s_]) * (s_ < 4 ? inv_n1 : inv_n2);
    float var[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int slo = G::seg(32 * j), shi = G::seg(32 * j + 31);
      const bool lo = slo == shi || lane < G::seg_begin(shi) - 32 * j;
      const float mu = lo ? mean[slo] : mean[shi];
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) { v[j][i] -= mu; q = fmaf(v[j][i], v[j][i], q); }
      if (slo == shi) {
        var[slo] += q;
      } else {
        var[slo] += lo ? q : 0.f;
        var[shi] += lo ? 0.f : q;
      }
    }
    float S = 0.f;
#pragma unroll
    for (int s_ = 0; s_ < 6; ++s_) S += (s_ < 4 ? inv_n1 : 0.5f * inv_n2) * warp_sum(var[s_]);
    const float rstd = sqrtf(8.0f / (S + eps));       // std = (sqrt2/4) * sqrt(S + eps)
    if (stats != nullptr && lane == 0) {
#pragma unroll
      for (int s_ = 0; s_ < 6; ++s_) stats[t * 8 + s_] = mean[s_];
      stats[t * 8 + 6] = rstd;
    }
    TY* yr = y + t * ldy + lane * 4;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      float a[4], o[4];
      Vec<float, 4>::load(alpha + lane * 4 + 128 * j, a);
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = v[j][i] * rstd * a[i];
      // beta lives on the A1 columns [0, C) = chunks [0, S1)
      if (beta != nullptr && 32 * j < G::S1 && lane + 32 * j < G::S1) {
        float bb[4];
        Vec<float, 4>::load(beta + lane * 4 + 128 * j, bb);
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] += bb[i];
      }
      Vec<TY, 4>::store(yr + 128 * j, o);
    }
  }
}

// Generic one-warp-per-token kernel (any segment layout; used for nn.LayerNorm, where it needs only 80 registers).
template <typename TY, bool D8, int NCH>
__global__ void __launch_bounds__(256) layernorm_fwd_plain_kernel(const float* __restrict__ x, long ldx,
                                                            const float* __restrict__ alpha,
                                                            const float* __restrict__ beta, float eps,
                                                            TY* __restrict__ y, long ldy, float* __restrict__ stats,
                                                            long T_rows, int D) {
  constexpr int NSEG = D8 ? 6 : 1;
  const int lane = threadIdx.x & 31;
  const int C = D / 8;
  const int nchunks = D / 4;
  const long warp_global = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const long nwarps = (static_cast<long>(gridDim.x) * blockDim.x) >> 5;
  for (long t = warp_global; t < T_rows; t += nwarps) {
    float v[NCH][4];
    int seg[NCH];
    float sum[NSEG];
#pragma unroll
    for (int s = 0; s < NSEG; ++s) sum[s] = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int ch = lane + 32 * j;
      seg[j] = -1;
      if (ch < nchunks) {
        Vec<float, 4>::load(x + t * ldx + ch * 4, v[j]);
        seg[j] = seg_of<D8>(ch * 4, C);
        const float s4 = (v[j][0] + v[j][1]) + (v[j][2] + v[j][3]);
#pragma unroll
        for (int s = 0; s < NSEG; ++s) sum[s] += (seg[j] == s) ? s4 : 0.f;
      }
    }
    float mean[NSEG], var[NSEG];
#pragma unroll
    for (int s = 0; s < NSEG; ++s) {
      const float n = D8 ? (s < 4 ? C : 2 * C) : D;
      mean[s] = warp_sum(sum[s]) / n;
      var[s] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      if (seg[j] >= 0) {
        float mu = 0.f;
#pragma unroll
        for (int s = 0; s < NSEG; ++s) mu = (seg[j] == s) ? mean[s] : mu;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[j][i] -= mu; q += v[j][i] * v[j][i]; }
#pragma unroll
        for (int s = 0; s < NSEG; ++s) var[s] += (seg[j] == s) ? q : 0.f;
      }
    }
    float S = 0.f;
#pragma unroll
    for (int s = 0; s < NSEG; ++s) {
      const float n = D8 ? (s < 4 ? C : 2 * C) : D;
      const float w = D8 ? (s < 4 ? 1.0f : 0.5f) : 1.0f;
      S += w * (warp_sum(var[s]) / n);
    }
    // D8: std = (sqrt2/4) * sqrt(S + eps)  ->  rstd = sqrt(8 / (S + eps));   plain: rstd = 1/sqrt(var + eps)
    const float rstd = D8 ? sqrtf(8.0f / (S + eps)) : rsqrtf(S + eps);
    if (stats != nullptr && lane == 0) {
      if (D8) {
#pragma unroll
        for (int s = 0; s < NSEG; ++s) stats[t * 8 + s] = mean[s];
        stats[t * 8 + 6] = rstd;
      } else {
        stats[t * 2] = mean[0];
        stats[t * 2 + 1] = rstd;
      }
    }
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      if (seg[j] >= 0) {
        const int col = (lane + 32 * j) * 4;
        float a[4], o[4];
        Vec<float, 4>::load(alpha + col, a);
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = v[j][i] * rstd * a[i];
        if (beta != nullptr && (!D8 || col < C)) {
          float b[4];
          Vec<float, 4>::load(beta + col, b);
#pragma unroll
          for (int i = 0; i < 4; ++i) o[i] += b[i];
        }
        Vec<TY, 4>::store(y + t * ldy + col, o);
      }
    }
  }
}

// Backward.  With yhat = (x - mu_g) * r, dyh = alpha * dout, Q = sum(dyh * yhat), m_g = mean_g(dyh):
//   dx = r * (dyh - m_g - coef_g * yhat * Q),  coef_g = w_g / (8 n_g)  (D8)  or 1/D (plain)   [SURVEY A.7]
// One CTA owns a token range; thread i owns float4 column chunk i for every token (so dalpha / dbeta accumulate in
// 8 registers and flush with one red.add per column per CTA).  Tokens are processed G at a time: each thread forms
// its partial Q and partial segment sum, a warp-shuffle + smem reduction makes the 7 per-token scalars available to
// all threads, then dx is written.  ~60 registers, so many CTAs are resident and the loads of G tokens are in
// flight together.
constexpr int kLnG = 4;   // tokens per reduction round
template <typename TDY, bool D8>
__global__ void __maxnreg__(96) layernorm_bwd_kernel(const TDY* __restrict__ dy, long lddy,
                                                             const float* __restrict__ x, long ldx,
                                                             const float* __restrict__ stats,
                                                             const float* __restrict__ alpha,
                                                             const float* __restrict__ dx_in, float* __restrict__ dx_out,
                                                             long lddx, float* __restrict__ dalpha,
                                                             float* __restrict__ dbeta, long T_rows, int D,
                                                             int tokens_per_block, __nv_bfloat16* __restrict__ dx_bf16,
                                                             float* __restrict__ dx_colsum) {
  constexpr int NSEG = D8 ? 6 : 1;
  constexpr int NRED = NSEG + 1;                     // Q + per-segment sums
  __shared__ float part[16][kLnG * NRED];            // per-warp partials
  __shared__ float tot[kLnG * NRED];
  const int C = D / 8;
  const int nchunks = D / 4;
  const int ch = threadIdx.x;
  const bool active = ch < nchunks;
  const int col = ch * 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int seg = active ? seg_of<D8>(col, C) : -1;
  const float nseg = D8 ? (seg < 4 ? C : 2 * C) : D;
  const float coef = D8 ? (seg < 4 ? 1.0f / (8.0f * C) : 0.5f / (16.0f * C)) : 1.0f / D;
  float a4[4] = {0.f, 0.f, 0.f, 0.f};
  if (active) Vec<float, 4>::load(alpha + col, a4);
  float acc_a[4] = {0.f, 0.f, 0.f, 0.f}, acc_b[4] = {0.f, 0.f, 0.f, 0.f};
  float acc_c[4] = {0.f, 0.f, 0.f, 0.f};             // column sums of dx (bias gradient of the layer below)

  (void)tokens_per_block;
  const long t1 = T_rows;
  for (long tb = static_cast<long>(blockIdx.x) * kLnG; tb < t1; tb += static_cast<long>(gridDim.x) * kLnG) {
    float yh[kLnG][4], dyh[kLnG][4], rstd[kLnG], q[kLnG], s4[kLnG], din[kLnG][4];
    // the skip-gradient rows are fetched together with x / dy (one DRAM round trip per round, not two)
#pragma unroll
    for (int u = 0; u < kLnG; ++u) {
      if (dx_in != nullptr) Vec<float, 4>::load(dx_in + min(tb + u, t1 - 1) * lddx + (active ? col : 0), din[u]);
      else { din[u][0] = 0.f; din[u][1] = 0.f; din[u][2] = 0.f; din[u][3] = 0.f; }
    }
#pragma unroll
    for (int u = 0; u < kLnG; ++u) {
      const long t = tb + u;
      rstd[u] = 0.f; q[u] = 0.f; s4[u] = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) { yh[u][i] = 0.f; dyh[u][i] = 0.f; }
      {
        // unconditional loads (indices clamped) so that the loads of all kLnG tokens are issued back to back
        const long tc = min(t, t1 - 1);
        const int colc = active ? col : 0;
        const int segc = active ? seg : 0;
        const float valid = (active && t < t1) ? 1.0f : 0.0f;
        float xv[4], d[4];
        Vec<float, 4>::load(x + tc * ldx + colc, xv);
        Vec<TDY, 4>::load(dy + tc * lddy + colc, d);
        const float mu = D8 ? __ldg(stats + tc * 8 + segc) : __ldg(stats + tc * 2);
        rstd[u] = D8 ? __ldg(stats + tc * 8 + 6) : __ldg(stats + tc * 2 + 1);
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] *= valid;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          yh[u][i] = (xv[i] - mu) * rstd[u];
          dyh[u][i] = a4[i] * d[i];
          acc_a[i] += d[i] * yh[u][i];
          acc_b[i] += d[i];
          q[u] += dyh[u][i] * yh[u][i];
          s4[u] += dyh[u][i];
        }
      }
    }
    // block reduction of kLnG * NRED scalars; a warp only shuffles for the segments it actually touches
#pragma unroll
    for (int sg = 0; sg < NSEG; ++sg) {
      const bool mine = (seg == sg);
      if (__any_sync(0xffffffffu, mine)) {
#pragma unroll
        for (int u = 0; u < kLnG; ++u) {
          const float v = warp_sum(mine ? s4[u] : 0.f);
          if (lane == 0) part[warp][u * NRED + 1 + sg] = v;
        }
      } else if (lane == 0) {
#pragma unroll
        for (int u = 0; u < kLnG; ++u) part[warp][u * NRED + 1 + sg] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < kLnG; ++u) {
      const float v = warp_sum(q[u]);
      if (lane == 0) part[warp][u * NRED] = v;
    }
    __syncthreads();
    if (threadIdx.x < kLnG * NRED) {
      float v = 0.f;
      for (int w = 0; w < nwarps; ++w) v += part[w][threadIdx.x];
      tot[threadIdx.x] = v;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kLnG; ++u) {
      const long t = tb + u;
      if (active && t < t1) {
        const float Q = tot[u * NRED];
        const float m = tot[u * NRED + 1 + (D8 ? seg : 0)] / nseg;
        float o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = rstd[u] * (dyh[u][i] - m - coef * yh[u][i] * Q);
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] += din[u][i];
        Vec<float, 4>::store(dx_out + t * lddx + col, o);
        if (dx_bf16 != nullptr) {
          // bf16 copy of the residual-stream gradient: the A operand of the dgrad / wgrad GEMMs of the branch below
          Vec<__nv_bfloat16, 4>::store(dx_bf16 + t * lddx + col, o);
#pragma unroll
          for (int i = 0; i < 4; ++i) acc_c[i] += o[i];
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (dalpha != nullptr) atomicAdd(dalpha + col + i, acc_a[i]);
      if (dbeta != nullptr && (!D8 || col < C)) atomicAdd(dbeta + col + i, acc_b[i]);
      if (dx_colsum != nullptr) atomicAdd(dx_colsum + col + i, acc_c[i]);
    }
  }
}

// Gamma-folded layer-scale backward, last step (see functional.py "gamma-folded backward"): the wgrad GEMM produced
// dW_raw = bf16(dres)^T x; here, one warp per weight row n:
//   dgamma[n] += <W[n, :], dW_raw[n, :]> + bias[n] * cs[n]      (= sum_t dres[t, n] * branch[t, n])
//   dW[n, :]   = gamma[n] * dW_raw[n, :]   (in place),   dbias[n] = gamma[n] * cs[n]
struct LsFinSeg {
  float* dw; const float* w; int N, K; const float* gamma; const float* bias; const float* cs; float* dgamma; float* dbias;
  float* dw_acc;   // optional: dw_acc += gamma * dW_raw (gradient accumulation fused), dW_raw left untouched
};
struct LsFinParams { int nseg; int row_begin[9]; LsFinSeg seg[8]; };
__global__ void __launch_bounds__(128) layerscale_finalize_kernel(const __grid_constant__ LsFinParams p) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int row = warp; row < p.row_begin[p.nseg]; row += nwarps) {
    int sg = 0;
    for (int i = 1; i < p.nseg; ++i) sg = row >= p.row_begin[i] ? i : sg;
    const LsFinSeg& S = p.seg[sg];
    const int n = row - p.row_begin[sg];
    float* dwr = S.dw + static_cast<long>(n) * S.K;
    float* accr = S.dw_acc != nullptr ? S.dw_acc + static_cast<long>(n) * S.K : nullptr;
    const float* wr = S.w + static_cast<long>(n) * S.K;
    const float g = S.gamma[n];
    float dot = 0.f;
    if ((S.K & 3) == 0 && ((reinterpret_cast<uintptr_t>(dwr) | reinterpret_cast<uintptr_t>(wr) | reinterpret_cast<uintptr_t>(accr)) & 15) == 0) {
      // four float4 column groups in flight per lane
      const int k4 = S.K >> 2;
      float4* d4 = reinterpret_cast<float4*>(dwr);
      const float4* w4 = reinterpret_cast<const float4*>(wr);
      for (int k0 = lane; k0 < k4; k0 += 128) {
        float4 dv[4], wv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int k = k0 + 32 * u;
          if (k < k4) { dv[u] = d4[k]; wv[u] = w4[k]; }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int k = k0 + 32 * u;
          if (k < k4) {
            dot = fmaf(wv[u].x, dv[u].x, fmaf(wv[u].y, dv[u].y, fmaf(wv[u].z, dv[u].z, fmaf(wv[u].w, dv[u].w, dot))));
            if (accr != nullptr) {
              float4* a4 = reinterpret_cast<float4*>(accr) + k;
              const float4 a = *a4;
              *a4 = make_float4(fmaf(g, dv[u].x, a.x), fmaf(g, dv[u].y, a.y), fmaf(g, dv[u].z, a.z), fmaf(g, dv[u].w, a.w));
            } else {
              d4[k] = make_float4(g * dv[u].x, g * dv[u].y, g * dv[u].z, g * dv[u].w);
            }
          }
        }
      }
    } else {
      for (int k = lane; k < S.K; k += 32) {
        const float d = dwr[k];
        dot = fmaf(wr[k], d, dot);
        if (accr != nullptr) accr[k] = fmaf(g, d, accr[k]);
        else dwr[k] = g * d;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (lane == 0) {
      const float c = S.cs != nullptr ? S.cs[n] : 0.f;
      if (S.dgamma != nullptr) S.dgamma[n] += dot + (S.bias != nullptr ? S.bias[n] * c : 0.f);
      if (S.dbias != nullptr) S.dbias[n] = g * c;
    }
  }
}

// --------------------------------------------- layer-scale backward ---------------------------------------------
// Threads own float4 column groups, CTAs own row ranges: dy = gamma * s * dres (bf16), dgamma += dres * s * branch,
// colsum += dy.
template <int NV, bool HAS_BRANCH>   // NV float4 groups per thread (D/4 <= blockDim.x * NV; the launcher sizes the block
                                     // so that every thread owns a column group: no idle second pass)
__global__ void __launch_bounds__(512) layerscale_bwd_kernel(const float* __restrict__ dres, long lddres,
                                                             const __nv_bfloat16* __restrict__ branch, long ldbr,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ row_scale, int rows_per_sample,
                                                             __nv_bfloat16* __restrict__ dy, long lddy,
                                                             float* __restrict__ dgamma, float* __restrict__ colsum,
                                                             long T_rows, int D, int rows_per_block) {
  const int nchunks = D / 4;
  constexpr int RU = 4;   // rows in flight per thread
  // persistent CTAs: row groups are dealt round-robin, so the grid can be sized to exactly fill the machine and
  // each CTA flushes its column sums once
  const long r1 = T_rows;
  const long rstride = static_cast<long>(gridDim.x) * RU;
  (void)rows_per_block;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int ch = threadIdx.x + blockDim.x * j;
    if (ch >= nchunks) continue;
    const int col = ch * 4;
    float g[4] = {1.f, 1.f, 1.f, 1.f}, ag[4] = {0.f, 0.f, 0.f, 0.f}, as[4] = {0.f, 0.f, 0.f, 0.f};
    if (gamma != nullptr) Vec<float, 4>::load(gamma + col, g);
    for (long rb = static_cast<long>(blockIdx.x) * RU; rb < r1; rb += rstride) {
      float d[RU][4], b[RU][4], sc[RU];
      // loads are unconditional (row index clamped) so the compiler can issue all of them before the first use
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        const long r = min(rb + u, r1 - 1);
        Vec<float, 4>::load(dres + r * lddres + col, d[u]);
        if (HAS_BRANCH) Vec<__nv_bfloat16, 4>::load(branch + r * ldbr + col, b[u]);
        sc[u] = row_scale != nullptr ? __ldg(row_scale + r / rows_per_sample) : 1.0f;
      }
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        if (rb + u < r1) {
          float o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (HAS_BRANCH) ag[i] += d[u][i] * sc[u] * b[u][i];
            o[i] = g[i] * sc[u] * d[u][i];
            // what the bias gradient sees is the bf16-rounded dy that also feeds dgrad/wgrad
            as[i] += __bfloat162float(__float2bfloat16(o[i]));
          }
          Vec<__nv_bfloat16, 4>::store(dy + (rb + u) * lddy + col, o);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (HAS_BRANCH && dgamma != nullptr) atomicAdd(dgamma + col + i, ag[i]);
      if (colsum != nullptr) atomicAdd(colsum + col + i, as[i]);
    }
  }
}

// --------------------------------------------- invariant / bridge ---------------------------------------------
__global__ void __launch_bounds__(256) power_spectrum_fwd_kernel(const float* __restrict__ x, long ldx,
                                                                 __nv_bfloat16* __restrict__ y, long ldy, long T_rows,
                                                                 int C) {
  const int nv = 6 * C / 4;
  const long total = T_rows * nv;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const long t = idx / nv;
    const int col = static_cast<int>(idx - t * nv) * 4;   // output column
    float a[4], o[4];
    Vec<float, 4>::load(x + t * ldx + col, a);
    if (col < C) {
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = a[i];
    } else if (col < 4 * C) {
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = fabsf(a[i]);
    } else {
      float b[4];
      Vec<float, 4>::load(x + t * ldx + col + 2 * C, b);
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = sqrtf(a[i] * a[i] + b[i] * b[i]);
    }
    Vec<__nv_bfloat16, 4>::store(y + t * ldy + col, o);
  }
}

template <typename TDY>
__global__ void __launch_bounds__(256) power_spectrum_bwd_kernel(const TDY* __restrict__ dy, long lddy,
                                                                 const float* __restrict__ x, long ldx,
                                                                 float* __restrict__ dx, long lddx, long T_rows, int C) {
  const int nv = 6 * C / 4;
  const long total = T_rows * nv;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const long t = idx / nv;
    const int col = static_cast<int>(idx - t * nv) * 4;
    float a[4], g[4], o[4];
    Vec<float, 4>::load(x + t * ldx + col, a);
    Vec<TDY, 4>::load(dy + t * lddy + col, g);
    if (col < C) {
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = g[i];
      Vec<float, 4>::store(dx + t * lddx + col, o);
    } else if (col < 4 * C) {
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = a[i] > 0.f ? g[i] : (a[i] < 0.f ? -g[i] : 0.f);
      Vec<float, 4>::store(dx + t * lddx + col, o);
    } else {
      float b[4], o2[4];
      Vec<float, 4>::load(x + t * ldx + col + 2 * C, b);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float n = sqrtf(a[i] * a[i] + b[i] * b[i]);
        const float inv = n > 0.f ? g[i] / n : 0.f;   // torch.norm backward: 0 at the origin
        o[i] = a[i] * inv;
        o2[i] = b[i] * inv;
      }
      Vec<float, 4>::store(dx + t * lddx + col, o);
      Vec<float, 4>::store(dx + t * lddx + col + 2 * C, o2);
    }
  }
}

// packed row -> cat(convert_5tuple_to_8tuple(xs)) order: swap column blocks [5C,6C) <-> [6C,7C)
__global__ void __launch_bounds__(256) bridge_permute_kernel(const float* __restrict__ x, long ldx, float* __restrict__ y,
                                                             long ldy, long T_rows, int C) {
  const int nv = 8 * C / 4;
  const long total = T_rows * nv;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const long t = idx / nv;
    const int col = static_cast<int>(idx - t * nv) * 4;
    int src = col;
    if (col >= 5 * C && col < 6 * C) src = col + C;
    else if (col >= 6 * C && col < 7 * C) src = col - C;
    float a[4];
    Vec<float, 4>::load(x + t * ldx + src, a);
    Vec<float, 4>::store(y + t * ldy + col, a);
  }
}

// --------------------------------------------------- im2col ---------------------------------------------------
__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ img, int B, int Cin, int Himg, int Wimg,
                                                     int p, __nv_bfloat16* __restrict__ out, long ldo, int Kp) {
  const int gh = Himg / p, gw = Wimg / p;
  const long rows = static_cast<long>(B) * gh * gw;
  const int K = Cin * p * p;
  const long total = rows * Kp;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const long row = idx / Kp;
    const int k = static_cast<int>(idx - row * Kp);
    float v = 0.f;
    if (k < K) {
      const int c = k / (p * p), ij = k - c * p * p, i = ij / p, j = ij - i * p;
      const int b = static_cast<int>(row / (gh * gw));
      const int pr = static_cast<int>(row - static_cast<long>(b) * gh * gw);
      const int gy = pr / gw, gx = pr - gy * gw;
      v = img[((static_cast<long>(b) * Cin + c) * Himg + gy * p + i) * Wimg + gx * p + j];
    }
    out[row * ldo + k] = __float2bfloat16(v);
  }
}

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ x, long ldx,
                                                            __nv_bfloat16* __restrict__ y, long ldy, long rows, int cols) {
  const int nv = cols / 4;
  const long total = rows * nv;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = idx / nv;
    const int c = static_cast<int>(idx - r * nv) * 4;
    float a[4];
    Vec<float, 4>::load(x + r * ldx + c, a);
    Vec<__nv_bfloat16, 4>::store(y + r * ldy + c, a);
  }
}

// ------------------------------------------------- host helpers -------------------------------------------------
static inline int grid_for(long work_items, int block = 256, int max_blocks = 148 * 16) {
  long b = (work_items + block - 1) / block;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return static_cast<int>(b);
}
static inline int last_err() { return cudaGetLastError() == cudaSuccess ? OCTIC_OK : OCTIC_ERR_CUDA; }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace octic

using namespace octic;

extern "C" {

int octic_colsum_bf16(const void* x, long ldx, long T, int n_cols, float* out, void* stream) {
  if (!x || !out || T < 0 || n_cols <= 0 || (n_cols & 1) || (ldx & 1)) return OCTIC_ERR_ARG;
  if (T == 0) return OCTIC_OK;
  const int rows_per_block = 512;
  dim3 grid((n_cols + 63) / 64, static_cast<unsigned>((T + rows_per_block - 1) / rows_per_block)), block(32, 8);
  colsum_bf16_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), ldx, T, n_cols, out, rows_per_block);
  return last_err();
}

int octic_gelu_d8_fwd(const void* x, long ldx, void* y, long ldy, long T, int C, int dtype, void* stream) {
  if (!x || !y || T < 0 || C <= 0) return OCTIC_ERR_ARG;
  if (T == 0) return OCTIC_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == OCTIC_BF16) {
    const bool v8 = (C % 8 == 0) && (ldx % 8 == 0) && (ldy % 8 == 0) && aligned16(x) && aligned16(y);
    if (v8)
      gelu_d8_fwd_kernel<__nv_bfloat16, 8><<<grid_for(T * (C / 8)), 256, 0, s>>>(
          static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(y), ldy, T, C);
    else
      gelu_d8_fwd_kernel<__nv_bfloat16, 1><<<grid_for(T * C), 256, 0, s>>>(
          static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(y), ldy, T, C);
  } else if (dtype == OCTIC_F32) {
    const bool v4 = (C % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && aligned16(x) && aligned16(y);
    if (v4)
      gelu_d8_fwd_kernel<float, 4><<<grid_for(T * (C / 4)), 256, 0, s>>>(static_cast<const float*>(x), ldx,
                                                                         static_cast<float*>(y), ldy, T, C);
    else
      gelu_d8_fwd_kernel<float, 1><<<grid_for(T * C), 256, 0, s>>>(static_cast<const float*>(x), ldx,
                                                                   static_cast<float*>(y), ldy, T, C);
  } else {
    return OCTIC_ERR_ARG;
  }
  return last_err();
}

int octic_gelu_d8_bwd(const void* g, long ldg, const void* x, long ldx, void* gin, long ldgin, long T, int C,
                      int dtype, float* colsum, void* stream) {
  if (!g || !x || !gin || T < 0 || C <= 0) return OCTIC_ERR_ARG;
  if (T == 0) return OCTIC_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == OCTIC_BF16) {
    const bool v8 = (C % 8 == 0) && (ldx % 8 == 0) && (ldg % 8 == 0) && (ldgin % 8 == 0) && aligned16(x) &&
                    aligned16(g) && aligned16(gin);
    if (v8)
      gelu_d8_bwd_kernel<__nv_bfloat16, 8><<<grid_for(T * (C / 8)), 256, 0, s>>>(
          static_cast<const __nv_bfloat16*>(g), ldg, static_cast<const __nv_bfloat16*>(x), ldx,
          static_cast<__nv_bfloat16*>(gin), ldgin, T, C);
    else
      gelu_d8_bwd_kernel<__nv_bfloat16, 1><<<grid_for(T * C), 256, 0, s>>>(
          static_cast<const __nv_bfloat16*>(g), ldg, static_cast<const __nv_bfloat16*>(x), ldx,
          static_cast<__nv_bfloat16*>(gin), ldgin, T, C);
    int rc = last_err();
    if (rc) return rc;
    // bias gradient of the preceding LinearD8 (A1 block only carries a bias, but the sum is cheap: C columns)
    if (colsum != nullptr) return octic_colsum_bf16(gin, ldgin, T, C, colsum, stream);
    return OCTIC_OK;
  } else if (dtype == OCTIC_F32) {
    if (colsum != nullptr) return OCTIC_ERR_ARG;
    const bool v4 = (C % 4 == 0) && (ldx % 4 == 0) && (ldg % 4 == 0) && (ldgin % 4 == 0) && aligned16(x) &&
                    aligned16(g) && aligned16(gin);
    if (v4)
      gelu_d8_bwd_kernel<float, 4><<<grid_for(T * (C / 4)), 256, 0, s>>>(
          static_cast<const float*>(g), ldg, static_cast<const float*>(x), ldx, static_cast<float*>(gin), ldgin, T, C);
    else
      gelu_d8_bwd_kernel<float, 1><<<grid_for(T * C), 256, 0, s>>>(
          static_cast<const float*>(g), ldg, static_cast<const float*>(x), ldx, static_cast<float*>(gin), ldgin, T, C);
    return last_err();
  }
  return OCTIC_ERR_ARG;
}

int octic_gelu_bwd(const void* g, const void* x, void* gin, long n_rows, int n_cols, float* colsum, void* stream) {
  if (!g || !x || !gin || n_rows < 0 || n_cols <= 0 || (n_cols % 8)) return OCTIC_ERR_ARG;
  if (n_rows == 0) return OCTIC_OK;
  const long n8 = n_rows * n_cols / 8;
  gelu_bwd_kernel<<<grid_for(n8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(g), static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(gin), n8);
  int rc = last_err();
  if (rc) return rc;
  if (colsum != nullptr) return octic_colsum_bf16(gin, n_cols, n_rows, n_cols, colsum, stream);
  return OCTIC_OK;
}

#define OCTIC_LN_DISPATCH(KERNEL, TY, D8FLAG, ...)                                          \
  do {                                                                                      \
    const int nch = (D / 4 + 31) / 32;                                                      \
    if (nch <= 3) KERNEL<TY, D8FLAG, 3><<<grid, 256, smem, s>>>(__VA_ARGS__);               \
    else if (nch <= 6) KERNEL<TY, D8FLAG, 6><<<grid, 256, smem, s>>>(__VA_ARGS__);          \
    else if (nch <= 8) KERNEL<TY, D8FLAG, 8><<<grid, 256, smem, s>>>(__VA_ARGS__);          \
    else if (nch <= 10) KERNEL<TY, D8FLAG, 10><<<grid, 256, smem, s>>>(__VA_ARGS__);        \
    else if (nch <= 16) KERNEL<TY, D8FLAG, 16><<<grid, 256, smem, s>>>(__VA_ARGS__);        \
    else return OCTIC_ERR_ARG;                                                              \
  } while (0)

static int ln_fwd_common(bool d8, const float* x, long ldx, const float* alpha, const float* beta, float eps, void* y,
                         long ldy, int y_dtype, float* stats, long T, int D, void* stream) {
  if (!x || !alpha || !y || T < 0 || D <= 0) return OCTIC_ERR_ARG;
  if (d8 ? (D % 32 != 0) : (D % 4 != 0)) return OCTIC_ERR_ARG;   // C = D/8 must be a multiple of 4
  if ((ldx % 4) || (ldy % 4) || !aligned16(x) || !aligned16(y) || !aligned16(alpha) || (beta && !aligned16(beta)))
    return OCTIC_ERR_ALIGN;
  if (T == 0) return OCTIC_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = grid_for(T * 32, 256, 148 * 8);
  const int smem = 0;
  // lane-contiguous fast path: C % 16 == 0 with C/16 chunks per lane in {3, 8, 10} (ViT-S / L / H)
  static int ln_generic = -1;      // OCTIC_LN_GENERIC=1: lane-interleaved generic kernel instead of the lane-contiguous one (A/B)
  if (ln_generic < 0) {
    const char* e = getenv("OCTIC_LN_GENERIC");
    ln_generic = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  const int lc_cpl = (!ln_generic && d8 && (D % 128) == 0 && (ldy % 8) == 0) ? D / 128 : 0;
  if (y_dtype == OCTIC_BF16) {
    __nv_bfloat16* yy = static_cast<__nv_bfloat16*>(y);
    static int ln_lc = -1;           // OCTIC_LN_LC=1: the lane-contiguous kernel also for ViT-L / ViT-H widths (A/B)
    if (ln_lc < 0) {
      const char* e = getenv("OCTIC_LN_LC");
      ln_lc = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    if (d8 && lc_cpl == 10 && !ln_lc) layernorm_d8_fwd_il_kernel<__nv_bfloat16, 10><<<grid, 256, 0, s>>>(x, ldx, alpha, beta, eps, yy, ldy, stats, T, D / 8);
    else if (d8 && lc_cpl == 8 && !ln_lc) layernorm_d8_fwd_il_kernel<__nv_bfloat16, 8><<<grid, 256, 0, s>>>(x, ldx, alpha, beta, eps, yy, ldy, stats, T, D / 8);
    else if (d8 && lc_cpl == 10) layernorm_d8_fwd_lc_kernel<__nv_bfloat16, 10><<<grid, 256, 0, s>>>(x, ldx, alpha, beta, eps, yy, ldy, stats, T, D / 8);
    else if (d8 && lc_cpl == 8) layernorm_d8_fwd_lc_kernel<__nv_bfloat16, 8><<<grid, 256, 0, s>>>(x, ldx, alpha, beta, eps, yy, ldy, stats, T, D / 8);
    else if (d8 && lc_cpl == 3) layernorm_d8_fwd_lc_kernel<__nv_bfloat16, 3><<<grid, 256, 0, s>>>(x, ldx, alpha, beta, eps, yy, ldy, stats, T, D / 8);
    else if (d8) OCTIC_LN_DISPATCH(layernorm_fwd_kernel, __nv_bfloat16, true, x, ldx, alpha, beta, eps, yy, ldy, stats, T, D);
    else OCTIC_LN_DISPATCH(layernorm_fwd_plain_kernel, __nv_bfloat16, false, x, ldx, alpha, beta, eps, yy, ldy, stats, T, D);
  } else if (y_dtype == OCTIC_F32) {
    float* yy = static_cast<float*>(y);
    if (d8 && lc_cpl == 10) layernorm_d8_fwd_lc_kernel<float, 10><<<grid, 256, 0, s>>>(x, ldx, alpha, beta, eps, yy, ldy, stats, T, D / 8);
    else if (d8 && lc_cpl == 8) layernorm_d8_fwd_lc_kernel<float, 8><<<grid, 256, 0, s>>>(x, ldx, alpha, beta, eps, yy, ldy, stats, T, D / 8);
    else if (d8 && lc_cpl == 3) layernorm_d8_fwd_lc_kernel<float, 3><<<grid, 256, 0, s>>>(x, ldx, alpha, beta, eps, yy, ldy, stats, T, D / 8);
    else if (d8) OCTIC_LN_DISPATCH(layernorm_fwd_kernel, float, true, x, ldx, alpha, beta, eps, yy, ldy, stats, T, D);
    else OCTIC_LN_DISPATCH(layernorm_fwd_plain_kernel, float, false, x, ldx, alpha, beta, eps, yy, ldy, stats, T, D);
  } else {
    return OCTIC_ERR_ARG;
  }
  return last_err();
}

static int ln_bwd_common(bool d8, const void* dy, long lddy, int dy_dtype, const float* x, long ldx, const float* stats,
                         const float* alpha, const float* dx_in, float* dx_out, long lddx, float* dalpha, float* dbeta,
                         long T, int D, void* dx_bf16, float* dx_colsum, void* stream) {
  if (!dy || !x || !stats || !alpha || !dx_out || T < 0 || D <= 0) return OCTIC_ERR_ARG;
  if (d8 ? (D % 32 != 0) : (D % 4 != 0)) return OCTIC_ERR_ARG;
  if ((ldx % 4) || (lddy % 4) || (lddx % 4) || !aligned16(x) || !aligned16(dy) || !aligned16(dx_out) ||
      !aligned16(alpha) || (dx_in && !aligned16(dx_in)))
    return OCTIC_ERR_ALIGN;
  if (T == 0) return OCTIC_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int nthreads = ((D / 4) + 31) / 32 * 32;
  if (nthreads > 512) return OCTIC_ERR_ARG;   // D <= 2048
  const int tokens_per_block = kLnG;
  const long groups = (T + kLnG - 1) / kLnG;
  const int per_sm = nthreads <= 128 ? 8 : (nthreads <= 256 ? 4 : 2);
  const int grid = static_cast<int>(groups < 148L * per_sm ? groups : 148L * per_sm);
  if (dy_dtype == OCTIC_BF16) {
    const __nv_bfloat16* d = static_cast<const __nv_bfloat16*>(dy);
    if (d8) layernorm_bwd_kernel<__nv_bfloat16, true><<<grid, nthreads, 0, s>>>(d, lddy, x, ldx, stats, alpha, dx_in, dx_out, lddx, dalpha, dbeta, T, D, tokens_per_block, static_cast<__nv_bfloat16*>(dx_bf16), dx_colsum);
    else layernorm_bwd_kernel<__nv_bfloat16, false><<<grid, nthreads, 0, s>>>(d, lddy, x, ldx, stats, alpha, dx_in, dx_out, lddx, dalpha, dbeta, T, D, tokens_per_block, static_cast<__nv_bfloat16*>(dx_bf16), dx_colsum);
  } else if (dy_dtype == OCTIC_F32) {
    const float* d = static_cast<const float*>(dy);
    if (d8) layernorm_bwd_kernel<float, true><<<grid, nthreads, 0, s>>>(d, lddy, x, ldx, stats, alpha, dx_in, dx_out, lddx, dalpha, dbeta, T, D, tokens_per_block, static_cast<__nv_bfloat16*>(dx_bf16), dx_colsum);
    else layernorm_bwd_kernel<float, false><<<grid, nthreads, 0, s>>>(d, lddy, x, ldx, stats, alpha, dx_in, dx_out, lddx, dalpha, dbeta, T, D, tokens_per_block, static_cast<__nv_bfloat16*>(dx_bf16), dx_colsum);
  } else {
    return OCTIC_ERR_ARG;
  }
  return last_err();
}

int octic_layernorm_d8_fwd(const float* x, long ldx, const float* alpha, const float* beta, float eps, void* y,
                           long ldy, int y_dtype, float* stats, long T, int D, void* stream) {
  return ln_fwd_common(true, x, ldx, alpha, beta, eps, y, ldy, y_dtype, stats, T, D, stream);
}
int octic_layernorm_d8_bwd(const void* dy, long lddy, int dy_dtype, const float* x, long ldx, const float* stats,
                           const float* alpha, const float* dx_in, float* dx_out, long lddx, float* dalpha,
                           float* dbeta, long T, int D, void* dx_bf16, float* dx_colsum, void* stream) {
  return ln_bwd_common(true, dy, lddy, dy_dtype, x, ldx, stats, alpha, dx_in, dx_out, lddx, dalpha, dbeta, T, D, dx_bf16,
                       dx_colsum, stream);
}
int octic_layernorm_fwd(const float* x, long ldx, const float* w, const float* b, float eps, void* y, long ldy,
                        int y_dtype, float* stats, long T, int D, void* stream) {
  return ln_fwd_common(false, x, ldx, w, b, eps, y, ldy, y_dtype, stats, T, D, stream);
}
int octic_layernorm_bwd(const void* dy, long lddy, int dy_dtype, const float* x, long ldx, const float* stats,
                        const float* w, const float* dx_in, float* dx_out, long lddx, float* dw, float* db, long T,
                        int D, void* dx_bf16, float* dx_colsum, void* stream) {
  return ln_bwd_common(false, dy, lddy, dy_dtype, x, ldx, stats, w, dx_in, dx_out, lddx, dw, db, T, D, dx_bf16, dx_colsum,
                       stream);
}

int octic_layerscale_wgrad_finalize(const octic_lsfin_seg* segs, int nseg, void* stream) {
  if (segs == nullptr || nseg < 1 || nseg > 8) return OCTIC_ERR_ARG;
  LsFinParams p;
  memset(&p, 0, sizeof(p));
  p.nseg = nseg;
  int rows = 0;
  for (int i = 0; i < nseg; ++i) {
    if (!segs[i].dw || !segs[i].w || !segs[i].gamma || segs[i].N <= 0 || segs[i].K <= 0) return OCTIC_ERR_ARG;
    p.row_begin[i] = rows;
    rows += segs[i].N;
    LsFinSeg& S = p.seg[i];
    S.dw = segs[i].dw; S.w = segs[i].w; S.N = segs[i].N; S.K = segs[i].K; S.gamma = segs[i].gamma;
    S.bias = segs[i].bias; S.cs = segs[i].cs; S.dgamma = segs[i].dgamma; S.dbias = segs[i].dbias;
    S.dw_acc = segs[i].dw_acc;
  }
  p.row_begin[nseg] = rows;
  const int grid = (rows + 3) / 4 < 148 * 16 ? (rows + 3) / 4 : 148 * 16;
  layerscale_finalize_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return last_err();
}

int octic_layerscale_bwd(const float* dres, long lddres, const void* branch, long ldbr, const float* gamma,
                         const float* row_scale, int rows_per_sample, void* dy, long lddy, float* dgamma,
                         float* colsum, long T, int D, void* stream) {
  if (!dres || !dy || T < 0 || D <= 0 || (D % 4)) return OCTIC_ERR_ARG;
  if ((lddres % 4) || (lddy % 4) || (branch && (ldbr % 4)) || !aligned16(dres) || !aligned16(dy) ||
      (gamma && !aligned16(gamma)))
    return OCTIC_ERR_ALIGN;
  if (T == 0) return OCTIC_OK;
  if (rows_per_sample <= 0) rows_per_sample = 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int rows_per_block = 4;
  long want = (T + rows_per_block - 1) / rows_per_block;
  const int nchunks = D / 4;
  const int nv = (nchunks + 511) / 512;                              // column groups per thread
  const int block = ((nchunks + nv - 1) / nv + 31) / 32 * 32;        // D = 1280 -> 320 threads, one float4 column each
  const int per_sm = block <= 128 ? 8 : (block <= 256 ? 4 : (block <= 384 ? 3 : 2));
  const int grid = static_cast<int>(want < 148L * per_sm ? want : 148L * per_sm);
  const __nv_bfloat16* br = static_cast<const __nv_bfloat16*>(branch);
  __nv_bfloat16* d = static_cast<__nv_bfloat16*>(dy);
#define OCTIC_LS_LAUNCH(NVV)                                                                                        \
  do {                                                                                                            \
    if (br != nullptr)                                                                                            \
      layerscale_bwd_kernel<NVV, true><<<grid, block, 0, s>>>(dres, lddres, br, ldbr, gamma, row_scale,           \
                                                              rows_per_sample, d, lddy, dgamma, colsum, T, D,      \
                                                              rows_per_block);                                     \
    else                                                                                                          \
      layerscale_bwd_kernel<NVV, false><<<grid, block, 0, s>>>(dres, lddres, br, ldbr, gamma, row_scale,          \
                                                               rows_per_sample, d, lddy, dgamma, colsum, T, D,     \
                                                               rows_per_block);                                    \
  } while (0)
  if (nv == 1) OCTIC_LS_LAUNCH(1);
  else if (nv == 2) OCTIC_LS_LAUNCH(2);
  else if (nv <= 4) OCTIC_LS_LAUNCH(4);
  else return OCTIC_ERR_ARG;
  return last_err();
}

int octic_power_spectrum_fwd(const float* x, long ldx, void* y, long ldy, long T, int C, void* stream) {
  if (!x || !y || T < 0 || C <= 0 || (C % 4)) return OCTIC_ERR_ARG;
  if ((ldx % 4) || (ldy % 4) || !aligned16(x) || !aligned16(y)) return OCTIC_ERR_ALIGN;
  if (T == 0) return OCTIC_OK;
  power_spectrum_fwd_kernel<<<grid_for(T * (6 * C / 4)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, ldx, static_cast<__nv_bfloat16*>(y), ldy, T, C);
  return last_err();
}

int octic_power_spectrum_bwd(const void* dy, long lddy, int dy_dtype, const float* x, long ldx, float* dx,
                             long lddx, long T, int C, void* stream) {
  if (!dy || !x || !dx || T < 0 || C <= 0 || (C % 4)) return OCTIC_ERR_ARG;
  if ((ldx % 4) || (lddy % 4) || (lddx % 4) || !aligned16(x) || !aligned16(dy) || !aligned16(dx)) return OCTIC_ERR_ALIGN;
  if (T == 0) return OCTIC_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = grid_for(T * (6 * C / 4));
  if (dy_dtype == OCTIC_BF16)
    power_spectrum_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(dy), lddy, x, ldx, dx, lddx, T, C);
  else if (dy_dtype == OCTIC_F32)
    power_spectrum_bwd_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(dy), lddy, x, ldx, dx, lddx, T, C);
  else
    return OCTIC_ERR_ARG;
  return last_err();
}

int octic_bridge_permute(const float* x, long ldx, float* y, long ldy, long T, int C, void* stream) {
  if (!x || !y || x == y || T < 0 || C <= 0 || (C % 4)) return OCTIC_ERR_ARG;
  if ((ldx % 4) || (ldy % 4) || !aligned16(x) || !aligned16(y)) return OCTIC_ERR_ALIGN;
  if (T == 0) return OCTIC_OK;
  bridge_permute_kernel<<<grid_for(T * (8 * C / 4)), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ldx, y, ldy, T, C);
  return last_err();
}

int octic_im2col_patches(const float* img, int B, int Cin, int Himg, int Wimg, int p, void* out, long ldo,
                         void* stream) {
  if (!img || !out || B <= 0 || Cin <= 0 || p <= 0 || Himg % p || Wimg % p) return OCTIC_ERR_ARG;
  const int Kp = roundup64(Cin * p * p);
  if (ldo < Kp) return OCTIC_ERR_ARG;
  const long rows = static_cast<long>(B) * (Himg / p) * (Wimg / p);
  im2col_kernel<<<grid_for(rows * Kp), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      img, B, Cin, Himg, Wimg, p, static_cast<__nv_bfloat16*>(out), ldo, Kp);
  return last_err();
}

int octic_cast_f32_to_bf16(const float* x, long ldx, void* y, long ldy, long rows, int cols, void* stream) {
  if (!x || !y || rows < 0 || cols <= 0 || (cols % 4)) return OCTIC_ERR_ARG;
  if ((ldx % 4) || (ldy % 4) || !aligned16(x) || (reinterpret_cast<uintptr_t>(y) & 7)) return OCTIC_ERR_ALIGN;
  if (rows == 0) return OCTIC_OK;
  cast_f32_bf16_kernel<<<grid_for(rows * (cols / 4)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, ldx, static_cast<__nv_bfloat16*>(y), ldy, rows, cols);
  return last_err();
}

int octic_sparse_rowmap(const float* in, long ld_in, float* out, long ld_out, long rows, int n_out, int K, const int* idx,
                        const float* coef, int accumulate, void* stream) {
  if (!in || !out || !idx || !coef || rows < 0 || n_out <= 0 || K <= 0 || K > 16) return OCTIC_ERR_ARG;
  if (rows == 0) return OCTIC_OK;
  sparse_rowmap_kernel<<<grid_for(rows * n_out), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, ld_in, out, ld_out, rows, n_out,
                                                                                           K, idx, coef, accumulate);
  return last_err();
}

int octic_sparse_posmap(const float* in, long ld_in, float* out, long ld_out, int n_out, int cols, int K, const int* idx,
                        const float* coef, int accumulate, void* stream) {
  if (!in || !out || !idx || !coef || n_out <= 0 || cols <= 0 || K <= 0 || K > 16) return OCTIC_ERR_ARG;
  sparse_posmap_kernel<<<grid_for(static_cast<long>(n_out) * cols), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, ld_in, out, ld_out, n_out, cols, K, idx, coef, accumulate);
  return last_err();
}

}  // extern "C"
