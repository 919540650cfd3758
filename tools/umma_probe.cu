// umma_probe -- one-MMA experiments that pin down tcgen05 shared-memory descriptor conventions (swizzle modes,
// K-major / MN-major, A operand from TMEM) before the attention kernels depend on them.  Test tool, not product code.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o build/umma_probe tools/umma_probe.cu
//   build/umma_probe <case>      (one case per process: an illegal descriptor may kill the context)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../octic_vits_b200/csrc/sm100_ptx.cuh"

using namespace octic;

// ---------------------------------------------------------------------------------------------------
struct ProbeCfg {
  int M, N, K;          // D[M=128][N] = A[M][K] * B[N][K]
  int a_src;            // 0: smem K-major, 1: TMEM, 2: smem MN-major (A stored [K][M])
  int b_major;          // 0: K-major (B stored [N][K]), 1: MN-major (B stored [K][N])
  int sw_a, sw_b;       // swizzle bytes 0 / 32 / 64 / 128
  int swap_none;        // SWIZZLE_NONE: swap the LBO/SBO roles
  int row0_a, row0_b;   // first row (K-major) used by the MMA inside a taller smem tile (multiple of 8)
  int n_split;          // >0: issue the MMA as two N pieces [0,n_split) and [n_split,N)
};

__host__ __device__ inline uint32_t layout_code(int sw) { return sw == 128 ? 2u : sw == 64 ? 4u : sw == 32 ? 6u : 0u; }

// byte offset of element (r, c) of a row-major [R][C] bf16 matrix stored in "compact swizzled atom columns"
__host__ __device__ inline uint32_t tile_off(int r, int c, int R, int sw) {
  if (sw == 0) return (c >> 3) * (R * 16) + r * 16 + (c & 7) * 2;                   // core matrices 8 x 16 B, [c/8][r]
  const int wc = sw / 2;                                                             // elements per atom row
  const int ac = c / wc, cb = (c % wc) * 2;                                          // atom column, byte in row
  const int x = sw == 128 ? (r & 7) : sw == 64 ? ((r >> 1) & 3) : ((r >> 2) & 1);
  return ac * (R * sw) + r * sw + (((cb >> 4) ^ x) << 4) + (cb & 15);
}

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t code) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(code) << 61;
  return d;
}
// operand stored row-major [R][C]; contraction runs along the columns (K-major); K-step ks covers columns 16ks..16ks+15
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t base, int R, int sw, int row0, int ks, int swap_none) {
  if (sw == 0) {
    const uint32_t a = base + (ks * 2) * (R * 16) + row0 * 16;
    const uint32_t kstride = R * 16, mstride = 128;
    return swap_none ? make_desc(a, mstride, kstride, 0) : make_desc(a, kstride, mstride, 0);
  }
  const int wc = sw / 2;
  const uint32_t a = base + ((ks * 16) / wc) * (R * sw) + row0 * sw + ((ks * 16) % wc) * 2;
  return make_desc(a, 0, 8 * sw, layout_code(sw));
}
// operand stored row-major [R = K rows][C = MN columns] (MN-major); K-step ks covers rows 16ks..16ks+15, MN starts at col0
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t base, int R, int sw, int col0, int ks, int swap_none) {
  if (sw == 0) {
    const uint32_t a = base + (col0 >> 3) * (R * 16) + (ks * 16) * 16;
    const uint32_t mnstride = R * 16, kstride = 128;
    return swap_none ? make_desc(a, kstride, mnstride, 0) : make_desc(a, mnstride, kstride, 0);
  }
  const int wc = sw / 2;
  const uint32_t a = base + (col0 / wc) * (R * sw) + (ks * 16) * sw;
  return make_desc(a, R * sw, 8 * sw, layout_code(sw));
}

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// A: global row-major [Ra][K] (K-major / TMEM) or [K][Ra... M] (MN-major).  B: [Rb][K] (K-major) or [K][N] (MN-major).
__global__ void __launch_bounds__(128, 1) probe_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                                       float* __restrict__ D, ProbeCfg c, int Ra, int Rb, long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  // smem tiles
  const int a_rows = c.a_src == 2 ? c.K : Ra, a_cols = c.a_src == 2 ? c.M : c.K;
  const int b_rows = c.b_major ? c.K : Rb, b_cols = c.b_major ? c.N : c.K;
  auto padded = [](int cols, int sw) { const int wc = sw ? sw / 2 : 8; return (cols + wc - 1) / wc * wc; };
  const uint32_t a_bytes = (c.a_src == 1) ? 0 : ((a_rows * padded(a_cols, c.sw_a) * 2 + 1023) & ~1023);
  uint8_t* sa = smem;
  uint8_t* sb = smem + a_bytes;
  const uint32_t b_bytes = b_rows * padded(b_cols, c.sw_b) * 2;
  for (uint32_t i = tid; i < (a_bytes + b_bytes) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  __syncthreads();
  if (c.a_src != 1)
    for (int i = tid; i < a_rows * a_cols; i += 128) {
      const int r = i / a_cols, cc = i % a_cols;
      *reinterpret_cast<__nv_bfloat16*>(sa + tile_off(r, cc, a_rows, c.sw_a)) = A[i];
    }
  for (int i = tid; i < b_rows * b_cols; i += 128) {
    const int r = i / b_cols, cc = i % b_cols;
    *reinterpret_cast<__nv_bfloat16*>(sb + tile_off(r, cc, b_rows, c.sw_b)) = B[i];
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_ptr, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_ptr;
  const uint32_t t_d = tbase, t_a = tbase + 256;
  if (c.a_src == 1) {
    // thread = row (lane): pack (k, k+1) bf16 pairs into 32-bit TMEM columns
    const int row = c.row0_a + tid;
    for (int k0 = 0; k0 < c.K; k0 += 16) {
      uint32_t r[8];
      for (int j = 0; j < 8; ++j) {
        const __nv_bfloat16 lo = row < Ra ? A[row * c.K + k0 + 2 * j] : __float2bfloat16(0.f);
        const __nv_bfloat16 hi = row < Ra ? A[row * c.K + k0 + 2 * j + 1] : __float2bfloat16(0.f);
        r[j] = static_cast<uint32_t>(__bfloat16_as_ushort(lo)) | (static_cast<uint32_t>(__bfloat16_as_ushort(hi)) << 16);
      }
      tmem_st_32x8(t_a + (static_cast<uint32_t>(warp * 32) << 16) + k0 / 2, r);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  long t0 = 0, t1 = 0;
  if (tid == 0) {
    t0 = clock64();
    const int npieces = c.n_split > 0 ? 2 : 1;
    for (int piece = 0; piece < npieces; ++piece) {
      const int n0 = piece == 0 ? 0 : c.n_split;
      const int nn = npieces == 1 ? c.N : (piece == 0 ? c.n_split : c.N - c.n_split);
      const uint32_t idesc = make_idesc_bf16(128, nn, c.a_src == 2 ? 1 : 0, c.b_major);
      for (int ks = 0; ks < c.K / 16; ++ks) {
        const uint64_t db = c.b_major ? desc_mnmajor(smem_u32(sb), b_rows, c.sw_b, n0, ks, c.swap_none)
                                      : desc_kmajor(smem_u32(sb), b_rows, c.sw_b, c.row0_b + n0, ks, c.swap_none);
        if (c.a_src == 1) {
          umma_bf16_ts(t_d + n0, t_a + ks * 8, db, idesc, ks != 0);
        } else {
          const uint64_t da = c.a_src == 2 ? desc_mnmajor(smem_u32(sa), a_rows, c.sw_a, c.row0_a, ks, c.swap_none)
                                           : desc_kmajor(smem_u32(sa), a_rows, c.sw_a, c.row0_a, ks, c.swap_none);
          umma_bf16(t_d + n0, da, db, idesc, ks != 0);
        }
      }
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  if (tid == 0) { t1 = clock64(); *cycles = t1 - t0; }
  for (int c0 = 0; c0 < c.N; c0 += 16) {
    uint32_t r[16];
    tmem_ld_32x16(t_d + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[tid * c.N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main(int argc, char** argv) {
  const int id = argc > 1 ? atoi(argv[1]) : 0;
  //             M    N    K  a_src b_major sw_a sw_b swap row0a row0b nsplit
  const ProbeCfg cases[] = {
      {128, 80, 80, 0, 0, 128, 128, 0, 0, 0, 0},    // 0 baseline: SS, both K-major, SW128 (what the GEMM kernel uses)
      {128, 80, 80, 0, 0, 64, 64, 0, 0, 0, 0},      // 1 K-major SW64
      {128, 80, 80, 0, 0, 32, 32, 0, 0, 0, 0},      // 2 K-major SW32
      {128, 80, 80, 0, 0, 0, 0, 0, 0, 0, 0},        // 3 K-major no swizzle (LBO = K stride, SBO = 8-row stride)
      {128, 80, 80, 0, 0, 0, 0, 1, 0, 0, 0},        // 4 K-major no swizzle, roles swapped
      {128, 80, 144, 0, 1, 32, 32, 0, 0, 0, 0},     // 5 B MN-major SW32 (PV-like: K = 144 keys, N = hd 80)
      {128, 80, 144, 0, 1, 64, 64, 0, 0, 0, 0},     // 6 B MN-major SW64
      {128, 80, 144, 0, 1, 128, 128, 0, 0, 0, 0},   // 7 B MN-major SW128
      {128, 80, 144, 0, 1, 0, 0, 0, 0, 0, 0},       // 8 B MN-major no swizzle
      {128, 80, 144, 0, 1, 0, 0, 1, 0, 0, 0},       // 9 B MN-major no swizzle, swapped
      {128, 80, 144, 1, 1, 32, 32, 0, 0, 0, 0},     // 10 A from TMEM, B MN-major SW32
      {128, 80, 144, 1, 1, 128, 128, 0, 0, 0, 0},   // 11 A from TMEM, B MN-major SW128
      {128, 144, 80, 0, 0, 32, 32, 0, 8, 136, 0},   // 12 K-major SW32 with row offsets inside taller tiles (Ra=272, Rb=280)
      {128, 80, 144, 2, 1, 32, 32, 0, 0, 0, 0},     // 13 A MN-major SW32 (stored [K][M]), B MN-major SW32
      {128, 80, 144, 2, 1, 128, 128, 0, 0, 0, 0},   // 14 A MN-major SW128, B MN-major SW128
      {128, 80, 80, 1, 0, 32, 32, 0, 0, 0, 0},      // 15 A from TMEM, B K-major SW32
      {128, 256, 80, 0, 0, 32, 32, 0, 0, 0, 0},     // 16 big N timing, SW32
      {128, 256, 80, 0, 0, 128, 128, 0, 0, 0, 0},   // 17 big N timing, SW128
      {128, 80, 272, 1, 1, 32, 32, 0, 0, 0, 0},     // 18 TS long K, SW32 (timing)
      {128, 80, 272, 1, 1, 128, 128, 0, 0, 0, 0},   // 19 TS long K, SW128 (timing)
      {128, 80, 144, 1, 1, 64, 64, 0, 0, 0, 0},     // 20 A from TMEM, B MN-major SW64
      {128, 80, 144, 1, 1, 128, 128, 0, 0, 0, 64},  // 21 TS, B MN-major SW128, N split 64 + 16
  };
  const int ncases = sizeof(cases) / sizeof(cases[0]);
  if (id < 0 || id >= ncases) { printf("case id out of range (0..%d)\n", ncases - 1); return 2; }
  ProbeCfg c = cases[id];
  const int Ra = (c.a_src == 2) ? c.M : (c.row0_a + 128 + (id == 12 ? 136 : 0));   // rows of the A tile in smem / gmem
  const int Rb = c.b_major ? c.K : (c.row0_b + c.N);
  const int a_elems = (c.a_src == 2) ? c.K * c.M : Ra * c.K;
  const int b_elems = c.b_major ? c.K * c.N : Rb * c.K;
  std::vector<__nv_bfloat16> hA(a_elems), hB(b_elems);
  std::vector<float> fA(a_elems), fB(b_elems);
  srand(1234 + id);
  for (int i = 0; i < a_elems; ++i) { fA[i] = bf((rand() % 2001 - 1000) / 500.f); hA[i] = __float2bfloat16(fA[i]); }
  for (int i = 0; i < b_elems; ++i) { fB[i] = bf((rand() % 2001 - 1000) / 500.f); hB[i] = __float2bfloat16(fB[i]); }
  __nv_bfloat16 *dA, *dB;
  float* dD;
  long* dcyc;
  cudaMalloc(&dA, a_elems * 2); cudaMalloc(&dB, b_elems * 2); cudaMalloc(&dD, 128 * c.N * 4); cudaMalloc(&dcyc, 8);
  cudaMemcpy(dA, hA.data(), a_elems * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), b_elems * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, 128 * c.N * 4);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe_kernel<<<1, 128, 200 * 1024>>>(dA, dB, dD, c, Ra, Rb, dcyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("case %2d: CUDA error %s\n", id, cudaGetErrorString(e)); return 1; }
  std::vector<float> hD(128 * c.N);
  long cyc = 0;
  cudaMemcpy(hD.data(), dD, 128 * c.N * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
  double max_err = 0, max_ref = 0;
  int bad = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < c.N; ++n) {
      double ref = 0;
      for (int k = 0; k < c.K; ++k) {
        const float a = c.a_src == 2 ? fA[k * c.M + m] : fA[(c.row0_a + m) * c.K + k];
        const float b = c.b_major ? fB[k * c.N + n] : fB[(c.row0_b + n) * c.K + k];
        ref += static_cast<double>(a) * b;
      }
      const double err = fabs(ref - hD[m * c.N + n]);
      if (err > max_err) max_err = err;
      if (fabs(ref) > max_ref) max_ref = fabs(ref);
      if (err > 1e-2 * (1 + fabs(ref))) ++bad;
    }
  printf("case %2d: M=%d N=%d K=%d a_src=%d b_major=%d sw=%d/%d swap=%d  max_err=%.4g (max |ref| %.3g) bad=%d  %s  issue->done %ld cycles\n",
         id, c.M, c.N, c.K, c.a_src, c.b_major, c.sw_a, c.sw_b, c.swap_none, max_err, max_ref, bad, bad == 0 ? "PASS" : "FAIL", cyc);
  return 0;
}
