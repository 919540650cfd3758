// epilogue_probe -- how fast can ONE CTA per SM drain 128 x 256 fp32 accumulator tiles from TMEM to a bf16 matrix in
// HBM?  Stand-alone experiment for the round-2 target "octic GEMMs are paced by their epilogue" (DESIGN.md section 3:
// octic fc1 runs 144 us where its DRAM traffic needs 57 us; tensor pipe 31 %, issue slots 41 %).  No MMA is issued:
// the accumulators are whatever TMEM holds, the output shape is octic fc1's (32 896 x 5120 bf16 = 337 MB, 5140 tiles
// over 148 persistent CTAs), so the numbers are upper bounds for the epilogue alone.  Test tool, not product code.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o build/epilogue_probe tools/epilogue_probe.cu
//   build/epilogue_probe <variant> [warps = 16] [reps = 5]
//
// variants
//   0  staged   tcgen05.ld.32x32b.x32 -> st.shared (rows 144 B apart) -> ld.shared -> 16-byte st.global, 8 rows x 64 B
//               per instruction (the product kernel's vector path, bias / mode logic stripped)
//   1  tma      tcgen05.ld.x32 (two chunks) -> bf16 -> st.shared into a 32-row x 64-column SWIZZLE_128B box ->
//               fence.proxy.async -> one elected lane issues cp.async.bulk.tensor.2d (TMA store); two boxes per warp
//               so that the store of one overlaps the fill of the other
//   2  ldonly   TMEM reads only (what tcgen05.ld sustains with this many warps)
//   3  stonly   the staged stores without TMEM reads (what the store side sustains)
// Prints clocks per tile (CTA 0), the time of the launch, and GB/s of output written.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../octic_vits_b200/csrc/sm100_ptx.cuh"

using namespace octic;

constexpr int kRows = 32896, kCols = 5120, kTileM = 128, kTileN = 256;
constexpr int kTilesN = kCols / kTileN, kTilesM = kRows / kTileM;     // 20 x 257
constexpr int kTiles = kTilesM * kTilesN;
constexpr int kMaxWarps = 16;
constexpr int kStageBytes = 32 * 144;                                  // staged path: per warp, 32 rows x 144 B
constexpr int kBoxBytes = 32 * 128;                                    // TMA path: per box, 32 rows x 64 bf16

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds16(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t smem_src, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

template <int VARIANT>
__global__ void __launch_bounds__(kMaxWarps * 32 + 32, 1)
probe_kernel(__nv_bfloat16* __restrict__ out, const __grid_constant__ CUtensorMap omap, int warps, long long* clocks) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);          // two 256-column accumulator stages, as in the GEMM
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  long long t0 = 0;
  if (threadIdx.x == 32) t0 = clock64();
  if (warp >= 1 && warp <= warps) {
    // a warp may only touch the TMEM lane quarter (warp id % 4); the four warps of a quarter interleave chunks / boxes
    const int ew = warp - 1, q = warp & 3, first = ew >> 2, step = warps >> 2;
    const uint32_t sbase = (smem_u32(smem) + 1023u) & ~1023u;      // SWIZZLE_128B boxes need 1024-byte alignment
    int it = 0;
    if (VARIANT == 1) {
      // boxes of 64 columns: box b of the tile covers columns 64b..64b+63; this warp takes boxes first, first+step, ...
      const uint32_t box0 = sbase + (ew * 2) * kBoxBytes;
      int parity = 0;
      for (int tile = blockIdx.x; tile < kTiles; tile += gridDim.x, ++it) {
        const int mb = tile / kTilesN, nb = tile % kTilesN;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (it & 1) * 256;
        for (int b = first; b < kTileN / 64; b += step) {
          uint32_t r0[32], r1[32];
          tmem_ld_32x32(taddr + b * 64, r0);
          tmem_ld_32x32(taddr + b * 64 + 32, r1);
          const uint32_t box = box0 + parity * kBoxBytes;
          if (lane == 0) bulk_wait_read<1>();            // the store that used this box two fills ago has read it
          __syncwarp();
          tmem_ld_wait();
          // row = lane: 128 B = eight 16-byte pieces; SWIZZLE_128B puts piece j of row r at piece (j ^ (r & 7))
          const uint32_t rowp = box + lane * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts16(rowp + ((j ^ (lane & 7)) << 4), pack2(__uint_as_float(r0[8 * j]), __uint_as_float(r0[8 * j + 1])),
                  pack2(__uint_as_float(r0[8 * j + 2]), __uint_as_float(r0[8 * j + 3])),
                  pack2(__uint_as_float(r0[8 * j + 4]), __uint_as_float(r0[8 * j + 5])),
                  pack2(__uint_as_float(r0[8 * j + 6]), __uint_as_float(r0[8 * j + 7])));
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts16(rowp + (((j + 4) ^ (lane & 7)) << 4), pack2(__uint_as_float(r1[8 * j]), __uint_as_float(r1[8 * j + 1])),
                  pack2(__uint_as_float(r1[8 * j + 2]), __uint_as_float(r1[8 * j + 3])),
                  pack2(__uint_as_float(r1[8 * j + 4]), __uint_as_float(r1[8 * j + 5])),
                  pack2(__uint_as_float(r1[8 * j + 6]), __uint_as_float(r1[8 * j + 7])));
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&omap, box, nb * kTileN + b * 64, mb * kTileM + q * 32);
            bulk_commit();
          }
          parity ^= 1;
        }
      }
      if (lane == 0) bulk_wait_read<0>();
    } else {
      const uint32_t stg = sbase + ew * kStageBytes;
      const int cq = lane & 3, rsub = lane >> 2;
      for (int tile = blockIdx.x; tile < kTiles; tile += gridDim.x, ++it) {
        const int mb = tile / kTilesN, nb = tile % kTilesN;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (it & 1) * 256;
        for (int c = first; c < kTileN / 32; c += step) {
          uint32_t r[32];
          if (VARIANT != 3) {
            tmem_ld_32x32(taddr + c * 32, r);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(static_cast<float>(tile + j + lane));
          }
          if (VARIANT == 2) {
            uint32_t x = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) x ^= r[j];
            if (x == 0x12345678u) out[lane] = __float2bfloat16(1.f);     // keeps the loads alive, never true in practice
            continue;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) sts16(stg + lane * 144 + j * 16, r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          __syncwarp();
          __nv_bfloat16* o = out + static_cast<long>(mb * kTileM + q * 32 + rsub) * kCols + nb * kTileN + c * 32 + cq * 8;
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const uint32_t a = stg + (h * 8 + rsub) * 144 + cq * 32;
            const uint4 x0 = lds16(a), x1 = lds16(a + 16);
            uint4 v;
            v.x = pack2(__uint_as_float(x0.x), __uint_as_float(x0.y));
            v.y = pack2(__uint_as_float(x0.z), __uint_as_float(x0.w));
            v.z = pack2(__uint_as_float(x1.x), __uint_as_float(x1.y));
            v.w = pack2(__uint_as_float(x1.z), __uint_as_float(x1.w));
            *reinterpret_cast<uint4*>(o + static_cast<long>(h) * 8 * kCols) = v;
          }
          __syncwarp();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 32 && clocks != nullptr) clocks[blockIdx.x] = clock64() - t0;
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

template <int V>
static int run(__nv_bfloat16* out, const CUtensorMap& map, int warps, int reps, int sms) {
  const size_t smem = (V == 1 ? static_cast<size_t>(warps) * 2 * kBoxBytes : static_cast<size_t>(warps) * kStageBytes) + 1024;
  CK(cudaFuncSetAttribute(probe_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  long long* clocks = nullptr;
  CK(cudaMalloc(&clocks, sizeof(long long) * sms));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int threads = (warps + 1) * 32;
  probe_kernel<V><<<sms, threads, smem>>>(out, map, warps, clocks);      // warm-up
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) probe_kernel<V><<<sms, threads, smem>>>(out, map, warps, clocks);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= reps;
  long long c0 = 0;
  CK(cudaMemcpy(&c0, clocks, sizeof(long long), cudaMemcpyDeviceToHost));
  const double tiles_per_cta = static_cast<double>(kTiles) / sms;
  const double bytes = V == 2 ? 0.0 : static_cast<double>(kRows) * kCols * 2;
  printf("variant %d warps %2d: %8.1f us per launch, %7.0f clk per tile (CTA 0, %.1f tiles), %6.0f GB/s written\n", V, warps,
         ms * 1e3, c0 / tiles_per_cta, tiles_per_cta, bytes / (ms * 1e-3) / 1e9);
  cudaFree(clocks);
  return 0;
}

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  int warps = argc > 2 ? atoi(argv[2]) : 16;
  const int reps = argc > 3 ? atoi(argv[3]) : 5;
  if (warps != 4 && warps != 8 && warps != 16) { printf("warps must be 4, 8 or 16\n"); return 1; }
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  __nv_bfloat16* out = nullptr;
  CK(cudaMalloc(&out, static_cast<size_t>(kRows) * kCols * 2));
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || ptr == nullptr) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  CUtensorMap map;
  cuuint64_t dims[2] = {kCols, kRows};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(kCols) * 2};
  cuuint32_t box[2] = {64, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(ptr)(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, dims, strides, box, estr,
                                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("tensor map encode failed: %d\n", static_cast<int>(r)); return 1; }
  switch (variant) {
    case 0: return run<0>(out, map, warps, reps, sms);
    case 1: return run<1>(out, map, warps, reps, sms);
    case 2: return run<2>(out, map, warps, reps, sms);
    case 3: return run<3>(out, map, warps, reps, sms);
    default: printf("variant 0..3\n"); return 1;
  }
}
