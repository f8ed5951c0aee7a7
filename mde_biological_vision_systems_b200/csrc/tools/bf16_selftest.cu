// Stand-alone bring-up tool (not part of the library): TMA-fed tcgen05.mma.kind::f16 with BF16 K-major operands in the
// two shared-memory layouts the bf16x3 kernels use -- SWIZZLE_128B (64-element K chunks: chain, patch embedding) and
// SWIZZLE_64B (32-element K chunks: conv3x3) -- including the row-shifted descriptor start addresses the implicit-GEMM
// convolution relies on for its dy taps.  Prints max |error| against the exact product of the bf16 inputs.
//   build:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o bf16_selftest bf16_selftest.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#include "../tc_common.cuh"

using namespace mde::tc;

constexpr int ROWS_A = 160, N = 128, K = 64;

__global__ void __launch_bounds__(128, 1)
    bf16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int kc, int row_off,
                uint32_t swz, uint32_t sbo, float* __restrict__ d_out) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* g = smem_dyn + (base - smem_u32(smem_dyn));
  const int chunks = K / kc;
  const uint32_t row_bytes = kc * 2;
  const uint32_t a_chunk = ROWS_A * row_bytes, b_chunk = N * row_bytes;
  const uint32_t s_a = base, s_b = base + 32768;
  const uint32_t bar_full = base + 65536, bar_done = bar_full + 8;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(g + 65536 + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_full, 1);
    mbar_init(bar_done, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32((const void*)slot), 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar_full, chunks * (a_chunk + b_chunk));
    for (int c = 0; c < chunks; ++c) {
      tma_load_2d(s_a + c * a_chunk, &map_a, bar_full, c * kc, 0);
      tma_load_2d(s_b + c * b_chunk, &map_b, bar_full, c * kc, 0);
    }
    mbar_wait(bar_full, 0, 1);
    tc_fence_after();
    const uint32_t idesc = make_idesc(FMT_BF16, 128, N, 0, 0);
    uint32_t acc = 0;
    for (int c = 0; c < chunks; ++c)
      for (int j = 0; j < kc / 16; ++j) {
        const uint64_t ad = make_smem_desc(s_a + c * a_chunk + row_off * row_bytes + j * 32, 16, sbo, swz);
        const uint64_t bd = make_smem_desc(s_b + c * b_chunk + j * 32, 16, sbo, swz);
        umma_f16_ss(tmem, ad, bd, idesc, acc);
        acc = 1;
      }
    umma_commit(bar_done);
  }
  __syncwarp();
  mbar_wait(bar_done, 0, 2);
  tc_fence_after();
  for (uint32_t c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32(tmem + c0 + ((uint32_t)(warp * 32) << 16), r);
    tmem_ld_wait();
    const int row = warp * 32 + lane;
#pragma unroll
    for (int i = 0; i < 32; ++i) d_out[(size_t)row * N + c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

static uint16_t f2bf(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  u += 0x7FFF + ((u >> 16) & 1);
  return (uint16_t)(u >> 16);
}
static float bf2f(uint16_t b) {
  uint32_t u = (uint32_t)b << 16;
  float v;
  memcpy(&v, &u, 4);
  return v;
}

int main() {
  std::vector<uint16_t> ha(ROWS_A * K), hb(N * K);
  srand(1);
  for (auto& v : ha) v = f2bf((float)rand() / RAND_MAX - 0.5f);
  for (auto& v : hb) v = f2bf((float)rand() / RAND_MAX - 0.5f);
  uint16_t *da, *db;
  float* dd;
  cudaMalloc(&da, ha.size() * 2);
  cudaMalloc(&db, hb.size() * 2);
  cudaMalloc(&dd, 128 * N * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  int fails = 0;
  struct Case { int kc; uint32_t swz; CUtensorMapSwizzle tswz; uint32_t sbo; const char* name; };
  const Case cases[] = {{64, SWZ_128B, CU_TENSOR_MAP_SWIZZLE_128B, 1024, "SW128 kc=64 sbo=1024"},
                        {32, SWZ_64B, CU_TENSOR_MAP_SWIZZLE_64B, 512, "SW64  kc=32 sbo=512"},
                        {32, SWZ_64B, CU_TENSOR_MAP_SWIZZLE_64B, 1024, "SW64  kc=32 sbo=1024 (expected wrong)"}};
  for (const Case& cs : cases) {
    CUtensorMap ma, mb;
    const uint64_t dims_a[2] = {K, ROWS_A}, dims_b[2] = {K, N};
    const uint64_t strides[1] = {K * 2};
    const uint32_t box_a[2] = {(uint32_t)cs.kc, ROWS_A}, box_b[2] = {(uint32_t)cs.kc, N};
    if (!encode_bf16(&ma, da, 2, dims_a, strides, box_a, cs.tswz) || !encode_bf16(&mb, db, 2, dims_b, strides, box_b, cs.tswz)) {
      printf("encode failed\n");
      return 1;
    }
    for (int row_off : {0, 8, 16, 24}) {
      cudaMemset(dd, 0, 128 * N * 4);
      bf16_kernel<<<1, 128, 80 * 1024>>>(ma, mb, cs.kc, row_off, cs.swz, cs.sbo, dd);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("%s row_off %d: CUDA error %s\n", cs.name, row_off, cudaGetErrorString(e));
        return 2;
      }
      std::vector<float> hd(128 * N);
      cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int k = 0; k < K; ++k) ref += (double)bf2f(ha[(m + row_off) * K + k]) * bf2f(hb[n * K + k]);
          maxerr = fmax(maxerr, fabs(ref - hd[m * N + n]));
        }
      const bool ok = maxerr < 1e-4;
      printf("%-40s row_off %2d: max err %.3e %s\n", cs.name, row_off, maxerr, ok ? "OK" : "MISMATCH");
      if (!ok && cs.sbo != 1024 + 0 * cs.kc && cs.kc == 32 && cs.sbo == 512) ++fails;
      if (!ok && cs.kc == 64) ++fails;
    }
  }
  printf(fails ? "SELFTEST FAILED (%d)\n" : "SELFTEST PASSED\n", fails);
  return fails ? 3 : 0;
}
