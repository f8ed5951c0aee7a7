// Stand-alone probe of the CTA-pair (cta_group::2) tcgen05 path: cluster of 2 CTAs, each holding 128 rows of A and N/2
// rows of B (K-major, SWIZZLE_128B images built on the host), the leader issues M = 256 MMAs, both CTAs read their half
// of the accumulator from their own TMEM.  Not part of the library.
//   build:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tc2_selftest tc2_selftest.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../tc_common.cuh"

using namespace mde::tc;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_tf32_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
    probe_kernel(const unsigned char* __restrict__ a_img, const unsigned char* __restrict__ b_img, float* __restrict__ d_out,
                 int N, uint32_t* __restrict__ info) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* g = smem_dyn + (base - smem_u32(smem_dyn));
  unsigned char* sa = g;          // 128 rows x 128 B
  unsigned char* sb = g + 16384;  // N/2 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(g + 16384 + 16384);
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(g + 16384 + 16384 + 64);
  const uint32_t rank = cluster_ctarank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b_bytes = (N / 2) * 128;
  for (int i = threadIdx.x * 4; i < 16384; i += blockDim.x * 4)
    *reinterpret_cast<uint32_t*>(sa + i) = *reinterpret_cast<const uint32_t*>(a_img + rank * 16384 + i);
  for (int i = threadIdx.x * 4; i < b_bytes; i += blockDim.x * 4)
    *reinterpret_cast<uint32_t*>(sb + i) = *reinterpret_cast<const uint32_t*>(b_img + rank * b_bytes + i);
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc2(smem_u32((const void*)slot), 256);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) info[rank] = tmem;
  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(FMT_TF32, 256, (uint32_t)N, 0, 0);
    for (int j = 0; j < 4; ++j) {
      const uint64_t adesc = make_smem_desc(smem_u32(sa) + j * 32, 16, 1024, SWZ_128B);
      const uint64_t bdesc = make_smem_desc(smem_u32(sb) + j * 32, 16, 1024, SWZ_128B);
      umma2_tf32_ss(tmem, adesc, bdesc, idesc, j > 0);
    }
    umma2_commit_mc(smem_u32(bar), 3);
  }
  mbar_wait(smem_u32(bar), 0, 41);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32(tmem + c0 + ((uint32_t)(warp * 32) << 16), r);
    tmem_ld_wait();
    float* dst = d_out + (size_t)(rank * 128 + warp * 32 + lane) * N + c0;
    for (int i = 0; i < 32; ++i) dst[i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc2(tmem, 256);
  }
}

static void kmajor_sw128_image(const std::vector<float>& m, int rows, std::vector<unsigned char>& img) {
  // m: [rows][32] floats; image: row r at r*128 B, 16-byte chunk c stored at position c ^ (r & 7)
  img.assign((size_t)rows * 128, 0);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < 8; ++c)
      for (int e = 0; e < 4; ++e)
        reinterpret_cast<float*>(img.data() + (size_t)r * 128 + ((c ^ (r & 7)) << 4))[e] = m[(size_t)r * 32 + c * 4 + e];
}

int main() {
  for (int N : {128, 256, 64}) {
    std::vector<float> A(256 * 32), B((size_t)N * 32);
    srand(7);
    auto rnd = [] { return (float)((rand() % 17) - 8) / 8.0f; };  // exactly representable in TF32
    for (auto& v : A) v = rnd();
    for (auto& v : B) v = rnd();
    std::vector<unsigned char> a_img, a0, a1, b_img, b0, b1;
    kmajor_sw128_image(std::vector<float>(A.begin(), A.begin() + 128 * 32), 128, a0);
    kmajor_sw128_image(std::vector<float>(A.begin() + 128 * 32, A.end()), 128, a1);
    kmajor_sw128_image(std::vector<float>(B.begin(), B.begin() + (size_t)(N / 2) * 32), N / 2, b0);
    kmajor_sw128_image(std::vector<float>(B.begin() + (size_t)(N / 2) * 32, B.end()), N / 2, b1);
    a_img = a0; a_img.insert(a_img.end(), a1.begin(), a1.end());
    b_img = b0; b_img.insert(b_img.end(), b1.begin(), b1.end());
    unsigned char *da, *db;
    float* dd;
    uint32_t* di;
    cudaMalloc(&da, a_img.size());
    cudaMalloc(&db, b_img.size());
    cudaMalloc(&dd, sizeof(float) * 256 * N);
    cudaMalloc(&di, 16);
    cudaMemcpy(da, a_img.data(), a_img.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(db, b_img.data(), b_img.size(), cudaMemcpyHostToDevice);
    cudaMemset(dd, 0xff, sizeof(float) * 256 * N);
    cudaMemset(di, 0, 16);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    probe_kernel<<<2, 128, 40 * 1024>>>(da, db, dd, N, di);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> D((size_t)256 * N);
    uint32_t info[4];
    cudaMemcpy(D.data(), dd, sizeof(float) * D.size(), cudaMemcpyDeviceToHost);
    cudaMemcpy(info, di, 16, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    int bad = 0;
    for (int m = 0; m < 256; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < 32; ++k) ref += (double)A[m * 32 + k] * B[(size_t)n * 32 + k];
        const double err = fabs(ref - D[(size_t)m * N + n]);
        if (!(err < 1e-4)) ++bad;
        if (err > maxerr || err != err) maxerr = err;
      }
    printf("N=%d: cuda=%s tmem0=%u tmem1=%u max_err=%g bad=%d of %d   D[0][0]=%g D[128][0]=%g D[0][N/2]=%g\n", N,
           cudaGetErrorString(e), info[0], info[1], maxerr, bad, 256 * N, D[0], D[(size_t)128 * N], D[N / 2]);
    cudaFree(da); cudaFree(db); cudaFree(dd); cudaFree(di);
  }
  return 0;
}
