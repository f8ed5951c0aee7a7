// Stand-alone bring-up tool for the tcgen05 / TMEM / TMA building blocks of tc_common.cuh (not part of the library).
// The HOST builds the exact shared-memory byte image of each operand for a layout hypothesis; the kernel copies the
// images verbatim into 1024-aligned shared memory, issues the MMAs with the descriptor fields under test, reads the
// accumulator back from TMEM and the host compares with the exact product.  One process sweeps many hypotheses.
//   build:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tc_selftest tc_selftest.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../tc_common.cuh"

using namespace mde::tc;

struct Params {
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo, swz_a, swz_b, version;
  uint32_t a_mn_major, b_mn_major, N, ksteps, a_step_bytes, b_step_bytes, a_bytes, b_bytes;
};

__global__ void __launch_bounds__(128, 1)
    selftest_kernel(Params p, const unsigned char* __restrict__ a_img, const unsigned char* __restrict__ b_img,
                    float* __restrict__ d_out, uint32_t* __restrict__ info) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* g = smem_dyn + (base - smem_u32(smem_dyn));
  unsigned char* sa = g;                 // up to 64 KB
  unsigned char* sb = g + 65536;         // up to 128 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(g + 65536 + 131072);
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(g + 65536 + 131072 + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x * 4; i < p.a_bytes; i += blockDim.x * 4)
    *reinterpret_cast<uint32_t*>(sa + i) = *reinterpret_cast<const uint32_t*>(a_img + i);
  for (uint32_t i = threadIdx.x * 4; i < p.b_bytes; i += blockDim.x * 4)
    *reinterpret_cast<uint32_t*>(sb + i) = *reinterpret_cast<const uint32_t*>(b_img + i);
  fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32((const void*)slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) info[0] = tmem;

  // ---- TMEM st/ld round trip at columns [384, 416) ----
  {
    uint32_t w[32], r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) w[i] = (uint32_t)(threadIdx.x * 1000 + i);
    const uint32_t taddr = tmem + 384 + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), "r"(w[9]),
        "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]), "r"(w[16]), "r"(w[17]), "r"(w[18]),
        "r"(w[19]), "r"(w[20]), "r"(w[21]), "r"(w[22]), "r"(w[23]), "r"(w[24]), "r"(w[25]), "r"(w[26]), "r"(w[27]),
        "r"(w[28]), "r"(w[29]), "r"(w[30]), "r"(w[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tmem_ld_32x32(taddr, r);
    tmem_ld_wait();
    uint32_t bad = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) bad += (r[i] != w[i]);
    if (bad) atomicAdd(&info[1], bad);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(FMT_TF32, 128, p.N, p.a_mn_major, p.b_mn_major);
    info[2] = idesc;
    for (uint32_t j = 0; j < p.ksteps; ++j) {
      const uint64_t ad = make_smem_desc(smem_u32(sa) + j * p.a_step_bytes, p.a_lbo, p.a_sbo, p.swz_a, p.version);
      const uint64_t bd = make_smem_desc(smem_u32(sb) + j * p.b_step_bytes, p.b_lbo, p.b_sbo, p.swz_b, p.version);
      if (j == 0) {
        info[4] = (uint32_t)ad; info[5] = (uint32_t)(ad >> 32); info[6] = (uint32_t)bd; info[7] = (uint32_t)(bd >> 32);
      }
      umma_tf32_ss(tmem, ad, bd, idesc, j != 0);
    }
    umma_commit(smem_u32(bar));
  }
  __syncwarp();
  mbar_wait(smem_u32(bar), 0, 9);
  tc_fence_after();
  for (uint32_t c0 = 0; c0 < p.N; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32(tmem + c0 + ((uint32_t)(warp * 32) << 16), r);
    tmem_ld_wait();
    const int row = warp * 32 + lane;
#pragma unroll
    for (int i = 0; i < 32; ++i) d_out[(size_t)row * p.N + c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// TMA dump: load one 32 x 32 fp32 box (SWIZZLE_128B) of a [rows][cols] matrix and copy the 4096 smem bytes out
__global__ void tma_dump_kernel(const __grid_constant__ CUtensorMap map, int c0, int c1, float* out) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* g = smem_dyn + (base - smem_u32(smem_dyn));
  uint64_t* bar = reinterpret_cast<uint64_t*>(g + 8192);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(bar), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(smem_u32(bar), 4096);
    tma_load_2d(base, &map, smem_u32(bar), c0, c1);
  }
  mbar_wait(smem_u32(bar), 0, 8);
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = reinterpret_cast<float*>(g)[i];
}

// ---------------------------------------------------------------------------------------------------------------
// host-side layout images
static inline float aval(int m, int k) { return (float)((m * 3 + k * 7) % 11 - 5); }   // small exact integers
static inline float bval(int n, int k) { return (float)((n * 5 + k * 3) % 13 - 6); }

// K-major, no swizzle: core matrix = 8 rows x 16 B; k-core stride kc_stride, 8-row-group stride grp_stride
static void img_kmajor_none(std::vector<unsigned char>& img, int rows, int K, uint32_t kc_stride, uint32_t grp_stride,
                            float (*f)(int, int)) {
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < K; ++k) {
      const size_t off = (size_t)(r / 8) * grp_stride + (size_t)(k / 4) * kc_stride + (r % 8) * 16 + (k % 4) * 4;
      if (off + 4 > img.size()) img.resize(off + 4);
      const float v = f(r, k);
      memcpy(&img[off], &v, 4);
    }
}
// K-major, SWIZZLE_128B: row r = 128 B (32 fp32 of K), 16-byte chunk index XOR (r % 8); 8-row atoms 1024 B apart
static void img_kmajor_sw128(std::vector<unsigned char>& img, int rows, int K, float (*f)(int, int)) {
  img.assign((size_t)rows * 128 * ((K + 31) / 32), 0);
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < K; ++k) {
      const int kc = k / 32, kk = k % 32;
      const size_t off = (size_t)kc * rows * 128 + (size_t)r * 128 + (((kk / 4) ^ (r % 8)) * 16) + (kk % 4) * 4;
      const float v = f(r, k);
      memcpy(&img[off], &v, 4);
    }
}
// MN-major, SWIZZLE_128B: one "row" = 128 B = 32 consecutive m for a fixed k; chunk index XOR (k % 8);
// mn-atom (32 m) stride mn_stride, k-atom (8 k) stride k_stride
static void img_mnmajor_sw128(std::vector<unsigned char>& img, int M, int K, uint32_t mn_stride, uint32_t k_stride,
                              float (*f)(int, int)) {
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) {
      const size_t off = (size_t)(m / 32) * mn_stride + (size_t)(k / 8) * k_stride + (k % 8) * 128 +
                         ((((m % 32) / 4) ^ (k % 8)) * 16) + (m % 4) * 4;
      if (off + 4 > img.size()) img.resize(off + 4);
      const float v = f(m, k);
      memcpy(&img[off], &v, 4);
    }
}

// MN-major, SWIZZLE_128B_BASE32B (descriptor layout type 1): a row = 128 B = 32 consecutive m for a fixed k;
// 32-byte chunk index XOR (k % 4); k-atom = 4 rows (512 B)
static void img_mnmajor_sw128_32b(std::vector<unsigned char>& img, int M, int K, uint32_t mn_stride, uint32_t k_stride,
                                  float (*f)(int, int)) {
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) {
      const size_t off = (size_t)(m / 32) * mn_stride + (size_t)(k / 4) * k_stride + (k % 4) * 128 +
                         ((((m % 32) / 8) ^ (k % 4)) * 32) + (m % 8) * 4;
      if (off + 4 > img.size()) img.resize(off + 4);
      const float v = f(m, k);
      memcpy(&img[off], &v, 4);
    }
}
// MN-major, no swizzle: core matrix = 8 k-rows x 16 B (4 consecutive m); m-group stride mg_stride, k-group stride kg_stride
static void img_mnmajor_none(std::vector<unsigned char>& img, int M, int K, uint32_t mg_stride, uint32_t kg_stride,
                             float (*f)(int, int)) {
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) {
      const size_t off = (size_t)(m / 4) * mg_stride + (size_t)(k / 8) * kg_stride + (k % 8) * 16 + (m % 4) * 4;
      if (off + 4 > img.size()) img.resize(off + 4);
      const float v = f(m, k);
      memcpy(&img[off], &v, 4);
    }
}

static int run(const char* name, Params p, const std::vector<unsigned char>& a, const std::vector<unsigned char>& b, int K) {
  unsigned char *da, *db;
  float* dd;
  uint32_t* di;
  p.a_bytes = (uint32_t)((a.size() + 3) & ~3u);
  p.b_bytes = (uint32_t)((b.size() + 3) & ~3u);
  if (p.a_bytes > 65536 || p.b_bytes > 131072) {
    printf("%-34s SKIP (image too large a=%u b=%u)\n", name, p.a_bytes, p.b_bytes);
    return 1;
  }
  cudaMalloc(&da, 65536);
  cudaMalloc(&db, 131072);
  cudaMalloc(&dd, 128 * 256 * 4);
  cudaMalloc(&di, 64);
  cudaMemset(da, 0, 65536);
  cudaMemset(db, 0, 131072);
  cudaMemset(dd, 0xff, 128 * 256 * 4);
  cudaMemset(di, 0, 64);
  cudaMemcpy(da, a.data(), a.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size(), cudaMemcpyHostToDevice);
  const int smem = 65536 + 131072 + 128 + 1024;
  cudaFuncSetAttribute(selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  selftest_kernel<<<1, 128, smem>>>(p, da, db, dd, di);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-34s CUDA ERROR %s\n", name, cudaGetErrorString(e));
    return -1;
  }
  std::vector<float> d(128 * p.N);
  uint32_t info[16];
  cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(info, di, 64, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  int nonzero = 0, nan = 0;
  for (int m = 0; m < 128; ++m)
    for (uint32_t n = 0; n < p.N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)aval(m, k) * bval(n, k);
      const float v = d[(size_t)m * p.N + n];
      if (v != v) { nan++; continue; }
      if (v != 0.f) nonzero++;
      const double er = fabs(v - ref);
      if (er > maxerr) maxerr = er;
    }
  printf("%-34s lbo/sbo A %5u/%5u B %5u/%5u ver %u : maxerr %-10.3g nonzero %5d nan %d | tmem %08x st/ld bad %u idesc %08x "
         "adesc %08x%08x bdesc %08x%08x | D[0][0..3] %g %g %g %g D[1][0] %g D[37][5] %g\n",
         name, p.a_lbo, p.a_sbo, p.b_lbo, p.b_sbo, p.version, maxerr, nonzero, nan, info[0], info[1], info[2], info[5],
         info[4], info[7], info[6], d[0], d[1], d[2], d[3], d[p.N], d[37 * p.N + 5]);
  {
    double r00 = 0, r10 = 0, r375 = 0;
    for (int k = 0; k < K; ++k) {
      r00 += (double)aval(0, k) * bval(0, k);
      r10 += (double)aval(1, k) * bval(0, k);
      r375 += (double)aval(37, k) * bval(5, k);
    }
    printf("%-34s expected D[0][0] %g D[1][0] %g D[37][5] %g\n", "", r00, r10, r375);
  }
  cudaFree(da); cudaFree(db); cudaFree(dd); cudaFree(di);
  return maxerr < 1e-3 ? 0 : 1;
}

int main() {
  int dev = 0;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, dev);
  printf("device %s cc %d.%d\n", prop.name, prop.major, prop.minor);
  std::vector<unsigned char> a, b;

  // V1: K-major both, no swizzle, K = 8 (one MMA).  LBO = k-core stride, SBO = 8-row-group stride (and swapped)
  for (int swap = 0; swap < 2; ++swap) {
    a.clear(); b.clear();
    img_kmajor_none(a, 128, 8, 128, 256, aval);
    img_kmajor_none(b, 128, 8, 128, 256, bval);
    Params p{};
    p.a_lbo = swap ? 256 : 128; p.a_sbo = swap ? 128 : 256; p.b_lbo = p.a_lbo; p.b_sbo = p.a_sbo;
    p.swz_a = SWZ_NONE; p.swz_b = SWZ_NONE; p.version = 1; p.N = 128; p.ksteps = 1;
    run(swap ? "V1 K/K none (lbo<->sbo swapped)" : "V1 K/K none", p, a, b, 8);
  }
  // V2: K-major both, SWIZZLE_128B, K = 32 (4 MMAs advancing the start address by 32 B)
  for (int ver = 1; ver >= 0; --ver) {
    img_kmajor_sw128(a, 128, 32, aval);
    img_kmajor_sw128(b, 128, 32, bval);
    Params p{};
    p.a_lbo = 16; p.a_sbo = 1024; p.b_lbo = 16; p.b_sbo = 1024; p.swz_a = SWZ_128B; p.swz_b = SWZ_128B; p.version = ver;
    p.N = 128; p.ksteps = 4; p.a_step_bytes = 32; p.b_step_bytes = 32;
    run(ver ? "V2 K/K sw128" : "V2 K/K sw128 (version 0)", p, a, b, 32);
  }
  // V3: A MN-major SW128 (mn-atom stride 4096, k-atom stride 1024, as the head-chain kernel), B K-major SW128, K = 32
  for (int swap = 0; swap < 2; ++swap) {
    a.clear();
    img_mnmajor_sw128(a, 128, 32, 4096, 1024, aval);
    img_kmajor_sw128(b, 128, 32, bval);
    Params p{};
    p.a_lbo = swap ? 1024 : 4096; p.a_sbo = swap ? 4096 : 1024; p.b_lbo = 16; p.b_sbo = 1024;
    p.swz_a = SWZ_128B; p.swz_b = SWZ_128B; p.version = 1; p.a_mn_major = 1; p.N = 128; p.ksteps = 4;
    p.a_step_bytes = 1024; p.b_step_bytes = 32;
    run(swap ? "V3 MN/K sw128 (A lbo<->sbo swapped)" : "V3 MN/K sw128", p, a, b, 32);
  }
  // V4: as V3 with N = 256 and the full K = 128 of the head-chain kernel (k-chunks of 32 are separate 16 KB stages)
  {
    // A: 4 stages of [4 mn-atoms][32 k rows][128 B]; stage s holds k in [32 s, 32 s + 32)
    a.assign(65536, 0);
    for (int m = 0; m < 128; ++m)
      for (int k = 0; k < 128; ++k) {
        const int s = k / 32, kk = k % 32;
        const size_t off = (size_t)s * 16384 + (size_t)(m / 32) * 4096 + (size_t)(kk / 8) * 1024 + (kk % 8) * 128 +
                           ((((m % 32) / 4) ^ (kk % 8)) * 16) + (m % 4) * 4;
        const float v = aval(m, k);
        memcpy(&a[off], &v, 4);
      }
    img_kmajor_sw128(b, 256, 128, bval);  // [4 k-chunks][256 rows][128 B]
    // 16 k-steps do not advance uniformly (stage stride 16384 after every 4 steps): run the 4 stages as 4 launches'
    // worth of descriptors is not possible with one stride, so check the first stage (K = 32) and the N = 256 path
    Params p{};
    p.a_lbo = 4096; p.a_sbo = 1024; p.b_lbo = 16; p.b_sbo = 1024; p.swz_a = SWZ_128B; p.swz_b = SWZ_128B; p.version = 1;
    p.a_mn_major = 1; p.N = 256; p.ksteps = 4; p.a_step_bytes = 1024; p.b_step_bytes = 32;
    run("V4 MN/K sw128 N=256 (K=32)", p, a, b, 32);
  }
  // V6: A MN-major with descriptor layout type 1 (SWIZZLE_128B_BASE32B), mn-atom stride 4096, k-atom (4 rows) stride 512
  for (int swap = 0; swap < 2; ++swap) {
    a.clear();
    img_mnmajor_sw128_32b(a, 128, 32, 4096, 512, aval);
    img_kmajor_sw128(b, 128, 32, bval);
    Params p{};
    p.a_lbo = swap ? 512 : 4096; p.a_sbo = swap ? 4096 : 512; p.b_lbo = 16; p.b_sbo = 1024;
    p.swz_a = 1; p.swz_b = SWZ_128B; p.version = 1; p.a_mn_major = 1; p.N = 128; p.ksteps = 4;
    p.a_step_bytes = 1024; p.b_step_bytes = 32;
    run(swap ? "V6 MN(128B_BASE32B)/K (swapped)" : "V6 MN(128B_BASE32B)/K", p, a, b, 32);
  }
  // V7: A MN-major without swizzle (core matrix 8 k x 4 m): m-group stride 128, k-group stride 4096, K = 8
  for (int swap = 0; swap < 2; ++swap) {
    a.clear();
    img_mnmajor_none(a, 128, 8, 128, 4096, aval);
    img_kmajor_sw128(b, 128, 32, bval);
    Params p{};
    p.a_lbo = swap ? 128 : 4096; p.a_sbo = swap ? 4096 : 128; p.b_lbo = 16; p.b_sbo = 1024;
    p.swz_a = SWZ_NONE; p.swz_b = SWZ_128B; p.version = 1; p.a_mn_major = 1; p.N = 128; p.ksteps = 1;
    run(swap ? "V7 MN(none)/K K=8 (swapped)" : "V7 MN(none)/K K=8", p, a, b, 8);
  }
  // V8: B MN-major (SW128 16B-atom and 32B-atom), A K-major sw128: is it only the A side?
  for (int mode = 0; mode < 2; ++mode) {
    img_kmajor_sw128(a, 128, 32, aval);
    b.clear();
    if (mode == 0) img_mnmajor_sw128(b, 128, 32, 4096, 1024, bval);
    else img_mnmajor_sw128_32b(b, 128, 32, 4096, 512, bval);
    Params p{};
    p.a_lbo = 16; p.a_sbo = 1024; p.b_lbo = 4096; p.b_sbo = mode ? 512 : 1024;
    p.swz_a = SWZ_128B; p.swz_b = mode ? 1 : SWZ_128B; p.version = 1; p.b_mn_major = 1; p.N = 128; p.ksteps = 4;
    p.a_step_bytes = 32; p.b_step_bytes = 1024;
    run(mode ? "V8 K/MN(128B_BASE32B)" : "V8 K/MN(sw128)", p, a, b, 32);
  }
  // V5: TMA SWIZZLE_128B dump of a 32 x 32 box: does the landed image match  chunk ^= (row % 8) ?
  {
    const int rows = 64, cols = 256;
    std::vector<float> h((size_t)rows * cols);
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols; ++c) h[(size_t)r * cols + c] = (float)(r * 1000 + c);
    float *dm, *dout;
    cudaMalloc(&dm, h.size() * 4);
    cudaMalloc(&dout, 4096);
    cudaMemcpy(dm, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap map;
    const uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows};
    const uint64_t strides[1] = {(uint64_t)cols * 4};
    const uint32_t box[2] = {32, 32};
    if (!encode_f32(&map, dm, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) {
      printf("V5 TMA: tensor map encode FAILED\n");
    } else {
      cudaFuncSetAttribute(tma_dump_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
      tma_dump_kernel<<<1, 128, 16384>>>(map, 64, 8, dout);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<float> o(1024);
      cudaMemcpy(o.data(), dout, 4096, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int r = 0; r < 32; ++r)
        for (int c = 0; c < 32; ++c) {
          const size_t off = (size_t)r * 32 + (((c / 4) ^ (r % 8)) * 4) + (c % 4);
          if (o[off] != (float)((8 + r) * 1000 + 64 + c)) bad++;
        }
      printf("V5 TMA sw128 box dump: %s, mismatches vs (chunk ^= row%%8) image: %d / 1024; o[0..3] %g %g %g %g o[32..35] %g %g %g %g\n",
             cudaGetErrorString(e), bad, o[0], o[1], o[2], o[3], o[32], o[33], o[34], o[35]);
    }
    CUtensorMap map2;
    if (!encode_f32(&map2, dm, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) {
      printf("V5b TMA ATOM_32B: tensor map encode FAILED\n");
    } else {
      tma_dump_kernel<<<1, 128, 16384>>>(map2, 64, 8, dout);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<float> o(1024);
      cudaMemcpy(o.data(), dout, 4096, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int r = 0; r < 32; ++r)
        for (int c = 0; c < 32; ++c) {
          const size_t off = (size_t)r * 32 + (((c / 8) ^ (r % 4)) * 8) + (c % 8);
          if (o[off] != (float)((8 + r) * 1000 + 64 + c)) bad++;
        }
      printf("V5b TMA 128B_ATOM_32B dump: %s, mismatches vs (chunk32 ^= row%%4) image: %d / 1024; o[32..47]", cudaGetErrorString(e), bad);
      for (int i = 32; i < 48; ++i) printf(" %g", o[i]);
      printf("\n");
    }
  }
  return 0;
}
