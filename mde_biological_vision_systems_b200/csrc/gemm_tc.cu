// Batched "NT" GEMM on tcgen05 (TF32 inputs, fp32 accumulation in TMEM), TMA-fed, optional split-K:
//     C[b][m][n] (+)= sum_k A[b][m][k] * B[b][n][k]
// Both operands are K-major (k contiguous), which is how every matrix of the head's backward pass lies in memory:
//   * d feat^T[k, p]  = sum_j W'^T[k, j] * gl[p, j]      (M = 128 channels, N = pixels, K = 256 bins)
//   * d W'[j, k]      = sum_p gl^T[j, p] * feat^T[k, p]  (M = 256 bins, N = 128 channels, K = pixels: split-K)
// i.e. the two GEMMs behind the reference's autograd of PixelWiseDotProduct + conv_out (models/layers.py:31-36,
// models/unet_adaptive_bins.py:286); also usable for any nn.Linear-shaped product whose inputs tolerate TF32.
// One CTA = one 128 x TN output tile of one batch entry and one K split: warp 0 TMA producer (A box {32, 128}, B box
// {32, TN}, SWIZZLE_128B, out-of-range rows / k zero-filled by the TMA), warp 1 MMA issuer, warp 2 TMEM allocator,
// warps 3-6 epilogue (tcgen05.ld -> plain store, or atomicAdd when the K range is split).
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace mde {
namespace tc {

constexpr int GM_THREADS = 224;
constexpr int GM_KC = 32;

struct GemmGeom {
  int M, N;
  long long ldc, c_batch;  // C row pitch / batch stride in floats
  int tn;                  // N tile (multiple of 16, <= 256)
  int n_tiles;
  int chunks_per_split;    // K chunks (of 32) per split
  int total_chunks;
  int nstages;
  int tmem_cols;
  int atomic;              // accumulate with atomicAdd (split-K)
  float alpha;
  const float* bias;       // [N] added once (by K split 0), may be null
  int act;                 // 0 none, 1 ReLU (not with split-K)
  int split_out;           // 1: write rows as [v | v - trunc_tf32(v) | v] (row pitch >= 3N): the A operand of a 3xTF32 GEMM
  long long plane_stride;  // > 0: K split s writes its partial product to C + s * plane_stride (no atomics; the consumer sums)
  int tap_wp;              // > 0: the "batch" index is a 3x3 filter tap (ky, kx) = (z / 3, z % 3) of a conv3x3 weight gradient:
                           // A is the kx-th of three pre-shifted copies of one matrix, B one matrix read at k + (ky - 1) * tap_wp
                           // (tap_wp % 4 == 0: the TMA needs 16-byte aligned box starts along the contiguous axis)
};

__device__ __forceinline__ float tf32_trunc(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

__global__ void __launch_bounds__(GM_THREADS, 1)
    gemm_nt_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   float* __restrict__ C, const GemmGeom g) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* gbase = smem_dyn + (base - smem_u32(smem_dyn));
  constexpr int A_BYTES = 128 * 128;
  const int b_bytes = g.tn * 128;
  const int stage_bytes = A_BYTES + b_bytes;
  const uint32_t s_bar = base + g.nstages * stage_bytes;
  const uint32_t bar_full = s_bar, bar_empty = s_bar + 8 * g.nstages, bar_acc = s_bar + 16 * g.nstages;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + (s_bar - base) + 16 * g.nstages + 16);
  float* stg_all = reinterpret_cast<float*>(gbase + (s_bar - base) + 16 * g.nstages + 64);  // [4 warps][32][33] transpose staging
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x / g.n_tiles, nt = blockIdx.x - mt * g.n_tiles;
  const int split = blockIdx.y, batch = blockIdx.z;
  const int m0 = mt * 128, n0 = nt * g.tn;
  const int c_begin = split * g.chunks_per_split;
  int c_end = c_begin + g.chunks_per_split;
  if (c_end > g.total_chunks) c_end = g.total_chunks;
  const int nchunks = c_end - c_begin;  // >= 1 by construction of the grid

  if (threadIdx.x == 0) {
    for (int i = 0; i < g.nstages; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    mbar_init(bar_acc, 1);
    fence_barrier_init();
    fence_proxy_async();
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 2) {
    tmem_alloc(smem_u32((const void*)tmem_slot), (uint32_t)g.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();  // the prologue above overlaps the previous kernel; operands and C are touched from here on

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const int za = g.tap_wp > 0 ? batch % 3 : batch, zb = g.tap_wp > 0 ? 0 : batch;
      const int kb_off = g.tap_wp > 0 ? (batch / 3 - 1) * g.tap_wp : 0;  // out-of-range k: TMA zero fill
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1, 31);
        mbar_expect_tx(bar_full + 8 * stage, (uint32_t)stage_bytes);
        const uint32_t dst = base + stage * stage_bytes;
        tma_load_3d(dst, &map_a, bar_full + 8 * stage, (c_begin + c) * GM_KC, m0, za);
        tma_load_3d(dst + A_BYTES, &map_b, bar_full + 8 * stage, (c_begin + c) * GM_KC + kb_off, n0, zb);
        if (++stage == (uint32_t)g.nstages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(FMT_TF32, 128, (uint32_t)g.tn, 0, 0);
      uint32_t stage = 0, phase = 0;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(bar_full + 8 * stage, phase, 32);
        tc_fence_after();
        const uint32_t a0 = base + stage * stage_bytes, b0 = a0 + A_BYTES;
#pragma unroll
        for (int j = 0; j < GM_KC / 8; ++j) {
          const uint64_t adesc = make_smem_desc(a0 + j * 32, 16, 1024, SWZ_128B);
          const uint64_t bdesc = make_smem_desc(b0 + j * 32, 16, 1024, SWZ_128B);
          umma_tf32_ss(tmem_base, adesc, bdesc, idesc, (c | j) != 0);
        }
        umma_commit(bar_empty + 8 * stage);
        if (++stage == (uint32_t)g.nstages) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(bar_acc);
    }
  } else if (warp >= 3) {
    const int quarter = warp & 3;
    const int m = m0 + quarter * 32 + lane;
    mbar_wait(bar_acc, 0, 33);
    tc_fence_after();
    C += (long long)split * g.plane_stride;
    float* crow = C + (long long)batch * g.c_batch + (long long)m * g.ldc + n0;
    const bool row_ok = m < g.M;
#pragma unroll 1
    for (int c0 = 0; c0 < g.tn; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + c0 + ((uint32_t)(quarter * 32) << 16), r);
      tmem_ld_wait();
      const int nvalid = min(32, min(g.tn - c0, g.N - n0 - c0));
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float t = g.alpha * __uint_as_float(r[i]);
        if (g.bias != nullptr && split == 0 && i < nvalid) t += __ldg(g.bias + n0 + c0 + i);
        v[i] = g.act == 1 ? fmaxf(t, 0.f) : t;
      }
      if (g.atomic) {
        if (row_ok)
          for (int i = 0; i < nvalid; ++i) atomicAdd(crow + c0 + i, v[i]);
        continue;
      }
      // transpose the warp's 32 rows x 32 columns through shared memory so that 8 lanes write one row's 128 bytes:
      // every store instruction covers 4 whole lines instead of 32 partial ones
      float* stg = stg_all + quarter * (32 * 33);
#pragma unroll
      for (int i = 0; i < 32; ++i) stg[lane * 33 + i] = v[i];
      __syncwarp();
      const int cc = (lane & 7) * 4;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int rr = it * 4 + (lane >> 3);
        const int mm = m0 + quarter * 32 + rr;
        if (mm >= g.M) continue;
        const float* sp = stg + rr * 33 + cc;
        const float4 o = make_float4(sp[0], sp[1], sp[2], sp[3]);
        float* dst = C + (long long)batch * g.c_batch + (long long)mm * g.ldc + n0 + c0 + cc;
        if (cc + 3 < nvalid && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && (!g.split_out || (g.N & 3) == 0)) {
          *reinterpret_cast<float4*>(dst) = o;
          if (g.split_out) {
            *reinterpret_cast<float4*>(dst + g.N) =
                make_float4(o.x - tf32_trunc(o.x), o.y - tf32_trunc(o.y), o.z - tf32_trunc(o.z), o.w - tf32_trunc(o.w));
            *reinterpret_cast<float4*>(dst + 2 * g.N) = o;
          }
        } else {
          const float ov[4] = {o.x, o.y, o.z, o.w};
          for (int q = 0; q < 4; ++q) {
            if (cc + q >= nvalid) break;
            dst[q] = ov[q];
            if (g.split_out) {
              dst[g.N + q] = ov[q] - tf32_trunc(ov[q]);
              dst[2 * g.N + q] = ov[q];
            }
          }
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols);
  }
}

}  // namespace tc
}  // namespace mde

using namespace mde;

extern "C" {

int mde_gemm_nt_tf32(const float* A, int64_t lda, int64_t a_batch, const float* B, int64_t ldb, int64_t b_batch, float* C,
                     int64_t ldc, int64_t c_batch, int batch, int M, int N, int K, int splits, float alpha,
                     mde_stream_t stream) {
  return mde_gemm_nt_tf32_ex(A, lda, a_batch, B, ldb, b_batch, C, ldc, c_batch, batch, M, N, K, splits, alpha, nullptr, 0, 0,
                             stream);
}

int mde_gemm_nt_tf32_ex(const float* A, int64_t lda, int64_t a_batch, const float* B, int64_t ldb, int64_t b_batch, float* C,
                        int64_t ldc, int64_t c_batch, int batch, int M, int N, int K, int splits, float alpha,
                        const float* bias, int act, int split_out, mde_stream_t stream) {
  return mde_gemm_nt_tf32_planes(A, lda, a_batch, B, ldb, b_batch, C, ldc, c_batch, batch, M, N, K, splits, alpha, bias, act,
                                 split_out, 0, stream);
}

static int gemm_nt_launch(const float* A, int64_t lda, int64_t a_batch, const float* B, int64_t ldb, int64_t b_batch,
                          float* C, int64_t ldc, int64_t c_batch, int batch, int M, int N, int K, int splits, float alpha,
                          const float* bias, int act, int split_out, int64_t plane_stride, int tap_wp, mde_stream_t stream);

int mde_gemm_nt_tf32_planes(const float* A, int64_t lda, int64_t a_batch, const float* B, int64_t ldb, int64_t b_batch,
                            float* C, int64_t ldc, int64_t c_batch, int batch, int M, int N, int K, int splits, float alpha,
                            const float* bias, int act, int split_out, int64_t plane_stride, mde_stream_t stream) {
  return gemm_nt_launch(A, lda, a_batch, B, ldb, b_batch, C, ldc, c_batch, batch, M, N, K, splits, alpha, bias, act, split_out,
                        plane_stride, 0, stream);
}

// Weight gradient of a 3x3 / stride 1 / pad 1 convolution as ONE launch of the NT GEMM with the nine filter taps on the grid's
// z axis:  dW9[ky*3+kx][co][ci] = sum_k dyT3[kx][co][k] * xT[ci][k + (ky-1)*Wp],  k running over the zero-PADDED pixel axis
// (b, y+1, x+1) of pitch Wp (>= W+2, a multiple of 4).  mde_nhwc_to_cpad_tf32 builds both operands (TF32-rounded): xT as is,
// dyT3 as three copies shifted by kx-1 along k (dyT3[kx][co][k] = dy_pad[co][k - (kx-1)]), so that the horizontal tap is
// baked into the data and the vertical tap is a 16-byte aligned shift of the TMA's K coordinate -- the TMA cannot start a
// box at an unaligned element of the contiguous axis.  Shifts that leave the matrix are the TMA's zero fill.
int mde_conv3x3_wgrad_tf32(const float* dyT3, const float* xT, float* dW9, int Cout, int Cin, int64_t Kp, int64_t ld, int Wp,
                           int splits, mde_stream_t stream) {
  if (!dyT3 || !xT || !dW9) return MDE_ERR_BAD_POINTER;
  if (Cout <= 0 || Cin <= 0 || Kp <= 0 || Kp > 0x7fffffffLL || ld < Kp || Wp < 4 || Wp % 4 != 0) return MDE_ERR_BAD_SHAPE;
  if (splits > 1) cudaMemsetAsync(dW9, 0, sizeof(float) * 9 * (size_t)Cout * Cin, (cudaStream_t)stream);
  return gemm_nt_launch(dyT3, ld, (int64_t)Cout * ld, xT, ld, 0, dW9, Cin, (int64_t)Cout * Cin, 9, Cout, Cin, (int)Kp, splits,
                        1.0f, nullptr, 0, 0, 0, Wp, stream);
}

static int gemm_nt_launch(const float* A, int64_t lda, int64_t a_batch, const float* B, int64_t ldb, int64_t b_batch,
                          float* C, int64_t ldc, int64_t c_batch, int batch, int M, int N, int K, int splits, float alpha,
                          const float* bias, int act, int split_out, int64_t plane_stride, int tap_wp, mde_stream_t stream) {
  if (!A || !B || !C) return MDE_ERR_BAD_POINTER;
  if (batch <= 0 || batch > 65535 || M <= 0 || N <= 0 || K <= 0 || splits <= 0 || splits > 65535) return MDE_ERR_BAD_SHAPE;
  if (act < 0 || act > 1 || (splits > 1 && (act != 0 || split_out)) || (split_out && ldc < 3 * (int64_t)N)) return MDE_ERR_BAD_SHAPE;
  if (lda % 4 != 0 || ldb % 4 != 0 || a_batch % 4 != 0 || b_batch % 4 != 0 || !aligned(A, 16) || !aligned(B, 16))
    return MDE_ERR_UNSUPPORTED;  // TMA: 16-byte global strides
  tc::GemmGeom g;
  g.M = M;
  g.N = N;
  g.ldc = ldc;
  g.c_batch = c_batch;
  g.total_chunks = (K + tc::GM_KC - 1) / tc::GM_KC;
  if (splits > g.total_chunks) splits = g.total_chunks;
  // N tile (measured on B200 with scripts/time_gemms.py): 256-wide tiles when they still give >= 2 waves of CTAs (fewer
  // re-reads of the other operand), otherwise 128 -- narrower tiles did not pay (fixed cost per CTA ~4.5 us)
  {
    const long long m_tiles_ = (M + 127) / 128;
    const int n_cap = N >= 256 ? 256 : ((N + 15) / 16) * 16;
    int tn = n_cap;
    if (n_cap == 256 && m_tiles_ * ((N + 255) / 256) * splits * batch < 2 * MDE_NUM_SMS) tn = 128;
    const char* force = getenv("MDE_GEMM_TN");
    if (force && atoi(force) >= 16 && atoi(force) <= n_cap) tn = atoi(force) / 16 * 16;
    g.tn = tn;
  }
  g.n_tiles = (N + g.tn - 1) / g.tn;
  g.chunks_per_split = (g.total_chunks + splits - 1) / splits;
  splits = (g.total_chunks + g.chunks_per_split - 1) / g.chunks_per_split;  // no empty split
  g.plane_stride = plane_stride > 0 ? plane_stride : 0;
  g.atomic = (splits > 1 && plane_stride <= 0) ? 1 : 0;
  g.alpha = alpha;
  g.bias = bias;
  g.act = act;
  g.split_out = split_out;
  g.tap_wp = tap_wp;
  g.tmem_cols = g.tn <= 32 ? 32 : g.tn <= 64 ? 64 : g.tn <= 128 ? 128 : 256;
  const int stage_bytes = 128 * 128 + g.tn * 128;
  g.nstages = (184 * 1024) / stage_bytes;
  if (g.nstages > 8) g.nstages = 8;
  if (g.nstages > g.chunks_per_split) g.nstages = g.chunks_per_split;
  const int smem = g.nstages * stage_bytes + 16 * g.nstages + 64 + 4 * 32 * 33 * 4 + 1024;
  const int m_tiles = (M + 127) / 128;
  if ((long long)m_tiles * g.n_tiles > 0x7fffffffLL) return MDE_ERR_BAD_SHAPE;

  CUtensorMap ma, mb;
  {
    const uint64_t dims[3] = {(uint64_t)K, (uint64_t)M, (uint64_t)(tap_wp > 0 ? 3 : batch)};
    const uint64_t strides[2] = {(uint64_t)lda * 4, (uint64_t)(batch > 1 ? a_batch : lda * M) * 4};
    const uint32_t box[3] = {(uint32_t)tc::GM_KC, 128, 1};
    if (!tc::encode_f32(&ma, A, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return MDE_ERR_DRIVER;
  }
  {
    const uint64_t dims[3] = {(uint64_t)K, (uint64_t)N, (uint64_t)(tap_wp > 0 ? 1 : batch)};
    const uint64_t strides[2] = {(uint64_t)ldb * 4, (uint64_t)((batch > 1 && tap_wp == 0) ? b_batch : ldb * N) * 4};
    const uint32_t box[3] = {(uint32_t)tc::GM_KC, (uint32_t)g.tn, 1};
    if (!tc::encode_f32(&mb, B, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return MDE_ERR_DRIVER;
  }
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(tc::gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess)
      return MDE_ERR_LAUNCH;
    attr = true;
  }
  dim3 grid((unsigned)(m_tiles * g.n_tiles), (unsigned)splits, (unsigned)batch);
  if (splits > 1 && g.atomic) {  // preceded by the caller's memset of C: an ordinary launch
    tc::gemm_nt_kernel<<<grid, tc::GM_THREADS, smem, (cudaStream_t)stream>>>(ma, mb, C, g);
  } else {
    launch_pdl(PDL_CHAIN, tc::gemm_nt_kernel, grid, dim3(tc::GM_THREADS), smem, (cudaStream_t)stream, ma, mb, C, g);
  }
  return check_launch();
}

}  // extern "C"
