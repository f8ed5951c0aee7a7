// K1c -- 3x3 / stride 1 / pad 1 convolution on channels_last (NHWC) activations as a TMA-fed tcgen05 implicit GEMM
// (fp32 accumulation in TMEM) with a fused per-channel affine + LeakyReLU epilogue.  Two operand precisions:
//   PREC_X3   (default of the model): activations and filter travel as split-bf16 PAIRS (hi, mid planes, tc_common.cuh);
//             every K step issues the three products hi*hi + mid*hi + hi*mid (kind::f16) -- fp32-grade results (~2^-17
//             per product) at 1.5x the MMA time of a TF32 pass.  This is what keeps `pred` within 1e-3 of the fp32
//             reference on every pixel.  Output: fp32 NHWC or a split-bf16 pair for the next consumer.
//   PREC_TF32: fp32 operands read as TF32 (single pass; operands must be pre-rounded, RNA, by their producers).
//
// Reference: mViT.conv3x3 (models/miniViT.py:16,27: Conv2d(128,128,3,padding=1), 16.7 GFLOP/img = 75 % of the head's
// FLOPs) and the DecoderBN blocks (models/unet_adaptive_bins.py:39-49: Conv2d 3x3 -> BatchNorm2d -> LeakyReLU, twice per
// up-sampling step; eval-mode BatchNorm is the per-channel affine of the epilogue), and conv3 (:73).
//
// GEMM view: M = pixels, N = C_out, K = 9 * C.  One CTA works on a "super tile" of NT vertically stacked 128-pixel
// patches (TH x TW pixels each, TH * TW = 128) and an N tile of <= 256 output channels.  A pipeline stage is one
// (32-channel chunk, dx) pair:
//   * A: ONE TMA box {32 ch, TW px, NT*TH + 2 rows, 1 img (, 2 planes)} at (c0, x0 + dx - 1, y0 - 1, b): the x shift is
//     the TMA coordinate, the zero padding is the TMA's out-of-bounds fill, and the three dy taps of every stacked patch
//     are the SAME shared-memory bytes addressed by UMMA descriptors whose start address moves by dy * TW rows (a whole
//     number of swizzle atoms, so the swizzle phase is unchanged).  That cuts the activation traffic from 9 to
//     3 * (NT*TH + 2) / (NT*TH) tile reads per chunk.
//   * B: one box {32 ch, N tile, 3 (dy), 1 (dx) (, 2 planes)} of the filter pre-laid-out as [dx][dy][C_out][C].
//   Shared-memory rows are 128 B (32 fp32, SWIZZLE_128B) for TF32 and 64 B (32 bf16, SWIZZLE_64B) for the pairs, so a
//   stage holds the same number of bytes in both precisions.
// Accumulators: NT x N-tile fp32 columns per buffer, two buffers in TMEM, so the epilogue of one super tile overlaps the
// MMAs of the next.  Epilogue: tcgen05.ld -> affine / LeakyReLU -> swizzled shared staging -> TMA tensor store (edge
// tiles are clipped by the hardware; no predicates anywhere).  Persistent, 1 CTA / SM, warp-specialised: warp 0 TMA
// producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 3-6 epilogue.
// PREC_BF16 (model.precision = "bf16"): the hi planes only, one bf16 product per K step; the TMA boxes carry one plane.
// Measured (B200, B = 16): head conv 128 -> 128 at 208 x 272, x3 form: 0.516 ms = 1.55 PFLOP/s of issued bf16 products = 0.96 of
// the measured bf16 peak (the TF32 form of the same kernel: 0.327 ms).  The decoder convs reach 0.57-0.87: the 26 x 34 / 52 x 68
// maps pad to whole 128-pixel tiles (69 % / 79 % useful), C_out = 80 runs an N = 80 MMA at the shared-memory operand rate, and
// operand rows must start on 64-byte boundaries (callers pad the channel pitch to 32: a 680- or 344-channel pitch put every other
// row at a 16-byte offset and cost 25-30 % of the kernel).
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace mde {
namespace tc {

constexpr int CV_THREADS = 224;
constexpr int CV_KC = 32;                // channels per K chunk
constexpr int CV_STG_BYTES = 128 * 128;  // one staged output group: 128 pixels x 32 channels (fp32, or 2 bf16 planes)
enum { PREC_TF32 = 0, PREC_X3 = 1, PREC_BF16 = 2 };  // BF16: the hi planes only, one bf16 product per K step (2e-2 mode)
enum { OUT_F32 = 0, OUT_TF32 = 1, OUT_PAIR = 2 };  // epilogue output: fp32, fp32 rounded to TF32, split-bf16 pair

struct ConvGeom {
  int B, H, W, C, Cout;
  int tiles_x, tiles_y, tiles_n;  // super-tile grid
  int total;                      // tiles_x * tiles_y * tiles_n * B
  int chunks;                     // ceil(C / 32)
  int n_tile;                     // output channels per CTA tile (multiple of 16, <= 256)
  int nstages;
  int tmem_cols;
  float slope;                    // LeakyReLU slope (1.0f = identity)
  int out_mode;
};

template <int NT, int TW, int PREC>
__global__ void __launch_bounds__(CV_THREADS, 1)
    conv3x3_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                   const __grid_constant__ CUtensorMap map_y, const float* __restrict__ scale,
                   const float* __restrict__ shift, const ConvGeom g) {
  constexpr bool X3 = PREC == PREC_X3;
  constexpr bool BF = PREC != PREC_TF32;       // bf16 operands (pair planes); BF && !X3: plane 0 (hi) only
  constexpr int TH = 128 / TW;
  constexpr int SR = NT * TH;                  // pixel rows per super tile
  constexpr int ROWB = BF ? 64 : 128;          // bytes of one shared-memory operand row (32 channels)
  constexpr int PLANES = X3 ? 2 : 1;
  constexpr int A_PLANE = (SR + 2) * TW * ROWB;  // halo box of one stage, one plane
  constexpr int A_BYTES = PLANES * A_PLANE;
  constexpr uint32_t SWZ = BF ? SWZ_64B : SWZ_128B;
  constexpr uint32_t SBO = 8 * ROWB;           // 8-row swizzle atom
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* gbase = smem_dyn + (base - smem_u32(smem_dyn));
  const int nb = g.n_tile;                     // filter rows per stage
  const int b_plane = 3 * nb * ROWB;
  const int b_bytes = PLANES * b_plane;
  const int stage_bytes = A_BYTES + b_bytes;
  const uint32_t s_stg = (base + g.nstages * stage_bytes + 1023u) & ~1023u;  // two staging buffers (1024-aligned: SWIZZLE_128B)
  const uint32_t s_bar = s_stg + 2 * CV_STG_BYTES;
  const uint32_t bar_full = s_bar, bar_empty = s_bar + 8 * g.nstages, bar_acc_full = s_bar + 16 * g.nstages,
                 bar_acc_empty = bar_acc_full + 16;
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(gbase + (s_bar - base) + 16 * g.nstages + 32);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit0 = blockIdx.x, ustep = gridDim.x;

  if (threadIdx.x == 0) {
    for (int i = 0; i < g.nstages; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_acc_full + 8 * i, 1);
      mbar_init(bar_acc_empty + 8 * i, 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
    fence_proxy_async();
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
    tma_prefetch_desc(&map_y);
  }
  if (warp == 2) {
    tmem_alloc(smem_u32((const void*)tmem_slot), (uint32_t)g.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();  // the prologue above overlaps the previous kernel of the stream (common.cuh); global memory from here on

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = unit0; tile < g.total; tile += ustep) {
        int r = tile;
        const int ni = r % g.tiles_n; r /= g.tiles_n;
        const int txi = r % g.tiles_x; r /= g.tiles_x;
        const int tyi = r % g.tiles_y;
        const int b = r / g.tiles_y;
        const int x0 = txi * TW, y0 = tyi * SR, n0 = ni * g.n_tile;
        for (int c = 0; c < g.chunks; ++c) {
          for (int dx = 0; dx < 3; ++dx) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1, 21);
            mbar_expect_tx(bar_full + 8 * stage, (uint32_t)stage_bytes);
            const uint32_t dst = base + stage * stage_bytes;
            if constexpr (BF) {  // the planes of the pair this precision reads in one box (outermost box dimension)
              tma_load_5d(dst, &map_x, bar_full + 8 * stage, c * CV_KC, x0 + dx - 1, y0 - 1, b, 0);
              tma_load_5d(dst + A_BYTES, &map_w, bar_full + 8 * stage, c * CV_KC, n0, 0, dx, 0);
            } else {
              tma_load_4d(dst, &map_x, bar_full + 8 * stage, c * CV_KC, x0 + dx - 1, y0 - 1, b);
              tma_load_4d(dst + A_BYTES, &map_w, bar_full + 8 * stage, c * CV_KC, n0, 0, dx);
            }
            if (++stage == (uint32_t)g.nstages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BF ? FMT_BF16 : FMT_TF32, 128, (uint32_t)g.n_tile, 0, 0);
      uint32_t stage = 0, phase = 0, it = 0;
      for (int tile = unit0; tile < g.total; tile += ustep, ++it) {
        const uint32_t buf = it & 1, aphase = (it >> 1) & 1;
        mbar_wait(bar_acc_empty + 8 * buf, aphase ^ 1, 22);
        tc_fence_after();
        const uint32_t d0 = tmem_base + buf * NT * g.n_tile;
        uint32_t first = 1;
        for (int c = 0; c < g.chunks; ++c) {
          for (int dx = 0; dx < 3; ++dx) {
            mbar_wait(bar_full + 8 * stage, phase, 23);
            tc_fence_after();
            const uint32_t a0 = base + stage * stage_bytes;
            const uint32_t b0 = a0 + A_BYTES;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              if constexpr (X3) {
#pragma unroll
                for (int j = 0; j < CV_KC / 16; ++j) {  // K = 16 bf16 = 32 B per instruction
                  const uint32_t boff = b0 + dy * nb * ROWB + j * 32;
                  const uint64_t bhi = make_smem_desc(boff, 16, SBO, SWZ), bmid = make_smem_desc(boff + b_plane, 16, SBO, SWZ);
#pragma unroll
                  for (int t = 0; t < NT; ++t) {
                    const uint32_t aoff = a0 + (t * TH + dy) * TW * ROWB + j * 32;
                    const uint64_t ahi = make_smem_desc(aoff, 16, SBO, SWZ), amid = make_smem_desc(aoff + A_PLANE, 16, SBO, SWZ);
                    umma_f16_ss(d0 + t * g.n_tile, ahi, bhi, idesc, first ^ 1);
                    umma_f16_ss(d0 + t * g.n_tile, amid, bhi, idesc, 1);
                    umma_f16_ss(d0 + t * g.n_tile, ahi, bmid, idesc, 1);
                  }
                  first = 0;
                }
              } else if constexpr (BF) {
#pragma unroll
                for (int j = 0; j < CV_KC / 16; ++j) {
                  const uint64_t bdesc = make_smem_desc(b0 + dy * nb * ROWB + j * 32, 16, SBO, SWZ);
#pragma unroll
                  for (int t = 0; t < NT; ++t) {
                    const uint64_t adesc = make_smem_desc(a0 + (t * TH + dy) * TW * ROWB + j * 32, 16, SBO, SWZ);
                    umma_f16_ss(d0 + t * g.n_tile, adesc, bdesc, idesc, first ^ 1);
                  }
                  first = 0;
                }
              } else {
#pragma unroll
                for (int j = 0; j < CV_KC / 8; ++j) {
                  const uint64_t bdesc = make_smem_desc(b0 + dy * nb * ROWB + j * 32, 16, SBO, SWZ);
#pragma unroll
                  for (int t = 0; t < NT; ++t) {
                    const uint64_t adesc = make_smem_desc(a0 + (t * TH + dy) * TW * ROWB + j * 32, 16, SBO, SWZ);
                    umma_tf32_ss(d0 + t * g.n_tile, adesc, bdesc, idesc, first ^ 1);
                  }
                  first = 0;
                }
              }
            }
            umma_commit(bar_empty + 8 * stage);
            if (++stage == (uint32_t)g.nstages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        umma_commit(bar_acc_full + 8 * buf);
      }
    }
  } else if (warp >= 3) {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;  // accumulator row = pixel (ty, tx) = (row / TW, row % TW) of the patch
    const int etid = threadIdx.x - 96;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int groups = (g.n_tile + 31) / 32;
    uint32_t it = 0, nstore = 0;
    for (int tile = unit0; tile < g.total; tile += ustep, ++it) {
      int r = tile;
      const int ni = r % g.tiles_n; r /= g.tiles_n;
      const int txi = r % g.tiles_x; r /= g.tiles_x;
      const int tyi = r % g.tiles_y;
      const int b = r / g.tiles_y;
      const int x0 = txi * TW, y0 = tyi * SR, n0 = ni * g.n_tile;
      const uint32_t buf = it & 1, aphase = (it >> 1) & 1;
      mbar_wait(bar_acc_full + 8 * buf, aphase, 24);
      tc_fence_after();
#pragma unroll 1
      for (int t = 0; t < NT; ++t) {
#pragma unroll 1
        for (int cg = 0; cg < groups; ++cg) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (buf * NT + t) * g.n_tile + cg * 32 + lane_addr, v);
          const int ch0 = n0 + cg * 32;
          // the group's 32 scale / shift values as 16-byte loads issued before the accumulator arrives (C_out % 4 == 0, ch0 % 32
          // == 0): the element loop below is then pure register arithmetic the compiler interleaves freely
          float sc[32], sh[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const bool in = ch0 + 4 * q < g.Cout;
            const float4 a = (scale != nullptr && in) ? __ldg(reinterpret_cast<const float4*>(scale + ch0) + q)
                                                      : make_float4(1.f, 1.f, 1.f, 1.f);
            const float4 c = (shift != nullptr && in) ? __ldg(reinterpret_cast<const float4*>(shift + ch0) + q)
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
            sc[4 * q] = a.x; sc[4 * q + 1] = a.y; sc[4 * q + 2] = a.z; sc[4 * q + 3] = a.w;
            sh[4 * q] = c.x; sh[4 * q + 1] = c.y; sh[4 * q + 2] = c.z; sh[4 * q + 3] = c.w;
          }
          tmem_ld_wait();
          float o[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float y = fmaf(__uint_as_float(v[i]), sc[i], sh[i]);
            o[i] = y > 0.0f ? y : y * g.slope;
          }
          if (g.out_mode == OUT_TF32) {
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = tf32_round(o[i]);
          }
          const uint32_t sbuf = s_stg + (nstore & 1) * CV_STG_BYTES;
          if (etid == 0) tma_store_wait_read<1>();  // the store that last read this staging buffer has drained
          named_barrier(1, 128);
          if (g.out_mode == OUT_PAIR) {
            // two bf16 planes of [128 px][32 ch] = 64-byte rows, SWIZZLE_64B: 16-byte chunk q of row r sits at q ^ ((r >> 1) & 3)
            const uint32_t rowaddr = sbuf + row * 64;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint32_t hi[4], mid[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) split_bf16x2(o[8 * q + 2 * e], o[8 * q + 2 * e + 1], hi[e], mid[e]);
              const uint32_t addr = rowaddr + ((q ^ ((row >> 1) & 3)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3])
                           : "memory");
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + CV_STG_BYTES / 2), "r"(mid[0]), "r"(mid[1]),
                           "r"(mid[2]), "r"(mid[3])
                           : "memory");
            }
          } else {
            const uint32_t rowaddr = sbuf + row * 128;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const uint32_t addr = rowaddr + ((q ^ (row & 7)) << 4);
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(o[4 * q]), "f"(o[4 * q + 1]),
                           "f"(o[4 * q + 2]), "f"(o[4 * q + 3])
                           : "memory");
            }
          }
          fence_proxy_async();
          named_barrier(1, 128);
          if (etid == 0) {
            if (g.out_mode == OUT_PAIR) tma_store_5d(&map_y, sbuf, ch0, x0, y0 + t * TH, b, 0);
            else tma_store_4d(&map_y, sbuf, ch0, x0, y0 + t * TH, b);
            tma_store_commit();
          }
          ++nstore;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty + 8 * buf);  // hand the accumulator buffer back to the MMA issuer
    }
    if (etid == 0) tma_store_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols);
  }
}

// filter [Cout, C, 3, 3] (torch layout) -> [dx][dy][Cout][C], TF32-rounded (RNA)
__global__ void __launch_bounds__(256) conv3x3_prep_weight_kernel(const float* __restrict__ w, float* __restrict__ out,
                                                                  uint16_t* __restrict__ out_pair, int Cout, int C,
                                                                  float scale) {
  const long long n = (long long)Cout * C * 9;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % C);
  long long r = i / C;
  const int co = (int)(r % Cout);
  r /= Cout;
  const int dy = (int)(r % 3), dx = (int)(r / 3);
  const float v = w[(((long long)co * C + c) * 3 + dy) * 3 + dx] * scale;
  if (out_pair) {
    uint16_t hi, mid;
    split_bf16(v, hi, mid);
    out_pair[i] = hi;
    out_pair[n + i] = mid;
  } else {
    out[i] = tf32_round(v);
  }
}

// Direct fp32 convolution for a handful of output channels (the noAdaBins decoder's conv3: C -> 1,
// models/unet_adaptive_bins.py:78-80): ONE THREAD per output pixel, 16-byte loads along the channel axis of the nine taps
// (neighbouring pixels' taps overlap, so L1 serves 8 of 9 reads), the 9*C*Cout filter in shared memory (every lane reads the
// same filter word: a broadcast).  Exact fp32; ~0.7 kFLOP per pixel, bound by L1 load issue.
template <int COUT_MAX>
__global__ void __launch_bounds__(256) conv3x3_small_kernel(const float* __restrict__ x, const float* __restrict__ w_oihw,
                                                            const float* __restrict__ bias, float* __restrict__ y, int B,
                                                            int H, int W, int C, int Cout) {
  extern __shared__ __align__(16) float sw[];  // [Cout][9][C]
  for (int i = threadIdx.x; i < Cout * 9 * C; i += blockDim.x) {
    const int c = i % C, tap = (i / C) % 9, co = i / (9 * C);
    sw[i] = w_oihw[((long long)co * C + c) * 9 + tap];
  }
  __syncthreads();
  const long long npix = (long long)B * H * W;
  const int c4 = C >> 2;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(p % W), yy = (int)((p / W) % H);
    const long long b = p / ((long long)W * H);
    float acc[COUT_MAX];
#pragma unroll
    for (int co = 0; co < COUT_MAX; ++co) acc[co] = 0.f;
    for (int tap = 0; tap < 9; ++tap) {
      const int sy = yy + tap / 3 - 1, sx = xx + tap % 3 - 1;
      if (sy < 0 || sy >= H || sx < 0 || sx >= W) continue;
      const float4* src = reinterpret_cast<const float4*>(x + ((b * H + sy) * W + sx) * C);
#pragma unroll 4
      for (int c = 0; c < c4; ++c) {
        const float4 v = __ldg(src + c);
#pragma unroll
        for (int co = 0; co < COUT_MAX; ++co) {
          if (co < Cout) {
            const float4 wv = *reinterpret_cast<const float4*>(sw + (co * 9 + tap) * C + 4 * c);
            acc[co] = fmaf(v.x, wv.x, fmaf(v.y, wv.y, fmaf(v.z, wv.z, fmaf(v.w, wv.w, acc[co]))));
          }
        }
      }
    }
#pragma unroll
    for (int co = 0; co < COUT_MAX; ++co)
      if (co < Cout) y[p * Cout + co] = acc[co] + (bias ? bias[co] : 0.f);
  }
}

}  // namespace tc
}  // namespace mde

using namespace mde;

namespace {

// shared host side of the two precisions: tiling, tensor maps, launch
template <int PREC>
int conv3x3_launch(const void* x, const void* w_prep, const float* scale, const float* shift, void* y, int B, int H, int W,
                   int C, int Cout, float lrelu_slope, int out_mode, cudaStream_t st) {
  constexpr bool X3 = PREC == tc::PREC_X3;
  constexpr bool BF = PREC != tc::PREC_TF32;
  const int in_align = BF ? 8 : 4;                                   // TMA: 16-byte global strides
  const int out_align = out_mode == tc::OUT_PAIR ? 8 : 4;
  if (C % in_align != 0 || Cout % out_align != 0 || !aligned(x, 16) || !aligned(w_prep, 16) || !aligned(y, 16) ||
      (scale && !aligned(scale, 16)) || (shift && !aligned(shift, 16)))
    return MDE_ERR_UNSUPPORTED;
  // N tile: the whole C_out when it fits one instruction (<= 256, multiple of 16), else the largest divisor of C_out
  // that is a multiple of 32 (so that 32-channel store groups never straddle two N tiles) -- the first candidate, in that
  // order, whose stage (halo box + filter taps) fits the shared-memory ring at least twice.  Measured on B200: choosing
  // narrower tiles so that two stacked patches share the filter (C_out = 640 as 5 x 128, 320 as 5 x 64) was slower
  // (up1 + up2: 1.54 vs 1.37 ms).
  const int rowb = BF ? 64 : 128, planes = X3 ? 2 : 1;
  const int budget = 222 * 1024 - 2 * tc::CV_STG_BYTES - 2048 - 512;
  int forced = 0;
  {
    const char* force = getenv("MDE_CONV_NTILE");  // tuning aid
    if (force && atoi(force) > 0 && Cout % atoi(force) == 0 && atoi(force) % 32 == 0 && atoi(force) <= 256) forced = atoi(force);
  }
  int n_tile = 0, nt = 0, tw = 0, stage_bytes = 0;
  for (int pass = 0; pass < 11 && n_tile == 0; ++pass) {
    int cand;
    if (forced) {
      if (pass > 0) break;
      cand = forced;
    } else if (pass == 0) {
      if (!(Cout <= 256 && Cout % 16 == 0)) continue;
      cand = Cout;
    } else if (pass <= 8) {
      cand = 256 - 32 * (pass - 1);
      if (Cout % cand != 0 || cand == Cout) continue;
    } else {
      // no exact tiling (e.g. the 1392 / 680 / 344 input widths of the decoder, which are the OUTPUT widths of the input-
      // gradient convolutions): ragged last tile -- filter rows and output channels beyond C_out are the TMA's zero fill /
      // store clipping
      cand = pass == 9 ? 128 : 64;
    }
    const int cnt = (4 * cand <= 512) ? 2 : 1;
    if (2 * cnt * cand + 16 > 512 && (cand % 32) != 0) continue;
    // patch shape: 8 rows x 16 px or 16 rows x 8 px, whichever wastes fewer padded pixels
    auto padded = [&](int w_) {
      const int th_ = 128 / w_, sr_ = cnt * th_;
      return (long long)((W + w_ - 1) / w_) * w_ * ((H + sr_ - 1) / sr_) * sr_;
    };
    const int ctw = padded(16) <= padded(8) ? 16 : 8;
    const int csr = cnt * (128 / ctw);
    const int sb = planes * (csr + 2) * ctw * rowb + planes * 3 * cand * rowb;
    if (budget / sb < 2) continue;
    n_tile = cand; nt = cnt; tw = ctw; stage_bytes = sb;
  }
  if (n_tile == 0) return MDE_ERR_UNSUPPORTED;
  const int th = 128 / tw, sr = nt * th;

  tc::ConvGeom g;
  g.B = B; g.H = H; g.W = W; g.C = C; g.Cout = Cout;
  g.tiles_x = (W + tw - 1) / tw;
  g.tiles_y = (H + sr - 1) / sr;
  g.tiles_n = (Cout + n_tile - 1) / n_tile;
  g.total = g.tiles_x * g.tiles_y * g.tiles_n * B;
  g.chunks = (C + tc::CV_KC - 1) / tc::CV_KC;
  g.n_tile = n_tile;
  g.slope = lrelu_slope;
  g.out_mode = out_mode;
  int cols = 2 * nt * n_tile + ((n_tile % 32) ? 16 : 0);
  g.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
  g.nstages = budget / stage_bytes;
  if (g.nstages > 6) g.nstages = 6;
  const int smem = g.nstages * stage_bytes + 2 * tc::CV_STG_BYTES + 32 * g.nstages + 128 + 2048;

  CUtensorMap mx, mw, my;
  const uint64_t es = BF ? 2 : 4;  // operand element size
  if (BF) {
    const uint32_t pl = X3 ? 2 : 1;  // planes per box: the hi plane alone in the single-product mode
    const uint64_t xplane = (uint64_t)B * H * W * C * es, wplane = (uint64_t)9 * Cout * C * es;
    const uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B, 2};
    const uint64_t strides[4] = {(uint64_t)C * es, (uint64_t)W * C * es, (uint64_t)H * W * C * es, xplane};
    const uint32_t box[5] = {(uint32_t)tc::CV_KC, (uint32_t)tw, (uint32_t)(sr + 2), 1, pl};
    if (!tc::encode_bf16(&mx, x, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B)) return MDE_ERR_DRIVER;
    const uint64_t wdims[5] = {(uint64_t)C, (uint64_t)Cout, 3, 3, 2};
    const uint64_t wstrides[4] = {(uint64_t)C * es, (uint64_t)Cout * C * es, (uint64_t)3 * Cout * C * es, wplane};
    const uint32_t wbox[5] = {(uint32_t)tc::CV_KC, (uint32_t)n_tile, 3, 1, pl};
    if (!tc::encode_bf16(&mw, w_prep, 5, wdims, wstrides, wbox, CU_TENSOR_MAP_SWIZZLE_64B)) return MDE_ERR_DRIVER;
  } else {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)C * 4, (uint64_t)W * C * 4, (uint64_t)H * W * C * 4};
    const uint32_t box[4] = {(uint32_t)tc::CV_KC, (uint32_t)tw, (uint32_t)(sr + 2), 1};
    if (!tc::encode_f32(&mx, x, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return MDE_ERR_DRIVER;
    const uint64_t wdims[4] = {(uint64_t)C, (uint64_t)Cout, 3, 3};
    const uint64_t wstrides[3] = {(uint64_t)C * 4, (uint64_t)Cout * C * 4, (uint64_t)3 * Cout * C * 4};
    const uint32_t wbox[4] = {(uint32_t)tc::CV_KC, (uint32_t)n_tile, 3, 1};
    if (!tc::encode_f32(&mw, w_prep, 4, wdims, wstrides, wbox, CU_TENSOR_MAP_SWIZZLE_128B)) return MDE_ERR_DRIVER;
  }
  if (out_mode == tc::OUT_PAIR) {
    const uint64_t dims[5] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)B, 2};
    const uint64_t strides[4] = {(uint64_t)Cout * 2, (uint64_t)W * Cout * 2, (uint64_t)H * W * Cout * 2,
                                 (uint64_t)B * H * W * Cout * 2};
    const uint32_t box[5] = {32, (uint32_t)tw, (uint32_t)th, 1, 2};
    if (!tc::encode_bf16(&my, y, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B)) return MDE_ERR_DRIVER;
  } else {
    const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)Cout * 4, (uint64_t)W * Cout * 4, (uint64_t)H * W * Cout * 4};
    const uint32_t box[4] = {32, (uint32_t)tw, (uint32_t)th, 1};
    if (!tc::encode_f32(&my, y, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return MDE_ERR_DRIVER;
  }
  const int grid = g.total < MDE_NUM_SMS ? g.total : MDE_NUM_SMS;
#define MDE_CV_LAUNCH(NT, TW)                                                                                          \
  {                                                                                                                    \
    static bool attr = false;                                                                                          \
    if (!attr) {                                                                                                       \
      if (cudaFuncSetAttribute(tc::conv3x3_kernel<NT, TW, PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                               226 * 1024) != cudaSuccess)                                                             \
        return MDE_ERR_LAUNCH;                                                                                         \
      attr = true;                                                                                                     \
    }                                                                                                                  \
    launch_pdl(PDL_TC, tc::conv3x3_kernel<NT, TW, PREC>, dim3(grid), dim3(tc::CV_THREADS), smem, st, mx, mw, my, scale, shift, g); \
  }
  switch (nt * 100 + tw) {
    case 216: MDE_CV_LAUNCH(2, 16) break;
    case 208: MDE_CV_LAUNCH(2, 8) break;
    case 116: MDE_CV_LAUNCH(1, 16) break;
    default: MDE_CV_LAUNCH(1, 8) break;
  }
#undef MDE_CV_LAUNCH
  return check_launch();
}

}  // namespace

extern "C" {

int mde_conv3x3_prep_weight(const float* w_oihw, float* w_prep, int Cout, int C, float operand_scale,
                            mde_stream_t stream) {
  if (!w_oihw || !w_prep) return MDE_ERR_BAD_POINTER;
  if (Cout <= 0 || C <= 0) return MDE_ERR_BAD_SHAPE;
  const long long n = (long long)Cout * C * 9;
  tc::conv3x3_prep_weight_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w_oihw, w_prep, nullptr, Cout,
                                                                                                C, operand_scale);
  return check_launch();
}

int mde_conv3x3_prep_weight_x3(const float* w_oihw, uint16_t* w_pair, int Cout, int C, mde_stream_t stream) {
  if (!w_oihw || !w_pair) return MDE_ERR_BAD_POINTER;
  if (Cout <= 0 || C <= 0) return MDE_ERR_BAD_SHAPE;
  const long long n = (long long)Cout * C * 9;
  tc::conv3x3_prep_weight_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w_oihw, nullptr, w_pair, Cout,
                                                                                                C, 1.0f);
  return check_launch();
}

int mde_conv3x3_nhwc_fwd(const float* x_nhwc, const float* w_prep, const float* scale, const float* shift, float* y_nhwc,
                         int B, int H, int W, int C, int Cout, float lrelu_slope, int round_tf32, mde_stream_t stream) {
  if (!x_nhwc || !w_prep || !y_nhwc) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || Cout <= 0) return MDE_ERR_BAD_SHAPE;
  return conv3x3_launch<tc::PREC_TF32>(x_nhwc, w_prep, scale, shift, y_nhwc, B, H, W, C, Cout, lrelu_slope,
                                       round_tf32 ? tc::OUT_TF32 : tc::OUT_F32, (cudaStream_t)stream);
}

int mde_conv3x3_nhwc_x3_fwd(const uint16_t* x_pair, const uint16_t* w_pair, const float* scale, const float* shift, void* y,
                            int y_is_pair, int B, int H, int W, int C, int Cout, float lrelu_slope, int products,
                            mde_stream_t stream) {
  if (!x_pair || !w_pair || !y) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || Cout <= 0 || (products != 3 && products != 1)) return MDE_ERR_BAD_SHAPE;
  if (products == 1)
    return conv3x3_launch<tc::PREC_BF16>(x_pair, w_pair, scale, shift, y, B, H, W, C, Cout, lrelu_slope,
                                         y_is_pair ? tc::OUT_PAIR : tc::OUT_F32, (cudaStream_t)stream);
  return conv3x3_launch<tc::PREC_X3>(x_pair, w_pair, scale, shift, y, B, H, W, C, Cout, lrelu_slope,
                                     y_is_pair ? tc::OUT_PAIR : tc::OUT_F32, (cudaStream_t)stream);
}

int mde_conv3x3_small_nhwc_fwd(const float* x_nhwc, const float* w_oihw, const float* bias, float* y_nhwc, int B, int H, int W,
                               int C, int Cout, mde_stream_t stream) {
  if (!x_nhwc || !w_oihw || !y_nhwc) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || Cout <= 0) return MDE_ERR_BAD_SHAPE;
  if (Cout > 4 || C % 4 != 0 || !aligned(x_nhwc, 16) || (size_t)Cout * 9 * C * sizeof(float) > 200 * 1024) return MDE_ERR_UNSUPPORTED;
  const size_t sm = (size_t)Cout * 9 * C * sizeof(float);
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(tc::conv3x3_small_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
      return MDE_ERR_LAUNCH;
    attr = true;
  }
  const long long npix = (long long)B * H * W;
  long long grid = (npix + 255) / 256;
  if (grid > MDE_NUM_SMS * 8) grid = MDE_NUM_SMS * 8;
  tc::conv3x3_small_kernel<4><<<(unsigned)grid, 256, sm, (cudaStream_t)stream>>>(x_nhwc, w_oihw, bias, y_nhwc, B, H, W, C, Cout);
  return check_launch();
}

}  // extern "C"
