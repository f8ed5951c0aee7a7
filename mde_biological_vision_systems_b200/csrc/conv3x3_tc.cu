// K1c -- 3x3 / stride 1 / pad 1 convolution on channels_last (NHWC) fp32 activations as a TMA-fed tcgen05 implicit
// GEMM (TF32 inputs, fp32 accumulation in TMEM) with a fused per-channel affine + LeakyReLU (+ TF32 rounding) epilogue.
//
// Reference: mViT.conv3x3 (models/miniViT.py:16,27: Conv2d(128,128,3,padding=1), 16.7 GFLOP/img = 75 % of the head's
// FLOPs) and the DecoderBN blocks (models/unet_adaptive_bins.py:39-49: Conv2d 3x3 -> BatchNorm2d -> LeakyReLU, twice per
// up-sampling step; eval-mode BatchNorm is the per-channel affine of the epilogue), and conv3 (:73).
//
// GEMM view: M = pixels, N = C_out, K = 9 * C.  One CTA works on a "super tile" of NT vertically stacked 128-pixel
// patches (TH x TW pixels each, TH * TW = 128) and an N tile of <= 256 output channels.  A pipeline stage is one
// (32-channel chunk, dx) pair:
//   * A: ONE 4-D TMA box {32 ch, TW px, NT*TH + 2 rows, 1 img} at (c0, x0 + dx - 1, y0 - 1, b): the x shift is the TMA
//     coordinate, the zero padding is the TMA's out-of-bounds fill, and the three dy taps of every stacked patch are
//     the SAME shared-memory bytes addressed by UMMA descriptors whose start address moves by dy * TW rows (TW * 128 B
//     is a whole number of 1024-byte swizzle atoms, so the 128B-swizzle phase is unchanged).  That cuts the activation
//     traffic from 9 to 3 * (NT*TH + 2) / (NT*TH) tile reads per chunk -- operand delivery from L2 (~42 B/clk/SM), not
//     the tensor pipe, is what bounds an fp32-operand implicit GEMM.
//   * B: one box {32 ch, N tile, 3 (dy), 1 (dx)} of the filter pre-laid-out as [dx][dy][C_out][C] (TF32-rounded, scaled
//     by MDE_TF32_TRUNC_COMP): three K-major 128B-swizzled operand tiles shared by all NT patches.
// Measured (B200, head conv 128 -> 128 at B = 16, 208 x 272): 0.327 ms = 816 TFLOP/s = 72 % of the 1.125 PFLOP/s TF32
// nominal peak, tensor pipe 80 % active under ncu (cuDNN: 0.73 ms).  Two attempts to go further were measured and did
// not pay: a variant with 3-4 stacked patches, separate A/B rings and single-buffered accumulators (fewer operand bytes
// from L2): 0.36-0.42 ms; CTA pairs (cta_group::2, half the filter rows per CTA, kept below as an opt-in): 0.338 ms.
// Accumulators: NT x N-tile fp32 columns per buffer, two buffers in TMEM, so the epilogue of one super tile overlaps the
// MMAs of the next.  Epilogue: tcgen05.ld -> affine / LeakyReLU / optional TF32 rounding -> swizzled shared staging ->
// TMA tensor store (edge tiles are clipped by the hardware; no predicates anywhere).  Persistent, 1 CTA / SM,
// warp-specialised: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 3-6 epilogue.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace mde {
namespace tc {

constexpr int CV_THREADS = 224;
constexpr int CV_KC = 32;          // channels per K chunk (one 128-byte swizzle row)
constexpr int CV_STG_BYTES = 128 * 128;  // one staged output group: 128 pixels x 32 channels

struct ConvGeom {
  int B, H, W, C, Cout;
  int tiles_x, tiles_y, tiles_n;  // super-tile grid
  int total;                      // tiles_x * tiles_y * tiles_n * B
  int chunks;                     // ceil(C / 32)
  int n_tile;                     // output channels per CTA tile (multiple of 16, <= 256)
  int nstages;
  int tmem_cols;
  float slope;                    // LeakyReLU slope (1.0f = identity)
  int round_tf32;
  int pair_x, pair_y;             // CTA-pair mode: the two CTAs of a pair take adjacent super tiles along x (2,1) or y (1,2)
};

// CTAS = 2: a CTA pair (cluster of two, tcgen05 cta_group::2) works on two adjacent super tiles with ONE M = 256 MMA stream
// issued by the leader; each CTA stages its own halo box and only HALF of the filter rows (n_tile / 2 output channels), so
// the shared-memory operand reads per MMA drop from 8 KB to 6 KB per SM at N = 128 (the measured ceiling of the 1-CTA
// form) and the filter traffic from L2 halves.  tiles_x / tiles_y then count PAIRS along the paired axis.
template <int NT, int TW, int CTAS = 1>
__global__ void __launch_bounds__(CV_THREADS, 1)
    conv3x3_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                   const __grid_constant__ CUtensorMap map_y, const float* __restrict__ scale,
                   const float* __restrict__ shift, const ConvGeom g) {
  constexpr int TH = 128 / TW;
  constexpr int SR = NT * TH;                  // pixel rows per super tile
  constexpr int A_BYTES = (SR + 2) * TW * 128; // halo box of one stage
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* gbase = smem_dyn + (base - smem_u32(smem_dyn));
  const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;  // 0 = leader (issues the MMAs of the pair)
  const int nbh = g.n_tile / CTAS;                            // filter rows staged by this CTA
  const int b_bytes = 3 * nbh * 128;
  const int stage_bytes = A_BYTES + b_bytes;
  const uint32_t s_stg = base + g.nstages * stage_bytes;  // two staging buffers
  const uint32_t s_bar = s_stg + 2 * CV_STG_BYTES;
  const uint32_t bar_full = s_bar, bar_empty = s_bar + 8 * g.nstages, bar_acc_full = s_bar + 16 * g.nstages,
                 bar_acc_empty = bar_acc_full + 16;
  const uint32_t bar_peerfull = bar_acc_empty + 16 + 16;  // [nstages] leader only: the peer's stage has landed
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(gbase + (s_bar - base) + 16 * g.nstages + 32);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit0 = blockIdx.x / CTAS, ustep = gridDim.x / CTAS;  // work units = super tiles or super-tile pairs
  const int rx = (CTAS == 2 && g.pair_x == 2) ? (int)rank : 0, ry = (CTAS == 2 && g.pair_y == 2) ? (int)rank : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < g.nstages; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_acc_full + 8 * i, 1);
      mbar_init(bar_acc_empty + 8 * i, 4 * CTAS);  // one arrive per epilogue warp (of both CTAs of a pair)
    }
    if (CTAS == 2)
      for (int i = 0; i < g.nstages; ++i) mbar_init(bar_peerfull + 8 * i, 1);
    fence_barrier_init();
    fence_proxy_async();
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
    tma_prefetch_desc(&map_y);
  }
  if (warp == 2) {
    if (CTAS == 2) {
      tmem_alloc2(smem_u32((const void*)tmem_slot), (uint32_t)g.tmem_cols);
      tmem_relinquish2();
    } else {
      tmem_alloc(smem_u32((const void*)tmem_slot), (uint32_t)g.tmem_cols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all();  // barrier inits visible to the peer before any remote arrive / multicast commit
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = unit0; tile < g.total; tile += ustep) {
        int r = tile;
        const int ni = r % g.tiles_n; r /= g.tiles_n;
        const int txi = r % g.tiles_x; r /= g.tiles_x;
        const int tyi = r % g.tiles_y;
        const int b = r / g.tiles_y;
        const int x0 = (txi * g.pair_x + rx) * TW, y0 = (tyi * g.pair_y + ry) * SR, n0 = ni * g.n_tile;
        for (int c = 0; c < g.chunks; ++c) {
          for (int dx = 0; dx < 3; ++dx) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1, 21);
            mbar_expect_tx(bar_full + 8 * stage, (uint32_t)stage_bytes);
            const uint32_t dst = base + stage * stage_bytes;
            tma_load_4d(dst, &map_x, bar_full + 8 * stage, c * CV_KC, x0 + dx - 1, y0 - 1, b);
            tma_load_4d(dst + A_BYTES, &map_w, bar_full + 8 * stage, c * CV_KC, n0 + (int)rank * nbh, 0, dx);
            if (++stage == (uint32_t)g.nstages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1 && CTAS == 2 && rank == 1) {
    // peer CTA: relay "my operands of this stage have landed" to the leader
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = unit0; tile < g.total; tile += ustep) {
        for (int c = 0; c < 3 * g.chunks; ++c) {
          mbar_wait(bar_full + 8 * stage, phase, 28);
          mbar_arrive_remote(bar_peerfull + 8 * stage, 0);
          if (++stage == (uint32_t)g.nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(FMT_TF32, 128 * CTAS, (uint32_t)g.n_tile, 0, 0);
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
            if (CTAS == 2) mbar_wait(bar_peerfull + 8 * stage, phase, 29);
            tc_fence_after();
            const uint32_t a0 = base + stage * stage_bytes;
            const uint32_t b0 = a0 + A_BYTES;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
              for (int j = 0; j < CV_KC / 8; ++j) {
                const uint64_t bdesc = make_smem_desc(b0 + dy * nbh * 128 + j * 32, 16, 1024, SWZ_128B);
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                  const uint64_t adesc = make_smem_desc(a0 + (t * TH + dy) * TW * 128 + j * 32, 16, 1024, SWZ_128B);
                  if (CTAS == 2) umma2_tf32_ss(d0 + t * g.n_tile, adesc, bdesc, idesc, first ^ 1);
                  else umma_tf32_ss(d0 + t * g.n_tile, adesc, bdesc, idesc, first ^ 1);
                }
                first = 0;
              }
            }
            if (CTAS == 2) umma2_commit_mc(bar_empty + 8 * stage);
            else umma_commit(bar_empty + 8 * stage);
            if (++stage == (uint32_t)g.nstages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        if (CTAS == 2) umma2_commit_mc(bar_acc_full + 8 * buf);
        else umma_commit(bar_acc_full + 8 * buf);
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
      const int x0 = (txi * g.pair_x + rx) * TW, y0 = (tyi * g.pair_y + ry) * SR, n0 = ni * g.n_tile;
      const uint32_t buf = it & 1, aphase = (it >> 1) & 1;
      mbar_wait(bar_acc_full + 8 * buf, aphase, 24);
      tc_fence_after();
#pragma unroll 1
      for (int t = 0; t < NT; ++t) {
#pragma unroll 1
        for (int cg = 0; cg < groups; ++cg) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (buf * NT + t) * g.n_tile + cg * 32 + lane_addr, v);
          tmem_ld_wait();
          const int ch0 = n0 + cg * 32;
          float o[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int ch = ch0 + i;
            const bool in = ch < g.Cout;
            const float sc = (scale != nullptr && in) ? __ldg(scale + ch) : 1.0f;
            const float sh = (shift != nullptr && in) ? __ldg(shift + ch) : 0.0f;
            float y = fmaf(__uint_as_float(v[i]), sc, sh);
            y = y > 0.0f ? y : y * g.slope;
            o[i] = g.round_tf32 ? tf32_round(y) : y;
          }
          const uint32_t sbuf = s_stg + (nstore & 1) * CV_STG_BYTES;
          if (etid == 0) tma_store_wait_read<1>();  // the store that last read this staging buffer has drained
          named_barrier(1, 128);
          const uint32_t rowaddr = sbuf + row * 128;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const uint32_t addr = rowaddr + ((q ^ (row & 7)) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(o[4 * q]), "f"(o[4 * q + 1]),
                         "f"(o[4 * q + 2]), "f"(o[4 * q + 3])
                         : "memory");
          }
          fence_proxy_async();
          named_barrier(1, 128);
          if (etid == 0) {
            tma_store_4d(&map_y, sbuf, ch0, x0, y0 + t * TH, b);
            tma_store_commit();
          }
          ++nstore;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {  // hand the accumulator buffer back to the MMA issuer (the leader's barrier; the peer arrives remotely)
        if (CTAS == 2 && rank != 0) mbar_arrive_remote(bar_acc_empty + 8 * buf, 0);
        else mbar_arrive(bar_acc_empty + 8 * buf);
      }
    }
    if (etid == 0) tma_store_wait<0>();
  }
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all();  // neither CTA may exit while the other can still signal its barriers
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (CTAS == 2) tmem_dealloc2(tmem_base, (uint32_t)g.tmem_cols);
    else tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols);
  }
}

// filter [Cout, C, 3, 3] (torch layout) -> [dx][dy][Cout][C], TF32-rounded after scaling
__global__ void __launch_bounds__(256) conv3x3_prep_weight_kernel(const float* __restrict__ w, float* __restrict__ out,
                                                                  int Cout, int C, float scale) {
  const long long n = (long long)Cout * C * 9;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % C);
  long long r = i / C;
  const int co = (int)(r % Cout);
  r /= Cout;
  const int dy = (int)(r % 3), dx = (int)(r / 3);
  out[i] = tf32_round(w[(((long long)co * C + c) * 3 + dy) * 3 + dx] * scale);
}

}  // namespace tc
}  // namespace mde

using namespace mde;

extern "C" {

int mde_conv3x3_prep_weight(const float* w_oihw, float* w_prep, int Cout, int C, float operand_scale,
                            mde_stream_t stream) {
  if (!w_oihw || !w_prep) return MDE_ERR_BAD_POINTER;
  if (Cout <= 0 || C <= 0) return MDE_ERR_BAD_SHAPE;
  const long long n = (long long)Cout * C * 9;
  tc::conv3x3_prep_weight_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w_oihw, w_prep, Cout, C,
                                                                                                operand_scale);
  return check_launch();
}

int mde_conv3x3_nhwc_fwd(const float* x_nhwc, const float* w_prep, const float* scale, const float* shift, float* y_nhwc,
                         int B, int H, int W, int C, int Cout, float lrelu_slope, int round_tf32, mde_stream_t stream) {
  if (!x_nhwc || !w_prep || !y_nhwc) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || Cout <= 0) return MDE_ERR_BAD_SHAPE;
  if (C % 4 != 0 || Cout % 4 != 0 || !aligned(x_nhwc, 16) || !aligned(w_prep, 16) || !aligned(y_nhwc, 16))
    return MDE_ERR_UNSUPPORTED;  // TMA: 16-byte global strides
  // N tile: the whole C_out when it fits one instruction (<= 256, multiple of 16), else the largest divisor of C_out
  // that is a multiple of 32 (so that 32-channel store groups never straddle two N tiles).  Measured on B200: choosing
  // narrower tiles so that two stacked patches share the filter (C_out = 640 as 5 x 128, 320 as 5 x 64) was slower
  // (up1 + up2: 1.54 vs 1.37 ms) -- the MMA's shared-memory operand reads, not the L2 traffic, set the pace.
  int n_tile = 0;
  if (Cout <= 256 && Cout % 16 == 0) {
    n_tile = Cout;
  } else {
    for (int cand = 256; cand >= 32; cand -= 32)
      if (Cout % cand == 0) {
        n_tile = cand;
        break;
      }
  }
  {
    const char* force = getenv("MDE_CONV_NTILE");  // tuning aid
    if (force && atoi(force) > 0 && Cout % atoi(force) == 0 && atoi(force) % 32 == 0 && atoi(force) <= 256) n_tile = atoi(force);
  }
  if (n_tile == 0) return MDE_ERR_UNSUPPORTED;
  const int nt = (4 * n_tile <= 512) ? 2 : 1;
  if (2 * nt * n_tile + 16 > 512 && (n_tile % 32) != 0) return MDE_ERR_UNSUPPORTED;
  // patch shape: 8 rows x 16 px or 16 rows x 8 px, whichever wastes fewer padded pixels
  auto padded = [&](int tw) {
    const int th = 128 / tw, sr = nt * th;
    return (long long)((W + tw - 1) / tw) * tw * ((H + sr - 1) / sr) * sr;
  };
  const int tw = padded(16) <= padded(8) ? 16 : 8;
  const int th = 128 / tw, sr = nt * th;

  // CTA pairs (cta_group::2) are opt-in (MDE_CONV_CTAS=2): parity-green, but measured on B200 they do not pay -- head conv
  // 0.338 ms vs 0.343 ms, whole step 10.99 vs 11.06 ms -- i.e. neither the shared-memory operand reads nor the filter
  // traffic from L2 is what holds the single-CTA kernel at ~80 % tensor-pipe activity.  The two CTAs of a pair take
  // adjacent super tiles along the axis that wastes fewer out-of-image tiles.
  int ctas = 1;
  {
    const char* force = getenv("MDE_CONV_CTAS");
    if (force && atoi(force) == 2 && (n_tile / 2) % 8 == 0) ctas = 2;
  }
  tc::ConvGeom g;
  g.B = B; g.H = H; g.W = W; g.C = C; g.Cout = Cout;
  g.tiles_x = (W + tw - 1) / tw;
  g.tiles_y = (H + sr - 1) / sr;
  g.pair_x = g.pair_y = 1;
  if (ctas == 2) {
    const int px = (g.tiles_x + 1) / 2, py = (g.tiles_y + 1) / 2;
    if (2 * px * g.tiles_y <= g.tiles_x * 2 * py) {
      g.pair_x = 2;
      g.tiles_x = px;
    } else {
      g.pair_y = 2;
      g.tiles_y = py;
    }
  }
  g.tiles_n = Cout / n_tile;
  g.total = g.tiles_x * g.tiles_y * g.tiles_n * B;
  g.chunks = (C + tc::CV_KC - 1) / tc::CV_KC;
  g.n_tile = n_tile;
  g.slope = lrelu_slope;
  g.round_tf32 = round_tf32;
  int cols = 2 * nt * n_tile + ((n_tile % 32) ? 16 : 0);
  g.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
  const int a_bytes = (sr + 2) * tw * 128, stage_bytes = a_bytes + 3 * (n_tile / ctas) * 128;
  const int budget = 222 * 1024 - 2 * tc::CV_STG_BYTES - 1024 - 512;
  g.nstages = budget / stage_bytes;
  if (g.nstages > 6) g.nstages = 6;
  if (g.nstages < 2) return MDE_ERR_UNSUPPORTED;
  const int smem = g.nstages * stage_bytes + 2 * tc::CV_STG_BYTES + 32 * g.nstages + 128 + 1024;

  CUtensorMap mx, mw, my;
  {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)C * 4, (uint64_t)W * C * 4, (uint64_t)H * W * C * 4};
    const uint32_t box[4] = {(uint32_t)tc::CV_KC, (uint32_t)tw, (uint32_t)(sr + 2), 1};
    if (!tc::encode_f32(&mx, x_nhwc, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return MDE_ERR_DRIVER;
  }
  {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)Cout, 3, 3};
    const uint64_t strides[3] = {(uint64_t)C * 4, (uint64_t)Cout * C * 4, (uint64_t)3 * Cout * C * 4};
    const uint32_t box[4] = {(uint32_t)tc::CV_KC, (uint32_t)(n_tile / ctas), 3, 1};  // a CTA of a pair stages half the rows
    if (!tc::encode_f32(&mw, w_prep, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return MDE_ERR_DRIVER;
  }
  {
    const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)Cout * 4, (uint64_t)W * Cout * 4, (uint64_t)H * W * Cout * 4};
    const uint32_t box[4] = {32, (uint32_t)tw, (uint32_t)th, 1};
    if (!tc::encode_f32(&my, y_nhwc, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return MDE_ERR_DRIVER;
  }
  const int max_units = MDE_NUM_SMS / ctas;
  const int grid = ctas * (g.total < max_units ? g.total : max_units);
  cudaStream_t st = (cudaStream_t)stream;
#define MDE_CV_LAUNCH(NT, TW, CT)                                                                                      \
  {                                                                                                                    \
    static bool attr = false;                                                                                          \
    if (!attr) {                                                                                                       \
      if (cudaFuncSetAttribute(tc::conv3x3_kernel<NT, TW, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize,            \
                               226 * 1024) != cudaSuccess)                                                             \
        return MDE_ERR_LAUNCH;                                                                                         \
      attr = true;                                                                                                     \
    }                                                                                                                  \
    if (CT == 1) {                                                                                                     \
      tc::conv3x3_kernel<NT, TW, CT><<<grid, tc::CV_THREADS, smem, st>>>(mx, mw, my, scale, shift, g);                 \
    } else {                                                                                                           \
      cudaLaunchConfig_t cfg = {};                                                                                     \
      cfg.gridDim = dim3((unsigned)grid);                                                                              \
      cfg.blockDim = dim3(tc::CV_THREADS);                                                                             \
      cfg.dynamicSmemBytes = smem;                                                                                     \
      cfg.stream = st;                                                                                                 \
      cudaLaunchAttribute lattr[1];                                                                                    \
      lattr[0].id = cudaLaunchAttributeClusterDimension;                                                               \
      lattr[0].val.clusterDim.x = 2;                                                                                   \
      lattr[0].val.clusterDim.y = 1;                                                                                   \
      lattr[0].val.clusterDim.z = 1;                                                                                   \
      cfg.attrs = lattr;                                                                                               \
      cfg.numAttrs = 1;                                                                                                \
      if (cudaLaunchKernelEx(&cfg, tc::conv3x3_kernel<NT, TW, CT>, mx, mw, my, scale, shift, g) != cudaSuccess)        \
        return MDE_ERR_LAUNCH;                                                                                         \
    }                                                                                                                  \
  }
  switch (ctas * 1000 + nt * 100 + tw) {
    case 1216: MDE_CV_LAUNCH(2, 16, 1) break;
    case 1208: MDE_CV_LAUNCH(2, 8, 1) break;
    case 1116: MDE_CV_LAUNCH(1, 16, 1) break;
    case 1108: MDE_CV_LAUNCH(1, 8, 1) break;
    case 2216: MDE_CV_LAUNCH(2, 16, 2) break;
    case 2208: MDE_CV_LAUNCH(2, 8, 2) break;
    case 2116: MDE_CV_LAUNCH(1, 16, 2) break;
    default: MDE_CV_LAUNCH(1, 8, 2) break;
  }
#undef MDE_CV_LAUNCH
  return check_launch();
}

}  // extern "C"
