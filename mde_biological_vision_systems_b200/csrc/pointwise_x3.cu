// 1x1 convolution on channels_last activations as a tcgen05 GEMM with fp32-grade accuracy:
//     y[m][n] = act( sum_k x[m][k] * w[n][k] + bias[n] ) (+ residual[m][n]),   m = (b, y, x) pixels, k = C_in, n = C_out.
//
// Reference: DecoderBN.conv2 (models/unet_adaptive_bins.py:61, Conv2d(bottleneck, features, kernel_size=1, padding=1): the
// caller pads the 13x17 input spatially, border pixels then come out as the bias, exactly like the padded conv) and -- through
// the same operator -- the point-wise convolutions of the EfficientNet passthrough body (expansion / projection / head conv with
// their folded BatchNorm bias, SiLU and residual in the epilogue), which the library otherwise runs either in TF32 (outside
// the 1e-3 contract for sharply peaked softmaxes) or on its slow legacy fp32 kernels (7.7 ms of an 18 ms step at config 2).
//
// Operands: the weights travel as a split-bf16 pair (prepared once per parameter version); the activations arrive as plain
// fp32 from whatever produced them, are staged by TMA as raw fp32 and are SPLIT IN SHARED MEMORY by four converter warps,
// in place: thread r owns row r of the 128-row tile, reads its 64 floats of the K chunk (two 128-byte swizzled rows) into
// registers and writes the hi bf16 row over the first box and the mid bf16 row over the second -- the same 256 bytes, now two
// K-major SWIZZLE_128B operand tiles.  Every K step then issues the three products hi*hi + mid*hi + hi*mid (kind::f16).
// No extra pass over HBM, no pair tensors in the callers.
// Persistent, 1 CTA / SM: warp 0 TMA producer (raw activation boxes + the weight pair), warp 1 MMA issuer, warp 2 TMEM
// allocator, warps 3-6 converter, warps 7-10 epilogue (tcgen05.ld -> bias / SiLU -> shared-memory transpose -> coalesced
// float4 stores, residual added there); two TMEM accumulator buffers, so the epilogue of one tile overlaps the next tile's
// loads, conversion and MMAs.  Options: a per-image channel gate multiplied into x by the converter warps (squeeze-excite), a
// zero-bordered output map (TensorFlow-SAME padding of a following stride-2 convolution).
// Measured (B200, config 2, 46 launches per step): 2.16 ms.  The large-M / small-K layers are bound by the epilogue (ONE warp per
// scheduler: 64 MUFU per 32 columns for SiLU is 512 clocks on the quarter-rate pipe before anything else), the large-K layers by
// the converter (~1.2 us per 64-channel chunk, again one warp per scheduler) and by a 2-3 stage ring; see DESIGN.md section 9.
#include "common.cuh"
#include "tc_common.cuh"

namespace mde {
namespace tc {

constexpr int PW_THREADS = 352;              // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3-6 converter, 7-10 epilogue
constexpr int PW_KC = 64;                    // K elements per chunk
constexpr int PW_A_BYTES = 128 * PW_KC * 4;  // raw fp32 tile = afterwards [hi tile 16 KB][mid tile 16 KB]

struct PwGeom {
  long long M;
  int N, K;
  long long ldc, ldr;
  int tn, n_tiles, chunks, nstages, tmem_cols, act;
  long long total_tiles;
  const float* bias;
  const float* residual;
  const float* gate;            // optional [M / rows_per_image][K] per-image channel scale applied to x before the product
  long long rows_per_image;
  int out_W, out_H, out_Wp, out_Hp, out_pt, out_pl;  // out_W > 0: row m = (b, y, x) is written at (b, y + pt, x + pl) of a padded map
};

// Persistent: a CTA walks output tiles (m tile major, n tile minor) with ONE continuous operand pipeline; two TMEM
// accumulator buffers let the epilogue of tile i overlap the loads / conversion / MMAs of tile i+1.
__global__ void __launch_bounds__(PW_THREADS, 1)
    pointwise_x3_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                        float* __restrict__ Y, const PwGeom g) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* gbase = smem_dyn + (base - smem_u32(smem_dyn));
  const int b_plane = g.tn * 128;
  const int stage_bytes = PW_A_BYTES + 2 * b_plane;
  const uint32_t s_bar = base + g.nstages * stage_bytes;
  const uint32_t bar_raw = s_bar, bar_ops = s_bar + 8 * g.nstages, bar_empty = s_bar + 16 * g.nstages,
                 bar_acc_full = s_bar + 24 * g.nstages, bar_acc_empty = bar_acc_full + 16;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + (s_bar - base) + 24 * g.nstages + 32);
  float* stg_all = reinterpret_cast<float*>(gbase + (s_bar - base) + ((24 * g.nstages + 64 + 15) & ~15));  // [4 warps][32][36], 16-byte aligned
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < g.nstages; ++i) {
      mbar_init(bar_raw + 8 * i, 1);
      mbar_init(bar_ops + 8 * i, 4);  // one arrive per converter warp
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
      // The ring is shallow (2-3 stages of 64-96 KB), far less than the ~90 KB per SM that must be in flight to cover the DRAM
      // latency at full bandwidth.  The activation boxes of the next PF chunks are therefore prefetched into L2
      // (cp.async.bulk.prefetch.tensor: no shared memory, no barrier), so the ring's own loads only pay the L2 latency.
      constexpr int PF = 6;
      long long pf_tile = blockIdx.x;
      int pf_c = 0;
      auto prefetch_next = [&]() {
        if (pf_tile >= g.total_tiles) return;
        const int pm0 = (int)((pf_tile / g.n_tiles) * 128);
        tma_prefetch_l2_2d(&map_x, pf_c * PW_KC, pm0);
        if (g.K - pf_c * PW_KC > 32) tma_prefetch_l2_2d(&map_x, pf_c * PW_KC + 32, pm0);
        if (++pf_c == g.chunks) {
          pf_c = 0;
          pf_tile += gridDim.x;
        }
      };
      for (int i = 0; i < PF; ++i) prefetch_next();
      for (long long tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
        const long long mt = tile / g.n_tiles;
        const int n0 = (int)(tile - mt * g.n_tiles) * g.tn;
        const int m0 = (int)(mt * 128);
        for (int c = 0; c < g.chunks; ++c) {
          const bool two = g.K - c * PW_KC > 32;  // the chunk's second 32-float box holds data (else it is skipped altogether)
          prefetch_next();
          mbar_wait(bar_empty + 8 * stage, phase ^ 1, 41);
          mbar_expect_tx(bar_raw + 8 * stage, (uint32_t)(stage_bytes - (two ? 0 : 16384)));
          const uint32_t dst = base + stage * stage_bytes;
          tma_load_2d(dst, &map_x, bar_raw + 8 * stage, c * PW_KC, m0);               // k [0, 32) of the chunk
          if (two) tma_load_2d(dst + 16384, &map_x, bar_raw + 8 * stage, c * PW_KC + 32, m0);  // k [32, 64)
          tma_load_3d(dst + PW_A_BYTES, &map_w, bar_raw + 8 * stage, c * PW_KC, n0, 0);
          tma_load_3d(dst + PW_A_BYTES + b_plane, &map_w, bar_raw + 8 * stage, c * PW_KC, n0, 1);
          if (++stage == (uint32_t)g.nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(FMT_BF16, 128, (uint32_t)g.tn, 0, 0);
      uint32_t stage = 0, phase = 0, it = 0;
      for (long long tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++it) {
        const uint32_t buf = it & 1;
        mbar_wait(bar_acc_empty + 8 * buf, ((it >> 1) & 1) ^ 1, 46);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * g.tn;
        for (int c = 0; c < g.chunks; ++c) {
          mbar_wait(bar_raw + 8 * stage, phase, 42);  // the weight pair (written by the TMA) ...
          mbar_wait(bar_ops + 8 * stage, phase, 43);  // ... and the converted activation tiles
          tc_fence_after();
          const uint32_t a0 = base + stage * stage_bytes, b0 = a0 + PW_A_BYTES;
          const int kvalid = min(PW_KC, g.K - c * PW_KC);
          const int nj = (kvalid + 15) >> 4;  // K = 16 steps that hold data (the rest of the chunk is zero)
#pragma unroll 1
          for (int j = 0; j < nj; ++j) {
            const uint64_t ahi = make_smem_desc(a0 + j * 32, 16, 1024, SWZ_128B);
            const uint64_t amid = make_smem_desc(a0 + 16384 + j * 32, 16, 1024, SWZ_128B);
            const uint64_t bhi = make_smem_desc(b0 + j * 32, 16, 1024, SWZ_128B);
            const uint64_t bmid = make_smem_desc(b0 + b_plane + j * 32, 16, 1024, SWZ_128B);
            umma_f16_ss(d_tmem, ahi, bhi, idesc, (c | j) != 0);
            umma_f16_ss(d_tmem, amid, bhi, idesc, 1);
            umma_f16_ss(d_tmem, ahi, bmid, idesc, 1);
          }
          umma_commit(bar_empty + 8 * stage);
          if (++stage == (uint32_t)g.nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(bar_acc_full + 8 * buf);
      }
    }
  } else if (warp >= 3 && warp < 7) {
    // ---- converter: fp32 -> (hi, mid) bf16, in place; thread r owns row r of the 128-row tile ----------------------
    const int r = (warp - 3) * 32 + lane;
    uint32_t stage = 0, phase = 0;
    const uint32_t sw = (uint32_t)(r & 7);
    for (long long tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
      const float* grow = nullptr;  // this row's squeeze-excite gate (rows of one image share it)
      if (g.gate != nullptr) {
        long long m = (tile / g.n_tiles) * 128 + r;
        if (m >= g.M) m = g.M - 1;
        grow = g.gate + (m / g.rows_per_image) * g.K;
      }
      for (int c = 0; c < g.chunks; ++c) {
        float4 gv[16];
        if (grow != nullptr) {  // issued before the wait: the loads fly while the TMA box lands
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int k = c * PW_KC + 4 * q;
            gv[q] = k < g.K ? __ldg(reinterpret_cast<const float4*>(grow + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        mbar_wait(bar_raw + 8 * stage, phase, 44);
        const uint32_t row0 = base + stage * stage_bytes + r * 128, row1 = row0 + 16384;
        const bool two = g.K - c * PW_KC > 32;
        float v[64];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(v[4 * q]), "=f"(v[4 * q + 1]), "=f"(v[4 * q + 2]), "=f"(v[4 * q + 3])
                       : "r"(row0 + ((q ^ sw) << 4)));
        }
        if (two) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(v[32 + 4 * q]), "=f"(v[33 + 4 * q]), "=f"(v[34 + 4 * q]), "=f"(v[35 + 4 * q])
                         : "r"(row1 + ((q ^ sw) << 4)));
        }
        if (grow != nullptr) {  // x * gate in fp32, exactly the product the reference's SqueezeExcite forms before the conv
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            if (q >= 8 && !two) break;
            v[4 * q] *= gv[q].x; v[4 * q + 1] *= gv[q].y; v[4 * q + 2] *= gv[q].z; v[4 * q + 3] *= gv[q].w;
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {  // 16-byte chunk j of the bf16 rows = k [8j, 8j + 8)
          if (j >= 4 && !two) break;   // k >= 32 of a one-box chunk is never read by the MMAs
          uint32_t hi[4], mid[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) split_bf16x2(v[8 * j + 2 * e], v[8 * j + 2 * e + 1], hi[e], mid[e]);
          const uint32_t off = (uint32_t)((j ^ sw) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row0 + off), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3])
                       : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row1 + off), "r"(mid[0]), "r"(mid[1]), "r"(mid[2]),
                       "r"(mid[3])
                       : "memory");
        }
        fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's (async proxy) operand reads
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_ops + 8 * stage);
        if (++stage == (uint32_t)g.nstages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp >= 7) {
    // ---- epilogue: tcgen05.ld -> bias / SiLU -> shared-memory transpose -> coalesced stores (+ residual) -------------
    // One warp per scheduler does this work, so it lives on instruction-level parallelism: the 32 elements of a thread's
    // row are handled by BRANCH-FREE code (bias as 8 float4 loads up front, the activation switch hoisted out of the
    // element loop, raw ex2 / rcp) that the compiler can interleave.  (With a bias test and an activation test per
    // element every element was its own load -> add -> ex2 -> rcp chain: 880 instructions and ~3200 clocks per 32 columns.)
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, +32) are accessible to this warp
    float* stg = stg_all + quarter * (32 * 36);  // [32 rows][36]: 144-byte rows keep the 16-byte accesses of both phases conflict-free
    const int cc = (lane & 7) * 4;
    uint32_t it = 0;
    for (long long tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++it) {
      const long long mt = tile / g.n_tiles;
      const int n0 = (int)(tile - mt * g.n_tiles) * g.tn;
      const long long m0 = mt * 128;
      const uint32_t buf = it & 1;
      // destination / residual pointers of the 8 tile rows this thread stores (rows k*4 + lane/8), once per tile
      float* drow[8];
      const float* rrow[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const long long mm = m0 + quarter * 32 + k * 4 + (lane >> 3);
        long long om = mm;
        if (g.out_W > 0 && mm < g.M) {  // position inside the zero-bordered map (M < 2^31: 32-bit divisions)
          const unsigned hw = (unsigned)(g.out_W * g.out_H);
          const unsigned img = (unsigned)mm / hw, rem = (unsigned)mm - img * hw;
          const unsigned yy = rem / (unsigned)g.out_W, xx = rem - yy * (unsigned)g.out_W;
          om = ((long long)img * g.out_Hp + yy + g.out_pt) * g.out_Wp + xx + g.out_pl;
        }
        drow[k] = mm < g.M ? Y + om * g.ldc + n0 + cc : nullptr;
        rrow[k] = (g.residual != nullptr && mm < g.M) ? g.residual + mm * g.ldr + n0 + cc : nullptr;
      }
      mbar_wait(bar_acc_full + 8 * buf, (it >> 1) & 1, 45);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < g.tn; c0 += 32) {
        uint32_t acc[32];
        tmem_ld_32x32(tmem_base + buf * g.tn + c0 + ((uint32_t)(quarter * 32) << 16), acc);
        const int nvalid = min(32, min(g.tn - c0, g.N - n0 - c0));  // a multiple of 4
        // the residual pieces this thread will add, all eight loads issued now (as dependent load -> add -> store chains inside
        // the store loop they cost eight L2 round trips per 32 columns)
        float4 rv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
          rv[k] = (rrow[k] != nullptr && cc < nvalid) ? __ldg(reinterpret_cast<const float4*>(rrow[k] + c0))
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
        float bv[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 b4 = (g.bias != nullptr && 4 * q < nvalid) ? __ldg(reinterpret_cast<const float4*>(g.bias + n0 + c0) + q)
                                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
          bv[4 * q] = b4.x; bv[4 * q + 1] = b4.y; bv[4 * q + 2] = b4.z; bv[4 * q + 3] = b4.w;
        }
        tmem_ld_wait();
        if (c0 + 32 >= g.tn) {  // accumulator fully read: hand the buffer back to the MMA issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acc_empty + 8 * buf);
        }
        float o[32];
        if (g.act == 1) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {  // SiLU: t / (1 + 2^(-t log2 e)); ex2 flushes to 0 / overflows to inf at the ends, both limits exact
            const float t = __uint_as_float(acc[i]) + bv[i];
            float e, r;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t * -1.4426950408889634f));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
            o[i] = t * r;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(acc[i]) + bv[i];
        }
        float4* srow = reinterpret_cast<float4*>(stg + lane * 36);
#pragma unroll
        for (int q = 0; q < 8; ++q) srow[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        __syncwarp();
        // 8 lanes write one row's 128 bytes: every store instruction covers 4 whole lines
        if (cc < nvalid) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (drow[k] == nullptr) continue;
            float4 v = *reinterpret_cast<const float4*>(stg + (k * 4 + (lane >> 3)) * 36 + cc);
            v.x += rv[k].x; v.y += rv[k].y; v.z += rv[k].z; v.w += rv[k].w;
            *reinterpret_cast<float4*>(drow[k] + c0) = v;
          }
        }
        __syncwarp();
      }
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

int mde_pointwise_x3_fwd(const float* x, const float* gate, int64_t rows_per_image, const uint16_t* w_pair, const float* bias,
                         int act, const float* residual, float* y, int64_t M, int N, int K, int64_t ldc, int64_t ldr,
                         const int* out_pad, mde_stream_t stream) {
  if (!x || !w_pair || !y) return MDE_ERR_BAD_POINTER;
  if (out_pad && (out_pad[0] <= 0 || out_pad[1] <= 0 || out_pad[2] < 0 || out_pad[3] < 0 || out_pad[4] < 0 || out_pad[5] < 0 ||
                  M % ((int64_t)out_pad[0] * out_pad[1]) != 0 || residual))
    return MDE_ERR_BAD_SHAPE;
  if (gate && (rows_per_image <= 0 || M % rows_per_image != 0)) return MDE_ERR_BAD_SHAPE;
  if (gate && !aligned(gate, 16)) return MDE_ERR_BAD_POINTER;
  if (M <= 0 || M > 0x7fffffffLL || N <= 0 || K <= 0 || ldc < N || (residual && ldr < N) || act < 0 || act > 1)
    return MDE_ERR_BAD_SHAPE;
  if (K % 8 != 0 || !aligned(x, 16) || !aligned(w_pair, 16)) return MDE_ERR_UNSUPPORTED;  // TMA: 16-byte global strides
  // 16-byte epilogue accesses: whole float4 groups of output channels, aligned rows
  if (N % 4 != 0 || ldc % 4 != 0 || !aligned(y, 16) || (bias && !aligned(bias, 16)) ||
      (residual && (ldr % 4 != 0 || !aligned(residual, 16))))
    return MDE_ERR_UNSUPPORTED;
  tc::PwGeom g;
  g.M = M; g.N = N; g.K = K; g.ldc = ldc; g.ldr = ldr; g.act = act; g.bias = bias; g.residual = residual;
  g.gate = gate; g.rows_per_image = gate ? rows_per_image : 1;
  g.out_W = 0; g.out_H = g.out_Wp = g.out_Hp = g.out_pt = g.out_pl = 0;
  if (out_pad) {  // {H, W, pad_top, pad_bottom, pad_left, pad_right}: the caller zeroes the border
    g.out_H = out_pad[0]; g.out_W = out_pad[1]; g.out_pt = out_pad[2]; g.out_pl = out_pad[4];
    g.out_Hp = out_pad[0] + out_pad[2] + out_pad[3]; g.out_Wp = out_pad[1] + out_pad[4] + out_pad[5];
  }
  // N tile: 256 columns when that still leaves >= 2 CTAs per SM (each N tile converts the activation tile again), else 128
  const long long m_tiles = (M + 127) / 128;
  int tn = N >= 256 ? 256 : ((N + 15) / 16) * 16;
  if (tn == 256 && m_tiles * ((N + 255) / 256) < 2 * MDE_NUM_SMS) tn = 128;
  g.tn = tn;
  g.n_tiles = (N + tn - 1) / tn;
  g.chunks = (K + tc::PW_KC - 1) / tc::PW_KC;
  g.tmem_cols = 2 * tn <= 32 ? 32 : 2 * tn <= 64 ? 64 : 2 * tn <= 128 ? 128 : 2 * tn <= 256 ? 256 : 512;  // two accumulator buffers
  g.total_tiles = m_tiles * g.n_tiles;
  const int stage_bytes = tc::PW_A_BYTES + 2 * tn * 128;
  // the ring runs ACROSS tiles (a CTA's chunk stream is continuous), so its depth is set by shared memory alone -- capping it
  // at the chunks of one tile (K <= 64: one) serialises load -> convert -> MMA per tile behind a full DRAM round trip
  g.nstages = (200 * 1024) / stage_bytes;
  if (g.nstages > 6) g.nstages = 6;
  const int smem = g.nstages * stage_bytes + 24 * g.nstages + 64 + 4 * 32 * 36 * 4 + 1024;
  CUtensorMap mx, mw;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)K * 4};
    const uint32_t box[2] = {32, 128};
    if (!tc::encode_f32(&mx, x, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return MDE_ERR_DRIVER;
  }
  {
    const uint64_t dims[3] = {(uint64_t)K, (uint64_t)N, 2};
    const uint64_t strides[2] = {(uint64_t)K * 2, (uint64_t)N * K * 2};
    const uint32_t box[3] = {(uint32_t)tc::PW_KC, (uint32_t)tn, 1};
    if (!tc::encode_bf16(&mw, w_pair, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return MDE_ERR_DRIVER;
  }
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(tc::pointwise_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024) != cudaSuccess)
      return MDE_ERR_LAUNCH;
    attr = true;
  }
  const long long grid = g.total_tiles < MDE_NUM_SMS ? g.total_tiles : MDE_NUM_SMS;
  launch_pdl(PDL_TC, tc::pointwise_x3_kernel, dim3((unsigned)grid), dim3(tc::PW_THREADS), smem, (cudaStream_t)stream, mx, mw, y, g);
  return check_launch();
}

}  // extern "C"
