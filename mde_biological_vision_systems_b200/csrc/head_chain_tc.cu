// K1d / K1d+K1e+K2 -- the range-attention contraction on tcgen05, stand-alone and fused with conv_out + softmax + bins.
//
// Reference: PixelWiseDotProduct (models/layers.py:31-36), conv_out + Softmax(dim=1)
// (models/unet_adaptive_bins.py:190-191,286) and the centre-weighted sum (:298-300).
//
// One persistent, warp-specialised kernel (1 CTA / SM, 352 threads):
//   warp 0      TMA producer for the activation tiles  x[b, k, p0:p0+128]  (NCHW, so the pixel axis is contiguous:
//               the A operand is MN-major).  A tile is 128 pixels x 128 channels fp32 = 64 KB, streamed as four
//               32-channel stages of 16 KB (four 32-pixel x 32-channel SWIZZLE_128B_ATOM_32B boxes each) through an
//               NS-deep ring.
//   warp 1      MMA issuer: one thread issues tcgen05.mma.kind::tf32 (M = 128 pixels, N = NB, K = 8 per instruction,
//               16 instructions per tile), accumulating in TMEM; two TMEM accumulator buffers so the epilogue of tile
//               i overlaps the MMAs of tile i+1.
//   warp 2      per-image weight loader (B operand, K-major, SWIZZLE_128B; re-loaded when the CTA crosses an image
//               boundary) + TMEM allocation / release.
//   warps 3-10  epilogue, 8 warps per tile (4 TMEM lane quarters x 2 column halves): tcgen05.ld 32 columns at a time
//               (thread = pixel row; the next chunk is prefetched while the current one is reduced), then either
//                 EPI_STORE   : write y[b, n, p]                      (stand-alone range attention, N = 128)
//                 EPI_SOFTMAX : online softmax over the NB = 256 logits and centre-weighted sum -> pred[b, p]
//               so in the fused form neither the range-attention maps (29 MB/img) nor the logits / softmax
//               (58 MB/img each) ever reach HBM: algorithmic traffic is 128*P*4 B in + P*4 B out per image.
// In the fused form the two 1x1 contractions are folded by associativity, W'_b = W_out @ Q_b (tiny, exact fp32,
// mde_fold_queries), pre-scaled by log2(e) and rounded to TF32 once, so the tensor cores run ONE K = 128 contraction
// per pixel against the per-image 256 x 128 operand; the bias enters the softmax as exp2(b_j) factors.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace mde {
namespace tc {

constexpr int TILE_M = 128;                      // pixels per tile (UMMA M)
constexpr int KDIM = 128;                        // contraction length (channels)
constexpr int KC = 32;                           // channels per stage = one 128-byte swizzle row per channel
constexpr int STAGE_BYTES = TILE_M * KC * 4;     // 16384
constexpr int BOX_BYTES = 32 * KC * 4;           // one 32-pixel x 32-channel box
constexpr int EPI_WARP0 = 3;                     // warps 0..2: TMA producer, MMA issuer, weight loader / TMEM allocator
// epilogue warps per tile = 4 TMEM lane quarters x ES column parts (ES = 2: 8 warps, 128 columns each; ES = 4: 16 warps,
// 64 columns each -- more warps per scheduler to keep the MUFU busy, at <= 104 registers per thread)
__host__ __device__ constexpr int chain_threads(int es) { return 32 * EPI_WARP0 + 128 * es; }
enum { EPI_STORE = 0, EPI_SOFTMAX = 1, EPI_BWD = 2 };

// extra pointers of the training forms (all may be null for inference):
//   EPI_SOFTMAX: stats [B,P,2] receives the per-pixel softmax state (max logit in log2 units, sum of 2^(z - max) * 2^bias)
//   EPI_BWD    : reads stats, pred (the forward output) and gpred [B,P]; writes gl [B,P,NB], glT [B,NB,P] (TF32-rounded
//                d loss / d logit) and accumulates gc [B,NB] (d loss / d centre) and gb [B,NB] (sum over pixels of gl)
struct ChainTrain {
  float* stats;
  const float* pred;
  const float* gpred;
  float* gl;
  float* glT;
  float* gc;
  float* gb;
};

// column sums over the 32 lanes of a warp of v[32] (lane = row): afterwards v[0] of lane l holds sum_rows v[l]
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = upper ? v[i] : v[i + n / 2];
      const float keep = upper ? v[i + n / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

struct DebugCfg {
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo, version;
  long long* prof;  // optional [grid][8] cycle counters: where each role waits (bring-up / tuning only)
};
static DebugCfg g_dbg = {4096, 512, 16, 1024, 1, nullptr};

#define MDE_TIMED_WAIT(acc, ...)        \
  {                                     \
    const long long t__ = clock64();    \
    __VA_ARGS__;                        \
    acc += clock64() - t__;             \
  }

// CTAS = 2: a CTA pair (cluster of two, tcgen05 cta_group::2) works on two consecutive 128-pixel tiles with ONE M = 256
// MMA stream issued by the leader: each CTA stages its own tile and only HALF of the per-image B operand (NB/2 bin rows),
// so the shared-memory operand reads per MMA drop from 12 KB to 8 KB per SM -- the measured ceiling of the 1-CTA form.
template <int NB, int CTAS = 1>
struct SmemPlan {
  static constexpr int NBH = NB / CTAS;           // B rows held by one CTA
  static constexpr int W_BYTES = NBH * KDIM * 4;  // per-image B operand: 4 K-chunks x [NBH rows][128 B]
  static constexpr int NS = (NB == 256) ? (CTAS == 2 ? 8 : 5) : 8;
  static constexpr int RING_BYTES = NS * STAGE_BYTES;
  static constexpr int CONST_BYTES = 2 * NB * 4 + 3 * 128 * 4 * 4;  // exp2(bias), exp2(bias)*centre per bin; merge slots
  static constexpr int BAR_BYTES = 512;
  static constexpr int TOTAL = W_BYTES + RING_BYTES + CONST_BYTES + BAR_BYTES + 1024;  // + alignment slack
};

// A_KMAJOR = false: activations NCHW (pixel axis contiguous, MN-major A, four 32-pixel boxes per stage)
// A_KMAJOR = true : activations NHWC / channels_last (channel axis contiguous, K-major A, one 128-row box per stage)
template <int NB, int EPI, bool A_KMAJOR, int CTAS = 1, int ES = 2>
__global__ void __launch_bounds__(chain_threads(ES), 1)
    head_chain_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                      const float* __restrict__ biasf, const float* __restrict__ centers, float* __restrict__ out,
                      int tiles_per_img, int total_tiles, long long P, DebugCfg dbg, ChainTrain tr) {
  using Plan = SmemPlan<NB, CTAS>;
  constexpr int NS = Plan::NS;
  constexpr int NBH = Plan::NBH;
  const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;  // 0 = leader (issues the MMAs of the pair)
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* gbase = smem_dyn + (base - smem_u32(smem_dyn));
  const uint32_t s_w = base;
  const uint32_t s_ring = s_w + Plan::W_BYTES;
  float* c_all = reinterpret_cast<float*>(gbase + Plan::W_BYTES + Plan::RING_BYTES);  // [group][2][NB]
  const uint32_t s_bar = s_ring + Plan::RING_BYTES + Plan::CONST_BYTES;
  // barrier slots (8 B each)
  const uint32_t bar_full = s_bar;                 // [NS]
  const uint32_t bar_empty = s_bar + 8 * NS;       // [NS]
  const uint32_t bar_wfull = s_bar + 16 * NS;      // [1]
  const uint32_t bar_wempty = bar_wfull + 8;       // [1]
  const uint32_t bar_accfull = bar_wempty + 8;     // [2]
  const uint32_t bar_accempty = bar_accfull + 16;  // [2]
  const uint32_t s_tmem_slot = bar_accempty + 16;  // uint32
  const uint32_t bar_peerfull = s_tmem_slot + 8;   // [NS]  leader only: the peer CTA's stage has landed (remote arrive)
  const uint32_t bar_peerwfull = bar_peerfull + 8 * NS;  // [1]
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + (s_tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // contiguous range of tiles (CTAS = 2: tile PAIRS; `tiles_per_img` / `total_tiles` then count pairs) of this CTA /
  // CTA pair, so that it crosses at most one image boundary
  const int unit = blockIdx.x / CTAS, nunits = gridDim.x / CTAS;
  const int t_begin = (int)(((long long)total_tiles * unit) / nunits);
  const int t_end = (int)(((long long)total_tiles * (unit + 1)) / nunits);

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    mbar_init(bar_wfull, 1);
    mbar_init(bar_wempty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_accfull + 8 * i, 1);
      mbar_init(bar_accempty + 8 * i, 4 * ES * CTAS);  // one arrive per epilogue warp (of both CTAs of a pair)
    }
    if (CTAS == 2) {
      for (int i = 0; i < NS; ++i) mbar_init(bar_peerfull + 8 * i, 1);
      mbar_init(bar_peerwfull, 1);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
  }
  constexpr uint32_t TMEM_COLS = (2 * NB <= 32) ? 32 : (2 * NB <= 64) ? 64 : (2 * NB <= 128) ? 128 : (2 * NB <= 256) ? 256 : 512;
  if (warp == 2) {
    if (CTAS == 2) {
      tmem_alloc2(s_tmem_slot, TMEM_COLS);
      tmem_relinquish2();
    } else {
      tmem_alloc(s_tmem_slot, TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all();  // barrier inits visible to the peer before any remote arrive / multicast commit
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= activation producer =================
    if (lane == 0) {
      // Optional L2 prefetch PF_DIST tiles ahead (cp.async.bulk.prefetch.tensor).  Measured on B200 (scripts/
      // chain_waits.py): the MMA thread waits only ~490 of ~3800 cycles/tile for activations and the prefetch did not
      // reduce that (116 vs 109 us), so it is off; the kernel is bound by the SS-mode tensor pipe reading 12 KB of
      // shared memory per 128x256x8 TF32 MMA (~190 instead of 128 cycles each) -- the fix is cta_group::2.
      constexpr int PF_DIST = 0;
      auto prefetch_tile = [&](int t) {
        const int img = t / tiles_per_img;
        const int p0 = ((t - img * tiles_per_img) * CTAS + (int)rank) * TILE_M;
        for (int kc = 0; kc < KDIM / KC; ++kc) {
          if (A_KMAJOR) {
            tma_prefetch_l2_2d(&map_x, kc * KC, (int)((long long)img * P + p0));
          } else {
#pragma unroll
            for (int m = 0; m < 4; ++m) tma_prefetch_l2_3d(&map_x, p0 + 32 * m, kc * KC, img);
          }
        }
      };
      if (PF_DIST > 0)
        for (int d = 1; d < PF_DIST && t_begin + d < t_end; ++d) prefetch_tile(t_begin + d);
      uint32_t stage = 0, phase = 0;
      long long w_empty = 0;
      const long long t_start = clock64();
      for (int t = t_begin; t < t_end; ++t) {
        const int img = t / tiles_per_img;
        const int p0 = ((t - img * tiles_per_img) * CTAS + (int)rank) * TILE_M;
        if (PF_DIST > 0 && t + PF_DIST < t_end) prefetch_tile(t + PF_DIST);
        for (int kc = 0; kc < KDIM / KC; ++kc) {
          MDE_TIMED_WAIT(w_empty, mbar_wait(bar_empty + 8 * stage, phase ^ 1, 1))
          mbar_expect_tx(bar_full + 8 * stage, STAGE_BYTES);
          const uint32_t dst = s_ring + stage * STAGE_BYTES;
          if (A_KMAJOR) {
            tma_load_2d(dst, &map_x, bar_full + 8 * stage, kc * KC, (int)((long long)img * P + p0));
          } else {
#pragma unroll
            for (int m = 0; m < 4; ++m)
              tma_load_3d(dst + m * BOX_BYTES, &map_x, bar_full + 8 * stage, p0 + 32 * m, kc * KC, img);
          }
          if (++stage == NS) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      if (dbg.prof) {
        dbg.prof[blockIdx.x * 8 + 0] = w_empty;
        dbg.prof[blockIdx.x * 8 + 1] = clock64() - t_start;
      }
    }
  } else if (warp == 2) {
    // ================= per-image weight loader =================
    if (lane == 0) {
      int cur = -1;
      uint32_t wphase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        const int img = t / tiles_per_img;
        if (img == cur) continue;
        cur = img;
        mbar_wait(bar_wempty, wphase ^ 1, 2);  // previous image's MMAs have drained
        mbar_expect_tx(bar_wfull, Plan::W_BYTES);
#pragma unroll
        for (int kc = 0; kc < KDIM / KC; ++kc)
          tma_load_3d(s_w + kc * (NBH * 128), &map_w, bar_wfull, kc * KC, (int)rank * NBH, img);
        wphase ^= 1;
      }
    }
  } else if (warp == 1 && CTAS == 2 && rank == 1) {
    // ================= peer CTA: relay "my operands have landed" to the leader's barriers =================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, wphase = 0;
      int cur = -1;
      for (int t = t_begin; t < t_end; ++t) {
        const int img = t / tiles_per_img;
        if (img != cur) {
          mbar_wait(bar_wfull, wphase, 7);
          mbar_arrive_remote(bar_peerwfull, 0);
          wphase ^= 1;
          cur = img;
        }
        for (int kc = 0; kc < KDIM / KC; ++kc) {
          mbar_wait(bar_full + 8 * stage, phase, 8);
          mbar_arrive_remote(bar_peerfull + 8 * stage, 0);
          if (++stage == NS) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc =
          make_idesc(FMT_TF32, TILE_M * CTAS, NB, /*A MN-major?*/ A_KMAJOR ? 0 : 1, /*B K-major*/ 0);
      uint32_t stage = 0, phase = 0, wphase = 0;
      int cur = -1;
      int it = 0;
      long long w_full = 0, w_acc = 0, w_w = 0;
      const long long t_start = clock64();
      for (int t = t_begin; t < t_end; ++t, ++it) {
        const int img = t / tiles_per_img;
        if (img != cur) {
          if (cur >= 0) {  // fires when every MMA that read the old weights is done
            if (CTAS == 2) umma2_commit_mc(bar_wempty);
            else umma_commit(bar_wempty);
          }
          MDE_TIMED_WAIT(w_w, mbar_wait(bar_wfull, wphase, 3))
          if (CTAS == 2) mbar_wait_cluster(bar_peerwfull, wphase, 9);
          wphase ^= 1;
          cur = img;
        }
        const uint32_t buf = it & 1;
        if (CTAS == 2) {
          MDE_TIMED_WAIT(w_acc, mbar_wait_cluster(bar_accempty + 8 * buf, ((it >> 1) & 1) ^ 1, 4))
        } else {
          MDE_TIMED_WAIT(w_acc, mbar_wait(bar_accempty + 8 * buf, ((it >> 1) & 1) ^ 1, 4))
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * NB;
        for (int kc = 0; kc < KDIM / KC; ++kc) {
          MDE_TIMED_WAIT(w_full, mbar_wait(bar_full + 8 * stage, phase, 5))
          if (CTAS == 2) mbar_wait_cluster(bar_peerfull + 8 * stage, phase, 10);
          tc_fence_after();
          const uint32_t a_base = s_ring + stage * STAGE_BYTES;
          const uint32_t b_base = s_w + kc * (NBH * 128);
#pragma unroll
          for (int j = 0; j < KC / 8; ++j) {
            // A (MN-major fp32/TF32 => "128B swizzle, 32B atom" layout, descriptor type 1): rows = channels (128 B =
            // 32 pixels each), K-atom = 4 rows = 512 B (SBO), 32-pixel MN-atoms 4096 B apart (LBO); one MMA eats K = 8
            // channels = 1024 B.  (Plain SWIZZLE_128B with an MN-major 32-bit operand silently yields zeros.)
            const uint64_t adesc = A_KMAJOR
                                       ? make_smem_desc(a_base + j * 32, dbg.b_lbo, dbg.b_sbo, SWZ_128B, dbg.version)
                                       : make_smem_desc(a_base + j * 1024, dbg.a_lbo, dbg.a_sbo, SWZ_128B_32B, dbg.version);
            // B (K-major, SW128): 8 tf32 = 32 B along the 128-B swizzle row; 8-row atoms 1024 B apart
            const uint64_t bdesc = make_smem_desc(b_base + j * 32, dbg.b_lbo, dbg.b_sbo, SWZ_128B, dbg.version);
            if (CTAS == 2) umma2_tf32_ss(d_tmem, adesc, bdesc, idesc, (kc | j) != 0);
            else umma_tf32_ss(d_tmem, adesc, bdesc, idesc, (kc | j) != 0);
          }
          // frees the stage (in both CTAs of a pair) when these MMAs complete
          if (CTAS == 2) umma2_commit_mc(bar_empty + 8 * stage);
          else umma_commit(bar_empty + 8 * stage);
          if (++stage == NS) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (CTAS == 2) umma2_commit_mc(bar_accfull + 8 * buf);
        else umma_commit(bar_accfull + 8 * buf);
      }
      if (dbg.prof) {
        dbg.prof[blockIdx.x * 8 + 2] = w_full;
        dbg.prof[blockIdx.x * 8 + 3] = w_acc;
        dbg.prof[blockIdx.x * 8 + 4] = w_w;
        dbg.prof[blockIdx.x * 8 + 5] = clock64() - t_start;
      }
    }
  } else {
    // ================= epilogue: warps 3..10, all on the same tile =================
    // warp -> (lane quarter = warp & 3, column half = (warp - 3) / 4).  Splitting the 256 accumulator columns over two
    // warps per lane quarter halves the time a TMEM buffer is held, so the MMAs of tile i+1 (other buffer) fully
    // overlap the epilogue of tile i; the two partial softmax states are merged through shared memory.
    const int half = (warp - EPI_WARP0) >> 2;
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32) are accessible to this warp
    const int etid = threadIdx.x - EPI_WARP0 * 32;
    float* c_fac = c_all;           // [NB] exp2(bias)
    float* c_cen = c_all + NB;      // [NB] exp2(bias)*centre
    float4* merge = reinterpret_cast<float4*>(c_all + 2 * NB);  // [128] (m, s, ws, -) of the upper column half
    constexpr int COLS = NB / ES;  // columns per warp
    constexpr int EPI_THREADS = 128 * ES;
    // hand an accumulator buffer back to the MMA issuer (the leader CTA's barrier; the peer arrives remotely)
    auto acc_release = [&](uint32_t bar) {
      if (CTAS == 2 && rank != 0) mbar_arrive_remote(bar, 0);
      else mbar_arrive(bar);
    };
    float acc_gb[4] = {0.f, 0.f, 0.f, 0.f}, acc_gc[4] = {0.f, 0.f, 0.f, 0.f};  // EPI_BWD: per-lane bin sums
    int cur = -1;
    int it = 0;
    long long w_accfull = 0;
    const long long t_start = clock64();
    for (int t = t_begin; t < t_end; ++t, ++it) {
      const int img = t / tiles_per_img;
      const int p0 = ((t - img * tiles_per_img) * CTAS + (int)rank) * TILE_M;
      if (EPI == EPI_BWD && img != cur && cur >= 0) {
        // per-bin sums of the image just finished: lane l of this warp owns bin half*COLS + 32*chunk + l
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          atomicAdd(tr.gb + (long long)cur * NB + half * (NB / ES) + 32 * c + lane, acc_gb[c]);
          atomicAdd(tr.gc + (long long)cur * NB + half * (NB / ES) + 32 * c + lane, acc_gc[c]);
          acc_gb[c] = 0.f;
          acc_gc[c] = 0.f;
        }
      }
      if (EPI != EPI_STORE && img != cur) {
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");  // everyone finished reading the old constants
        for (int j = etid; j < NB; j += EPI_THREADS) {
          const float f = exp2f(biasf[(long long)img * NB + j]);
          c_fac[j] = f;
          c_cen[j] = f * centers[(long long)img * NB + j];
        }
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
      }
      cur = img;
      const uint32_t buf = it & 1;
      MDE_TIMED_WAIT(w_accfull, mbar_wait(bar_accfull + 8 * buf, (it >> 1) & 1, 6))
      tc_fence_after();
      const uint32_t taddr = tmem_base + buf * NB + half * COLS + ((uint32_t)(quarter * 32) << 16);
      const int row = quarter * 32 + lane;
      const long long pix = (long long)p0 + row;
      if constexpr (EPI == EPI_SOFTMAX) {
        float m = -INFINITY;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
        uint32_t ra[32], rb[32];
        tmem_ld_32x32(taddr, ra);
        // one 32-column chunk: r = registers of this chunk, nxt = registers to prefetch the next chunk into
#define MDE_CHUNK(r, nxt, c0, last)                                                                                    \
  {                                                                                                                    \
    tmem_ld_wait();                                                                                                    \
    if (last) { /* accumulator fully read: hand the TMEM buffer back to the MMA warp */                                \
      tc_fence_before();                                                                                               \
      __syncwarp();                                                                                                    \
      if (lane == 0) acc_release(bar_accempty + 8 * buf);                                                              \
    } else {                                                                                                           \
      tmem_ld_32x32(taddr + (c0) + 32, nxt); /* prefetch: overlaps the math below */                                   \
    }                                                                                                                  \
    float t8[8];                                                                                                       \
    _Pragma("unroll") for (int i = 0; i < 8; ++i) t8[i] =                                                              \
        fmaxf(fmaxf(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1])),                                         \
              fmaxf(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])));                                    \
    const float cm =                                                                                                   \
        fmaxf(fmaxf(fmaxf(t8[0], t8[1]), fmaxf(t8[2], t8[3])), fmaxf(fmaxf(t8[4], t8[5]), fmaxf(t8[6], t8[7])));       \
    if (cm > m) {                                                                                                      \
      const float sc = ex2_approx(m - cm); /* 0 on the first chunk (m = -inf) */                                       \
      s0 *= sc; s1 *= sc; s2 *= sc; s3 *= sc;                                                                          \
      w0 *= sc; w1 *= sc; w2 *= sc; w3 *= sc;                                                                          \
      m = cm;                                                                                                          \
    }                                                                                                                  \
    _Pragma("unroll") for (int i = 0; i < 32; i += 4) {                                                                \
      const float4 f = *reinterpret_cast<const float4*>(c_fac + half * COLS + (c0) + i);                               \
      const float4 g = *reinterpret_cast<const float4*>(c_cen + half * COLS + (c0) + i);                               \
      const float e0 = ex2_approx(__uint_as_float(r[i + 0]) - m);                                                      \
      const float e1 = ex2_approx(__uint_as_float(r[i + 1]) - m);                                                      \
      const float e2 = ex2_approx(__uint_as_float(r[i + 2]) - m);                                                      \
      const float e3 = ex2_approx(__uint_as_float(r[i + 3]) - m);                                                      \
      s0 = fmaf(e0, f.x, s0); w0 = fmaf(e0, g.x, w0);                                                                  \
      s1 = fmaf(e1, f.y, s1); w1 = fmaf(e1, g.y, w1);                                                                  \
      s2 = fmaf(e2, f.z, s2); w2 = fmaf(e2, g.z, w2);                                                                  \
      s3 = fmaf(e3, f.w, s3); w3 = fmaf(e3, g.w, w3);                                                                  \
    }                                                                                                                  \
  }
        static_assert(COLS == 128 || COLS == 64, "the unrolled epilogue covers 4 or 2 chunks of 32 columns per warp");
        if constexpr (COLS == 128) {
          MDE_CHUNK(ra, rb, 0, false)
          MDE_CHUNK(rb, ra, 32, false)
          MDE_CHUNK(ra, rb, 64, false)
          MDE_CHUNK(rb, ra, 96, true)
        } else {
          MDE_CHUNK(ra, rb, 0, false)
          MDE_CHUNK(rb, ra, 32, true)
        }
#undef MDE_CHUNK
        const float s = (s0 + s1) + (s2 + s3), ws = (w0 + w1) + (w2 + w3);
        // merge the ES column parts of each pixel row
        if (half != 0) merge[(half - 1) * 128 + row] = make_float4(m, s, ws, 0.f);
        asm volatile("bar.sync 2, %0;" ::"n"(EPI_THREADS) : "memory");
        if (half == 0) {
          float mm = m, stot = s, wtot = ws;
#pragma unroll
          for (int k = 0; k < ES - 1; ++k) {
            const float4 o = merge[k * 128 + row];
            const float nm = fmaxf(mm, o.x);
            const float a = ex2_approx(mm - nm), bsc = ex2_approx(o.x - nm);
            stot = stot * a + o.y * bsc;
            wtot = wtot * a + o.z * bsc;
            mm = nm;
          }
          out[(long long)img * P + pix] = wtot / stot;
          if (tr.stats) *reinterpret_cast<float2*>(tr.stats + 2 * ((long long)img * P + pix)) = make_float2(mm, stot);
        }
        asm volatile("bar.sync 3, %0;" ::"n"(EPI_THREADS) : "memory");  // merge slots free for the next tile
      } else if constexpr (EPI == EPI_BWD) {
        // single pass given the forward's softmax state: p_j = 2^(z_j - m) f_j / S;  u_j = p_j g;  gl_j = u_j (c_j - pred)
        static_assert(COLS == 128 && ES == 2, "4 chunks of 32 columns per warp");
        const long long gp = (long long)img * P + pix;
        const float2 st = *reinterpret_cast<const float2*>(tr.stats + 2 * gp);
        const float predv = tr.pred[gp];
        const float inv = tr.gpred[gp] / st.y;
        float* gl_row = tr.gl + gp * NB + half * COLS;
        float* glt_col = tr.glT + ((long long)img * NB + half * COLS) * P + pix;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + 32 * c, r);
          tmem_ld_wait();
          if (c == 3) {  // accumulator fully read
            tc_fence_before();
            __syncwarp();
            if (lane == 0) acc_release(bar_accempty + 8 * buf);
          }
          float e[32], gv[32];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 f = *reinterpret_cast<const float4*>(c_fac + half * COLS + 32 * c + i);
            const float4 cc = *reinterpret_cast<const float4*>(c_cen + half * COLS + 32 * c + i);
            const float fa[4] = {f.x, f.y, f.z, f.w}, ca[4] = {cc.x, cc.y, cc.z, cc.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float ev = ex2_approx(__uint_as_float(r[i + q]) - st.x) * inv;
              gv[i + q] = tf32_round(ev * (ca[q] - fa[q] * predv));
              e[i + q] = ev * fa[q];  // u_j
            }
            *reinterpret_cast<float4*>(gl_row + 32 * c + i) = make_float4(gv[i], gv[i + 1], gv[i + 2], gv[i + 3]);
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) glt_col[(long long)(32 * c + i) * P] = gv[i];  // 128 B per warp per bin row
          acc_gb[c] += warp_transpose_sum(gv, lane);
          acc_gc[c] += warp_transpose_sum(e, lane);
        }
      } else {
#pragma unroll 1
        for (int c0 = 0; c0 < COLS; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c0, r);
          tmem_ld_wait();
          if (c0 + 32 == COLS) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) acc_release(bar_accempty + 8 * buf);
          }
          float* dst = out + ((long long)img * NB + half * COLS + c0) * P + pix;
#pragma unroll
          for (int i = 0; i < 32; ++i) dst[(long long)i * P] = __uint_as_float(r[i]);  // 128 B per warp per row
        }
      }
    }
    if (EPI == EPI_BWD && cur >= 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        atomicAdd(tr.gb + (long long)cur * NB + half * (NB / ES) + 32 * c + lane, acc_gb[c]);
        atomicAdd(tr.gc + (long long)cur * NB + half * (NB / ES) + 32 * c + lane, acc_gc[c]);
      }
    }
    if (dbg.prof && etid == 0) {
      dbg.prof[blockIdx.x * 8 + 6] = w_accfull;
      dbg.prof[blockIdx.x * 8 + 7] = clock64() - t_start;
    }
  }
  // ---- teardown ----
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all();  // neither CTA may exit while the other can still signal its barriers
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (CTAS == 2) tmem_dealloc2(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int NB, int EPI, bool A_KMAJOR, int CTAS, int ES = 2>
static int launch_chain_impl(const float* x, const float* w, const float* biasf, const float* centers, float* out, int B,
                             long long P, cudaStream_t st, ChainTrain tr) {
  using Plan = SmemPlan<NB, CTAS>;
  if (P % (TILE_M * CTAS) != 0) return MDE_ERR_BAD_SHAPE;
  if (!aligned(x, 16) || !aligned(w, 16)) return MDE_ERR_BAD_POINTER;
  CUtensorMap mx, mw;
  if (A_KMAJOR) {
    if ((long long)B * P > 0x7fffffffLL) return MDE_ERR_BAD_SHAPE;
    const uint64_t dims[2] = {(uint64_t)KDIM, (uint64_t)B * (uint64_t)P};
    const uint64_t strides[1] = {(uint64_t)KDIM * 4};
    const uint32_t box[2] = {KC, TILE_M};
    if (!encode_f32(&mx, x, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return MDE_ERR_DRIVER;
  } else {
    const uint64_t dims[3] = {(uint64_t)P, (uint64_t)KDIM, (uint64_t)B};
    const uint64_t strides[2] = {(uint64_t)P * 4, (uint64_t)P * KDIM * 4};
    const uint32_t box[3] = {32, KC, 1};
    if (!encode_f32(&mx, x, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return MDE_ERR_DRIVER;
  }
  {
    const uint64_t dims[3] = {(uint64_t)KDIM, (uint64_t)NB, (uint64_t)B};
    const uint64_t strides[2] = {(uint64_t)KDIM * 4, (uint64_t)NB * KDIM * 4};
    const uint32_t box[3] = {KC, (uint32_t)(NB / CTAS), 1};  // a CTA of a pair stages its half of the bin rows
    if (!encode_f32(&mw, w, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return MDE_ERR_DRIVER;
  }
  const int units_per_img = (int)(P / (TILE_M * CTAS));  // tiles (CTAS = 1) or tile pairs (CTAS = 2) per image
  const long long total = (long long)units_per_img * B;
  if (total > 0x7fffffffLL) return MDE_ERR_BAD_SHAPE;
  const int max_units = MDE_NUM_SMS / CTAS;
  const int grid = CTAS * (int)(total < max_units ? total : max_units);
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(head_chain_kernel<NB, EPI, A_KMAJOR, CTAS, ES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             Plan::TOTAL) != cudaSuccess)
      return MDE_ERR_LAUNCH;
    attr_set = true;
  }
  if (CTAS == 1) {
    head_chain_kernel<NB, EPI, A_KMAJOR, CTAS, ES><<<grid, chain_threads(ES), Plan::TOTAL, st>>>(
        mx, mw, biasf, centers, out, units_per_img, (int)total, P, g_dbg, tr);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(chain_threads(ES));
    cfg.dynamicSmemBytes = Plan::TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CTAS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, head_chain_kernel<NB, EPI, A_KMAJOR, CTAS, ES>, mx, mw, biasf, centers, out, units_per_img,
                           (int)total, P, g_dbg, tr) != cudaSuccess)
      return MDE_ERR_LAUNCH;
  }
  return check_launch();
}

// MDE_CHAIN_CTAS=2 selects the CTA-pair (cta_group::2) form when the per-image tile count is even.  It is parity-green but
// NOT the default: measured on B200 at config 2 it runs in 113 us against 107 us for the single-CTA form -- halving the
// B-operand shared-memory reads does not help because the epilogue (8 warps, 2 per scheduler: 256 warp-wide MUFU.EX2 per
// tile and scheduler = 2048 clocks of the ~3500 per tile, 28 % of all stall samples) paces the kernel, and the pair adds a
// relay hop per stage and lock-steps two tiles.
template <int NB, int EPI, bool A_KMAJOR>
static int launch_chain(const float* x, const float* w, const float* biasf, const float* centers, float* out, int B,
                        long long P, cudaStream_t st, ChainTrain tr = ChainTrain{}) {
  if constexpr (NB == 256 && EPI != EPI_STORE) {
    const char* force = getenv("MDE_CHAIN_CTAS");
    const bool pair = (P % (2 * TILE_M) == 0) && (force && atoi(force) == 2) && g_dbg.prof == nullptr;
    if constexpr (EPI == EPI_SOFTMAX) {
      const char* es = getenv("MDE_CHAIN_ES");  // tuning aid: 2 = eight 128-column epilogue warps, 4 = sixteen 64-column ones
      const bool wide = es && atoi(es) == 4;
      if (pair && wide) return launch_chain_impl<NB, EPI, A_KMAJOR, 2, 4>(x, w, biasf, centers, out, B, P, st, tr);
      if (wide) return launch_chain_impl<NB, EPI, A_KMAJOR, 1, 4>(x, w, biasf, centers, out, B, P, st, tr);
    }
    if (pair) return launch_chain_impl<NB, EPI, A_KMAJOR, 2>(x, w, biasf, centers, out, B, P, st, tr);
  }
  return launch_chain_impl<NB, EPI, A_KMAJOR, 1>(x, w, biasf, centers, out, B, P, st, tr);
}

__global__ void round_tf32_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, float scale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned int r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(in[i] * scale));
    out[i] = __uint_as_float(r);
  }
}

}  // namespace tc
}  // namespace mde

using namespace mde;

extern "C" {

int mde_head_chain_fwd(const float* x, int x_channels_last, const float* wf, const float* biasf, const float* centers,
                       float* pred, int B, int n_bins, int64_t P, mde_stream_t stream) {
  if (!x || !wf || !biasf || !centers || !pred) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || P <= 0) return MDE_ERR_BAD_SHAPE;
  if (n_bins != 256) return MDE_ERR_UNSUPPORTED;
  if (x_channels_last)
    return tc::launch_chain<256, tc::EPI_SOFTMAX, true>(x, wf, biasf, centers, pred, B, P, (cudaStream_t)stream);
  return tc::launch_chain<256, tc::EPI_SOFTMAX, false>(x, wf, biasf, centers, pred, B, P, (cudaStream_t)stream);
}

// training forward: as mde_head_chain_fwd, also recording the per-pixel softmax state stats [B,P,2] for the backward
int mde_head_chain_fwd_train(const float* x, int x_channels_last, const float* wf, const float* biasf, const float* centers,
                             float* pred, float* stats, int B, int n_bins, int64_t P, mde_stream_t stream) {
  if (!x || !wf || !biasf || !centers || !pred || !stats) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || P <= 0) return MDE_ERR_BAD_SHAPE;
  if (n_bins != 256) return MDE_ERR_UNSUPPORTED;
  tc::ChainTrain tr{};
  tr.stats = stats;
  if (x_channels_last)
    return tc::launch_chain<256, tc::EPI_SOFTMAX, true>(x, wf, biasf, centers, pred, B, P, (cudaStream_t)stream, tr);
  return tc::launch_chain<256, tc::EPI_SOFTMAX, false>(x, wf, biasf, centers, pred, B, P, (cudaStream_t)stream, tr);
}

// backward, step 1: recompute the logits on the tensor cores and emit d loss / d logit in both layouts + per-bin sums
int mde_head_chain_bwd_logits(const float* x, int x_channels_last, const float* wf, const float* biasf,
                              const float* centers, const float* pred, const float* stats, const float* gpred, float* gl,
                              float* glT, float* gc, float* gb, int B, int n_bins, int64_t P, mde_stream_t stream) {
  if (!x || !wf || !biasf || !centers || !pred || !stats || !gpred || !gl || !glT || !gc || !gb) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || P <= 0) return MDE_ERR_BAD_SHAPE;
  if (n_bins != 256) return MDE_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(gc, 0, sizeof(float) * (size_t)B * n_bins, st);
  cudaMemsetAsync(gb, 0, sizeof(float) * (size_t)B * n_bins, st);
  tc::ChainTrain tr{};
  tr.stats = const_cast<float*>(stats);
  tr.pred = pred;
  tr.gpred = gpred;
  tr.gl = gl;
  tr.glT = glT;
  tr.gc = gc;
  tr.gb = gb;
  if (x_channels_last)
    return tc::launch_chain<256, tc::EPI_BWD, true>(x, wf, biasf, centers, nullptr, B, P, st, tr);
  return tc::launch_chain<256, tc::EPI_BWD, false>(x, wf, biasf, centers, nullptr, B, P, st, tr);
}

// stand-alone range attention on tensor cores (called by mde_range_attention(impl = 1)); q should be TF32-rounded
int mde_range_attention_tc(const float* x, const float* q, float* y, int B, int K, int N, int64_t P, cudaStream_t st) {
  if (K != 128 || N != 128) return MDE_ERR_UNSUPPORTED;
  return tc::launch_chain<128, tc::EPI_STORE, false>(x, q, nullptr, nullptr, y, B, P, st);
}

int mde_round_tf32(const float* in, float* out, int64_t n, float scale, mde_stream_t stream) {
  if (!in || !out) return MDE_ERR_BAD_POINTER;
  if (n <= 0) return n == 0 ? MDE_OK : MDE_ERR_BAD_SHAPE;
  long long g = (n + 255) / 256;
  if (g > MDE_NUM_SMS * 8) g = MDE_NUM_SMS * 8;
  tc::round_tf32_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(in, out, n, scale);
  return check_launch();
}

// debug/bring-up knobs for the UMMA shared-memory descriptors (bytes); version = descriptor version field
int mde_tc_debug_config(int a_lbo, int a_sbo, int b_lbo, int b_sbo, int version) {
  tc::g_dbg.a_lbo = (uint32_t)a_lbo;
  tc::g_dbg.a_sbo = (uint32_t)a_sbo;
  tc::g_dbg.b_lbo = (uint32_t)b_lbo;
  tc::g_dbg.b_sbo = (uint32_t)b_sbo;
  tc::g_dbg.version = (uint32_t)version;
  return MDE_OK;
}

// device buffer of [grid][8] int64 cycle counters filled by the next chain launches (NULL disables): per CTA
// {producer wait-empty, producer total, mma wait-full, mma wait-acc-empty, mma wait-weights, mma total,
//  epilogue wait-acc-full, epilogue total}
int mde_tc_debug_profile(long long* buf) {
  tc::g_dbg.prof = buf;
  return MDE_OK;
}

// last barrier-timeout code recorded by a tcgen05 kernel of this file (0 = none). Synchronises the device.
int mde_tc_last_error(void) {
  int v = 0;
  cudaMemcpyFromSymbol(&v, tc::g_tc_error, sizeof(int));
  return v;
}

}  // extern "C"
