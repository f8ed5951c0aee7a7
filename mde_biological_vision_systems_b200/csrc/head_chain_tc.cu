// K1d / K1d+K1e+K2 -- the range-attention contraction on tcgen05, stand-alone and fused with conv_out + softmax + bins.
//
// Reference: PixelWiseDotProduct (models/layers.py:31-36), conv_out + Softmax(dim=1)
// (models/unet_adaptive_bins.py:190-191,286) and the centre-weighted sum (:298-300).
//
// Operands travel as split-bf16 PAIRS (hi, mid planes; tc_common.cuh): the activations x[b, p, k] (K-major: channels
// contiguous, i.e. the NHWC feature map the conv3x3 epilogue writes) and the per-image operand W'_b[j, k].  Every K step
// issues the three products hi*hi + mid*hi + hi*mid (tcgen05.mma.kind::f16, fp32 accumulation in TMEM): the logits carry
// ~2^-17 relative error per product, which keeps `pred` within 1e-3 of the fp32 reference on every pixel even for logits
// of magnitude ~100 (a single TF32 pass is off by 2e-3 already at |logit| ~ 10; scripts/precision_study_head.py).
//
// One persistent, warp-specialised kernel (1 CTA / SM, 352 threads):
//   warp 0      TMA producer: a 128-pixel tile is 2 planes x 2 K-chunks of 64 channels = four 16 KB units
//               ([128 px][128 B], SWIZZLE_128B) streamed through an NS-deep ring, order (hi,k0) (mid,k0) (hi,k1) (mid,k1).
//   warp 1      MMA issuer: per hi unit 4 x (A_hi B_hi, A_hi B_mid), per mid unit 4 x (A_mid B_hi): 24 MMAs of
//               M = 128 pixels x N = NB x K = 16 per tile; two TMEM accumulator buffers so the epilogue of tile i
//               overlaps the MMAs of tile i+1.
//   warp 2      per-image weight loader (B operand pair, K-major, SWIZZLE_128B, 4 x NB x 128 B; re-loaded when the CTA
//               crosses an image boundary) + TMEM allocation / release.
//   warps 3-10  epilogue, 8 warps per tile (4 TMEM lane quarters x 2 column halves): tcgen05.ld 32 columns at a time
//               (thread = pixel row; the next chunk is prefetched while the current one is reduced), then
//                 EPI_STORE   : write y[b, n, p]                      (stand-alone range attention, N = 128)
//                 EPI_SOFTMAX : online softmax over the NB = 256 logits and centre-weighted sum -> pred[b, p]
//                 EPI_BWD     : d loss / d logit from the forward's softmax state (training backward)
//               so in the fused form neither the range-attention maps (29 MB/img) nor the logits / softmax
//               (58 MB/img each) ever reach HBM: algorithmic traffic is 128*P*4 B in + P*4 B out per image.
// In the fused form the two 1x1 contractions are folded by associativity, W'_b = W_out @ Q_b (tiny, exact fp32,
// mde_fold_queries), pre-scaled by log2(e) and split once, so the tensor cores run ONE K = 128 contraction per pixel
// against the per-image 256 x 128 operand; the bias enters the softmax as exp2(b_j) factors.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace mde {
namespace tc {

constexpr int TILE_M = 128;                      // pixels per tile (UMMA M)
constexpr int KDIM = 128;                        // contraction length (channels)
constexpr int KC = 64;                           // channels per unit = one 128-byte swizzle row of bf16
constexpr int UNIT_BYTES = TILE_M * KC * 2;      // 16384
constexpr int UNITS_PER_TILE = 2 * (KDIM / KC);  // planes x K-chunks
constexpr int EPI_WARP0 = 3;                     // warps 0..2: TMA producer, MMA issuer, weight loader / TMEM allocator
constexpr int ES = 2;                            // column parts per lane quarter: 8 epilogue warps, NB / 2 columns each
constexpr int CHAIN_THREADS = 32 * EPI_WARP0 + 128 * ES;
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
  int single;  // != 0: the hi planes only, one bf16 product per K step (the 2e-2 bf16 mode): mid units are neither loaded nor multiplied
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

// optional [grid][8] cycle counters: where each role waits (tuning only; compiled in for the PROF instantiation alone)
static long long* g_prof = nullptr;

template <bool PROF>
struct WaitClock {
  long long acc = 0;
  __device__ __forceinline__ long long begin() const { return PROF ? clock64() : 0; }
  __device__ __forceinline__ void end(long long t0) {
    if (PROF) acc += clock64() - t0;
  }
};

template <int NB>
struct SmemPlan {
  static constexpr int W_UNIT = NB * KC * 2;          // one (plane, K-chunk) block of the per-image operand
  static constexpr int W_BYTES = UNITS_PER_TILE * W_UNIT;
  static constexpr int NS = (NB == 256) ? 5 : 8;
  static constexpr int RING_BYTES = NS * UNIT_BYTES;
  static constexpr int CONST_BYTES = 2 * NB * 4 + 128 * 4 * 4;  // exp2(bias), exp2(bias)*centre per bin; merge slots
  static constexpr int BAR_BYTES = 512;
  static constexpr int TOTAL = W_BYTES + RING_BYTES + CONST_BYTES + BAR_BYTES + 1024;  // + alignment slack
};

template <int NB, int EPI, bool PROF>
__global__ void __launch_bounds__(CHAIN_THREADS, 1)
    head_chain_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                      const float* __restrict__ biasf, const float* __restrict__ centers, float* __restrict__ out,
                      int tiles_per_img, int total_tiles, long long P, long long* prof, ChainTrain tr) {
  using Plan = SmemPlan<NB>;
  constexpr int NS = Plan::NS;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* gbase = smem_dyn + (base - smem_u32(smem_dyn));
  const uint32_t s_w = base;
  const uint32_t s_ring = s_w + Plan::W_BYTES;
  float* c_all = reinterpret_cast<float*>(gbase + Plan::W_BYTES + Plan::RING_BYTES);  // [2][NB] + merge slots
  const uint32_t s_bar = s_ring + Plan::RING_BYTES + Plan::CONST_BYTES;
  const uint32_t bar_full = s_bar;                 // [NS]
  const uint32_t bar_empty = s_bar + 8 * NS;       // [NS]
  const uint32_t bar_wfull = s_bar + 16 * NS;      // [1]
  const uint32_t bar_wempty = bar_wfull + 8;       // [1]
  const uint32_t bar_accfull = bar_wempty + 8;     // [2]
  const uint32_t bar_accempty = bar_accfull + 16;  // [2]
  const uint32_t s_tmem_slot = bar_accempty + 16;  // uint32
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + (s_tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // contiguous range of tiles of this CTA, so that it crosses at most one image boundary
  const int t_begin = (int)(((long long)total_tiles * blockIdx.x) / gridDim.x);
  const int t_end = (int)(((long long)total_tiles * (blockIdx.x + 1)) / gridDim.x);

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    mbar_init(bar_wfull, 1);
    mbar_init(bar_wempty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_accfull + 8 * i, 1);
      mbar_init(bar_accempty + 8 * i, 4 * ES);  // one arrive per epilogue warp
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
  }
  constexpr uint32_t TMEM_COLS = (2 * NB <= 256) ? 256 : 512;
  if (warp == 2) {
    tmem_alloc(s_tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();  // the prologue above overlaps the previous kernel of the stream (common.cuh); global memory from here on

  if (warp == 0) {
    // ================= activation producer =================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      WaitClock<PROF> w_empty;
      const long long t_start = PROF ? clock64() : 0;
      for (int t = t_begin; t < t_end; ++t) {
        const int img = t / tiles_per_img;
        const int p0 = (t - img * tiles_per_img) * TILE_M;
        const int row0 = (int)((long long)img * P + p0);
#pragma unroll
        for (int u = 0; u < UNITS_PER_TILE; ++u) {  // u = 2 * kc + plane
          if (tr.single && (u & 1)) continue;
          const long long c0 = w_empty.begin();
          mbar_wait(bar_empty + 8 * stage, phase ^ 1, 1);
          w_empty.end(c0);
          mbar_expect_tx(bar_full + 8 * stage, UNIT_BYTES);
          tma_load_3d(s_ring + stage * UNIT_BYTES, &map_x, bar_full + 8 * stage, (u >> 1) * KC, row0, u & 1);
          if (++stage == NS) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      if (PROF && prof) {
        prof[blockIdx.x * 8 + 0] = w_empty.acc;
        prof[blockIdx.x * 8 + 1] = clock64() - t_start;
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
        mbar_expect_tx(bar_wfull, tr.single ? Plan::W_BYTES / 2 : Plan::W_BYTES);
#pragma unroll
        for (int u = 0; u < UNITS_PER_TILE; ++u)  // smem block u = 2 * kc + plane
          if (!(tr.single && (u & 1))) tma_load_4d(s_w + u * Plan::W_UNIT, &map_w, bar_wfull, (u >> 1) * KC, 0, img, u & 1);
        wphase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(FMT_BF16, TILE_M, NB, 0, 0);
      uint32_t stage = 0, phase = 0, wphase = 0;
      int cur = -1;
      int it = 0;
      WaitClock<PROF> w_full, w_acc, w_w;
      const long long t_start = PROF ? clock64() : 0;
      for (int t = t_begin; t < t_end; ++t, ++it) {
        const int img = t / tiles_per_img;
        if (img != cur) {
          if (cur >= 0) umma_commit(bar_wempty);  // fires when every MMA that read the old weights is done
          const long long c0 = w_w.begin();
          mbar_wait(bar_wfull, wphase, 3);
          w_w.end(c0);
          wphase ^= 1;
          cur = img;
        }
        const uint32_t buf = it & 1;
        {
          const long long c0 = w_acc.begin();
          mbar_wait(bar_accempty + 8 * buf, ((it >> 1) & 1) ^ 1, 4);
          w_acc.end(c0);
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * NB;
#pragma unroll
        for (int u = 0; u < UNITS_PER_TILE; ++u) {
          if (tr.single && (u & 1)) continue;
          {
            const long long c0 = w_full.begin();
            mbar_wait(bar_full + 8 * stage, phase, 5);
            w_full.end(c0);
          }
          tc_fence_after();
          const uint32_t a_base = s_ring + stage * UNIT_BYTES;
          const uint32_t b_hi = s_w + (u & ~1) * Plan::W_UNIT;  // (kc, plane 0)
          const uint32_t b_mid = b_hi + Plan::W_UNIT;           // (kc, plane 1)
#pragma unroll
          for (int j = 0; j < KC / 16; ++j) {
            // K-major SWIZZLE_128B operands: 16 bf16 = 32 B along the 128-byte swizzle row; 8-row atoms 1024 B apart
            const uint64_t adesc = make_smem_desc(a_base + j * 32, 16, 1024, SWZ_128B);
            const uint64_t bdesc = make_smem_desc(b_hi + j * 32, 16, 1024, SWZ_128B);
            umma_f16_ss(d_tmem, adesc, bdesc, idesc, (u | j) != 0);   // A_hi B_hi (u even) / A_mid B_hi (u odd)
            if ((u & 1) == 0 && !tr.single)
              umma_f16_ss(d_tmem, adesc, make_smem_desc(b_mid + j * 32, 16, 1024, SWZ_128B), idesc, 1);  // A_hi B_mid
          }
          umma_commit(bar_empty + 8 * stage);  // frees the unit when these MMAs complete
          if (++stage == NS) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(bar_accfull + 8 * buf);
      }
      if (PROF && prof) {
        prof[blockIdx.x * 8 + 2] = w_full.acc;
        prof[blockIdx.x * 8 + 3] = w_acc.acc;
        prof[blockIdx.x * 8 + 4] = w_w.acc;
        prof[blockIdx.x * 8 + 5] = clock64() - t_start;
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
    auto acc_release = [&](uint32_t bar) { mbar_arrive(bar); };  // hand an accumulator buffer back to the MMA issuer
    float acc_gb[4] = {0.f, 0.f, 0.f, 0.f}, acc_gc[4] = {0.f, 0.f, 0.f, 0.f};  // EPI_BWD: per-lane bin sums
    int cur = -1;
    int it = 0;
    WaitClock<PROF> w_accfull;
    const long long t_start = PROF ? clock64() : 0;
    for (int t = t_begin; t < t_end; ++t, ++it) {
      const int img = t / tiles_per_img;
      const int p0 = (t - img * tiles_per_img) * TILE_M;
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
      {
        const long long c0 = w_accfull.begin();
        mbar_wait(bar_accfull + 8 * buf, (it >> 1) & 1, 6);
        w_accfull.end(c0);
      }
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
        static_assert(COLS == 128, "the unrolled epilogue covers 4 chunks of 32 columns per warp");
        MDE_CHUNK(ra, rb, 0, false)
        MDE_CHUNK(rb, ra, 32, false)
        MDE_CHUNK(ra, rb, 64, false)
        MDE_CHUNK(rb, ra, 96, true)
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
    if (PROF && prof && etid == 0) {
      prof[blockIdx.x * 8 + 6] = w_accfull.acc;
      prof[blockIdx.x * 8 + 7] = clock64() - t_start;
    }
  }
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// x_pair: planes[2][B*P][128] bf16 (NHWC pair); w_pair: planes[2][B][NB][128] bf16
template <int NB, int EPI, bool PROF = false>
static int launch_chain(const uint16_t* x_pair, const uint16_t* w_pair, const float* biasf, const float* centers, float* out,
                        int B, long long P, cudaStream_t st, ChainTrain tr = ChainTrain{}) {
  using Plan = SmemPlan<NB>;
  if (P % TILE_M != 0) return MDE_ERR_BAD_SHAPE;
  if (!aligned(x_pair, 16) || !aligned(w_pair, 16)) return MDE_ERR_BAD_POINTER;
  if ((long long)B * P > 0x7fffffffLL) return MDE_ERR_BAD_SHAPE;
  CUtensorMap mx, mw;
  {
    const uint64_t rows = (uint64_t)B * (uint64_t)P;
    const uint64_t dims[3] = {(uint64_t)KDIM, rows, 2};
    const uint64_t strides[2] = {(uint64_t)KDIM * 2, rows * KDIM * 2};
    const uint32_t box[3] = {KC, TILE_M, 1};
    if (!encode_bf16(&mx, x_pair, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return MDE_ERR_DRIVER;
  }
  {
    const uint64_t dims[4] = {(uint64_t)KDIM, (uint64_t)NB, (uint64_t)B, 2};
    const uint64_t strides[3] = {(uint64_t)KDIM * 2, (uint64_t)NB * KDIM * 2, (uint64_t)B * NB * KDIM * 2};
    const uint32_t box[4] = {KC, (uint32_t)NB, 1, 1};
    if (!encode_bf16(&mw, w_pair, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return MDE_ERR_DRIVER;
  }
  const int tiles_per_img = (int)(P / TILE_M);
  const long long total = (long long)tiles_per_img * B;
  if (total > 0x7fffffffLL) return MDE_ERR_BAD_SHAPE;
  const int grid = (int)(total < MDE_NUM_SMS ? total : MDE_NUM_SMS);
  static bool attr_set = false;  // per instantiation; one process drives one device (common.cuh)
  if (!attr_set) {
    if (cudaFuncSetAttribute(head_chain_kernel<NB, EPI, PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize, Plan::TOTAL) !=
        cudaSuccess)
      return MDE_ERR_LAUNCH;
    attr_set = true;
  }
  launch_pdl(PDL_TC, head_chain_kernel<NB, EPI, PROF>, dim3(grid), dim3(CHAIN_THREADS), Plan::TOTAL, st, mx, mw, biasf, centers, out,
             tiles_per_img, (int)total, P, g_prof, tr);
  return check_launch();
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

// the bf16 mode of the fused head (north star: depth within 2e-2 of the fp32 reference): same operands, hi planes only
int mde_head_chain_bf16_fwd(const uint16_t* x_pair, const uint16_t* w_pair, const float* biasf, const float* centers,
                            float* pred, int B, int n_bins, int64_t P, mde_stream_t stream) {
  if (!x_pair || !w_pair || !biasf || !centers || !pred) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || P <= 0) return MDE_ERR_BAD_SHAPE;
  if (n_bins != 256) return MDE_ERR_UNSUPPORTED;
  tc::ChainTrain tr{};
  tr.single = 1;
  return tc::launch_chain<256, tc::EPI_SOFTMAX>(x_pair, w_pair, biasf, centers, pred, B, P, (cudaStream_t)stream, tr);
}

int mde_head_chain_fwd(const uint16_t* x_pair, const uint16_t* w_pair, const float* biasf, const float* centers, float* pred,
                       int B, int n_bins, int64_t P, mde_stream_t stream) {
  if (!x_pair || !w_pair || !biasf || !centers || !pred) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || P <= 0) return MDE_ERR_BAD_SHAPE;
  if (n_bins != 256) return MDE_ERR_UNSUPPORTED;
  if (tc::g_prof) return tc::launch_chain<256, tc::EPI_SOFTMAX, true>(x_pair, w_pair, biasf, centers, pred, B, P, (cudaStream_t)stream);
  return tc::launch_chain<256, tc::EPI_SOFTMAX>(x_pair, w_pair, biasf, centers, pred, B, P, (cudaStream_t)stream);
}

// training forward: as mde_head_chain_fwd, also recording the per-pixel softmax state stats [B,P,2] for the backward
int mde_head_chain_fwd_train(const uint16_t* x_pair, const uint16_t* w_pair, const float* biasf, const float* centers,
                             float* pred, float* stats, int B, int n_bins, int64_t P, mde_stream_t stream) {
  if (!x_pair || !w_pair || !biasf || !centers || !pred || !stats) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || P <= 0) return MDE_ERR_BAD_SHAPE;
  if (n_bins != 256) return MDE_ERR_UNSUPPORTED;
  tc::ChainTrain tr{};
  tr.stats = stats;
  return tc::launch_chain<256, tc::EPI_SOFTMAX>(x_pair, w_pair, biasf, centers, pred, B, P, (cudaStream_t)stream, tr);
}

// backward, step 1: recompute the logits on the tensor cores and emit d loss / d logit in both layouts + per-bin sums
int mde_head_chain_bwd_logits(const uint16_t* x_pair, const uint16_t* w_pair, const float* biasf, const float* centers,
                              const float* pred, const float* stats, const float* gpred, float* gl, float* glT, float* gc,
                              float* gb, int B, int n_bins, int64_t P, mde_stream_t stream) {
  if (!x_pair || !w_pair || !biasf || !centers || !pred || !stats || !gpred || !gl || !glT || !gc || !gb)
    return MDE_ERR_BAD_POINTER;
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
  return tc::launch_chain<256, tc::EPI_BWD>(x_pair, w_pair, biasf, centers, nullptr, B, P, st, tr);
}

// stand-alone range attention on the tensor cores: y[b, n, p] = sum_k x[b, p, k] q[b, n, k], operands as split-bf16 pairs
int mde_range_attention_tc(const uint16_t* x_pair, const uint16_t* q_pair, float* y, int B, int K, int N, int64_t P,
                           mde_stream_t stream) {
  if (!x_pair || !q_pair || !y) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || P <= 0) return MDE_ERR_BAD_SHAPE;
  if (K != 128 || N != 128) return MDE_ERR_UNSUPPORTED;
  return tc::launch_chain<128, tc::EPI_STORE>(x_pair, q_pair, nullptr, nullptr, y, B, P, (cudaStream_t)stream);
}

int mde_round_tf32(const float* in, float* out, int64_t n, float scale, mde_stream_t stream) {
  if (!in || !out) return MDE_ERR_BAD_POINTER;
  if (n <= 0) return n == 0 ? MDE_OK : MDE_ERR_BAD_SHAPE;
  long long g = (n + 255) / 256;
  if (g > MDE_NUM_SMS * 8) g = MDE_NUM_SMS * 8;
  tc::round_tf32_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(in, out, n, scale);
  return check_launch();
}

// device buffer of [grid][8] int64 cycle counters filled by the next mde_head_chain_fwd launches (NULL disables; the
// instrumented instantiation is a separate kernel, the default build carries no clock reads): per CTA
// {producer wait-empty, producer total, mma wait-full, mma wait-acc-empty, mma wait-weights, mma total,
//  epilogue wait-acc-full, epilogue total}
int mde_tc_debug_profile(long long* buf) {
  tc::g_prof = buf;
  return MDE_OK;
}

// last barrier-timeout code recorded by a tcgen05 kernel of this file (0 = none). Synchronises the device.
int mde_tc_last_error(void) {
  int v = 0;
  cudaMemcpyFromSymbol(&v, tc::g_tc_error, sizeof(int));
  return v;
}

}  // extern "C"
