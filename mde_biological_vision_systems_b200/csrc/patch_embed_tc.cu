// K1a -- patch-embedding convolution (kernel = stride = patch) as a TMA-fed tcgen05 split-K GEMM.
//
// Reference: PatchTransformerEncoder.forward (models/layers.py:16-19):
//   emb = Conv2d(C, E, kernel_size=p, stride=p)(x).flatten(2) + positional_encodings[:S].T ; tokens = emb.permute(2,0,1)
// With non-overlapping patches the conv is a GEMM  tokens[(b,py,px), e] = sum_{i,j,c} x[b, p*py+i, p*px+j, c] * W[e,i,j,c]
// (M = B*S, N = E = 128, K = p*p*C = 32768 for p = 16, C = 128): 1.85 GFLOP per image over a 29 MB activation read, i.e.
// HBM/L2-bound.  Input is the channels_last (NHWC) feature map as a split-bf16 pair (hi, mid planes, tc_common.cuh), so
// for a fixed kernel row i the 64-element K-chunk of every patch of a token row is one 128-byte line and a 5-D TMA box
// {64 bf16, w/p patches, 1 kernel row, PYT token rows, 1 image} lands a whole K-major, 128B-swizzled A tile per plane; the
// filter in NHWC order [E][i][j][c] (pair) is the K-major B operand as it lies in memory.  Every K step issues the three
// products hi*hi + mid*hi + hi*mid (kind::f16, fp32 accumulation): fp32-grade tokens (the K = 32768 sums feed the bin
// widths and the queries, where a TF32 pass costs 1e-4 .. 1e-3 of `pred`).  One CTA = one image x one K split: NT
// accumulator tiles (PYT token rows each) share every B stage, accumulate in TMEM, and the split-K partials are reduced
// (with bias + positional rows) by a second kernel into the reference's [S, B, E] token layout -- deterministic, no atomics.
#include "common.cuh"
#include "tc_common.cuh"

namespace mde {
namespace tc {

constexpr int PE_THREADS = 224;  // warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warps 3-6 epilogue
constexpr int PE_KC = 64;        // bf16 elements per K-chunk (one 128-byte swizzle row)
constexpr int PE_TILE_BYTES = 128 * 128;  // one operand tile per stage and plane: 128 rows x 128 B

template <int NT>
__global__ void __launch_bounds__(PE_THREADS, 1)
    patch_embed_kernel(const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xm,
                       const __grid_constant__ CUtensorMap map_w, float* __restrict__ part, int B, int hp, int wp, int pyt,
                       int chunks_per_row, int chunks_per_split, int nstages) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* gbase = smem_dyn + (base - smem_u32(smem_dyn));
  constexpr int STAGE = 2 * (NT + 1) * PE_TILE_BYTES;  // [A_hi x NT][A_mid x NT][B_hi][B_mid]
  const uint32_t s_bar = base + nstages * STAGE;
  const uint32_t bar_full = s_bar, bar_empty = s_bar + 8 * nstages, bar_acc = s_bar + 16 * nstages;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + nstages * STAGE + 16 * nstages + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x, split = blockIdx.y;
  const int g0 = split * chunks_per_split;

  if (threadIdx.x == 0) {
    for (int i = 0; i < nstages; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    mbar_init(bar_acc, 1);
    fence_barrier_init();
    fence_proxy_async();
    tma_prefetch_desc(&map_xh);
    tma_prefetch_desc(&map_xm);
    tma_prefetch_desc(&map_w);
  }
  constexpr uint32_t TMEM_COLS = NT * 128 <= 128 ? 128 : (NT * 128 <= 256 ? 256 : 512);
  if (warp == 2) {
    tmem_alloc(smem_u32((const void*)tmem_slot), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();  // the prologue above overlaps the previous kernel of the stream (common.cuh); global memory from here on
  const uint32_t box_bytes = (uint32_t)(PE_KC * 2 * wp * pyt);  // OOB rows are zero-filled but still counted

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int c = 0; c < chunks_per_split; ++c) {
        const int g = g0 + c;
        const int i = g / chunks_per_row, jc0 = (g - i * chunks_per_row) * PE_KC;
        mbar_wait(bar_empty + 8 * stage, phase ^ 1, 11);
        mbar_expect_tx(bar_full + 8 * stage, 2 * NT * box_bytes + 2 * PE_TILE_BYTES);
        const uint32_t dst = base + stage * STAGE;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          tma_load_5d(dst + t * PE_TILE_BYTES, &map_xh, bar_full + 8 * stage, jc0, 0, i, t * pyt, b);
          tma_load_5d(dst + (NT + t) * PE_TILE_BYTES, &map_xm, bar_full + 8 * stage, jc0, 0, i, t * pyt, b);
        }
        tma_load_3d(dst + 2 * NT * PE_TILE_BYTES, &map_w, bar_full + 8 * stage, g * PE_KC, 0, 0);
        tma_load_3d(dst + (2 * NT + 1) * PE_TILE_BYTES, &map_w, bar_full + 8 * stage, g * PE_KC, 0, 1);
        if (++stage == (uint32_t)nstages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(FMT_BF16, 128, 128, 0, 0);
      uint32_t stage = 0, phase = 0;
      for (int c = 0; c < chunks_per_split; ++c) {
        mbar_wait(bar_full + 8 * stage, phase, 12);
        tc_fence_after();
        const uint32_t a0 = base + stage * STAGE;
        const uint32_t b0 = a0 + 2 * NT * PE_TILE_BYTES;
#pragma unroll
        for (int j = 0; j < PE_KC / 16; ++j) {
          const uint64_t bhi = make_smem_desc(b0 + j * 32, 16, 1024, SWZ_128B);
          const uint64_t bmid = make_smem_desc(b0 + PE_TILE_BYTES + j * 32, 16, 1024, SWZ_128B);
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            const uint64_t ahi = make_smem_desc(a0 + t * PE_TILE_BYTES + j * 32, 16, 1024, SWZ_128B);
            const uint64_t amid = make_smem_desc(a0 + (NT + t) * PE_TILE_BYTES + j * 32, 16, 1024, SWZ_128B);
            umma_f16_ss(tmem_base + t * 128, ahi, bhi, idesc, (c | j) != 0);
            umma_f16_ss(tmem_base + t * 128, amid, bhi, idesc, 1);
            umma_f16_ss(tmem_base + t * 128, ahi, bmid, idesc, 1);
          }
        }
        umma_commit(bar_empty + 8 * stage);
        if (++stage == (uint32_t)nstages) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(bar_acc);
    }
  } else if (warp >= 3) {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;  // accumulator row = local patch index py_local * wp + px
    mbar_wait(bar_acc, 0, 13);
    tc_fence_after();
    const int pyl = row / wp, px = row - pyl * wp;
    const long long S = (long long)hp * wp;
#pragma unroll 1
    for (int t = 0; t < NT; ++t) {
      const int py = t * pyt + pyl;
      const bool valid = (pyl < pyt) && (py < hp);
      float* dst = part + (((long long)split * S + ((long long)py * wp + px)) * B + b) * 128;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + t * 128 + c0 + ((uint32_t)(quarter * 32) << 16), r);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(dst + c0 + i) = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                                                   __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// tokens[s, b, e] = bias[e] + pos[s, e] + sum_split part[split][s][b][e]
__global__ void __launch_bounds__(256) patch_reduce_kernel(const float* __restrict__ part, const float* __restrict__ bias,
                                                           const float* __restrict__ pos, float* __restrict__ tokens,
                                                           long long SB, int B, int splits) {
  const long long idx4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // float4 index over [S*B][128]
  pdl_sync();
  if (idx4 >= SB * 32) return;
  const long long rowi = idx4 >> 5;
  const int e4 = (int)(idx4 & 31);
  const long long s = rowi / B;
  float4 acc = reinterpret_cast<const float4*>(bias)[e4];
  const float4 p = reinterpret_cast<const float4*>(pos + s * 128)[e4];
  acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
  for (int k = 0; k < splits; ++k) {
    const float4 v = reinterpret_cast<const float4*>(part + (long long)k * SB * 128)[idx4];
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  reinterpret_cast<float4*>(tokens)[idx4] = acc;
}

static int pick_splits(int B, int total_chunks) {
  int s = 1;
  while (s * 2 * B <= MDE_NUM_SMS && s * 2 <= 64 && total_chunks % (s * 2) == 0) s *= 2;
  return s;
}

}  // namespace tc
}  // namespace mde

using namespace mde;

extern "C" {

int64_t mde_patch_embed_ws_floats(int B, int h, int w, int patch, int C) {
  if (B <= 0 || patch <= 0 || h < patch || w < patch) return 0;
  const int hp = h / patch, wp = w / patch;
  const int chunks = patch * patch * C / tc::PE_KC;
  return (int64_t)tc::pick_splits(B, chunks) * hp * wp * B * 128;
}

int mde_patch_embed_fwd(const uint16_t* x_pair, const uint16_t* w_pair, const float* bias, const float* pos, float* tokens,
                        float* ws, int B, int h, int w, int C, int patch, int E, mde_stream_t stream) {
  const uint16_t* x_nhwc = x_pair;
  const uint16_t* w_nhwc = w_pair;
  if (!x_nhwc || !w_nhwc || !bias || !pos || !tokens || !ws) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || B > 65535 || E != 128 || C <= 0 || patch <= 0 || h < patch || w < patch) return MDE_ERR_BAD_SHAPE;
  const int hp = h / patch, wp = w / patch;
  const int row_floats = patch * C;  // elements of one kernel row of one patch (contiguous in NHWC)
  if (row_floats % tc::PE_KC != 0 || wp > 128 || !aligned(x_nhwc, 16) || !aligned(w_nhwc, 16)) return MDE_ERR_UNSUPPORTED;
  const int pyt = 128 / wp;                   // token rows per accumulator tile
  const int nt = (hp + pyt - 1) / pyt;        // accumulator tiles per image
  if (nt > 4 || C % 8 != 0) return MDE_ERR_UNSUPPORTED;
  const int chunks_per_row = row_floats / tc::PE_KC;
  const int total_chunks = chunks_per_row * patch;
  const int splits = tc::pick_splits(B, total_chunks);
  const int chunks_per_split = total_chunks / splits;
  cudaStream_t st = (cudaStream_t)stream;

  CUtensorMap mxh, mxm, mw;
  {
    // x[b, y = patch*py + i, x = patch*px + j, c]  ->  dims {jc, px, i, py, b}; one map per plane of the pair (rank 5 is TMA's max)
    const uint64_t dims[5] = {(uint64_t)row_floats, (uint64_t)wp, (uint64_t)patch, (uint64_t)hp, (uint64_t)B};
    const uint64_t strides[4] = {(uint64_t)row_floats * 2, (uint64_t)w * C * 2, (uint64_t)patch * w * C * 2,
                                 (uint64_t)h * w * C * 2};
    const uint32_t box[5] = {(uint32_t)tc::PE_KC, (uint32_t)wp, 1, (uint32_t)pyt, 1};
    const uint16_t* mid = x_nhwc + (size_t)B * h * w * C;
    if (!tc::encode_bf16(&mxh, x_nhwc, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return MDE_ERR_DRIVER;
    if (!tc::encode_bf16(&mxm, mid, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return MDE_ERR_DRIVER;
  }
  {
    const uint64_t K = (uint64_t)total_chunks * tc::PE_KC;
    const uint64_t dims[3] = {K, (uint64_t)E, 2};
    const uint64_t strides[2] = {K * 2, (uint64_t)E * K * 2};
    const uint32_t box[3] = {(uint32_t)tc::PE_KC, 128, 1};
    if (!tc::encode_bf16(&mw, w_nhwc, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return MDE_ERR_DRIVER;
  }
  const int stage_bytes = 2 * (nt + 1) * tc::PE_TILE_BYTES;
  int nstages = (200 * 1024) / stage_bytes;
  if (nstages > 8) nstages = 8;
  if (nstages > chunks_per_split) nstages = chunks_per_split;
  if (nstages < 1) return MDE_ERR_BAD_SHAPE;
  const int smem = nstages * stage_bytes + 16 * nstages + 64 + 1024;
  dim3 grid((unsigned)B, (unsigned)splits);
#define MDE_PE_LAUNCH(NT)                                                                                               \
  {                                                                                                                     \
    static bool attr = false;                                                                                           \
    if (!attr) {                                                                                                        \
      if (cudaFuncSetAttribute(tc::patch_embed_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) !=  \
          cudaSuccess)                                                                                                  \
        return MDE_ERR_LAUNCH;                                                                                          \
      attr = true;                                                                                                      \
    }                                                                                                                   \
    launch_pdl(PDL_TC, tc::patch_embed_kernel<NT>, grid, dim3(tc::PE_THREADS), smem, st, mxh, mxm, mw, ws, B, hp, wp, pyt,       \
               chunks_per_row, chunks_per_split, nstages);                                                             \
  }
  switch (nt) {
    case 1: MDE_PE_LAUNCH(1) break;
    case 2: MDE_PE_LAUNCH(2) break;
    case 3: MDE_PE_LAUNCH(3) break;
    default: MDE_PE_LAUNCH(4) break;
  }
#undef MDE_PE_LAUNCH
  int rc = check_launch();
  if (rc) return rc;
  const long long SB = (long long)hp * wp * B;
  launch_pdl(PDL_CHAIN, tc::patch_reduce_kernel, dim3((unsigned)((SB * 32 + 255) / 256)), dim3(256), 0, st, ws, bias, pos, tokens, SB, B, splits);
  return check_launch();
}

}  // extern "C"
