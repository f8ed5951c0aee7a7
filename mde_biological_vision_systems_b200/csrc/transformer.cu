// K1b -- PatchTransformerEncoder layers (post-LN nn.TransformerEncoderLayer, d_model 128, 4 heads, FF 1024, ReLU).
//
// Reference: models/layers.py:8-9,23 (nn.TransformerEncoder of 4 default nn.TransformerEncoderLayer), eval semantics
// (dropout is the identity).  Token layout is the reference's [S, N, E]: row index = s * N + n.
// The sequence is short (S = 13*17 = 221 tokens/image) and the whole encoder is 0.68 GFLOP/image (3 % of the head): the layer is
// latency-bound, not throughput-bound.  Two forms:
//   mde_encoder_layer_tc_fwd (default): the four nn.Linear products on the tcgen05 NT GEMM (gemm_tc.cu) in 3xTF32 on split-form
//                         operands (fp32-grade: the queries feed a softmax downstream, DESIGN.md "precision"), attention and
//                         LayerNorm on the kernels below;
//   mde_encoder_layer_fwd: everything exact fp32 SIMT (linear_kernel: 64x64x16 shared-memory tiles, 4x4 register blocking).
//   attention_kernel      softmax(Q K^T / sqrt(hd)) V   one CTA per (image, head, row range), K and V resident in shared memory,
//                                                    eight query rows per warp pass (see the kernel's comment)
//   add_layernorm128_kernel  LN(x + y) * g + b       one warp per token row (E = 128 -> float4 per lane), sums split-K planes
#include "common.cuh"

namespace mde {

// C[m, n] = act( sum_k A[m*lda + k] * W[n*ldw + k] + bias[n] )
template <int ACT>  // 0 none, 1 relu, 2 leaky relu (slope 0.01)
__global__ void __launch_bounds__(256) linear_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W,
                                                     int ldw, const float* __restrict__ bias, float* __restrict__ C,
                                                     int ldc, int M, int N, int K) {
  __shared__ float sa[16][64 + 4];
  __shared__ float sw[16][64 + 4];
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, each 4 (m) x 4 (n)
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int lr = threadIdx.x >> 2;        // 0..63 tile row
  const int lk = (threadIdx.x & 3) * 4;   // 0,4,8,12
  const bool vec = (K % 4 == 0) && (lda % 4 == 0) && (ldw % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
  for (int k0 = 0; k0 < K; k0 += 16) {
    float va[4] = {0.f, 0.f, 0.f, 0.f}, vw[4] = {0.f, 0.f, 0.f, 0.f};
    const int gm = m0 + lr, gn = n0 + lr, gk = k0 + lk;
    if (gm < M) {
      if (vec && gk + 3 < K) {
        const float4 t = *reinterpret_cast<const float4*>(A + (long long)gm * lda + gk);
        va[0] = t.x; va[1] = t.y; va[2] = t.z; va[3] = t.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (gk + i < K) va[i] = A[(long long)gm * lda + gk + i];
      }
    }
    if (gn < N) {
      if (vec && gk + 3 < K) {
        const float4 t = *reinterpret_cast<const float4*>(W + (long long)gn * ldw + gk);
        vw[0] = t.x; vw[1] = t.y; vw[2] = t.z; vw[3] = t.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (gk + i < K) vw[i] = W[(long long)gn * ldw + gk + i];
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      sa[lk + i][lr] = va[i];
      sw[lk + i][lr] = vw[i];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&sa[kk][ty * 4]);
      const float4 w = *reinterpret_cast<const float4*>(&sw[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (ACT == 1) v = fmaxf(v, 0.f);
      if (ACT == 2) v = v > 0.f ? v : 0.01f * v;
      C[(long long)m * ldc + n] = v;
    }
  }
}

// qkv [S*NB, 3E] (row = s*NB + n); out rows have pitch out_ld.  grid (NB*heads, y splits over the query rows), 4 / 8 / 16 warps.
// One CTA per (image, head, row range): K ([S][HD+1], conflict-free for lane == key) and V ([S][HD]) of the (image, head) are
// resident in shared memory; a warp owns EIGHT query rows per pass.  The kernel lives on the shared-memory pipe (one wavefront
// per clock and SM), so every phase is shaped to feed >= 4 FMAs per shared-memory instruction:
//   scores : lane = key, TWO key blocks (j, j + 32) per pass; per head-dim channel two 16-byte broadcasts of the eight rows' q
//            and two K words feed 16 FMAs;
//   softmax: probabilities kept interleaved [j][8 rows] in shared memory (un-normalised; 1 / sum is applied to the output);
//   P V    : a half-warp owns four of the rows, lane = a PAIR of head-dim channels: one 16-byte broadcast of the four rows' p_j
//            and one 8-byte V load feed 8 FMAs.
// (History: one row per warp -- 666 shared-memory loads per row; four rows with q hoisted into registers -- 162 registers, one
// 256-thread CTA per SM, and a grid of 320 CTAs = three waves on 148 SMs: 42 us per layer at B = 16 for 0.4 GFLOP.)
// split3 != 0 writes each output row as [v | v - trunc_tf32(v) | v] (the 3xTF32 A operand form).
__device__ __forceinline__ float tf32_lo(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

constexpr int ATT_ROWS = 8;  // query rows per warp pass

__host__ __device__ inline size_t attention_kv_floats(int S, int HD) { return (((size_t)S * (HD + 1) + 3) & ~(size_t)3) + (size_t)S * HD; }
__host__ __device__ inline size_t attention_warp_floats(int S, int HD) { return (size_t)ATT_ROWS * (HD + S); }

template <int HD>
__global__ void __launch_bounds__(512, 1) attention_kernel(const float* __restrict__ qkv, float* __restrict__ out, int S,
                                                        int NB, int E, int heads, float scale, int out_ld, int split3) {
  static_assert(HD == 32, "lane == head-dim channel when q is staged; a half-warp covers HD as channel pairs");
  extern __shared__ __align__(16) float smem_att[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* sk = smem_att;                                                    // [S][HD+1]
  float* sv = smem_att + (((size_t)S * (HD + 1) + 3) & ~(size_t)3);        // [S][HD]
  float* sq = smem_att + attention_kv_floats(S, HD) + (size_t)warp * attention_warp_floats(S, HD);  // [HD][8]
  float* sp = sq + ATT_ROWS * HD;                                          // [S][8] scores / exponentials of the eight rows
  const int n = blockIdx.x / heads, hh = blockIdx.x % heads;
  const int ld = 3 * E;
  pdl_sync();  // qkv comes from the previous kernel of the stream
  {
    // K and V rows of this (image, head): float4 loads, four rows in flight per thread before the dependent smem stores
    constexpr int V4 = HD / 4;
    const int total = S * V4;
    for (int i0 = threadIdx.x; i0 < total; i0 += 4 * blockDim.x) {
      float4 kv[4], vv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * blockDim.x;
        if (i < total) {
          const int s = i / V4, d4 = i - s * V4;
          const float* row = qkv + ((long long)s * NB + n) * ld + hh * HD + d4 * 4;
          kv[u] = __ldg(reinterpret_cast<const float4*>(row + E));
          vv[u] = __ldg(reinterpret_cast<const float4*>(row + 2 * E));
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * blockDim.x;
        if (i < total) {
          const int s = i / V4, d4 = i - s * V4;
          float* kd = sk + s * (HD + 1) + d4 * 4;
          kd[0] = kv[u].x; kd[1] = kv[u].y; kd[2] = kv[u].z; kd[3] = kv[u].w;
          *reinterpret_cast<float4*>(sv + s * HD + d4 * 4) = vv[u];
        }
      }
    }
  }
  __syncthreads();
  const int rows_per_cta = (S + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(S, r0 + rows_per_cta);
  const int half = lane >> 4, cp = lane & 15;
  for (int i0 = r0 + ATT_ROWS * warp; i0 < r1; i0 += ATT_ROWS * nwarps) {
    // (1) the eight (pre-scaled, as torch does: q / sqrt(hd)) query rows -> sq[d][r]; lane = channel
    {
      float t[ATT_ROWS];
#pragma unroll
      for (int r = 0; r < ATT_ROWS; ++r) {
        const int i = min(i0 + r, S - 1);
        t[r] = __ldg(qkv + ((long long)i * NB + n) * ld + hh * HD + lane) * scale;
      }
      reinterpret_cast<float4*>(sq)[2 * lane] = make_float4(t[0], t[1], t[2], t[3]);
      reinterpret_cast<float4*>(sq)[2 * lane + 1] = make_float4(t[4], t[5], t[6], t[7]);
    }
    __syncwarp();
    // (2) scores, lane = key, two key blocks per pass
    float mx[ATT_ROWS];
#pragma unroll
    for (int r = 0; r < ATT_ROWS; ++r) mx[r] = -INFINITY;
#pragma unroll 1
    for (int j0 = 0; j0 < S; j0 += 64) {
      const int ja = j0 + lane, jb = ja + 32;
      const float* ka = sk + (size_t)min(ja, S - 1) * (HD + 1);
      const float* kb = sk + (size_t)min(jb, S - 1) * (HD + 1);
      float a[ATT_ROWS], b[ATT_ROWS];
#pragma unroll
      for (int r = 0; r < ATT_ROWS; ++r) a[r] = b[r] = 0.f;
      // q is re-read from shared memory (broadcast loads) in every key-block pass: hoisting the 256 loop-invariant words out of
      // this loop is what the compiler would do otherwise -- and spill them.  The empty asm makes the address opaque per pass.
      int opaque = 0;
      asm volatile("" : "+r"(opaque));  // (an offset, not the pointer: the loads must stay ld.shared)
      const float4* sq4 = reinterpret_cast<const float4*>(sq) + opaque;
#pragma unroll
      for (int d = 0; d < HD; ++d) {
        const float4 q0 = sq4[2 * d], q1 = sq4[2 * d + 1];
        const float q[ATT_ROWS] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
        const float kda = ka[d], kdb = kb[d];
#pragma unroll
        for (int r = 0; r < ATT_ROWS; ++r) {
          a[r] = fmaf(q[r], kda, a[r]);
          b[r] = fmaf(q[r], kdb, b[r]);
        }
      }
      if (ja < S) {
        reinterpret_cast<float4*>(sp)[2 * ja] = make_float4(a[0], a[1], a[2], a[3]);
        reinterpret_cast<float4*>(sp)[2 * ja + 1] = make_float4(a[4], a[5], a[6], a[7]);
#pragma unroll
        for (int r = 0; r < ATT_ROWS; ++r) mx[r] = fmaxf(mx[r], a[r]);
      }
      if (jb < S) {
        reinterpret_cast<float4*>(sp)[2 * jb] = make_float4(b[0], b[1], b[2], b[3]);
        reinterpret_cast<float4*>(sp)[2 * jb + 1] = make_float4(b[4], b[5], b[6], b[7]);
#pragma unroll
        for (int r = 0; r < ATT_ROWS; ++r) mx[r] = fmaxf(mx[r], b[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < ATT_ROWS; ++r) mx[r] = warp_max(mx[r]);
    __syncwarp();
    // (3) exponentials and their sums
    float sum[ATT_ROWS];
#pragma unroll
    for (int r = 0; r < ATT_ROWS; ++r) sum[r] = 0.f;
    for (int j = lane; j < S; j += 32) {
      float4 u = reinterpret_cast<float4*>(sp)[2 * j], v = reinterpret_cast<float4*>(sp)[2 * j + 1];
      u.x = expf(u.x - mx[0]); u.y = expf(u.y - mx[1]); u.z = expf(u.z - mx[2]); u.w = expf(u.w - mx[3]);
      v.x = expf(v.x - mx[4]); v.y = expf(v.y - mx[5]); v.z = expf(v.z - mx[6]); v.w = expf(v.w - mx[7]);
      sum[0] += u.x; sum[1] += u.y; sum[2] += u.z; sum[3] += u.w;
      sum[4] += v.x; sum[5] += v.y; sum[6] += v.z; sum[7] += v.w;
      reinterpret_cast<float4*>(sp)[2 * j] = u;
      reinterpret_cast<float4*>(sp)[2 * j + 1] = v;
    }
#pragma unroll
    for (int r = 0; r < ATT_ROWS; ++r) sum[r] = warp_sum(sum[r]);
    __syncwarp();
    // (4) P V: this half-warp's four rows x this lane's two channels
    float o[4][2];
#pragma unroll
    for (int r = 0; r < 4; ++r) o[r][0] = o[r][1] = 0.f;
    const float* pj = sp + 4 * half;
    const float* vj = sv + 2 * cp;
#pragma unroll 4
    for (int j = 0; j < S; ++j) {
      const float4 p4 = *reinterpret_cast<const float4*>(pj + (size_t)j * ATT_ROWS);
      const float2 v2 = *reinterpret_cast<const float2*>(vj + (size_t)j * HD);
      o[0][0] = fmaf(p4.x, v2.x, o[0][0]); o[0][1] = fmaf(p4.x, v2.y, o[0][1]);
      o[1][0] = fmaf(p4.y, v2.x, o[1][0]); o[1][1] = fmaf(p4.y, v2.y, o[1][1]);
      o[2][0] = fmaf(p4.z, v2.x, o[2][0]); o[2][1] = fmaf(p4.z, v2.y, o[2][1]);
      o[3][0] = fmaf(p4.w, v2.x, o[3][0]); o[3][1] = fmaf(p4.w, v2.y, o[3][1]);
    }
    const float den[4] = {half ? sum[4] : sum[0], half ? sum[5] : sum[1], half ? sum[6] : sum[2], half ? sum[7] : sum[3]};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = i0 + 4 * half + r;
      if (i < r1) {
        const float v0 = o[r][0] / den[r], v1 = o[r][1] / den[r];
        float* orow = out + ((long long)i * NB + n) * out_ld + hh * HD + 2 * cp;
        *reinterpret_cast<float2*>(orow) = make_float2(v0, v1);
        if (split3) {
          *reinterpret_cast<float2*>(orow + E) = make_float2(tf32_lo(v0), tf32_lo(v1));
          *reinterpret_cast<float2*>(orow + 2 * E) = make_float2(v0, v1);
        }
      }
    }
    __syncwarp();
  }
}

// Launch geometry.  One CTA is resident per SM (shared memory), so the grid should be ONE wave and every warp should make one
// pass: the query rows of an (image, head) -- groups of ATT_ROWS rows, one per warp pass -- are split over
// ysplit = SMs / (images * heads) CTAs (at least four groups each, so that the K / V load of a CTA is amortised), and the CTA gets
// the smallest of 4 / 8 / 16 warps that covers its groups in one pass and fits the shared memory.  B = 16, S = 221: 28 groups per
// (image, head), ysplit 2, 14 groups per CTA -> 128 CTAs of 16 warps, one pass.  (The first version of this heuristic preferred
// more, narrower CTAs: 128 CTAs of 8 warps making two passes at 11 % of the SM's warp slots -- 28.6 us per layer under ncu.)
static int launch_attention(const float* qkv, float* out, int S, int NB, int E, int heads, int out_ld, int split3,
                            cudaStream_t st) {
  constexpr int HD = 32;
  const size_t limit = 227 * 1024;
  auto smem_of = [&](int nw) { return sizeof(float) * (attention_kv_floats(S, HD) + (size_t)nw * attention_warp_floats(S, HD)); };
  const int groups = (S + ATT_ROWS - 1) / ATT_ROWS;
  int ysplit = MDE_NUM_SMS / (NB * heads);
  if (ysplit > (groups + 3) / 4) ysplit = (groups + 3) / 4;
  if (ysplit < 1) ysplit = 1;
  const int rows_per_cta = (S + ysplit - 1) / ysplit;
  const int groups_per_cta = (rows_per_cta + ATT_ROWS - 1) / ATT_ROWS;
  int nw = groups_per_cta <= 4 ? 4 : groups_per_cta <= 8 ? 8 : 16;
  while (nw > 4 && smem_of(nw) > limit) nw >>= 1;
  if (smem_of(nw) > limit) return MDE_ERR_BAD_SHAPE;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(attention_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit) != cudaSuccess)
      return MDE_ERR_LAUNCH;
    attr = true;
  }
  launch_pdl(PDL_CHAIN, attention_kernel<HD>, dim3(NB * heads, ysplit), dim3(32 * nw), smem_of(nw), st, qkv, out, S, NB, E, heads,
             1.0f / sqrtf((float)HD), out_ld, split3);
  return check_launch();
}

// out[m,:] = LayerNorm(x[m,:] + y[m,:]) * g + b ;  E == 128 (one float4 per lane)
__global__ void __launch_bounds__(256) add_layernorm128_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                               const float* __restrict__ g, const float* __restrict__ b,
                                                               float* __restrict__ out, int M, float eps, int x_ld = 128,
                                                               int y_ld = 128, int out_ld = 128, int split3 = 0,
                                                               int y_planes = 1, long long y_plane_stride = 0) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  pdl_sync();
  if (row >= M) return;
  const float4 a = reinterpret_cast<const float4*>(x + (long long)row * x_ld)[lane];
  float4 c = reinterpret_cast<const float4*>(y + (long long)row * y_ld)[lane];
  for (int pl = 1; pl < y_planes; ++pl) {  // split-K partial products of the producing GEMM
    const float4 t = reinterpret_cast<const float4*>(y + pl * y_plane_stride + (long long)row * y_ld)[lane];
    c.x += t.x; c.y += t.y; c.z += t.z; c.w += t.w;
  }
  float v[4] = {a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w};
  float mean = warp_sum(v[0] + v[1] + v[2] + v[3]) * (1.f / 128.f);
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i] -= mean;
    var = fmaf(v[i], v[i], var);
  }
  var = warp_sum(var) * (1.f / 128.f);
  const float rstd = rsqrtf(var + eps);
  const float4 gg = reinterpret_cast<const float4*>(g)[lane];
  const float4 bb = reinterpret_cast<const float4*>(b)[lane];
  float4 o;
  o.x = v[0] * rstd * gg.x + bb.x;
  o.y = v[1] * rstd * gg.y + bb.y;
  o.z = v[2] * rstd * gg.z + bb.z;
  o.w = v[3] * rstd * gg.w + bb.w;
  float* orow = out + (long long)row * out_ld;
  reinterpret_cast<float4*>(orow)[lane] = o;
  if (split3) {
    reinterpret_cast<float4*>(orow + 128)[lane] = make_float4(tf32_lo(o.x), tf32_lo(o.y), tf32_lo(o.z), tf32_lo(o.w));
    reinterpret_cast<float4*>(orow + 256)[lane] = o;
  }
}

// [rows][E] -> [rows][3E] = [v | v - trunc_tf32(v) | v]
__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ in, float* __restrict__ out, long long rows,
                                                     int E) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  pdl_sync();
  if (i >= rows * E) return;
  const long long r = i / E;
  const int c = (int)(i - r * E);
  const float v = in[i];
  float* o = out + r * 3 * E + c;
  o[0] = v;
  o[E] = tf32_lo(v);
  o[2 * E] = v;
}

static int launch_linear(const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc, int M,
                         int N, int K, int act, cudaStream_t st) {
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  if (act == 1) linear_kernel<1><<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, C, ldc, M, N, K);
  else if (act == 2) linear_kernel<2><<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, C, ldc, M, N, K);
  else linear_kernel<0><<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, C, ldc, M, N, K);
  return check_launch();
}

}  // namespace mde

using namespace mde;

extern "C" {

int mde_linear_fwd(const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc, int M, int N,
                   int K, int act, mde_stream_t stream) {
  if (!A || !W || !C) return MDE_ERR_BAD_POINTER;
  if (M <= 0 || N <= 0 || K <= 0 || lda < K || ldw < K || ldc < N || act < 0 || act > 2) return MDE_ERR_BAD_SHAPE;
  return launch_linear(A, lda, W, ldw, bias, C, ldc, M, N, K, act, (cudaStream_t)stream);
}

int64_t mde_encoder_layer_ws_floats(int S, int NB, int E, int FF) {
  const int64_t M = (int64_t)S * NB;
  return M * (3 * E + E + E + FF);
}

int mde_encoder_layer_fwd(const float* x, float* y, const float* in_w, const float* in_b, const float* out_w,
                          const float* out_b, const float* ln1_w, const float* ln1_b, const float* l1_w, const float* l1_b,
                          const float* l2_w, const float* l2_b, const float* ln2_w, const float* ln2_b, float* ws, int S,
                          int NB, int E, int heads, int FF, float eps, mde_stream_t stream) {
  if (!x || !y || !in_w || !in_b || !out_w || !out_b || !ln1_w || !ln1_b || !l1_w || !l1_b || !l2_w || !l2_b || !ln2_w ||
      !ln2_b || !ws)
    return MDE_ERR_BAD_POINTER;
  if (S <= 0 || NB <= 0 || E != 128 || heads <= 0 || E % heads != 0 || E / heads != 32 || FF <= 0) return MDE_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int M = S * NB;
  float* qkv = ws;
  float* att = qkv + (size_t)M * 3 * E;
  float* tmp = att + (size_t)M * E;
  float* ff = tmp + (size_t)M * E;
  int rc;
  if ((rc = launch_linear(x, E, in_w, E, in_b, qkv, 3 * E, M, 3 * E, E, 0, st))) return rc;
  if ((rc = launch_attention(qkv, att, S, NB, E, heads, E, 0, st))) return rc;
  if ((rc = launch_linear(att, E, out_w, E, out_b, tmp, E, M, E, E, 0, st))) return rc;
  launch_pdl(PDL_CHAIN, add_layernorm128_kernel, dim3((M + 7) / 8), dim3(256), 0, st, x, tmp, ln1_w, ln1_b, att, M, eps, 128, 128, 128, 0, 1,
             0LL);  // att := LN1(x + sa)
  if ((rc = check_launch())) return rc;
  if ((rc = launch_linear(att, E, l1_w, E, l1_b, ff, FF, M, FF, E, 1, st))) return rc;
  if ((rc = launch_linear(ff, FF, l2_w, FF, l2_b, tmp, E, M, E, FF, 0, st))) return rc;
  launch_pdl(PDL_CHAIN, add_layernorm128_kernel, dim3((M + 7) / 8), dim3(256), 0, st, att, tmp, ln2_w, ln2_b, y, M, eps, 128, 128, 128, 0, 1, 0LL);
  return check_launch();
}

// ---- tensor-core variant: the four nn.Linear products on the tcgen05 NT GEMM in 3xTF32 --------------------------------
// Token matrices travel in "split form" rows [v | v_lo | v] (pitch 3E, v_lo = v - trunc_tf32(v)) and the weights as
// [w_hi | w_hi | w_lo] along K (prepared once by the caller), so ONE K' = 3K TF32 GEMM evaluates v_hi*w_hi + v_lo*w_hi +
// v_hi*w_lo: fp32-grade accuracy (the queries feed a softmax downstream) at tensor-core speed.
int64_t mde_encoder_layer_tc_ws_floats(int S, int NB, int E, int FF) {
  const int64_t M = (int64_t)S * NB;
  return M * (3 * E /*qkv*/ + 3 * E /*att3*/ + 4 * E /*tmp: up to 4 split-K planes*/ + 3 * E /*x1_3*/ + 3 * (int64_t)FF /*h3*/);
}

int mde_split3_tf32(const float* in, float* out, int64_t rows, int E, mde_stream_t stream) {
  if (!in || !out) return MDE_ERR_BAD_POINTER;
  if (rows <= 0 || E <= 0) return MDE_ERR_BAD_SHAPE;
  const long long n = rows * E;
  launch_pdl(PDL_CHAIN, split3_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, in, out, (long long)rows, E);
  return check_launch();
}

int mde_encoder_layer_tc_fwd(const float* x3, float* y, int y_split, const float* in_w3, const float* in_b,
                             const float* out_w3, const float* out_b, const float* ln1_w, const float* ln1_b,
                             const float* l1_w3, const float* l1_b, const float* l2_w3, const float* l2_b, const float* ln2_w,
                             const float* ln2_b, float* ws, int S, int NB, int E, int heads, int FF, float eps,
                             mde_stream_t stream) {
  if (!x3 || !y || !in_w3 || !in_b || !out_w3 || !out_b || !ln1_w || !ln1_b || !l1_w3 || !l1_b || !l2_w3 || !l2_b || !ln2_w ||
      !ln2_b || !ws)
    return MDE_ERR_BAD_POINTER;
  if (S <= 0 || NB <= 0 || E != 128 || heads <= 0 || E % heads != 0 || E / heads != 32 || FF <= 0 || FF % 4 != 0)
    return MDE_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int M = S * NB, E3 = 3 * E;
  float* qkv = ws;                        // [M][3E] plain q | k | v
  float* att3 = qkv + (size_t)M * E3;     // [M][3E] split form
  float* tmp = att3 + (size_t)M * E3;     // [M][E]
  float* x13 = tmp + (size_t)4 * M * E;   // [M][3E] split form of LN1 output
  float* h3 = x13 + (size_t)M * E3;       // [M][3FF] split form of relu(linear1)
  int rc;
  // qkv = x W_in^T + b_in
  if ((rc = mde_gemm_nt_tf32_ex(x3, E3, 0, in_w3, E3, 0, qkv, E3, 0, 1, M, E3, E3, 1, 1.0f, in_b, 0, 0, stream))) return rc;
  if ((rc = launch_attention(qkv, att3, S, NB, E, heads, E3, 1, st))) return rc;
  // tmp = att W_out^T + b_out ; x1 = LN1(x + tmp)
  if ((rc = mde_gemm_nt_tf32_ex(att3, E3, 0, out_w3, E3, 0, tmp, E, 0, 1, M, E, E3, 1, 1.0f, out_b, 0, 0, stream))) return rc;
  launch_pdl(PDL_CHAIN, add_layernorm128_kernel, dim3((M + 7) / 8), dim3(256), 0, st, x3, tmp, ln1_w, ln1_b, x13, M, eps, E3, E, E3, 1, 1, 0LL);
  if ((rc = check_launch())) return rc;
  // h = relu(x1 W_1^T + b_1), written in split form; tmp = h W_2^T + b_2: K' = 3 FF is split over 4 CTAs per output
  // tile, each writing its own partial plane (one CTA per tile: 31.7 us, atomics: 39 us); LN2 sums the planes
  if ((rc = mde_gemm_nt_tf32_ex(x13, E3, 0, l1_w3, E3, 0, h3, 3 * (int64_t)FF, 0, 1, M, FF, E3, 1, 1.0f, l1_b, 1, 1, stream)))
    return rc;
  const int kchunks = (3 * FF + 31) / 32;
  const int planes = kchunks >= 4 ? 4 : 1;
  if ((rc = mde_gemm_nt_tf32_planes(h3, 3 * (int64_t)FF, 0, l2_w3, 3 * (int64_t)FF, 0, tmp, E, 0, 1, M, E, 3 * FF, planes, 1.0f,
                                    l2_b, 0, 0, (int64_t)M * E, stream)))
    return rc;
  const int cps = (kchunks + planes - 1) / planes;
  const int real_planes = (kchunks + cps - 1) / cps;
  launch_pdl(PDL_CHAIN, add_layernorm128_kernel, dim3((M + 7) / 8), dim3(256), 0, st, x13, tmp, ln2_w, ln2_b, y, M, eps, E3, E,
             y_split ? E3 : E, y_split ? 1 : 0, real_planes, (long long)M * E);
  return check_launch();
}

}  // extern "C"
