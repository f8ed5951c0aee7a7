// K1b -- PatchTransformerEncoder layers (post-LN nn.TransformerEncoderLayer, d_model 128, 4 heads, FF 1024, ReLU).
//
// Reference: models/layers.py:8-9,23 (nn.TransformerEncoder of 4 default nn.TransformerEncoderLayer), eval semantics
// (dropout is the identity).  Token layout is the reference's [S, N, E]: row index = s * N + n.
// The sequence is short (S = 13*17 = 221 tokens/image) and the whole encoder is 0.68 GFLOP/image (3 % of the head),
// so these are exact-fp32 SIMT kernels (the queries feed a softmax downstream, see DESIGN.md "precision"):
//   linear_kernel      C = act(A W^T + b)            64x64x16 shared-memory tiles, 4x4 register blocking
//   attention_kernel   softmax(Q K^T / sqrt(hd)) V   one CTA per (image, head), K and V resident in shared memory,
//                                                    one warp per query row, lane == head-dim channel (hd = 32)
//   add_layernorm_kernel  LN(x + y) * g + b          one warp per token row (E = 128 -> float4 per lane)
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

// qkv [S*NB, 3E] (row = s*NB + n); out rows have pitch out_ld.  grid (NB*heads, y splits over the query rows).
// One CTA per (image, head, row range): K and V of the (image, head) resident in shared memory ([S][HD+1], conflict-free);
// each warp processes FOUR query rows per pass so that every K / V word read from shared memory feeds four FMAs (the
// one-row-per-warp form was bound by the shared-memory port: 666 warp-wide loads per row, now 222):
//   scores : lane = key j, q of the four rows broadcast as one float4 per head-dim channel
//   softmax: probabilities kept interleaved [j][4 rows] in shared memory
//   P V    : lane = head-dim channel, one float4 broadcast of the four rows' p_j + one V word per key
// split3 != 0 writes each output row as [v | v - trunc_tf32(v) | v] (the 3xTF32 A operand form).
__device__ __forceinline__ float tf32_lo(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

template <int HD>
__global__ void __launch_bounds__(256) attention_kernel(const float* __restrict__ qkv, float* __restrict__ out, int S,
                                                        int NB, int E, int heads, float scale, int out_ld, int split3) {
  static_assert(HD == 32, "lane == head-dim channel");
  extern __shared__ __align__(16) float smem_att[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* sk = smem_att;                               // [S][HD+1]
  float* sv = sk + (size_t)S * (HD + 1);              // [S][HD+1]
  size_t off = (size_t)2 * S * (HD + 1);
  off = (off + 3) & ~(size_t)3;                       // float4 alignment of the per-warp areas
  float* sq = smem_att + off + (size_t)warp * (4 * HD + 4 * S);  // [HD][4]  q of the four rows, channel-major
  float* sp = sq + 4 * HD;                                       // [S][4]   scores / probabilities of the four rows
  const int n = blockIdx.x / heads, hh = blockIdx.x % heads;
  const int ld = 3 * E;
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
          float* vd = sv + s * (HD + 1) + d4 * 4;
          kd[0] = kv[u].x; kd[1] = kv[u].y; kd[2] = kv[u].z; kd[3] = kv[u].w;
          vd[0] = vv[u].x; vd[1] = vv[u].y; vd[2] = vv[u].z; vd[3] = vv[u].w;
        }
      }
    }
  }
  __syncthreads();
  const int rows_per_cta = (S + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(S, r0 + rows_per_cta);
  for (int i0 = r0 + 4 * warp; i0 < r1; i0 += 4 * nwarps) {
    // (1) the four (pre-scaled, as torch does: q / sqrt(hd)) query rows -> sq[d][r]
    float4 qv;
    {
      float t[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = min(i0 + r, S - 1);
        t[r] = qkv[((long long)i * NB + n) * ld + hh * HD + lane] * scale;
      }
      qv = make_float4(t[0], t[1], t[2], t[3]);
    }
    reinterpret_cast<float4*>(sq)[lane] = qv;
    __syncwarp();
    // (2) scores, lane = key
    float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int j0 = 0; j0 < S; j0 += 32) {
      const int j = j0 + lane;
      const float* kr = sk + (size_t)min(j, S - 1) * (HD + 1);
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) {
        const float4 q4 = reinterpret_cast<const float4*>(sq)[d];
        const float kd = kr[d];
        a0 = fmaf(q4.x, kd, a0); a1 = fmaf(q4.y, kd, a1); a2 = fmaf(q4.z, kd, a2); a3 = fmaf(q4.w, kd, a3);
      }
      if (j < S) {
        reinterpret_cast<float4*>(sp)[j] = make_float4(a0, a1, a2, a3);
        mx[0] = fmaxf(mx[0], a0); mx[1] = fmaxf(mx[1], a1); mx[2] = fmaxf(mx[2], a2); mx[3] = fmaxf(mx[3], a3);
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) mx[r] = warp_max(mx[r]);
    // (3) exponentials and their sums
    float sum[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = lane; j < S; j += 32) {
      float4 a = reinterpret_cast<float4*>(sp)[j];
      a.x = expf(a.x - mx[0]); a.y = expf(a.y - mx[1]); a.z = expf(a.z - mx[2]); a.w = expf(a.w - mx[3]);
      sum[0] += a.x; sum[1] += a.y; sum[2] += a.z; sum[3] += a.w;
      reinterpret_cast<float4*>(sp)[j] = a;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) sum[r] = warp_sum(sum[r]);
    __syncwarp();
    // (4) P V, lane = head-dim channel
    float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
    for (int j = 0; j < S; ++j) {
      const float4 p4 = reinterpret_cast<const float4*>(sp)[j];
      const float vj = sv[(size_t)j * (HD + 1) + lane];
      o0 = fmaf(p4.x, vj, o0); o1 = fmaf(p4.y, vj, o1); o2 = fmaf(p4.z, vj, o2); o3 = fmaf(p4.w, vj, o3);
    }
    const float ov[4] = {o0 / sum[0], o1 / sum[1], o2 / sum[2], o3 / sum[3]};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = i0 + r;
      if (i < r1) {
        float* orow = out + ((long long)i * NB + n) * out_ld + hh * HD + lane;
        orow[0] = ov[r];
        if (split3) {
          orow[E] = tf32_lo(ov[r]);
          orow[2 * E] = ov[r];
        }
      }
    }
    __syncwarp();
  }
}

// out[m,:] = LayerNorm(x[m,:] + y[m,:]) * g + b ;  E == 128 (one float4 per lane)
__global__ void __launch_bounds__(256) add_layernorm128_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                               const float* __restrict__ g, const float* __restrict__ b,
                                                               float* __restrict__ out, int M, float eps, int x_ld = 128,
                                                               int y_ld = 128, int out_ld = 128, int split3 = 0,
                                                               int y_planes = 1, long long y_plane_stride = 0) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
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
  {
    const int hd = E / heads;
    const size_t sm = sizeof(float) * ((((size_t)2 * S * (hd + 1) + 3) & ~(size_t)3) + (size_t)8 * (4 * hd + 4 * S));
    if (sm > 200 * 1024) return MDE_ERR_BAD_SHAPE;
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(attention_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      attr = true;
    }
    int ysplit = (2 * MDE_NUM_SMS + NB * heads - 1) / (NB * heads);
    if (ysplit < 1) ysplit = 1;
    if (ysplit > (S + 7) / 8) ysplit = (S + 7) / 8;
    attention_kernel<32><<<dim3(NB * heads, ysplit), 256, sm, st>>>(qkv, att, S, NB, E, heads, 1.0f / sqrtf((float)hd), E, 0);
    if ((rc = check_launch())) return rc;
  }
  if ((rc = launch_linear(att, E, out_w, E, out_b, tmp, E, M, E, E, 0, st))) return rc;
  add_layernorm128_kernel<<<(M + 7) / 8, 256, 0, st>>>(x, tmp, ln1_w, ln1_b, att, M, eps);  // att := LN1(x + sa)
  if ((rc = check_launch())) return rc;
  if ((rc = launch_linear(att, E, l1_w, E, l1_b, ff, FF, M, FF, E, 1, st))) return rc;
  if ((rc = launch_linear(ff, FF, l2_w, FF, l2_b, tmp, E, M, E, FF, 0, st))) return rc;
  add_layernorm128_kernel<<<(M + 7) / 8, 256, 0, st>>>(att, tmp, ln2_w, ln2_b, y, M, eps);
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
  split3_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, out, rows, E);
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
  {
    const int hd = E / heads;
    const size_t sm = sizeof(float) * ((((size_t)2 * S * (hd + 1) + 3) & ~(size_t)3) + (size_t)8 * (4 * hd + 4 * S));
    if (sm > 200 * 1024) return MDE_ERR_BAD_SHAPE;
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(attention_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      attr = true;
    }
    int ysplit = (2 * MDE_NUM_SMS + NB * heads - 1) / (NB * heads);
    if (ysplit < 1) ysplit = 1;
    if (ysplit > (S + 7) / 8) ysplit = (S + 7) / 8;
    attention_kernel<32><<<dim3(NB * heads, ysplit), 256, sm, st>>>(qkv, att3, S, NB, E, heads, 1.0f / sqrtf((float)hd), E3, 1);
    if ((rc = check_launch())) return rc;
  }
  // tmp = att W_out^T + b_out ; x1 = LN1(x + tmp)
  if ((rc = mde_gemm_nt_tf32_ex(att3, E3, 0, out_w3, E3, 0, tmp, E, 0, 1, M, E, E3, 1, 1.0f, out_b, 0, 0, stream))) return rc;
  add_layernorm128_kernel<<<(M + 7) / 8, 256, 0, st>>>(x3, tmp, ln1_w, ln1_b, x13, M, eps, E3, E, E3, 1);
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
  add_layernorm128_kernel<<<(M + 7) / 8, 256, 0, st>>>(x13, tmp, ln2_w, ln2_b, y, M, eps, E3, E, y_split ? E3 : E,
                                                       y_split ? 1 : 0, real_planes, (long long)M * E);
  return check_launch();
}

}  // extern "C"
