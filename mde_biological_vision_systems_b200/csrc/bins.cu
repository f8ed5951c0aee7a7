// K2 and friends -- the per-pixel bin pipeline.
//   * regressor_bins : bin-width regressor MLP -> normalisation -> widths -> cumsum edges -> centres  (one CTA / image)
//   * bins_pred      : streaming softmax over n_bins + centre-weighted sum  (vectorised, coalesced HBM streaming)
//   * pixel_gemm     : SIMT fp32  y[b,n,p] = sum_k W[b?][n,k] * x[b,k,p] (+bias)  -- exact-fp32 range attention and
//                      conv_out, the un-fused companions of the tcgen05 chain in head_chain_tc.cu
//   * fold_queries   : wf[b] = tf32_round( (W_out @ Q_b) * log2 e )   for the fused chain
// Reference: models/miniViT.py:17-21,35-45; models/layers.py:31-36; models/unet_adaptive_bins.py:190-191,286-300.
#include "common.cuh"

namespace mde {

// ------------------------------------------------------------------------------------------------------------
// regressor + bins: one 1024-thread CTA per image.  Four output rows per warp pass, 16-byte coalesced weight reads.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dense_layer(const float* __restrict__ W, const float* __restrict__ bias,
                                            const float* in, float* out, int n_out, int n_in, bool leaky) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if ((n_in & 127) == 0 && ((reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(in)) & 15) == 0) {
    // FOUR output rows per warp pass, 16-byte weight loads: the layer is a chain of L2 round trips (640 KB of weights, 16 CTAs),
    // so what counts is how many loads a warp has in flight -- the one-row, 4-byte form had 8 dependent-free but un-unrolled
    // loads and a 5-step shuffle tree per row, 32 rows per warp: 73 us for the whole regressor at B = 16
    for (int r0 = warp * 4; r0 < n_out; r0 += nw * 4) {
      float a[4] = {0.f, 0.f, 0.f, 0.f};
      for (int k = lane * 4; k < n_in; k += 128) {
        const float4 x = *reinterpret_cast<const float4*>(in + k);
        float4 w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)  // rows beyond n_out re-read the last row (discarded below): no branch around the loads
          w[j] = __ldg(reinterpret_cast<const float4*>(W + (long long)min(r0 + j, n_out - 1) * n_in + k));
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] = fmaf(w[j].x, x.x, fmaf(w[j].y, x.y, fmaf(w[j].z, x.z, fmaf(w[j].w, x.w, a[j]))));
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) a[j] = warp_sum(a[j]);
      if (lane < 4 && r0 + lane < n_out) {
        float v = lane == 0 ? a[0] : lane == 1 ? a[1] : lane == 2 ? a[2] : a[3];
        v += bias[r0 + lane];
        out[r0 + lane] = leaky ? (v > 0.f ? v : 0.01f * v) : v;
      }
    }
    __syncthreads();
    return;
  }
  for (int r = warp; r < n_out; r += nw) {
    const float* wr = W + (long long)r * n_in;
    float a = 0.f;
    for (int k = lane; k < n_in; k += 32) a = fmaf(wr[k], in[k], a);
    a = warp_sum(a);
    if (lane == 0) {
      a += bias[r];
      out[r] = leaky ? (a > 0.f ? a : 0.01f * a) : a;
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(1024) regressor_bins_kernel(const float* __restrict__ t0, long long t0_stride,
                                                             const float* __restrict__ w1, const float* __restrict__ b1,
                                                             const float* __restrict__ w2, const float* __restrict__ b2,
                                                             const float* __restrict__ w3, const float* __restrict__ b3,
                                                             int E, int H, int n_bins, int norm_mode, float min_val,
                                                             float max_val, float* __restrict__ y_raw,
                                                             float* __restrict__ widths_normed, float* __restrict__ edges,
                                                             float* __restrict__ centers) {
  extern __shared__ __align__(16) double smd[];  // scan[n_bins+2] (double) | in[E] | h1[H] | h2[H] | y[n_bins]
  double* scan = smd;
  float* sin = reinterpret_cast<float*>(smd + n_bins + 2);
  float* h1 = sin + E;
  float* h2 = h1 + H;
  float* y = h2 + H;
  const int b = blockIdx.x;
  pdl_sync();
  if (w1 != nullptr) {
    for (int i = threadIdx.x; i < E; i += blockDim.x) sin[i] = t0[(long long)b * t0_stride + i];
    __syncthreads();
    dense_layer(w1, b1, sin, h1, H, E, true);
    dense_layer(w2, b2, h1, h2, H, H, true);
    dense_layer(w3, b3, h2, y, n_bins, H, false);
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x) y_raw[(long long)b * n_bins + i] = y[i];
  } else {  // finalize-only mode: the regressor output was produced by mde_linear_fwd
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x) y[i] = y_raw[(long long)b * n_bins + i];
    __syncthreads();
  }
  // normalisation (miniViT.py:36-44)
  __shared__ float red[32];
  __shared__ float bcast;
  if (norm_mode == MDE_NORM_SOFTMAX) {
    float m = -INFINITY;
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x) m = fmaxf(m, y[i]);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      float mm = red[0];
      for (int i = 1; i < (int)(blockDim.x >> 5); ++i) mm = fmaxf(mm, red[i]);
      bcast = mm;
    }
    __syncthreads();
    const float mm = bcast;
    __syncthreads();
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x) y[i] = expf(y[i] - mm);
  } else if (norm_mode == MDE_NORM_LINEAR) {
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x) y[i] = fmaxf(y[i], 0.f) + 0.1f;
  } else {
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x) y[i] = 1.f / (1.f + expf(-y[i]));
  }
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < n_bins; i += blockDim.x) s += y[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    bcast = t;
  }
  __syncthreads();
  const float total = bcast;
  for (int i = threadIdx.x; i < n_bins; i += blockDim.x) {
    const float wn = y[i] / total;
    y[i] = wn;
    widths_normed[(long long)b * n_bins + i] = wn;
  }
  __syncthreads();
  // edges = cumsum(pad(widths, (1,0), min_val)); float64 running sum, rounded to fp32 per element.  One warp: every lane sums
  // its run of consecutive widths, a shuffle scan turns the lane totals into offsets (a single thread walking all n_bins
  // dependent fp64 adds was a third of this kernel)
  if (threadIdx.x < 32) {
    const float range = max_val - min_val;
    const int per = (n_bins + 31) >> 5;
    const int i0 = min((int)threadIdx.x * per, n_bins), i1 = min(i0 + per, n_bins);
    double run = 0.0;
    for (int i = i0; i < i1; ++i) run += (double)(range * y[i]);
    double incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double u = __shfl_up_sync(0xffffffffu, incl, o);
      if ((int)threadIdx.x >= o) incl += u;
    }
    run = (double)min_val + (incl - run);  // exclusive offset of this lane's run
    if (threadIdx.x == 0) scan[0] = run;
    for (int i = i0; i < i1; ++i) {
      run += (double)(range * y[i]);
      scan[i + 1] = run;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= n_bins; i += blockDim.x) edges[(long long)b * (n_bins + 1) + i] = (float)scan[i];
  for (int i = threadIdx.x; i < n_bins; i += blockDim.x)
    centers[(long long)b * n_bins + i] = 0.5f * ((float)scan[i] + (float)scan[i + 1]);
}

// ------------------------------------------------------------------------------------------------------------
// bins_pred: pred[b,p] = sum_j softmax_j(logits[b,j,p]) * c[b,j].   logits [B,n,P] is read exactly once.
// CTA = 256 threads = 8 warps; the CTA owns VEC*32 consecutive pixels; warp w streams bins [w*n/8, (w+1)*n/8) with
// one 128-bit load per lane per bin (a warp reads 512 contiguous bytes per bin row), keeps an online softmax
// (running max, sum, centre-weighted sum) per pixel in registers, and the 8 partial states are merged through
// shared memory -- no second pass, no materialised softmax (the reference makes three 58 MB/img passes).
// ------------------------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256) bins_pred_kernel(const float* __restrict__ logits, const float* __restrict__ centers,
                                                        float* __restrict__ pred, int n, long long P) {
  extern __shared__ float sm2[];  // centres[n] | m[8][VEC*32] | s[8][..] | ws[8][..]
  constexpr int PX = VEC * 32;
  float* sc = sm2;
  float* pm = sc + n;
  float* ps = pm + 8 * PX;
  float* pw = ps + 8 * PX;
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < n; i += blockDim.x) sc[i] = centers[(long long)b * n + i];
  __syncthreads();
  const long long p0 = (long long)blockIdx.x * PX + (long long)lane * VEC;
  const bool active = p0 < P;
  const int per = (n + 7) / 8;
  const int j0 = warp * per, j1 = min(n, j0 + per);
  float m[VEC], s[VEC], ws[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    m[v] = -INFINITY;
    s[v] = 0.f;
    ws[v] = 0.f;
  }
  if (active) {
    const float* base = logits + ((long long)b * n) * P + p0;
    constexpr int U = 8;
    for (int j = j0; j < j1; j += U) {
      float x[U][VEC];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (j + u < j1) {
          if (VEC == 4) {
            const float4 t = ldg_stream(reinterpret_cast<const float4*>(base + (long long)(j + u) * P));
            x[u][0] = t.x; x[u][1] = t.y; x[u][2] = t.z; x[u][3] = t.w;
          } else {
            x[u][0] = __ldg(base + (long long)(j + u) * P);
          }
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) x[u][v] = -INFINITY;
        }
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float cm = x[0][v];
#pragma unroll
        for (int u = 1; u < U; ++u) cm = fmaxf(cm, x[u][v]);
        const float nm = fmaxf(m[v], cm);
        const float r = __expf(m[v] - nm);  // exp(-inf) = 0 on the first chunk
        float ls = 0.f, lw = 0.f;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float e = __expf(x[u][v] - nm);
          ls += e;
          lw = fmaf(e, (j + u < j1) ? sc[j + u] : 0.f, lw);
        }
        s[v] = fmaf(s[v], r, ls);
        ws[v] = fmaf(ws[v], r, lw);
        m[v] = nm;
      }
    }
  }
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    pm[warp * PX + lane * VEC + v] = m[v];
    ps[warp * PX + lane * VEC + v] = s[v];
    pw[warp * PX + lane * VEC + v] = ws[v];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < PX; i += blockDim.x) {
    const long long p = (long long)blockIdx.x * PX + i;
    if (p >= P) continue;
    float mm = pm[i];
#pragma unroll
    for (int w = 1; w < 8; ++w) mm = fmaxf(mm, pm[w * PX + i]);
    float S = 0.f, Wsum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const float r = __expf(pm[w * PX + i] - mm);
      S = fmaf(ps[w * PX + i], r, S);
      Wsum = fmaf(pw[w * PX + i], r, Wsum);
    }
    pred[(long long)b * P + p] = Wsum / S;
  }
}

// ------------------------------------------------------------------------------------------------------------
// pixel_gemm (SIMT fp32): y[b,n,p] = bias[n] + sum_k W[b*wbs + n*K + k] * x[b,k,p]
// CTA tile 64 (n) x 128 (p), K step 16; 256 threads, each 8 n x 4 p.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float tf32_rna(float v) {
  unsigned int r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// EPI 0: y = acc + bias[n];  EPI 1: y = tf32_rna(acc * out_scale);  EPI 2: y = acc * out_scale  (operand preparation
// for the tensor-core chain)
template <int EPI>
__global__ void __launch_bounds__(256) pixel_gemm_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                         long long wbs, const float* __restrict__ bias,
                                                         float* __restrict__ y, int K, int N, long long P,
                                                         float out_scale) {
  __shared__ __align__(16) float sx[16][128];
  __shared__ float sw[16][64 + 1];
  pdl_sync();
  const int b = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 128;
  const int n0 = blockIdx.y * 64;
  const int tp = threadIdx.x & 31, tn = threadIdx.x >> 5;
  const float* xb = x + (long long)b * K * P;
  const float* wb = W + (long long)b * wbs;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    // x tile: 16 x 128 floats = 512 float4 -> 2 per thread
    for (int i = threadIdx.x; i < 16 * 32; i += 256) {
      const int kk = i >> 5, c4 = i & 31;
      const long long p = p0 + c4 * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + kk < K) {
        const float* src = xb + (long long)(k0 + kk) * P + p;
        if (p + 3 < P && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) v = *reinterpret_cast<const float4*>(src);
        else {
          if (p < P) v.x = src[0];
          if (p + 1 < P) v.y = src[1];
          if (p + 2 < P) v.z = src[2];
          if (p + 3 < P) v.w = src[3];
        }
      }
      *reinterpret_cast<float4*>(&sx[kk][c4 * 4]) = v;
    }
    // W tile: 64 n x 16 k
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int nn = i >> 4, kk = i & 15;
      sw[kk][nn] = (n0 + nn < N && k0 + kk < K) ? wb[(long long)(n0 + nn) * K + k0 + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 xv = *reinterpret_cast<const float4*>(&sx[kk][tp * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float wv = sw[kk][tn * 8 + i];
        acc[i][0] = fmaf(wv, xv.x, acc[i][0]);
        acc[i][1] = fmaf(wv, xv.y, acc[i][1]);
        acc[i][2] = fmaf(wv, xv.z, acc[i][2]);
        acc[i][3] = fmaf(wv, xv.w, acc[i][3]);
      }
    }
    __syncthreads();
  }
  float* yb = y + (long long)b * N * P;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int n = n0 + tn * 8 + i;
    if (n >= N) continue;
    const float bv = (EPI == 0 && bias) ? bias[n] : 0.f;
    const long long p = p0 + tp * 4;
    float* dst = yb + (long long)n * P + p;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      o[j] = EPI == 0 ? acc[i][j] + bv : (EPI == 1 ? tf32_rna(acc[i][j] * out_scale) : acc[i][j] * out_scale);
    if (p + 3 < P && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
      *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (p + j < P) dst[j] = o[j];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// fold_queries: wf[b,j,k] = tf32_rna( log2e * scale * sum_n w_out[j,n] * q[b,n,k] ) is pixel_gemm<1> with x := q[b]
// viewed as [n][k] ("pixels" = k); biasf[j] = log2e * bias[j]
// ------------------------------------------------------------------------------------------------------------
// biasf[b,j] = log2e * bias[j] + (1/operand_scale) * sum_k wf[b,j,k] * feat_bias[k]
// (wf already carries log2e * operand_scale).  feat_bias is the bias of the conv that produced the activations: it is a
// per-channel constant at every pixel, so it moves through the contraction into a per-image, per-bin bias and the
// producer can run bias-free (saves a full read+write pass over the 29 MB/img feature map).
__global__ void __launch_bounds__(256) chain_bias_kernel(const float* __restrict__ bias, const float* __restrict__ wf,
                                                         const float* __restrict__ feat_bias, float* __restrict__ biasf,
                                                         int n_bins, int K, float log2e, float inv_scale) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  pdl_sync();
  if (j >= n_bins) return;
  float a = 0.f;
  if (feat_bias) {
    const float* row = wf + ((long long)b * n_bins + j) * K;
    for (int k = lane; k < K; k += 32) a = fmaf(row[k], feat_bias[k], a);
    a = warp_sum(a);
  }
  if (lane == 0) biasf[(long long)b * n_bins + j] = fmaf(a, inv_scale, bias[j] * log2e);
}

}  // namespace mde

using namespace mde;

extern "C" {

int mde_regressor_bins_fwd(const float* t0, int64_t t0_stride, const float* w1, const float* b1, const float* w2,
                           const float* b2, const float* w3, const float* b3, int B, int E, int H, int n_bins,
                           int norm_mode, float min_val, float max_val, float* y_raw, float* widths_normed,
                           float* edges, float* centers, mde_stream_t stream) {
  if (!t0 || !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !y_raw || !widths_normed || !edges || !centers)
    return MDE_ERR_BAD_POINTER;
  if (B <= 0 || E <= 0 || H <= 0 || n_bins <= 0 || n_bins > 4096 || E > 4096 || H > 4096) return MDE_ERR_BAD_SHAPE;
  if (norm_mode < 0 || norm_mode > 2) return MDE_ERR_UNSUPPORTED;
  const size_t sm = sizeof(float) * (size_t)(E + 2 * H + n_bins) + sizeof(double) * (size_t)(n_bins + 2);
  launch_pdl(PDL_CHAIN, regressor_bins_kernel, dim3(B), dim3(1024), sm, (cudaStream_t)stream, t0, (long long)t0_stride, w1, b1, w2, b2, w3, b3,
             E, H, n_bins, norm_mode, min_val, max_val, y_raw, widths_normed, edges, centers);
  return check_launch();
}

int mde_bins_finalize_fwd(float* y_raw, int B, int n_bins, int norm_mode, float min_val, float max_val,
                          float* widths_normed, float* edges, float* centers, mde_stream_t stream) {
  if (!y_raw || !widths_normed || !edges || !centers) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || n_bins <= 0 || n_bins > 4096) return MDE_ERR_BAD_SHAPE;
  if (norm_mode < 0 || norm_mode > 2) return MDE_ERR_UNSUPPORTED;
  const size_t sm = sizeof(float) * (size_t)(n_bins) + sizeof(double) * (size_t)(n_bins + 2);
  regressor_bins_kernel<<<B, 256, sm, (cudaStream_t)stream>>>(nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                              nullptr, 0, 0, n_bins, norm_mode, min_val, max_val, y_raw,
                                                              widths_normed, edges, centers);
  return check_launch();
}

int mde_bins_pred_fwd(const float* logits, const float* centers, float* pred, int B, int n_bins, int64_t P,
                      mde_stream_t stream) {
  if (!logits || !centers || !pred) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || B > 65535 || n_bins <= 0 || n_bins > 4096 || P <= 0) return MDE_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (P % 4 == 0) && aligned(logits, 16);
  if (vec) {
    const size_t sm = sizeof(float) * (size_t)(n_bins + 3 * 8 * 128);
    bins_pred_kernel<4><<<dim3((unsigned)((P + 127) / 128), (unsigned)B), 256, sm, st>>>(logits, centers, pred, n_bins, P);
  } else {
    const size_t sm = sizeof(float) * (size_t)(n_bins + 3 * 8 * 32);
    bins_pred_kernel<1><<<dim3((unsigned)((P + 31) / 32), (unsigned)B), 256, sm, st>>>(logits, centers, pred, n_bins, P);
  }
  return check_launch();
}

static int launch_pixel_gemm(const float* x, const float* W, int64_t wbs, const float* bias, float* y, int B, int K,
                             int N, int64_t P, cudaStream_t st) {
  if (B > 65535) return MDE_ERR_BAD_SHAPE;
  dim3 grid((unsigned)((P + 127) / 128), (unsigned)((N + 63) / 64), (unsigned)B);
  pixel_gemm_kernel<0><<<grid, 256, 0, st>>>(x, W, wbs, bias, y, K, N, P, 1.f);
  return check_launch();
}

int mde_conv1x1_fwd(const float* ram, const float* w, const float* bias, float* logits, int B, int K, int n_bins,
                    int64_t P, mde_stream_t stream) {
  if (!ram || !w || !logits) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || K <= 0 || n_bins <= 0 || P <= 0) return MDE_ERR_BAD_SHAPE;
  return launch_pixel_gemm(ram, w, 0, bias, logits, B, K, n_bins, P, (cudaStream_t)stream);
}

int mde_range_attention(const float* x, const float* q, float* y, int B, int K, int N, int64_t P, int impl,
                        mde_stream_t stream) {
  if (!x || !q || !y) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || K <= 0 || N <= 0 || P <= 0) return MDE_ERR_BAD_SHAPE;
  if (impl == 0) return launch_pixel_gemm(x, q, (int64_t)N * K, nullptr, y, B, K, N, P, (cudaStream_t)stream);
  return MDE_ERR_UNSUPPORTED;  // the tensor-core form takes split-bf16 pairs: mde_range_attention_tc
}

int mde_fold_queries(const float* w_out, const float* bias, const float* q, int64_t q_batch_stride,
                     const float* feat_bias, float* wf, float* biasf, int B, int n_bins, int N, int K,
                     float operand_scale, int round_tf32, mde_stream_t stream) {
  if (!w_out || !bias || !q || !wf || !biasf) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || B > 65535 || n_bins <= 0 || N <= 0 || K <= 0) return MDE_ERR_BAD_SHAPE;
  if (q_batch_stride != (int64_t)N * K) return MDE_ERR_BAD_SHAPE;  // q[b] must be a dense [N,K] block
  cudaStream_t st = (cudaStream_t)stream;
  const float LOG2E = 1.4426950408889634f;
  dim3 grid((unsigned)((K + 127) / 128), (unsigned)((n_bins + 63) / 64), (unsigned)B);
  if (round_tf32)
    launch_pdl(PDL_CHAIN, pixel_gemm_kernel<1>, grid, dim3(256), 0, st, q, w_out, 0LL, (const float*)nullptr, wf, N, n_bins, (long long)K,
               LOG2E * operand_scale);
  else
    launch_pdl(PDL_CHAIN, pixel_gemm_kernel<2>, grid, dim3(256), 0, st, q, w_out, 0LL, (const float*)nullptr, wf, N, n_bins, (long long)K,
               LOG2E * operand_scale);
  int rc = check_launch();
  if (rc) return rc;
  launch_pdl(PDL_CHAIN, chain_bias_kernel, dim3((unsigned)((n_bins + 7) / 8), (unsigned)B), dim3(256), 0, st, bias, wf, feat_bias, biasf,
             n_bins, K, LOG2E, 1.f / operand_scale);
  return check_launch();
}

}  // extern "C"
