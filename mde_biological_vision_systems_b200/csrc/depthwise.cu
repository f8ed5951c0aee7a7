// Depthwise k x k convolution (k = 3 or 5, stride 1 or 2) on channels_last activations with the folded-BatchNorm bias, SiLU
// and the squeeze-excite pooling in ONE pass -- the middle of every MBConv block of the EfficientNet passthrough body
// (geffnet InvertedResidual / DepthwiseSeparableConv: conv_dw -> bn -> act -> SqueezeExcite(mean over H, W)).
//
//   y[b, oy, ox, c] = act( bias[c] + sum_{dy, dx} w[dy, dx, c] * x[b, oy*s + dy - pad_t, ox*s + dx - pad_l, c] )
//   partial[b, slab, c] = sum over the slab's output pixels of y          (slabs of consecutive output pixels)
//
// A depthwise convolution has no channel reduction: it is k*k fused multiply-adds per output element in plain fp32 (exact in the
// sense of the 1e-3 contract whatever the library's TF32 switch says) and it is memory-bound -- one read and one write of the
// expanded tensor.  As three library/own passes (depthwise conv, bias + SiLU, spatial mean) the same tensor crossed HBM four times.
// Mapping: as bias_act_pool_nhwc_kernel (aux_mlp.cu) -- a block owns a slab of consecutive output pixels of one image; its
// threads are (channel group of 4, pixel lane), consecutive threads = consecutive float4s of a pixel row, every thread keeps ONE
// channel group so that its filter taps (k*k float4) live in registers and its running sum needs no indexing.  Zero padding is
// the bounds test (asymmetric TensorFlow-SAME padding = pad_t / pad_l offsets).  The k*k input taps of neighbouring output pixels
// overlap; they are served by L1/L2 (the block walks a compact slab).
#include "common.cuh"

namespace mde {

template <int K>
__global__ void __launch_bounds__(256) depthwise_bias_act_pool_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                      const float* __restrict__ bias, float* __restrict__ y,
                                                                      float* __restrict__ partial, int Hi, int Wi, int c4,
                                                                      int stride, int pad_t, int pad_l, int Ho, int Wo,
                                                                      int pix_per_slab, int act) {
  __shared__ float4 red[256];
  const int b = blockIdx.y, slab = blockIdx.x;
  const int P = Ho * Wo;
  const int p0 = slab * pix_per_slab;
  const int p1 = min(P, p0 + pix_per_slab);
  const float4* xb = reinterpret_cast<const float4*>(x) + (long long)b * Hi * Wi * c4;
  float4* yb = reinterpret_cast<float4*>(y) + (long long)b * P * c4;
  const int t = threadIdx.x;
  for (int cg0 = 0; cg0 < c4; cg0 += 256) {
    const int cw = min(256, c4 - cg0);  // channel groups of this pass
    const int ry = 256 / cw;            // pixels per sweep of the block
    const int cg = cg0 + t % cw, lanep = t / cw;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lanep < ry) {
      // k = 3: the nine taps of this thread's channel group stay in registers; k = 5: 25 float4 would cost 100 registers and
      // most of the occupancy that hides the input latency, so the taps are re-read through L1 (hot, 16 B per lane)
      constexpr bool WREG = K == 3;
      const float4* wg = reinterpret_cast<const float4*>(w) + cg;
      float4 wt[WREG ? K * K : 1];
      if (WREG) {
#pragma unroll
        for (int i = 0; i < K * K; ++i) wt[i] = __ldg(wg + (long long)i * c4);
      }
      const float4 bb = __ldg(reinterpret_cast<const float4*>(bias) + cg);
      for (int p = p0 + lanep; p < p1; p += ry) {
        const int oy = p / Wo, ox = p - oy * Wo;
        const int iy0 = oy * stride - pad_t, ix0 = ox * stride - pad_l;
        float4 o = bb;
        // branch-free taps: out-of-range taps read a clamped (valid) address and are multiplied by zero, so that the K loads
        // of a filter row issue back to back (a bounds `continue` per tap made every tap its own load -> fma round trip)
#pragma unroll
        for (int dy = 0; dy < K; ++dy) {
          const int iy = iy0 + dy;
          const int iyc = min(max(iy, 0), Hi - 1);
          const float my = iy == iyc ? 1.f : 0.f;
          const float4* row = xb + (long long)iyc * Wi * c4 + cg;
          float4 v[K];
          float m[K];
#pragma unroll
          for (int dx = 0; dx < K; ++dx) {
            const int ix = ix0 + dx;
            const int ixc = min(max(ix, 0), Wi - 1);
            m[dx] = ix == ixc ? my : 0.f;
            v[dx] = __ldg(row + (long long)ixc * c4);
          }
#pragma unroll
          for (int dx = 0; dx < K; ++dx) {
            const float4 q = WREG ? wt[WREG ? dy * K + dx : 0] : __ldg(wg + (long long)(dy * K + dx) * c4);
            o.x = fmaf(v[dx].x * m[dx], q.x, o.x); o.y = fmaf(v[dx].y * m[dx], q.y, o.y);
            o.z = fmaf(v[dx].z * m[dx], q.z, o.z); o.w = fmaf(v[dx].w * m[dx], q.w, o.w);
          }
        }
        if (act == 1) {
          o.x = __fdividef(o.x, 1.f + __expf(-o.x)); o.y = __fdividef(o.y, 1.f + __expf(-o.y));
          o.z = __fdividef(o.z, 1.f + __expf(-o.z)); o.w = __fdividef(o.w, 1.f + __expf(-o.w));
        }
        yb[(long long)p * c4 + cg] = o;
        acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
      }
    }
    red[t] = acc;
    __syncthreads();
    if (t < cw) {  // fixed order: pixel lanes 0 .. ry-1
      for (int k = 1; k < ry; ++k) {
        const float4 u = red[t + k * cw];
        acc.x += u.x; acc.y += u.y; acc.z += u.z; acc.w += u.w;
      }
      reinterpret_cast<float4*>(partial)[((long long)b * gridDim.x + slab) * c4 + cg] = acc;
    }
    __syncthreads();
  }
}

}  // namespace mde

extern "C" int mde_pool_slabs(int B, int64_t HW);

// x [B,Hi,Wi,C] fp32 NHWC; w [k][k][C] (the depthwise filter, channel innermost); bias [C]; y [B,Ho,Wo,C]; partial
// [B][mde_pool_slabs(B, Ho*Wo)][C] receives the per-slab channel sums of y.  C % 4 == 0, k in {3, 5}.
extern "C" int mde_depthwise_bias_act_pool_nhwc(const float* x, const float* w, const float* bias, float* y, float* partial, int B,
                                                int Hi, int Wi, int C, int k, int stride, int pad_top, int pad_left, int Ho,
                                                int Wo, int act, mde_stream_t stream) {
  using namespace mde;
  if (!x || !w || !bias || !y || !partial) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || B > 65535 || Hi <= 0 || Wi <= 0 || C <= 0 || Ho <= 0 || Wo <= 0 || stride < 1 || stride > 2 || pad_top < 0 ||
      pad_left < 0 || act < 0 || act > 1 || (long long)Ho * Wo > 0x7fffffffLL || (long long)Hi * Wi * (C / 4) > 0x7fffffffLL)
    return MDE_ERR_BAD_SHAPE;
  if ((k != 3 && k != 5) || C % 4 != 0 || !aligned(x, 16) || !aligned(y, 16) || !aligned(w, 16) || !aligned(bias, 16) ||
      !aligned(partial, 16))
    return MDE_ERR_UNSUPPORTED;
  if ((Ho - 1) * stride - pad_top >= Hi || (Wo - 1) * stride - pad_left >= Wi) return MDE_ERR_BAD_SHAPE;  // an output with no input tap
  const int slabs = mde_pool_slabs(B, (int64_t)Ho * Wo);
  const int pps = (Ho * Wo + slabs - 1) / slabs;
  const dim3 grid((unsigned)slabs, (unsigned)B);
  cudaStream_t st = (cudaStream_t)stream;
  if (k == 3)
    depthwise_bias_act_pool_kernel<3><<<grid, 256, 0, st>>>(x, w, bias, y, partial, Hi, Wi, C / 4, stride, pad_top, pad_left, Ho, Wo,
                                                            pps, act);
  else
    depthwise_bias_act_pool_kernel<5><<<grid, 256, 0, st>>>(x, w, bias, y, partial, Hi, Wi, C / 4, stride, pad_top, pad_left, Ho, Wo,
                                                            pps, act);
  return check_launch();
}
