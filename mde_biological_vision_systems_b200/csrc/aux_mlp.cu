// A3 -- per-pixel 1x1-conv MLP  C_in -> 10 -> 10 (ReLU after both), forward and backward.
//
// Reference: nn.Sequential(Conv2d(C_in,10,1), ReLU, Conv2d(10,10,1), ReLU) built at
// models/unet_adaptive_bins.py:146-174 and applied to area / size channels at :196-228.
// HBM-bound streaming kernel: 4*C_in bytes in, 40 bytes out per pixel; the 150 (C_in = 3) weights live in
// registers/shared memory; one thread owns 4 consecutive pixels (128-bit loads and stores per plane).
// The output pointer addresses a channel slice of the concatenated encoder input, so no torch.cat copy follows.
#include "common.cuh"

namespace mde {

constexpr int HID = 10;

template <int CIN>
struct MlpWeights {
  float w0[HID * CIN], b0[HID], w1[HID * HID], b1[HID];
};

template <int CIN>
__device__ __forceinline__ void load_weights(MlpWeights<CIN>& s, const float* w0, const float* b0, const float* w1,
                                             const float* b1) {
  for (int i = threadIdx.x; i < HID * CIN; i += blockDim.x) s.w0[i] = w0[i];
  for (int i = threadIdx.x; i < HID * HID; i += blockDim.x) s.w1[i] = w1[i];
  for (int i = threadIdx.x; i < HID; i += blockDim.x) {
    s.b0[i] = b0[i];
    s.b1[i] = b1[i];
  }
  __syncthreads();
}

template <int CIN, int VEC>
__global__ void __launch_bounds__(256) aux_mlp_fwd_kernel(const float* __restrict__ x, long long xbs,
                                                          const float* __restrict__ w0, const float* __restrict__ b0,
                                                          const float* __restrict__ w1, const float* __restrict__ b1,
                                                          float* __restrict__ out, long long obs, long long HW,
                                                          float in_scale) {
  __shared__ MlpWeights<CIN> sw;
  load_weights<CIN>(sw, w0, b0, w1, b1);
  const int b = blockIdx.y;
  const float* xb = x + (long long)b * xbs;
  float* ob = out + (long long)b * obs;
  const long long groups = HW / VEC;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups;
       g += (long long)gridDim.x * blockDim.x) {
    const long long p = g * VEC;
    float xin[CIN][VEC];
#pragma unroll
    for (int c = 0; c < CIN; ++c) {
      if (VEC == 4) {
        const float4 v = ldg_stream(reinterpret_cast<const float4*>(xb + (long long)c * HW + p));
        xin[c][0] = v.x / in_scale; xin[c][1] = v.y / in_scale; xin[c][2] = v.z / in_scale; xin[c][3] = v.w / in_scale;
      } else {
        xin[c][0] = xb[(long long)c * HW + p] / in_scale;
      }
    }
    float h[HID][VEC];
#pragma unroll
    for (int j = 0; j < HID; ++j)
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float a = sw.b0[j];
#pragma unroll
        for (int c = 0; c < CIN; ++c) a = fmaf(sw.w0[j * CIN + c], xin[c][v], a);
        h[j][v] = fmaxf(a, 0.f);
      }
#pragma unroll
    for (int i = 0; i < HID; ++i) {
      float o[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float a = sw.b1[i];
#pragma unroll
        for (int j = 0; j < HID; ++j) a = fmaf(sw.w1[i * HID + j], h[j][v], a);
        o[v] = fmaxf(a, 0.f);
      }
      if (VEC == 4) stg_stream(reinterpret_cast<float4*>(ob + (long long)i * HW + p), make_float4(o[0], o[1], o[2], o[3]));
      else ob[(long long)i * HW + p] = o[0];
    }
  }
}

// Backward: recompute the hidden layer, accumulate parameter gradients per thread, reduce per block, one atomic
// per parameter per block.
template <int CIN>
__global__ void __launch_bounds__(256) aux_mlp_bwd_kernel(const float* __restrict__ x, long long xbs,
                                                          const float* __restrict__ w0, const float* __restrict__ b0,
                                                          const float* __restrict__ w1, const float* __restrict__ b1,
                                                          const float* __restrict__ gout, long long gbs, float* gx,
                                                          float* gw0, float* gb0, float* gw1, float* gb1, int B,
                                                          long long HW, float in_scale) {
  __shared__ MlpWeights<CIN> sw;
  load_weights<CIN>(sw, w0, b0, w1, b1);
  float aw0[HID * CIN], ab0[HID], aw1[HID * HID], ab1[HID];
#pragma unroll
  for (int i = 0; i < HID * CIN; ++i) aw0[i] = 0.f;
#pragma unroll
  for (int i = 0; i < HID * HID; ++i) aw1[i] = 0.f;
#pragma unroll
  for (int i = 0; i < HID; ++i) ab0[i] = ab1[i] = 0.f;
  const long long total = (long long)B * HW;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long b = t / HW, p = t - b * HW;
    float xin[CIN];
#pragma unroll
    for (int c = 0; c < CIN; ++c) xin[c] = x[b * xbs + (long long)c * HW + p] / in_scale;
    float h[HID];
#pragma unroll
    for (int j = 0; j < HID; ++j) {
      float a = sw.b0[j];
#pragma unroll
      for (int c = 0; c < CIN; ++c) a = fmaf(sw.w0[j * CIN + c], xin[c], a);
      h[j] = fmaxf(a, 0.f);
    }
    float gh[HID];
#pragma unroll
    for (int j = 0; j < HID; ++j) gh[j] = 0.f;
#pragma unroll
    for (int i = 0; i < HID; ++i) {
      float a = sw.b1[i];
#pragma unroll
      for (int j = 0; j < HID; ++j) a = fmaf(sw.w1[i * HID + j], h[j], a);
      const float go = a > 0.f ? gout[b * gbs + (long long)i * HW + p] : 0.f;
      ab1[i] += go;
#pragma unroll
      for (int j = 0; j < HID; ++j) {
        aw1[i * HID + j] = fmaf(go, h[j], aw1[i * HID + j]);
        gh[j] = fmaf(sw.w1[i * HID + j], go, gh[j]);
      }
    }
    float gxin[CIN];
#pragma unroll
    for (int c = 0; c < CIN; ++c) gxin[c] = 0.f;
#pragma unroll
    for (int j = 0; j < HID; ++j) {
      const float g = h[j] > 0.f ? gh[j] : 0.f;
      ab0[j] += g;
#pragma unroll
      for (int c = 0; c < CIN; ++c) {
        aw0[j * CIN + c] = fmaf(g, xin[c], aw0[j * CIN + c]);
        gxin[c] = fmaf(sw.w0[j * CIN + c], g, gxin[c]);
      }
    }
    if (gx) {
#pragma unroll
      for (int c = 0; c < CIN; ++c) gx[b * xbs + (long long)c * HW + p] = gxin[c] / in_scale;
    }
  }
  // block reduction: warp shuffle, then shared, then one atomic per parameter
  __shared__ float red[8];
  auto block_add = [&](float v, float* dst) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
      if (s != 0.f) atomicAdd(dst, s);
    }
    __syncthreads();
  };
#pragma unroll
  for (int i = 0; i < HID * CIN; ++i) block_add(aw0[i], gw0 + i);
#pragma unroll
  for (int i = 0; i < HID; ++i) block_add(ab0[i], gb0 + i);
#pragma unroll
  for (int i = 0; i < HID * HID; ++i) block_add(aw1[i], gw1 + i);
#pragma unroll
  for (int i = 0; i < HID; ++i) block_add(ab1[i], gb1 + i);
}

}  // namespace mde

using namespace mde;

extern "C" {

int mde_aux_mlp_fwd(const float* x, int64_t xbs, const float* w0, const float* b0, const float* w1, const float* b1,
                    float* out, int64_t obs, int B, int C_in, int H1, int H2, int64_t HW, float in_scale,
                    mde_stream_t stream) {
  if (!x || !w0 || !b0 || !w1 || !b1 || !out) return MDE_ERR_BAD_POINTER;
  if (H1 != HID || H2 != HID || (C_in != 1 && C_in != 3)) return MDE_ERR_UNSUPPORTED;
  if (B <= 0 || HW <= 0 || B > 65535) return MDE_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (HW % 4 == 0) && (xbs % 4 == 0) && (obs % 4 == 0) && aligned(x, 16) && aligned(out, 16);
  const long long groups = vec ? HW / 4 : HW;
  long long gx = (groups + 255) / 256;
  if (gx > 4096) gx = 4096;
  dim3 grid((unsigned)gx, (unsigned)B);
  if (C_in == 1) {
    if (vec) aux_mlp_fwd_kernel<1, 4><<<grid, 256, 0, st>>>(x, xbs, w0, b0, w1, b1, out, obs, HW, in_scale);
    else aux_mlp_fwd_kernel<1, 1><<<grid, 256, 0, st>>>(x, xbs, w0, b0, w1, b1, out, obs, HW, in_scale);
  } else {
    if (vec) aux_mlp_fwd_kernel<3, 4><<<grid, 256, 0, st>>>(x, xbs, w0, b0, w1, b1, out, obs, HW, in_scale);
    else aux_mlp_fwd_kernel<3, 1><<<grid, 256, 0, st>>>(x, xbs, w0, b0, w1, b1, out, obs, HW, in_scale);
  }
  return check_launch();
}

int mde_aux_mlp_bwd(const float* x, int64_t xbs, const float* w0, const float* b0, const float* w1, const float* b1,
                    const float* gout, int64_t gbs, float* gx, float* gw0, float* gb0, float* gw1, float* gb1, int B,
                    int C_in, int H1, int H2, int64_t HW, float in_scale, mde_stream_t stream) {
  if (!x || !w0 || !b0 || !w1 || !b1 || !gout || !gw0 || !gb0 || !gw1 || !gb1) return MDE_ERR_BAD_POINTER;
  if (H1 != HID || H2 != HID || (C_in != 1 && C_in != 3)) return MDE_ERR_UNSUPPORTED;
  if (B <= 0 || HW <= 0) return MDE_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = MDE_NUM_SMS * 4;
  if (C_in == 1)
    aux_mlp_bwd_kernel<1><<<grid, 256, 0, st>>>(x, xbs, w0, b0, w1, b1, gout, gbs, gx, gw0, gb0, gw1, gb1, B, HW, in_scale);
  else
    aux_mlp_bwd_kernel<3><<<grid, 256, 0, st>>>(x, xbs, w0, b0, w1, b1, gout, gbs, gx, gw0, gb0, gw1, gb1, B, HW, in_scale);
  return check_launch();
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// Per-channel bias (+ SiLU) (+ residual) on channels_last activations, in place or out of place: the epilogue of a
// convolution whose eval-mode BatchNorm has been folded into its filter (models/efficientnet.py conv_bn).  Without it the
// folded bias costs one extra elementwise pass per convolution and the activation another.
// y[p, c] = act(x[p, c] + bias[c]) + (res ? res[p, c] : 0);   act: 0 identity, 1 SiLU
namespace mde {
__global__ void __launch_bounds__(256) bias_act_nhwc_kernel(const float* __restrict__ x, const float* __restrict__ bias,
                                                            const float* __restrict__ res, float* __restrict__ y,
                                                            long long total4, int c4, int act) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % c4);
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + cg);
    float o[4] = {v.x + b.x, v.y + b.y, v.z + b.z, v.w + b.w};
    if (act == 1) {
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = o[k] / (1.f + __expf(-o[k]));
    }
    if (res != nullptr) {
      const float4 r = reinterpret_cast<const float4*>(res)[i];
      o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w;
    }
    reinterpret_cast<float4*>(y)[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
}
}  // namespace mde

extern "C" int mde_bias_act_nhwc(const float* x, const float* bias, const float* residual, float* y, int64_t pixels, int C,
                                 int act, mde_stream_t stream) {
  using namespace mde;
  if (!x || !bias || !y) return MDE_ERR_BAD_POINTER;
  if (pixels <= 0 || C <= 0 || act < 0 || act > 1) return MDE_ERR_BAD_SHAPE;
  if (C % 4 != 0 || !aligned(x, 16) || !aligned(y, 16) || !aligned(bias, 16) || (residual && !aligned(residual, 16)))
    return MDE_ERR_UNSUPPORTED;
  const long long total4 = pixels * (C / 4);
  long long gx = (total4 + 255) / 256;
  if (gx > MDE_NUM_SMS * 16) gx = MDE_NUM_SMS * 16;
  bias_act_nhwc_kernel<<<(unsigned)gx, 256, 0, (cudaStream_t)stream>>>(x, bias, residual, y, total4, C / 4, act);
  return check_launch();
}

// Same epilogue, written into a zero-padded tensor: y [B, H + pt + pb, W + pl + pr, C] = pad(act(x + bias)).  Used when the
// consumer is a stride-2 TensorFlow-"SAME" convolution (asymmetric padding, models/efficientnet.py SamePadConv2d): the
// padding copy the reference's F.pad would make is folded into this pass.
namespace mde {
__global__ void __launch_bounds__(256) bias_act_pad_nhwc_kernel(const float* __restrict__ x, const float* __restrict__ bias,
                                                                float* __restrict__ y, int H, int W, int c4, int pt, int pl,
                                                                int Ho, int Wo, long long total4, int act) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % c4);
    long long p = i / c4;
    const int xo = (int)(p % Wo);
    p /= Wo;
    const int yo = (int)(p % Ho);
    const long long b = p / Ho;
    const int yi = yo - pt, xi = xo - pl;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (yi >= 0 && yi < H && xi >= 0 && xi < W) {
      const float4 v = reinterpret_cast<const float4*>(x)[((b * H + yi) * W + xi) * c4 + cg];
      const float4 bb = __ldg(reinterpret_cast<const float4*>(bias) + cg);
      float t[4] = {v.x + bb.x, v.y + bb.y, v.z + bb.z, v.w + bb.w};
      if (act == 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) t[k] = t[k] / (1.f + __expf(-t[k]));
      }
      o = make_float4(t[0], t[1], t[2], t[3]);
    }
    reinterpret_cast<float4*>(y)[i] = o;
  }
}
}  // namespace mde

extern "C" int mde_bias_act_pad_nhwc(const float* x, const float* bias, float* y, int B, int H, int W, int C, int pad_top,
                                     int pad_bottom, int pad_left, int pad_right, int act, mde_stream_t stream) {
  using namespace mde;
  if (!x || !bias || !y) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || pad_top < 0 || pad_bottom < 0 || pad_left < 0 || pad_right < 0 || act < 0 ||
      act > 1)
    return MDE_ERR_BAD_SHAPE;
  if (C % 4 != 0 || !aligned(x, 16) || !aligned(y, 16) || !aligned(bias, 16)) return MDE_ERR_UNSUPPORTED;
  const int Ho = H + pad_top + pad_bottom, Wo = W + pad_left + pad_right;
  const long long total4 = (long long)B * Ho * Wo * (C / 4);
  long long gx = (total4 + 255) / 256;
  if (gx > MDE_NUM_SMS * 16) gx = MDE_NUM_SMS * 16;
  bias_act_pad_nhwc_kernel<<<(unsigned)gx, 256, 0, (cudaStream_t)stream>>>(x, bias, y, H, W, C / 4, pad_top, pad_left, Ho, Wo,
                                                                          total4, act);
  return check_launch();
}

// ---------------------------------------------------------------------------------------------------------------
// Squeeze-excite of the MBConv blocks (geffnet SqueezeExcite: x * sigmoid(conv_expand(silu(conv_reduce(mean_hw(x)))))) in
// inference: the spatial mean rides on the bias + SiLU pass that follows the depthwise convolution, the two tiny 1x1
// convolutions run as one kernel, and the gate is multiplied into the activations inside the projection GEMM
// (mde_pointwise_x3_fwd) -- 4 launches and 1 read + 1 write of the expanded tensor per block instead of 10 launches and 4
// reads + 2 writes.
namespace mde {

// grid (slabs, B).  A slab is a contiguous run of rows, i.e. a contiguous run of rows * C/4 float4s.  The block's first
// c4 * RY threads (RY = 256 / c4 whole rows per pass) walk it with stride c4 * RY: consecutive threads touch consecutive
// float4s, every thread keeps ONE channel group (the stride is a multiple of c4), so its running sum needs no indexing.
// C > 1024 (c4 > 256): the channel groups are covered in passes of 256.
__global__ void __launch_bounds__(256) bias_act_pool_nhwc_kernel(const float* __restrict__ x, const float* __restrict__ bias,
                                                                 float* __restrict__ y, float* __restrict__ partial,
                                                                 long long HW, int c4, long long rows_per_slab, int act) {
  __shared__ float4 red[256];
  pdl_sync();
  const long long b = blockIdx.y, slab = blockIdx.x;
  const long long r0 = slab * rows_per_slab;
  const long long r1 = r0 + rows_per_slab < HW ? r0 + rows_per_slab : HW;
  const float4* xs = reinterpret_cast<const float4*>(x) + (b * HW + r0) * c4;
  float4* ys = reinterpret_cast<float4*>(y) + (b * HW + r0) * c4;
  const int t = threadIdx.x;
  for (int cg0 = 0; cg0 < c4; cg0 += 256) {
    const int cw = min(256, c4 - cg0);           // channel groups of this pass
    const int ry = 256 / cw;                     // rows per sweep of the block
    const int cg = cg0 + t % cw, row = t / cw;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < ry) {
      const float4 bb = __ldg(reinterpret_cast<const float4*>(bias) + cg);
#pragma unroll 4
      for (long long r = row; r < r1 - r0; r += ry) {
        const long long idx = r * c4 + cg;
        const float4 v = ldg_stream(xs + idx);
        float o[4] = {v.x + bb.x, v.y + bb.y, v.z + bb.z, v.w + bb.w};
        if (act == 1) {
#pragma unroll
          for (int k = 0; k < 4; ++k) o[k] = __fdividef(o[k], 1.f + __expf(-o[k]));
        }
        ys[idx] = make_float4(o[0], o[1], o[2], o[3]);
        acc.x += o[0]; acc.y += o[1]; acc.z += o[2]; acc.w += o[3];
      }
    }
    red[t] = acc;
    __syncthreads();
    if (t < cw) {  // fixed order: rows 0 .. ry-1 of the sweep
      for (int k = 1; k < ry; ++k) {
        const float4 u = red[t + k * cw];
        acc.x += u.x; acc.y += u.y; acc.z += u.z; acc.w += u.w;
      }
      reinterpret_cast<float4*>(partial)[(b * gridDim.x + slab) * c4 + cg] = acc;
    }
    __syncthreads();
  }
}

// grid (splits, B): every block recomputes the squeezed vector h (R <= 256 values), then fills its slice of the gate
__global__ void __launch_bounds__(256) se_gate_kernel(const float* __restrict__ partial, int slabs, float inv_hw,
                                                      const float* __restrict__ w1, const float* __restrict__ b1,
                                                      const float* __restrict__ w2, const float* __restrict__ b2,
                                                      float* __restrict__ gate, int C, int R) {
  extern __shared__ __align__(16) float se_sm[];  // mean[C] | h[R]
  pdl_sync();
  float* mean = se_sm;
  float* h = se_sm + C;
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float* pp = partial + (long long)b * slabs * C + c;
    float sacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // eight loads in flight; fixed association -> reproducible
    int i = 0;
    for (; i + 7 < slabs; i += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) sacc[u] += pp[(long long)(i + u) * C];
    }
    for (; i < slabs; ++i) sacc[0] += pp[(long long)i * C];
    mean[c] = (((sacc[0] + sacc[1]) + (sacc[2] + sacc[3])) + ((sacc[4] + sacc[5]) + (sacc[6] + sacc[7]))) * inv_hw;
  }
  __syncthreads();
  // squeeze: h[r] = silu(w1[r, :] . mean + b1[r]).  The rows of w1 come from L2 (~600 clocks): a warp works on TWO rows at once
  // with 16-byte loads, eight in flight per row, so a row costs one or two round trips instead of C / 32.
  const bool v4 = (C & 3) == 0 && (reinterpret_cast<uintptr_t>(w1) & 15) == 0;
  for (int r = warp; r < R; r += 16) {
    const int r2 = r + 8 < R ? r + 8 : r;
    float da = 0.f, db = 0.f;
    if (v4) {
      const float4* wa = reinterpret_cast<const float4*>(w1 + (long long)r * C);
      const float4* wb = reinterpret_cast<const float4*>(w1 + (long long)r2 * C);
      const float4* m4 = reinterpret_cast<const float4*>(mean);
#pragma unroll 8
      for (int c = lane; c < (C >> 2); c += 32) {
        const float4 a = __ldg(wa + c), bq = __ldg(wb + c), m = m4[c];
        da = fmaf(a.x, m.x, fmaf(a.y, m.y, fmaf(a.z, m.z, fmaf(a.w, m.w, da))));
        db = fmaf(bq.x, m.x, fmaf(bq.y, m.y, fmaf(bq.z, m.z, fmaf(bq.w, m.w, db))));
      }
    } else {
#pragma unroll 4
      for (int c = lane; c < C; c += 32) {
        da = fmaf(__ldg(w1 + (long long)r * C + c), mean[c], da);
        db = fmaf(__ldg(w1 + (long long)r2 * C + c), mean[c], db);
      }
    }
    da = warp_sum(da);
    db = warp_sum(db);
    if (lane == 0) {
      da += b1 ? b1[r] : 0.f;
      h[r] = da / (1.f + expf(-da));
      if (r2 != r) {
        db += b1 ? b1[r2] : 0.f;
        h[r2] = db / (1.f + expf(-db));
      }
    }
  }
  __syncthreads();
  const int per = (C + gridDim.x - 1) / gridDim.x;
  const int c_lo = blockIdx.x * per, c_hi = min(C, c_lo + per);
  const bool r4 = (R & 3) == 0 && (reinterpret_cast<uintptr_t>(w2) & 15) == 0;
  for (int c = c_lo + threadIdx.x; c < c_hi; c += blockDim.x) {
    const float* wr = w2 + (long long)c * R;
    float d0 = b2 ? b2[c] : 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
    if (r4) {
#pragma unroll 8
      for (int r = 0; r < R; r += 4) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(wr + r));
        d0 = fmaf(w.x, h[r], d0);
        d1 = fmaf(w.y, h[r + 1], d1);
        d2 = fmaf(w.z, h[r + 2], d2);
        d3 = fmaf(w.w, h[r + 3], d3);
      }
    } else {
#pragma unroll 8
      for (int r = 0; r < R; ++r) d0 = fmaf(__ldg(wr + r), h[r], d0);
    }
    const float d = (d0 + d1) + (d2 + d3);
    gate[(long long)b * C + c] = 1.f / (1.f + expf(-d));
  }
}
}  // namespace mde

extern "C" int mde_pool_slabs(int B, int64_t HW) {
  if (B <= 0 || HW <= 0) return 0;
  long long s = (MDE_NUM_SMS * 4 + B - 1) / B;  // ~4 blocks per SM over the batch
  const long long cap = (HW + 7) / 8;            // >= 8 rows per slab
  if (s > cap) s = cap;
  if (s < 1) s = 1;
  if (s > 256) s = 256;
  return (int)s;
}

extern "C" int mde_bias_act_pool_nhwc(const float* x, const float* bias, float* y, float* partial, int B, int64_t HW, int C,
                                      int act, mde_stream_t stream) {
  using namespace mde;
  if (!x || !bias || !y || !partial) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || B > 65535 || HW <= 0 || C <= 0 || act < 0 || act > 1) return MDE_ERR_BAD_SHAPE;
  if (C % 4 != 0 || !aligned(x, 16) || !aligned(y, 16) || !aligned(bias, 16) || !aligned(partial, 16)) return MDE_ERR_UNSUPPORTED;
  const int slabs = mde_pool_slabs(B, HW);
  const long long rps = (HW + slabs - 1) / slabs;
  launch_pdl(PDL_STREAM, bias_act_pool_nhwc_kernel, dim3((unsigned)slabs, (unsigned)B), dim3(256), 0, (cudaStream_t)stream, x, bias, y, partial,
             HW, C / 4, rps, act);
  return check_launch();
}

extern "C" int mde_se_gate(const float* partial, int slabs, float inv_hw, const float* w1, const float* b1, const float* w2,
                           const float* b2, float* gate, int B, int C, int R, mde_stream_t stream) {
  using namespace mde;
  if (!partial || !w1 || !w2 || !gate) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || B > 65535 || C <= 0 || R <= 0 || slabs <= 0) return MDE_ERR_BAD_SHAPE;
  const size_t sm = (size_t)(C + R) * sizeof(float);
  if (sm > 48 * 1024) return MDE_ERR_UNSUPPORTED;
  const int splits = C >= 512 ? 8 : (C >= 128 ? 4 : 1);
  launch_pdl(PDL_STREAM, se_gate_kernel, dim3((unsigned)splits, (unsigned)B), dim3(256), sm, (cudaStream_t)stream, partial, slabs, inv_hw, w1,
             b1, w2, b2, gate, C, R);
  return check_launch();
}
