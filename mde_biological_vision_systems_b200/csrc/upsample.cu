// "Next" row (f)1 -- DecoderBN's up-sampling step: bilinear (align_corners=True) resize of x to the skip tensor's size
// fused with the channel concatenation, forward and backward.
//
// Reference: UpSampleBN.forward (models/unet_adaptive_bins.py:51-54):
//   up_x = F.interpolate(x, size=skip.shape[-2:], mode='bilinear', align_corners=True);  f = cat([up_x, skip], 1)
// ATen's upsample_bilinear2d kernel takes 16.9 ms per step for the four decoder stages of config 2 (ncu launch list,
// profiles/r1_launches_step.txt) -- half of the whole forward step; this kernel writes the concatenated tensor once at
// HBM speed (one float4 store per 4 output pixels, the low-resolution source stays in L1/L2).
#include "common.cuh"

namespace mde {

__device__ __forceinline__ void up_src(int dst, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
  const float s = scale * (float)dst;  // ATen area_pixel_compute_source_index, align_corners=True
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = s - (float)i0;
  l0 = 1.f - l1;
}

// grid: (ceil(H*W / (VEC*256)), C1 + C2, B)
template <int VEC>
__global__ void __launch_bounds__(256) upsample_concat_kernel(const float* __restrict__ x, const float* __restrict__ skip,
                                                              float* __restrict__ out, int C1, int C2, int h, int w,
                                                              int H, int W, float sy, float sx) {
  const int c = blockIdx.y, b = blockIdx.z;
  const long long HW = (long long)H * W;
  const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (p >= HW) return;
  float* o = out + ((long long)b * (C1 + C2) + c) * HW + p;
  if (c >= C1) {
    const float* s = skip + ((long long)b * C2 + (c - C1)) * HW + p;
    if (VEC == 4) stg_stream(reinterpret_cast<float4*>(o), ldg_stream(reinterpret_cast<const float4*>(s)));
    else o[0] = s[0];
    return;
  }
  const float* src = x + ((long long)b * C1 + c) * h * w;
  const int y = (int)(p / W), x0p = (int)(p % W);
  int y0, y1;
  float ly0, ly1;
  up_src(y, sy, h, y0, y1, ly0, ly1);
  const float* r0 = src + y0 * w;
  const float* r1 = src + y1 * w;
  float v[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    int xa, xb;
    float lx0, lx1;
    up_src(x0p + i, sx, w, xa, xb, lx0, lx1);
    v[i] = ly0 * (lx0 * __ldg(r0 + xa) + lx1 * __ldg(r0 + xb)) + ly1 * (lx0 * __ldg(r1 + xa) + lx1 * __ldg(r1 + xb));
  }
  if (VEC == 4) stg_stream(reinterpret_cast<float4*>(o), make_float4(v[0], v[1], v[2], v[3]));
  else o[0] = v[0];
}

// Backward of the resize part, gather form (deterministic, no atomics): one thread per low-res pixel sums the
// contributions of the (<= ~3x3 for a 2x up-scale) high-res pixels whose bilinear taps include it.
// gout: [B, C1 + C2, H, W] (only the first C1 channels are read); gx: [B, C1, h, w]
__global__ void __launch_bounds__(256) upsample_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gx, int C1,
                                                           int Ctot, int h, int w, int H, int W, float sy, float sx,
                                                           float inv_sy, float inv_sx) {
  const int c = blockIdx.y, b = blockIdx.z;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= h * w) return;
  const int i = q / w, j = q % w;
  const float* g = gout + ((long long)b * Ctot + c) * H * W;
  // candidate output rows: those with source index in (i-1, i+1)
  int Y0 = (int)floorf((float)(i - 1) * inv_sy) - 1, Y1 = (int)ceilf((float)(i + 1) * inv_sy) + 1;
  int X0 = (int)floorf((float)(j - 1) * inv_sx) - 1, X1 = (int)ceilf((float)(j + 1) * inv_sx) + 1;
  if (sy == 0.f) { Y0 = 0; Y1 = H - 1; }
  if (sx == 0.f) { X0 = 0; X1 = W - 1; }
  Y0 = max(Y0, 0); Y1 = min(Y1, H - 1); X0 = max(X0, 0); X1 = min(X1, W - 1);
  float acc = 0.f;
  for (int Y = Y0; Y <= Y1; ++Y) {
    int y0, y1;
    float ly0, ly1;
    up_src(Y, sy, h, y0, y1, ly0, ly1);
    float wy = 0.f;
    if (y0 == i) wy += ly0;
    if (y1 == i) wy += ly1;
    if (wy == 0.f) continue;
    float racc = 0.f;
    for (int X = X0; X <= X1; ++X) {
      int xa, xb;
      float lx0, lx1;
      up_src(X, sx, w, xa, xb, lx0, lx1);
      float wx = 0.f;
      if (xa == j) wx += lx0;
      if (xb == j) wx += lx1;
      if (wx != 0.f) racc = fmaf(wx, g[(long long)Y * W + X], racc);
    }
    acc = fmaf(wy, racc, acc);
  }
  gx[((long long)b * C1 + c) * h * w + q] = acc;
}

}  // namespace mde

using namespace mde;

extern "C" {

static inline float up_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }

int mde_upsample_concat_fwd(const float* x, const float* skip, float* out, int B, int C1, int C2, int h, int w, int H,
                            int W, mde_stream_t stream) {
  if (!x || !out || (C2 > 0 && !skip)) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || C1 < 0 || C2 < 0 || C1 + C2 <= 0 || C1 + C2 > 65535 || B > 65535 || h <= 0 || w <= 0 || H <= 0 || W <= 0)
    return MDE_ERR_BAD_SHAPE;
  const float sy = up_scale(h, H), sx = up_scale(w, W);
  const long long HW = (long long)H * W;
  const bool vec = (W % 4 == 0) && aligned(out, 16) && (C2 == 0 || aligned(skip, 16));
  cudaStream_t st = (cudaStream_t)stream;
  if (vec) {
    dim3 grid((unsigned)((HW / 4 + 255) / 256), (unsigned)(C1 + C2), (unsigned)B);
    upsample_concat_kernel<4><<<grid, 256, 0, st>>>(x, skip, out, C1, C2, h, w, H, W, sy, sx);
  } else {
    dim3 grid((unsigned)((HW + 255) / 256), (unsigned)(C1 + C2), (unsigned)B);
    upsample_concat_kernel<1><<<grid, 256, 0, st>>>(x, skip, out, C1, C2, h, w, H, W, sy, sx);
  }
  return check_launch();
}

int mde_upsample_bwd(const float* gout, float* gx, int B, int C1, int Ctot, int h, int w, int H, int W,
                     mde_stream_t stream) {
  if (!gout || !gx) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || C1 <= 0 || C1 > Ctot || C1 > 65535 || B > 65535 || h <= 0 || w <= 0 || H <= 0 || W <= 0)
    return MDE_ERR_BAD_SHAPE;
  const float sy = up_scale(h, H), sx = up_scale(w, W);
  dim3 grid((unsigned)((h * w + 255) / 256), (unsigned)C1, (unsigned)B);
  upsample_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gout, gx, C1, Ctot, h, w, H, W, sy, sx,
                                                              sy > 0.f ? 1.f / sy : 0.f, sx > 0.f ? 1.f / sx : 0.f);
  return check_launch();
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// NCHW -> NHWC (channels_last) transpose through shared memory: reads are coalesced along pixels, writes along
// channels.  ATen's strided copy does this at ~1.8 TB/s (527 us for the 464 MB feature map of config 2); this tile
// kernel is a plain HBM stream.  grid (ceil(P/64), ceil(C/64), B), block 256.
namespace mde {
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int C,
                                                           long long P, int pitch) {
  __shared__ float tile[64][65];
  const int b = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 64;
  const int c0 = blockIdx.y * 64;
  const float* src = in + (long long)b * C * P;
  float* dst = out + (long long)b * pitch * P;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 x 4
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = c0 + ty + i * 4;
    const long long p = p0 + tx;
    tile[ty + i * 4][tx] = (c < C && p < P) ? src[(long long)c * P + p] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const long long p = p0 + ty + i * 4;
    const int c = c0 + tx;
    if (p < P && c < C) dst[p * pitch + c] = tile[tx][ty + i * 4];
  }
}
}  // namespace mde

extern "C" int mde_nchw_to_nhwc(const float* in, float* out, int B, int C, int64_t P, mde_stream_t stream) {
  if (!in || !out) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || C <= 0 || P <= 0 || B > 65535 || (C + 63) / 64 > 65535) return MDE_ERR_BAD_SHAPE;
  dim3 grid((unsigned)((P + 63) / 64), (unsigned)((C + 63) / 64), (unsigned)B);
  mde::nchw_to_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, C, P, C);
  return mde::check_launch();
}

// ---------------------------------------------------------------------------------------------------------------
// channels_last variant of the DecoderBN up-sampling step (feeds the tcgen05 conv3x3): x [B,h,w,C1] NHWC is resized
// (bilinear, align_corners=True) into channels [0,C1) of out [B,H,W,C1+C2]; the skip tensor (NHWC or NCHW) is copied /
// transposed into channels [C1, C1+C2).  One thread = 4 channels of one output pixel: float4 loads of the four taps,
// one float4 store; consecutive threads walk the channel axis, so every access is a full 128-byte line.
namespace mde {
__global__ void __launch_bounds__(256) upsample_nhwc_kernel(const float* __restrict__ x, float* __restrict__ out, int C1,
                                                            int Ctot, int h, int w, int H, int W, float sy, float sx,
                                                            long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c4 = C1 >> 2;
  const int cg = (int)(idx % c4);
  long long p = idx / c4;
  const int X = (int)(p % W);
  p /= W;
  const int Y = (int)(p % H);
  const int b = (int)(p / H);
  int y0, y1, xa, xb;
  float ly0, ly1, lx0, lx1;
  up_src(Y, sy, h, y0, y1, ly0, ly1);
  up_src(X, sx, w, xa, xb, lx0, lx1);
  const float4* src = reinterpret_cast<const float4*>(x + (long long)b * h * w * C1) + cg;
  const float4 v00 = __ldg(src + ((long long)y0 * w + xa) * c4), v01 = __ldg(src + ((long long)y0 * w + xb) * c4);
  const float4 v10 = __ldg(src + ((long long)y1 * w + xa) * c4), v11 = __ldg(src + ((long long)y1 * w + xb) * c4);
  float4 o;  // same association as ATen: ly0 * (lx0 * a + lx1 * b) + ly1 * (lx0 * c + lx1 * d)
  o.x = ly0 * (lx0 * v00.x + lx1 * v01.x) + ly1 * (lx0 * v10.x + lx1 * v11.x);
  o.y = ly0 * (lx0 * v00.y + lx1 * v01.y) + ly1 * (lx0 * v10.y + lx1 * v11.y);
  o.z = ly0 * (lx0 * v00.z + lx1 * v01.z) + ly1 * (lx0 * v10.z + lx1 * v11.z);
  o.w = ly0 * (lx0 * v00.w + lx1 * v01.w) + ly1 * (lx0 * v10.w + lx1 * v11.w);
  stg_stream(reinterpret_cast<float4*>(out + (((long long)b * H + Y) * W + X) * Ctot) + cg, o);
}

__global__ void __launch_bounds__(256) copy_channels_nhwc_kernel(const float* __restrict__ skip, float* __restrict__ out,
                                                                 int C2, int Ctot, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c4 = C2 >> 2;
  const int cg = (int)(idx % c4);
  const long long p = idx / c4;
  stg_stream(reinterpret_cast<float4*>(out + p * Ctot) + cg, ldg_stream(reinterpret_cast<const float4*>(skip + p * C2) + cg));
}
}  // namespace mde

extern "C" int mde_upsample_concat_nhwc_fwd(const float* x_nhwc, const float* skip, int skip_channels_last,
                                            float* out_nhwc, int B, int C1, int C2, int h, int w, int H, int W,
                                            mde_stream_t stream) {
  using namespace mde;
  if (!x_nhwc || !out_nhwc || (C2 > 0 && !skip)) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || C1 <= 0 || C2 < 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0 || B > 65535) return MDE_ERR_BAD_SHAPE;
  if (C1 % 4 != 0 || C2 % 4 != 0 || !aligned(x_nhwc, 16) || !aligned(out_nhwc, 16) || (C2 > 0 && !aligned(skip, 16)))
    return MDE_ERR_UNSUPPORTED;
  const int Ctot = C1 + C2;
  const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)B * H * W * (C1 / 4);
  upsample_nhwc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x_nhwc, out_nhwc, C1, Ctot, h, w, H, W, sy, sx,
                                                                        total);
  int rc = check_launch();
  if (rc || C2 == 0) return rc;
  const long long P = (long long)H * W;
  if (skip_channels_last) {
    const long long t2 = (long long)B * P * (C2 / 4);
    copy_channels_nhwc_kernel<<<(unsigned)((t2 + 255) / 256), 256, 0, st>>>(skip, out_nhwc + C1, C2, Ctot, t2);
  } else {
    dim3 grid((unsigned)((P + 63) / 64), (unsigned)((C2 + 63) / 64), (unsigned)B);
    nchw_to_nhwc_kernel<<<grid, 256, 0, st>>>(skip, out_nhwc + C1, C2, P, Ctot);
  }
  return check_launch();
}
