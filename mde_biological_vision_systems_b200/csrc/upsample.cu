// "Next" row (f)1 -- DecoderBN's up-sampling step: bilinear (align_corners=True) resize of x to the skip tensor's size
// fused with the channel concatenation, forward and backward.
//
// Reference: UpSampleBN.forward (models/unet_adaptive_bins.py:51-54):
//   up_x = F.interpolate(x, size=skip.shape[-2:], mode='bilinear', align_corners=True);  f = cat([up_x, skip], 1)
// ATen's upsample_bilinear2d kernel takes 16.9 ms per step for the four decoder stages of config 2 (ncu launch list,
// profiles/r1_launches_step.txt) -- half of the whole forward step; this kernel writes the concatenated tensor once at
// HBM speed (one float4 store per 4 output pixels, the low-resolution source stays in L1/L2).
#include "common.cuh"
#include "tc_common.cuh"

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

// Backward of the resize part, gather form (deterministic, no atomics).  A prologue kernel inverts the 1-D bilinear
// maps once per call: for every low-res index the (<= UP_TAPS) high-res indices whose taps touch it and their weights;
// the main kernels then sum wy * wx * gout over that small stencil.
constexpr int UP_TAPS = 8;
struct UpTap {
  int idx[UP_TAPS];   // high-res index, -1 = unused
  float wt[UP_TAPS];
};

__global__ void upsample_taps_kernel(UpTap* __restrict__ taps, int in_size, int out_size, float scale, float inv_scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= in_size) return;
  int lo = (int)floorf((float)(i - 1) * inv_scale) - 1, hi = (int)ceilf((float)(i + 1) * inv_scale) + 1;
  if (scale == 0.f) { lo = 0; hi = out_size - 1; }
  lo = max(lo, 0);
  hi = min(hi, out_size - 1);
  UpTap t;
  int n = 0;
#pragma unroll
  for (int k = 0; k < UP_TAPS; ++k) { t.idx[k] = -1; t.wt[k] = 0.f; }
  for (int Y = lo; Y <= hi; ++Y) {
    int a, b;
    float la, lb;
    up_src(Y, scale, in_size, a, b, la, lb);
    float wv = 0.f;
    if (a == i) wv += la;
    if (b == i) wv += lb;
    if (wv != 0.f && n < UP_TAPS) {
      t.idx[n] = Y;
      t.wt[n] = wv;
      ++n;
    }
  }
  taps[i] = t;
}

// gout: [B, Ctot, H, W] (only the first C1 channels are read); gx: [B, C1, h, w]
__global__ void __launch_bounds__(256) upsample_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gx, int C1,
                                                           int Ctot, int h, int w, int H, int W,
                                                           const UpTap* __restrict__ ty, const UpTap* __restrict__ tx) {
  const int c = blockIdx.y, b = blockIdx.z;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= h * w) return;
  const int i = q / w, j = q - i * w;
  const float* g = gout + ((long long)b * Ctot + c) * H * W;
  const UpTap a = ty[i], bt = tx[j];
  float acc = 0.f;
#pragma unroll
  for (int u = 0; u < UP_TAPS; ++u) {
    if (a.idx[u] < 0) break;
    const float* row = g + (long long)a.idx[u] * W;
    float racc = 0.f;
#pragma unroll
    for (int v = 0; v < UP_TAPS; ++v) {
      if (bt.idx[v] < 0) break;
      racc = fmaf(bt.wt[v], __ldg(row + bt.idx[v]), racc);
    }
    acc = fmaf(a.wt[u], racc, acc);
  }
  gx[((long long)b * C1 + c) * h * w + q] = acc;
}

// channels_last: gout [B, H, W, Ctot] (channels [0, C1) read), gx [B, h, w, C1]; one thread = 4 channels of one pixel
__global__ void __launch_bounds__(256) upsample_bwd_nhwc_kernel(const float* __restrict__ gout, float* __restrict__ gx,
                                                                int C1, int Ctot, int h, int w, int H, int W,
                                                                const UpTap* __restrict__ ty, const UpTap* __restrict__ tx,
                                                                long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c4 = C1 >> 2;
  const int cg = (int)(idx % c4);
  long long p = idx / c4;
  const int j = (int)(p % w);
  p /= w;
  const int i = (int)(p % h);
  const int b = (int)(p / h);
  const UpTap a = ty[i], bt = tx[j];
  const float* g = gout + (long long)b * H * W * Ctot + cg * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int u = 0; u < UP_TAPS; ++u) {
    if (a.idx[u] < 0) break;
    const float* row = g + (long long)a.idx[u] * W * Ctot;
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int v = 0; v < UP_TAPS; ++v) {
      if (bt.idx[v] < 0) break;
      const float4 gv = __ldg(reinterpret_cast<const float4*>(row + (long long)bt.idx[v] * Ctot));
      r.x = fmaf(bt.wt[v], gv.x, r.x); r.y = fmaf(bt.wt[v], gv.y, r.y);
      r.z = fmaf(bt.wt[v], gv.z, r.z); r.w = fmaf(bt.wt[v], gv.w, r.w);
    }
    acc.x = fmaf(a.wt[u], r.x, acc.x); acc.y = fmaf(a.wt[u], r.y, acc.y);
    acc.z = fmaf(a.wt[u], r.z, acc.z); acc.w = fmaf(a.wt[u], r.w, acc.w);
  }
  reinterpret_cast<float4*>(gx + (((long long)b * h + i) * w + j) * C1)[cg] = acc;
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

int64_t mde_upsample_bwd_ws_bytes(int h, int w) { return (int64_t)sizeof(UpTap) * ((int64_t)h + w); }

static int launch_taps(void* ws, int h, int w, int H, int W, cudaStream_t st) {
  const float sy = up_scale(h, H), sx = up_scale(w, W);
  UpTap* ty = reinterpret_cast<UpTap*>(ws);
  upsample_taps_kernel<<<(h + 127) / 128, 128, 0, st>>>(ty, h, H, sy, sy > 0.f ? 1.f / sy : 0.f);
  int rc = check_launch();
  if (rc) return rc;
  upsample_taps_kernel<<<(w + 127) / 128, 128, 0, st>>>(ty + h, w, W, sx, sx > 0.f ? 1.f / sx : 0.f);
  return check_launch();
}

int mde_upsample_bwd(const float* gout, float* gx, int channels_last, int B, int C1, int Ctot, int h, int w, int H, int W,
                     void* ws, mde_stream_t stream) {
  if (!gout || !gx || !ws) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || C1 <= 0 || C1 > Ctot || C1 > 65535 || B > 65535 || h <= 0 || w <= 0 || H <= 0 || W <= 0)
    return MDE_ERR_BAD_SHAPE;
  // a low-res index is touched by at most 2 * (H / h) + 1 high-res indices; the tap tables hold UP_TAPS
  if (2LL * H > (long long)(UP_TAPS - 1) * h || 2LL * W > (long long)(UP_TAPS - 1) * w) return MDE_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = launch_taps(ws, h, w, H, W, st);
  if (rc) return rc;
  const UpTap* ty = reinterpret_cast<const UpTap*>(ws);
  if (channels_last) {
    if (C1 % 4 != 0 || Ctot % 4 != 0 || !aligned(gout, 16) || !aligned(gx, 16)) return MDE_ERR_UNSUPPORTED;
    const long long total = (long long)B * h * w * (C1 / 4);
    upsample_bwd_nhwc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(gout, gx, C1, Ctot, h, w, H, W, ty, ty + h,
                                                                             total);
  } else {
    dim3 grid((unsigned)((h * w + 255) / 256), (unsigned)C1, (unsigned)B);
    upsample_bwd_kernel<<<grid, 256, 0, st>>>(gout, gx, C1, Ctot, h, w, H, W, ty, ty + h);
  }
  return check_launch();
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// NCHW -> NHWC (channels_last) transpose through shared memory: reads are coalesced along pixels, writes along
// channels.  ATen's strided copy does this at ~1.8 TB/s (527 us for the 464 MB feature map of config 2); this tile
// kernel is a plain HBM stream.  grid (ceil(P/64), ceil(C/64), B), block 256.
namespace mde {
// W > 0: the output is a (zero-)padded image of Wo x (Po / Wo) pixels and the source pixel (y, x) lands at (y + pt, x + pl)
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int C,
                                                           long long P, int pitch, int W = 0, int Wo = 0, long long Po = 0,
                                                           int pt = 0, int pl = 0) {
  __shared__ float tile[64][65];
  const int b = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 64;
  const int c0 = blockIdx.y * 64;
  const float* src = in + (long long)b * C * P;
  float* dst = out + (long long)b * pitch * (W > 0 ? Po : P);
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
    if (p < P && c < C) {
      const long long po = W > 0 ? ((p / W + pt) * Wo + (p % W + pl)) : p;
      dst[po * pitch + c] = tile[tx][ty + i * 4];
    }
  }
}
// C <= 4 (the RGB planes of the encoder input): one thread per pixel reads the C planes (coalesced across the warp) and writes C
// consecutive floats -- the 64 x 64 tile kernel above would idle 61 of its 64 channel lanes.
__global__ void __launch_bounds__(256) nchw_to_nhwc_smallc_kernel(const float* __restrict__ in, float* __restrict__ out, int C,
                                                                  long long P, int pitch, int W, int Wo, long long Po, int pt,
                                                                  int pl) {
  const int b = blockIdx.y;
  const float* src = in + (long long)b * C * P;
  float* dst = out + (long long)b * pitch * (W > 0 ? Po : P);
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const long long po = W > 0 ? ((p / W + pt) * Wo + (p % W + pl)) : p;
    for (int c = 0; c < C; ++c) dst[po * pitch + c] = __ldg(src + (long long)c * P + p);
  }
}

static int launch_nchw_to_nhwc(const float* in, float* out, int B, int C, long long P, int pitch, int W, int Wo, long long Po,
                               int pt, int pl, cudaStream_t st) {
  if (C <= 4) {
    long long gx = (P + 255) / 256;
    const long long cap = (MDE_NUM_SMS * 8 + B - 1) / B;
    if (gx > cap) gx = cap;
    nchw_to_nhwc_smallc_kernel<<<dim3((unsigned)gx, (unsigned)B), 256, 0, st>>>(in, out, C, P, pitch, W, Wo, Po, pt, pl);
  } else {
    dim3 grid((unsigned)((P + 63) / 64), (unsigned)((C + 63) / 64), (unsigned)B);
    nchw_to_nhwc_kernel<<<grid, 256, 0, st>>>(in, out, C, P, pitch, W, Wo, Po, pt, pl);
  }
  return check_launch();
}
}  // namespace mde

extern "C" int mde_nchw_to_nhwc(const float* in, float* out, int B, int C, int64_t P, mde_stream_t stream) {
  if (!in || !out) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || C <= 0 || P <= 0 || B > 65535 || (C + 63) / 64 > 65535) return MDE_ERR_BAD_SHAPE;
  return mde::launch_nchw_to_nhwc(in, out, B, C, P, C, 0, 0, 0, 0, 0, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------------------
// channels_last variant of the DecoderBN up-sampling step (feeds the tcgen05 conv3x3): x [B,h,w,C1] NHWC is resized
// (bilinear, align_corners=True) into channels [0,C1) of out [B,H,W,C1+C2]; the skip tensor (NHWC or NCHW) is copied /
// transposed into channels [C1, C1+C2).  One thread = 4 channels of one output pixel: float4 loads of the four taps,
// one float4 store; consecutive threads walk the channel axis, so every access is a full 128-byte line.
namespace mde {
// PAIR: the output is a split-bf16 pair (uint16 planes[2][B*H*W*Ctot], tc_common.cuh) for the bf16x3 conv3x3
template <bool PAIR>
__device__ __forceinline__ void store4(void* out, long long elem, long long plane_elems, const float4& o) {
  if constexpr (PAIR) {
    uint2 hi, mid;
    tc::split_bf16x2(o.x, o.y, hi.x, mid.x);
    tc::split_bf16x2(o.z, o.w, hi.y, mid.y);
    uint16_t* base = reinterpret_cast<uint16_t*>(out);
    *reinterpret_cast<uint2*>(base + elem) = hi;
    *reinterpret_cast<uint2*>(base + plane_elems + elem) = mid;
  } else {
    stg_stream(reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + elem), o);
  }
}

// eight channels at once: one 16-byte store per plane of a pair (two for fp32)
template <bool PAIR>
__device__ __forceinline__ void store8(void* out, long long elem, long long plane_elems, const float4& a, const float4& b) {
  if constexpr (PAIR) {
    uint4 hi, mid;
    tc::split_bf16x2(a.x, a.y, hi.x, mid.x);
    tc::split_bf16x2(a.z, a.w, hi.y, mid.y);
    tc::split_bf16x2(b.x, b.y, hi.z, mid.z);
    tc::split_bf16x2(b.z, b.w, hi.w, mid.w);
    uint16_t* base = reinterpret_cast<uint16_t*>(out);
    *reinterpret_cast<uint4*>(base + elem) = hi;
    *reinterpret_cast<uint4*>(base + plane_elems + elem) = mid;
  } else {
    stg_stream(reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + elem), a);
    stg_stream(reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + elem) + 1, b);
  }
}

__device__ __forceinline__ float4 bilerp4(const float4& v00, const float4& v01, const float4& v10, const float4& v11, float ly0,
                                          float ly1, float lx0, float lx1) {
  float4 o;  // same association as ATen: ly0 * (lx0 * a + lx1 * b) + ly1 * (lx0 * c + lx1 * d)
  o.x = ly0 * (lx0 * v00.x + lx1 * v01.x) + ly1 * (lx0 * v10.x + lx1 * v11.x);
  o.y = ly0 * (lx0 * v00.y + lx1 * v01.y) + ly1 * (lx0 * v10.y + lx1 * v11.y);
  o.z = ly0 * (lx0 * v00.z + lx1 * v01.z) + ly1 * (lx0 * v10.z + lx1 * v11.z);
  o.w = ly0 * (lx0 * v00.w + lx1 * v01.w) + ly1 * (lx0 * v10.w + lx1 * v11.w);
  return o;
}

// grid (ceil(W * C1/(4V) / 256), H, B): the row and the image come from the block index, so a thread needs one 32-bit division
// (the 64-bit index arithmetic of a flat launch cost more than the interpolation itself).  V = 2 (C1 % 8 == 0): a thread owns
// EIGHT channels of one output pixel -- eight 16-byte tap loads in flight and 16-byte stores into each plane of the pair (the
// four-channel form wrote 8 bytes per plane and thread and ran at 0.46-0.6 of the HBM rate).
template <bool PAIR, int V>
__global__ void __launch_bounds__(256) upsample_nhwc_kernel(const float* __restrict__ x, void* __restrict__ out, int C1,
                                                            int Ctot, int h, int w, int H, int W, float sy, float sx,
                                                            long long plane_elems) {
  const int c4 = C1 >> 2;        // float4 groups per pixel
  const int cgn = c4 / V;        // thread slots per pixel
  const unsigned t = blockIdx.x * 256u + threadIdx.x;
  pdl_sync();
  if (t >= (unsigned)(W * cgn)) return;
  const int X = (int)(t / (unsigned)cgn), cg = (int)(t - (unsigned)X * cgn) * V;
  const int Y = blockIdx.y, b = blockIdx.z;
  int y0, y1, xa, xb;
  float ly0, ly1, lx0, lx1;
  up_src(Y, sy, h, y0, y1, ly0, ly1);
  up_src(X, sx, w, xa, xb, lx0, lx1);
  const float4* src = reinterpret_cast<const float4*>(x + (long long)b * h * w * C1) + cg;
  const float4 *p00 = src + (y0 * w + xa) * c4, *p01 = src + (y0 * w + xb) * c4;
  const float4 *p10 = src + (y1 * w + xa) * c4, *p11 = src + (y1 * w + xb) * c4;
  const long long elem = (((long long)b * H + Y) * W + X) * Ctot + 4 * cg;
  if constexpr (V == 2) {
    const float4 a00 = __ldg(p00), a01 = __ldg(p01), a10 = __ldg(p10), a11 = __ldg(p11);
    const float4 b00 = __ldg(p00 + 1), b01 = __ldg(p01 + 1), b10 = __ldg(p10 + 1), b11 = __ldg(p11 + 1);
    store8<PAIR>(out, elem, plane_elems, bilerp4(a00, a01, a10, a11, ly0, ly1, lx0, lx1),
                 bilerp4(b00, b01, b10, b11, ly0, ly1, lx0, lx1));
  } else {
    store4<PAIR>(out, elem, plane_elems, bilerp4(__ldg(p00), __ldg(p01), __ldg(p10), __ldg(p11), ly0, ly1, lx0, lx1));
  }
}

// skip channels [C1, C1 + C2) of every pixel, then zeros up to the row pitch Ctot (>= C1 + C2: a pitch padded to whole
// 64-byte groups keeps the conv's TMA boxes sector-aligned).  grid (ceil(P * (Ctot - C1)/(4V) / 256), B)
template <bool PAIR, int V>
__global__ void __launch_bounds__(256) copy_channels_nhwc_kernel(const float* __restrict__ skip, void* __restrict__ out,
                                                                 int C1, int C2, int Ctot, unsigned per_image,
                                                                 long long P, long long plane_elems) {
  const unsigned t = blockIdx.x * 256u + threadIdx.x;
  pdl_sync();
  if (t >= per_image) return;
  const unsigned cgn = ((unsigned)(Ctot - C1) >> 2) / V;
  const unsigned pl = t / cgn, cg = (t - pl * cgn) * V;
  const long long p = (long long)blockIdx.y * P + pl;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* sp = reinterpret_cast<const float4*>(skip + p * C2) + cg;
  const float4 v = 4 * cg < (unsigned)C2 ? ldg_stream(sp) : z;
  if constexpr (V == 2) {
    const float4 u = 4 * (cg + 1) < (unsigned)C2 ? ldg_stream(sp + 1) : z;
    store8<PAIR>(out, p * Ctot + C1 + 4 * cg, plane_elems, v, u);
  } else {
    store4<PAIR>(out, p * Ctot + C1 + 4 * cg, plane_elems, v);
  }
}
}  // namespace mde

// the same transpose into a channel slice of a wider NHWC tensor: out points at the first channel of the slice, rows have
// out_pitch channels (concatenation of planar NCHW sources into one channels_last tensor without an intermediate copy)
extern "C" int mde_nchw_to_nhwc_slice(const float* in, float* out, int B, int C, int64_t P, int out_pitch,
                                      mde_stream_t stream) {
  if (!in || !out) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || C <= 0 || P <= 0 || out_pitch < C || B > 65535 || (C + 63) / 64 > 65535) return MDE_ERR_BAD_SHAPE;
  return mde::launch_nchw_to_nhwc(in, out, B, C, P, out_pitch, 0, 0, 0, 0, 0, (cudaStream_t)stream);
}

// ... and into a padded channels_last image [B, H + pad_top + pad_bottom, W + pad_left + pad_right, out_pitch] (the border
// is NOT written: the caller zeroes it), so that a following TensorFlow-"SAME" stride-2 convolution needs no F.pad copy
extern "C" int mde_nchw_to_nhwc_slice_padded(const float* in, float* out, int B, int C, int H, int W, int out_pitch,
                                             int pad_top, int pad_bottom, int pad_left, int pad_right, mde_stream_t stream) {
  if (!in || !out) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || out_pitch < C || B > 65535 || (C + 63) / 64 > 65535 || pad_top < 0 ||
      pad_bottom < 0 || pad_left < 0 || pad_right < 0)
    return MDE_ERR_BAD_SHAPE;
  const long long P = (long long)H * W;
  const int Wo = W + pad_left + pad_right;
  const long long Po = (long long)(H + pad_top + pad_bottom) * Wo;
  return mde::launch_nchw_to_nhwc(in, out, B, C, P, out_pitch, W, Wo, Po, pad_top, pad_left, (cudaStream_t)stream);
}

static int upsample_concat_nhwc_launch(const float* x_nhwc, const float* skip, int skip_channels_last, void* out, bool pair,
                                       int B, int C1, int C2, int Cpitch, int h, int w, int H, int W, cudaStream_t st) {
  using namespace mde;
  if (!x_nhwc || !out || (C2 > 0 && !skip)) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || C1 <= 0 || C2 < 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0 || B > 65535 || H > 65535 || Cpitch < C1 + C2)
    return MDE_ERR_BAD_SHAPE;
  if (C1 % 4 != 0 || C2 % 4 != 0 || Cpitch % 4 != 0 || !aligned(x_nhwc, 16) || !aligned(out, 16) || (C2 > 0 && !aligned(skip, 16)))
    return MDE_ERR_UNSUPPORTED;
  if (pair && C2 > 0 && !skip_channels_last) return MDE_ERR_UNSUPPORTED;  // pair output takes an NHWC skip
  if (Cpitch != C1 + C2 && !skip_channels_last) return MDE_ERR_UNSUPPORTED;
  const int Ctot = Cpitch;
  const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  const long long P = (long long)H * W;
  const long long plane = (long long)B * P * Ctot;
  if ((long long)W * (C1 / 4) > 0x7fffffffLL || (long long)h * w * (C1 / 4) > 0x7fffffffLL || P * ((Ctot - C1) / 4) > 0x7fffffffLL)
    return MDE_ERR_BAD_SHAPE;
  // eight channels per thread where the widths allow it (every DecoderBN step: 1280 / 640 / 320 / 160 up-sampled channels)
  const bool wide = C1 % 8 == 0 && Ctot % 8 == 0;
  const int vv = wide ? 2 : 1;
  const dim3 grid((unsigned)((W * (C1 / (4 * vv)) + 255) / 256), (unsigned)H, (unsigned)B);
  if (pair) {
    if (wide) launch_pdl(PDL_STREAM, upsample_nhwc_kernel<true, 2>, grid, dim3(256), 0, st, x_nhwc, out, C1, Ctot, h, w, H, W, sy, sx, plane);
    else launch_pdl(PDL_STREAM, upsample_nhwc_kernel<true, 1>, grid, dim3(256), 0, st, x_nhwc, out, C1, Ctot, h, w, H, W, sy, sx, plane);
  } else {
    if (wide) launch_pdl(PDL_STREAM, upsample_nhwc_kernel<false, 2>, grid, dim3(256), 0, st, x_nhwc, out, C1, Ctot, h, w, H, W, sy, sx, plane);
    else launch_pdl(PDL_STREAM, upsample_nhwc_kernel<false, 1>, grid, dim3(256), 0, st, x_nhwc, out, C1, Ctot, h, w, H, W, sy, sx, plane);
  }
  int rc = check_launch();
  if (rc || Ctot == C1) return rc;
  if (skip_channels_last) {
    const bool wide2 = (Ctot - C1) % 8 == 0 && C2 % 8 == 0 && C1 % 8 == 0;
    const unsigned per_image = (unsigned)(P * ((Ctot - C1) / (wide2 ? 8 : 4)));
    const dim3 g2((per_image + 255) / 256, (unsigned)B);
    if (pair) {
      if (wide2) launch_pdl(PDL_STREAM, copy_channels_nhwc_kernel<true, 2>, g2, dim3(256), 0, st, skip, out, C1, C2, Ctot, per_image, P, plane);
      else launch_pdl(PDL_STREAM, copy_channels_nhwc_kernel<true, 1>, g2, dim3(256), 0, st, skip, out, C1, C2, Ctot, per_image, P, plane);
    } else {
      if (wide2) launch_pdl(PDL_STREAM, copy_channels_nhwc_kernel<false, 2>, g2, dim3(256), 0, st, skip, out, C1, C2, Ctot, per_image, P, plane);
      else launch_pdl(PDL_STREAM, copy_channels_nhwc_kernel<false, 1>, g2, dim3(256), 0, st, skip, out, C1, C2, Ctot, per_image, P, plane);
    }
  } else {
    dim3 grid2((unsigned)((P + 63) / 64), (unsigned)((C2 + 63) / 64), (unsigned)B);
    nchw_to_nhwc_kernel<<<grid2, 256, 0, st>>>(skip, reinterpret_cast<float*>(out) + C1, C2, P, Ctot);
  }
  return check_launch();
}

extern "C" int mde_upsample_concat_nhwc_fwd(const float* x_nhwc, const float* skip, int skip_channels_last,
                                            float* out_nhwc, int B, int C1, int C2, int h, int w, int H, int W,
                                            mde_stream_t stream) {
  return upsample_concat_nhwc_launch(x_nhwc, skip, skip_channels_last, out_nhwc, false, B, C1, C2, C1 + C2, h, w, H, W,
                                     (cudaStream_t)stream);
}

// the same step writing a split-bf16 pair (planes[2][B,H,W,Cpitch], Cpitch >= C1 + C2, the channels beyond C1 + C2 zeroed) for
// mde_conv3x3_nhwc_x3_fwd; skip must be NHWC
extern "C" int mde_upsample_concat_nhwc_pair_fwd(const float* x_nhwc, const float* skip_nhwc, uint16_t* out_pair, int B, int C1,
                                                 int C2, int Cpitch, int h, int w, int H, int W, mde_stream_t stream) {
  if (Cpitch % 8 != 0) return MDE_ERR_UNSUPPORTED;
  return upsample_concat_nhwc_launch(x_nhwc, skip_nhwc, 1, out_pair, true, B, C1, C2, Cpitch, h, w, H, W, (cudaStream_t)stream);
}
