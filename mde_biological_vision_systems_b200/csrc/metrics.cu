// "Next" row (f)2 -- evaluation epilogue + depth metrics on the device, one launch per batch.
//
// Reference (evaluate.py:50-71,128-152 == train.py:543-568, utils.py:119-139), per image:
//   pred = interpolate(pred, gt.shape, 'bilinear', align_corners=True);  pred.clip(min_eval, max_eval), inf -> max,
//   nan -> min;  valid = (gt > min_eval) & (gt < max_eval) [& garg / eigen crop box];
//   compute_errors(gt[valid], pred[valid]) -> a1 a2 a3 abs_rel rmse log_10 rmse_log silog sq_rel
// The reference does this with .cpu().numpy() per image at batch 1; here the up-sampling is fused (never materialised),
// every image of the batch is reduced in the same launch (fp64 sums, warp shuffles, one atomic set per block) and the
// last block of each image writes its 9 metrics + the valid-pixel count.  HBM-bound: reads gt once (4 B/px).
#include "common.cuh"

namespace mde {

constexpr int MET_SUMS = 10;  // n, c1, c2, c3, abs_rel, sq_rel, sq, sqlog, log, log10
struct MetricsWs {
  double s[MET_SUMS];
  unsigned int ticket, pad;
};

__device__ __forceinline__ void met_src(int dst, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
  const float s = scale * (float)dst;  // ATen area_pixel_compute_source_index, align_corners=True
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = s - (float)i0;
  l0 = 1.f - l1;
}

// grid (blocks, B)
__global__ void __launch_bounds__(256) eval_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                           int h, int w, int H, int W, float sy, float sx, float min_eval,
                                                           float max_eval, int cy0, int cy1, int cx0, int cx1,
                                                           MetricsWs* ws, float* __restrict__ out) {
  const int b = blockIdx.y;
  const float* pb = pred + (long long)b * h * w;
  const float* gb = gt + (long long)b * H * W;
  double acc[MET_SUMS];
#pragma unroll
  for (int i = 0; i < MET_SUMS; ++i) acc[i] = 0.0;
  const int total = H * W;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int y = t / W, x = t - y * W;
    const float g = gb[t];
    if (!(g > min_eval && g < max_eval) || y < cy0 || y >= cy1 || x < cx0 || x >= cx1) continue;
    float p;
    {
      // always the 4-tap formula, also at equal sizes: the reference calls F.interpolate unconditionally, where a
      // zero-weight tap on a NaN / inf neighbour still poisons the pixel (0 * inf = NaN -> min_depth_eval)
      int y0, y1, x0, x1;
      float ly0, ly1, lx0, lx1;
      met_src(y, sy, h, y0, y1, ly0, ly1);
      met_src(x, sx, w, x0, x1, lx0, lx1);
      p = ly0 * (lx0 * __ldg(pb + y0 * w + x0) + lx1 * __ldg(pb + y0 * w + x1)) +
          ly1 * (lx0 * __ldg(pb + y1 * w + x0) + lx1 * __ldg(pb + y1 * w + x1));
    }
    if (p != p) p = min_eval;  // nan -> min; +-inf are handled by the clamp
    p = fminf(fmaxf(p, min_eval), max_eval);
    const float thresh = fmaxf(__fdiv_rn(g, p), __fdiv_rn(p, g));
    const float d = g - p;
    const float lg = logf(g), lp = logf(p);
    acc[0] += 1.0;
    acc[1] += thresh < 1.25f ? 1.0 : 0.0;
    acc[2] += thresh < 1.5625f ? 1.0 : 0.0;
    acc[3] += thresh < 1.953125f ? 1.0 : 0.0;
    acc[4] += (double)__fdiv_rn(fabsf(d), g);
    acc[5] += (double)__fdiv_rn(d * d, g);
    acc[6] += (double)(d * d);
    const float dl = lg - lp;
    acc[7] += (double)(dl * dl);
    acc[8] += (double)(lp - lg);
    acc[9] += (double)fabsf(log10f(g) - log10f(p));
  }
  __shared__ double red[8][MET_SUMS];
#pragma unroll
  for (int i = 0; i < MET_SUMS; ++i) acc[i] = warp_sum(acc[i]);
  const int wid = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int i = 0; i < MET_SUMS; ++i) red[wid][i] = acc[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    MetricsWs* wb = ws + b;
    for (int i = 0; i < MET_SUMS; ++i) {
      double a = 0.0;
      for (int k = 0; k < (int)(blockDim.x >> 5); ++k) a += red[k][i];
      atomicAdd(&wb->s[i], a);
    }
    __threadfence();
    if (atomicAdd(&wb->ticket, 1u) == gridDim.x - 1) {
      __threadfence();
      double s[MET_SUMS];
      for (int i = 0; i < MET_SUMS; ++i) s[i] = atomicAdd(&wb->s[i], 0.0);
      const double n = s[0];
      float* o = out + b * 10;
      // order of the reference's dict: a1 a2 a3 abs_rel rmse log_10 rmse_log silog sq_rel, then n_valid
      o[0] = (float)(s[1] / n);
      o[1] = (float)(s[2] / n);
      o[2] = (float)(s[3] / n);
      o[3] = (float)(s[4] / n);
      o[4] = (float)sqrt(s[6] / n);
      o[5] = (float)(s[9] / n);
      o[6] = (float)sqrt(s[7] / n);
      const double m = s[8] / n;
      o[7] = (float)(sqrt(fmax(s[7] / n - m * m, 0.0)) * 100.0);
      o[8] = (float)(s[5] / n);
      o[9] = (float)n;
    }
  }
}

// Mirror averaging of infer.py:108-118 at low resolution: out = 0.5 * (clip(a) + clip(flip_w(b)))
__global__ void __launch_bounds__(256) flip_average_kernel(const float* __restrict__ a, const float* __restrict__ bflip,
                                                           float* __restrict__ out, long long rows, int w, float lo,
                                                           float hi) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * w) return;
  const long long r = i / w;
  const int x = (int)(i - r * w);
  const float u = fminf(fmaxf(a[i], lo), hi);
  const float v = fminf(fmaxf(bflip[r * w + (w - 1 - x)], lo), hi);
  out[i] = 0.5f * (u + v);
}

}  // namespace mde

using namespace mde;

extern "C" {

int64_t mde_eval_metrics_ws_bytes(int B) { return (int64_t)sizeof(MetricsWs) * (B > 0 ? B : 0); }

int mde_eval_metrics_fwd(const float* pred, const float* gt, int B, int h, int w, int H, int W, float min_depth_eval,
                         float max_depth_eval, int crop_y0, int crop_y1, int crop_x0, int crop_x1, void* ws, float* out,
                         mde_stream_t stream) {
  if (!pred || !gt || !ws || !out) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || B > 65535 || h <= 0 || w <= 0 || H <= 0 || W <= 0 || (long long)H * W > 0x7fffffffLL) return MDE_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(ws, 0, sizeof(MetricsWs) * (size_t)B, st);
  const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  int gx = (H * W + 256 * 8 - 1) / (256 * 8);
  const int cap = (MDE_NUM_SMS * 8 + B - 1) / B;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  eval_metrics_kernel<<<dim3((unsigned)gx, (unsigned)B), 256, 0, st>>>(pred, gt, h, w, H, W, sy, sx, min_depth_eval,
                                                                      max_depth_eval, crop_y0, crop_y1, crop_x0, crop_x1,
                                                                      reinterpret_cast<MetricsWs*>(ws), out);
  return check_launch();
}

int mde_flip_average(const float* a, const float* b_flipped, float* out, int64_t rows, int w, float lo, float hi,
                     mde_stream_t stream) {
  if (!a || !b_flipped || !out) return MDE_ERR_BAD_POINTER;
  if (rows <= 0 || w <= 0) return MDE_ERR_BAD_SHAPE;
  const long long n = rows * w;
  flip_average_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, b_flipped, out, rows, w, lo, hi);
  return check_launch();
}

}  // extern "C"
