// K4 -- SILog loss: fused bilinear(align_corners=True) resample + mask + log-difference + variance, one pass.
//
// Reference: SILogLoss.forward (loss.py:12-25):
//   input = interpolate(input, target.shape[-2:], 'bilinear', align_corners=True); input,target = [mask]
//   g = log(input) - log(target);  Dg = var(g) + 0.15*mean(g)^2  (unbiased var over ALL masked pixels of the batch)
//   return 10*sqrt(Dg)
// HBM-bound: reads target (4 B/px) + mask (1 B/px) once; the quarter-size prediction stays in L2; the upsampled
// tensor is never materialised.  Sums are carried in float64 (sum g, sum g^2, n), block-reduced with warp shuffles,
// one atomic triple per block, and the last block to finish (ticket counter) writes the scalar -- one launch, no host
// sync (the reference needs nonzero() + several reductions).
#include "common.cuh"

namespace mde {

struct SilogWs {
  double sum, sumsq, count;
  unsigned int ticket;
  unsigned int pad;
};

// torch's area_pixel_compute_source_index for align_corners=True: src = dst * (in-1)/(out-1) in float
__device__ __forceinline__ void src_index(int dst, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
  const float s = scale * (float)dst;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = s - (float)i0;
  l0 = 1.f - l1;
}

template <bool INTERP, bool MASK>
__global__ void __launch_bounds__(256) silog_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                        const unsigned char* __restrict__ mask, int B, int h, int w,
                                                        int H, int W, float sy, float sx, SilogWs* ws, float* loss) {
  double s = 0.0, ss = 0.0;
  unsigned int n = 0;
  const long long total = (long long)B * H * W;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    if (MASK && !mask[t]) continue;
    const int x = (int)(t % W);
    const long long r = t / W;
    const int y = (int)(r % H);
    const int b = (int)(r / H);
    float v;
    if (INTERP) {
      int y0, y1, x0, x1;
      float ly0, ly1, lx0, lx1;
      src_index(y, sy, h, y0, y1, ly0, ly1);
      src_index(x, sx, w, x0, x1, lx0, lx1);
      const float* pb = pred + (long long)b * h * w;
      v = ly0 * (lx0 * __ldg(pb + y0 * w + x0) + lx1 * __ldg(pb + y0 * w + x1)) +
          ly1 * (lx0 * __ldg(pb + y1 * w + x0) + lx1 * __ldg(pb + y1 * w + x1));
    } else {
      v = pred[t];
    }
    const float g = logf(v) - logf(target[t]);
    s += (double)g;
    ss += (double)g * (double)g;
    ++n;
  }
  __shared__ double rs[8], rss[8];
  __shared__ unsigned int rn[8];
  __shared__ bool last;
  s = warp_sum(s);
  ss = warp_sum(ss);
  n = __reduce_add_sync(0xffffffffu, n);
  const int wid = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    rs[wid] = s;
    rss[wid] = ss;
    rn[wid] = n;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, c = 0;
    unsigned long long m = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
      a += rs[i];
      c += rss[i];
      m += rn[i];
    }
    atomicAdd(&ws->sum, a);
    atomicAdd(&ws->sumsq, c);
    atomicAdd(&ws->count, (double)m);
    __threadfence();
    last = atomicAdd(&ws->ticket, 1u) == gridDim.x - 1;
    if (last) {
      __threadfence();
      const double S = atomicAdd(&ws->sum, 0.0), SS = atomicAdd(&ws->sumsq, 0.0), N = atomicAdd(&ws->count, 0.0);
      const double mean = S / N;
      const double var = (SS - S * S / N) / (N - 1.0);  // torch.var default: unbiased
      const double dg = var + 0.15 * mean * mean;
      *loss = (float)(10.0 * sqrt(dg));
    }
  }
}

template <bool INTERP, bool MASK>
__global__ void __launch_bounds__(256) silog_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                        const unsigned char* __restrict__ mask, int B, int h, int w,
                                                        int H, int W, float sy, float sx, const SilogWs* ws,
                                                        const float* __restrict__ grad_loss, float* grad_pred) {
  const double S = ws->sum, SS = ws->sumsq, N = ws->count;
  const double mean = S / N;
  const double var = (SS - S * S / N) / (N - 1.0);
  const double root = sqrt(var + 0.15 * mean * mean);
  // d(10*sqrt(Dg))/dg_i = 5/sqrt(Dg) * ( 2*(g_i-mean)/(N-1) + 0.3*mean/N )
  const float c0 = (float)(5.0 / root * (double)grad_loss[0]);
  const float a = (float)(2.0 / (N - 1.0));
  const float bconst = (float)(0.3 * mean / N);
  const float fmean = (float)mean;
  const long long total = (long long)B * H * W;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    if (MASK && !mask[t]) continue;
    const int x = (int)(t % W);
    const long long r = t / W;
    const int y = (int)(r % H);
    const int b = (int)(r / H);
    if (INTERP) {
      int y0, y1, x0, x1;
      float ly0, ly1, lx0, lx1;
      src_index(y, sy, h, y0, y1, ly0, ly1);
      src_index(x, sx, w, x0, x1, lx0, lx1);
      const float* pb = pred + (long long)b * h * w;
      float* gb = grad_pred + (long long)b * h * w;
      const float v = ly0 * (lx0 * __ldg(pb + y0 * w + x0) + lx1 * __ldg(pb + y0 * w + x1)) +
                      ly1 * (lx0 * __ldg(pb + y1 * w + x0) + lx1 * __ldg(pb + y1 * w + x1));
      const float g = logf(v) - logf(target[t]);
      const float gv = c0 * (a * (g - fmean) + bconst) / v;
      atomicAdd(gb + y0 * w + x0, gv * ly0 * lx0);
      atomicAdd(gb + y0 * w + x1, gv * ly0 * lx1);
      atomicAdd(gb + y1 * w + x0, gv * ly1 * lx0);
      atomicAdd(gb + y1 * w + x1, gv * ly1 * lx1);
    } else {
      const float v = pred[t];
      const float g = logf(v) - logf(target[t]);
      grad_pred[t] = c0 * (a * (g - fmean) + bconst) / v;
    }
  }
}

}  // namespace mde

using namespace mde;

extern "C" {

int64_t mde_silog_ws_bytes(void) { return 64; }

static inline float scale_of(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }

int mde_silog_fwd(const float* pred, const float* target, const uint8_t* mask, int B, int h, int w, int H, int W,
                  int interpolate, void* ws, float* loss, mde_stream_t stream) {
  if (!pred || !target || !ws || !loss) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return MDE_ERR_BAD_SHAPE;
  if (!interpolate && (h != H || w != W)) return MDE_ERR_BAD_SHAPE;
  if (!aligned(ws, 8)) return MDE_ERR_BAD_POINTER;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(ws, 0, sizeof(SilogWs), st);
  const long long total = (long long)B * H * W;
  long long grid = (total + 256 * 8 - 1) / (256 * 8);
  if (grid > MDE_NUM_SMS * 8) grid = MDE_NUM_SMS * 8;
  if (grid < 1) grid = 1;
  const float sy = scale_of(h, H), sx = scale_of(w, W);
  SilogWs* W_ = reinterpret_cast<SilogWs*>(ws);
#define MDE_SILOG(I, M) \
  silog_fwd_kernel<I, M><<<(unsigned)grid, 256, 0, st>>>(pred, target, mask, B, h, w, H, W, sy, sx, W_, loss)
  if (interpolate && mask) MDE_SILOG(true, true);
  else if (interpolate) MDE_SILOG(true, false);
  else if (mask) MDE_SILOG(false, true);
  else MDE_SILOG(false, false);
#undef MDE_SILOG
  return check_launch();
}

int mde_silog_bwd(const float* pred, const float* target, const uint8_t* mask, int B, int h, int w, int H, int W,
                  int interpolate, const void* ws, const float* grad_loss, float* grad_pred, mde_stream_t stream) {
  if (!pred || !target || !ws || !grad_loss || !grad_pred) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return MDE_ERR_BAD_SHAPE;
  if (!interpolate && (h != H || w != W)) return MDE_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(grad_pred, 0, sizeof(float) * (size_t)B * h * w, st);
  const long long total = (long long)B * H * W;
  long long grid = (total + 256 * 8 - 1) / (256 * 8);
  if (grid > MDE_NUM_SMS * 8) grid = MDE_NUM_SMS * 8;
  if (grid < 1) grid = 1;
  const float sy = scale_of(h, H), sx = scale_of(w, W);
  const SilogWs* W_ = reinterpret_cast<const SilogWs*>(ws);
#define MDE_SILOG_B(I, M) \
  silog_bwd_kernel<I, M><<<(unsigned)grid, 256, 0, st>>>(pred, target, mask, B, h, w, H, W, sy, sx, W_, grad_loss, grad_pred)
  if (interpolate && mask) MDE_SILOG_B(true, true);
  else if (interpolate) MDE_SILOG_B(true, false);
  else if (mask) MDE_SILOG_B(false, true);
  else MDE_SILOG_B(false, false);
#undef MDE_SILOG_B
  return check_launch();
}

}  // extern "C"
