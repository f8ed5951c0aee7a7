// Section 8(e): SyncBatchNorm statistics for the DDP training path (train.py:296 converts every BatchNorm to
// nn.SyncBatchNorm; semantics = batch statistics over the GLOBAL batch).  Stock torch runs, per BN layer, collect_stats ->
// all_gather -> gather_stats_with_counts -> elemt (forward) and backward_reduce -> all_reduce -> backward_elemt plus a
// few hundred small copies / index kernels per step: +27 ms on a 67 ms step at 2 GPUs.  Here a layer is
//   forward : bn_stats (per-channel sum, sum of squares -> float64[2C])  -> ONE all_reduce of 2C+1 doubles -> bn_apply
//   backward: bn_bwd_reduce (sum dy, sum dy*xhat -> float64[2C])         -> ONE all_reduce of 2C doubles   -> bn_bwd_apply
// on channels_last (NHWC) activations: a thread owns 4 consecutive channels (float4) and strides over pixels, so every
// access is a full 128-byte line; partial sums are carried in fp32 per thread (a few hundred terms), promoted to fp64 for
// the block / grid reduction (fp64 atomics), so mean and variance come from exact-enough raw moments.
#include "common.cuh"

namespace mde {

// block = cl channel-group lanes x (256 / cl) pixel lanes; cl = min(32, next power of two >= C/4), so narrow layers
// (C = 16 ... 64) still keep every thread busy

// stats[0:C] += sum_p x[p,c];  stats[C:2C] += sum_p x[p,c]^2
__global__ void __launch_bounds__(256) bn_stats_nhwc_kernel(const float* __restrict__ x, long long N, int C,
                                                            double* __restrict__ stats, double* __restrict__ zero_next,
                                                            int cl_log2) {
  if (zero_next != nullptr && blockIdx.x == 0 && blockIdx.y == 0)  // recycle the buffer of the call after this one
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) zero_next[i] = 0.0;
  const int c4 = C >> 2;
  const int cl = 1 << cl_log2, npl = 256 >> cl_log2, ci = threadIdx.x & (cl - 1);
  const int cg = blockIdx.x * cl + ci;
  const int pl = threadIdx.x >> cl_log2;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = make_float4(0.f, 0.f, 0.f, 0.f);
  if (cg < c4) {
    for (long long p = (long long)blockIdx.y * npl + pl; p < N; p += (long long)gridDim.y * npl) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + p * C) + cg);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
    }
  }
  __shared__ double red[256][8];  // [pl][ci]
  double* r = red[pl * cl + ci];
  r[0] = s.x; r[1] = s.y; r[2] = s.z; r[3] = s.w; r[4] = q.x; r[5] = q.y; r[6] = q.z; r[7] = q.w;
  __syncthreads();
  if (pl == 0 && cg < c4) {
    double t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      t[i] = 0.0;
      for (int k = 0; k < npl; ++k) t[i] += red[k * cl + ci][i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      atomicAdd(stats + cg * 4 + i, t[i]);
      atomicAdd(stats + C + cg * 4 + i, t[4 + i]);
    }
  }
}

// y = (x - mean) * invstd * w + b with mean / var from the (globally reduced) raw moments; block (0,0) also writes
// save_mean / save_invstd and updates the running statistics (unbiased variance, like torch)
__global__ void __launch_bounds__(256) bn_apply_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, long long N,
                                                            int C, const double* __restrict__ stats, double count,
                                                            const float* __restrict__ weight, const float* __restrict__ bias,
                                                            float eps, float* __restrict__ save_mean,
                                                            float* __restrict__ save_invstd, float* running_mean,
                                                            float* running_var, float momentum, int cl_log2) {
  const int c4 = C >> 2;
  const int cl = 1 << cl_log2, npl = 256 >> cl_log2, ci = threadIdx.x & (cl - 1);
  const int cg = blockIdx.x * cl + ci;
  const int pl = threadIdx.x >> cl_log2;
  if (cg >= c4) return;
  float sc[4], sh[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = cg * 4 + i;
    const double mean = stats[c] / count;
    double var = stats[C + c] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float w = weight ? weight[c] : 1.f, b = bias ? bias[c] : 0.f;
    sc[i] = invstd * w;
    sh[i] = b - (float)mean * sc[i];
    if (blockIdx.y == 0 && pl == 0) {
      save_mean[c] = (float)mean;
      save_invstd[c] = invstd;
      if (running_mean) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
      }
    }
  }
  for (long long p = (long long)blockIdx.y * npl + pl; p < N; p += (long long)gridDim.y * npl) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + p * C) + cg);
    float4 o;
    o.x = fmaf(v.x, sc[0], sh[0]); o.y = fmaf(v.y, sc[1], sh[1]); o.z = fmaf(v.z, sc[2], sh[2]); o.w = fmaf(v.w, sc[3], sh[3]);
    reinterpret_cast<float4*>(y + p * C)[cg] = o;
  }
}

// sums[0:C] += sum_p dy;  sums[C:2C] += sum_p dy * (x - mean) * invstd
__global__ void __launch_bounds__(256) bn_bwd_reduce_nhwc_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                 long long N, int C, const float* __restrict__ mean,
                                                                 const float* __restrict__ invstd, double* __restrict__ sums,
                                                                 double* __restrict__ zero_next, int cl_log2) {
  if (zero_next != nullptr && blockIdx.x == 0 && blockIdx.y == 0)
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) zero_next[i] = 0.0;
  const int c4 = C >> 2;
  const int cl = 1 << cl_log2, npl = 256 >> cl_log2, ci = threadIdx.x & (cl - 1);
  const int cg = blockIdx.x * cl + ci;
  const int pl = threadIdx.x >> cl_log2;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = make_float4(0.f, 0.f, 0.f, 0.f);
  if (cg < c4) {
    const float4 m = reinterpret_cast<const float4*>(mean)[cg], is = reinterpret_cast<const float4*>(invstd)[cg];
    for (long long p = (long long)blockIdx.y * npl + pl; p < N; p += (long long)gridDim.y * npl) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + p * C) + cg);
      const float4 g = __ldg(reinterpret_cast<const float4*>(dy + p * C) + cg);
      s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
      q.x = fmaf(g.x, (v.x - m.x) * is.x, q.x); q.y = fmaf(g.y, (v.y - m.y) * is.y, q.y);
      q.z = fmaf(g.z, (v.z - m.z) * is.z, q.z); q.w = fmaf(g.w, (v.w - m.w) * is.w, q.w);
    }
  }
  __shared__ double red[256][8];  // [pl][ci]
  double* r = red[pl * cl + ci];
  r[0] = s.x; r[1] = s.y; r[2] = s.z; r[3] = s.w; r[4] = q.x; r[5] = q.y; r[6] = q.z; r[7] = q.w;
  __syncthreads();
  if (pl == 0 && cg < c4) {
    double t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      t[i] = 0.0;
      for (int k = 0; k < npl; ++k) t[i] += red[k * cl + ci][i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      atomicAdd(sums + cg * 4 + i, t[i]);
      atomicAdd(sums + C + cg * 4 + i, t[4 + i]);
    }
  }
}

// dx = (dy - sum_dy / count - xhat * sum_dy_xhat / count) * invstd * w     (sums reduced over the global batch)
__global__ void __launch_bounds__(256) bn_bwd_apply_nhwc_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                float* __restrict__ dx, long long N, int C,
                                                                const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                const float* __restrict__ weight, const double* __restrict__ sums,
                                                                double count, int cl_log2) {
  const int c4 = C >> 2;
  const int cl = 1 << cl_log2, npl = 256 >> cl_log2, ci = threadIdx.x & (cl - 1);
  const int cg = blockIdx.x * cl + ci;
  const int pl = threadIdx.x >> cl_log2;
  if (cg >= c4) return;
  float m[4], is[4], a[4], k1[4], k2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = cg * 4 + i;
    m[i] = mean[c];
    is[i] = invstd[c];
    a[i] = is[i] * (weight ? weight[c] : 1.f);
    k1[i] = (float)(sums[c] / count);
    k2[i] = (float)(sums[C + c] / count);
  }
  for (long long p = (long long)blockIdx.y * npl + pl; p < N; p += (long long)gridDim.y * npl) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + p * C) + cg);
    const float4 g = __ldg(reinterpret_cast<const float4*>(dy + p * C) + cg);
    float4 o;
    o.x = (g.x - k1[0] - (v.x - m[0]) * is[0] * k2[0]) * a[0];
    o.y = (g.y - k1[1] - (v.y - m[1]) * is[1] * k2[1]) * a[1];
    o.z = (g.z - k1[2] - (v.z - m[2]) * is[2] * k2[2]) * a[2];
    o.w = (g.w - k1[3] - (v.w - m[3]) * is[3] * k2[3]) * a[3];
    reinterpret_cast<float4*>(dx + p * C)[cg] = o;
  }
}

static int bn_cl_log2(int C) {
  int l = 0;
  while ((1 << l) < C / 4 && l < 5) ++l;
  return l;
}
static dim3 bn_grid(long long N, int C) {
  const int cl = 1 << bn_cl_log2(C), npl = 256 / cl;
  const int gx = (C / 4 + cl - 1) / cl;
  long long gy = (N + npl * 16 - 1) / (npl * 16);  // >= 16 pixels per thread
  const long long cap = (MDE_NUM_SMS * 8 + gx - 1) / gx;
  if (gy > cap) gy = cap;
  if (gy < 1) gy = 1;
  return dim3((unsigned)gx, (unsigned)gy);
}

}  // namespace mde

using namespace mde;

extern "C" {

int mde_bn_stats_nhwc(const float* x, int64_t N, int C, double* stats, double* zero_next, mde_stream_t stream) {
  if (!x || !stats) return MDE_ERR_BAD_POINTER;
  if (N <= 0 || C <= 0) return MDE_ERR_BAD_SHAPE;
  if (C % 4 != 0 || !aligned(x, 16)) return MDE_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  if (!zero_next) cudaMemsetAsync(stats, 0, sizeof(double) * 2 * (size_t)C, st);
  bn_stats_nhwc_kernel<<<bn_grid(N, C), 256, 0, st>>>(x, N, C, stats, zero_next, bn_cl_log2(C));
  return check_launch();
}

int mde_bn_apply_nhwc(const float* x, float* y, int64_t N, int C, const double* stats, double count, const float* weight,
                      const float* bias, float eps, float* save_mean, float* save_invstd, float* running_mean,
                      float* running_var, float momentum, mde_stream_t stream) {
  if (!x || !y || !stats || !save_mean || !save_invstd) return MDE_ERR_BAD_POINTER;
  if (N <= 0 || C <= 0 || count <= 0.0) return MDE_ERR_BAD_SHAPE;
  if (C % 4 != 0 || !aligned(x, 16) || !aligned(y, 16)) return MDE_ERR_UNSUPPORTED;
  bn_apply_nhwc_kernel<<<bn_grid(N, C), 256, 0, (cudaStream_t)stream>>>(x, y, N, C, stats, count, weight, bias, eps,
                                                                        save_mean, save_invstd, running_mean, running_var,
                                                                        momentum, bn_cl_log2(C));
  return check_launch();
}

int mde_bn_bwd_reduce_nhwc(const float* x, const float* dy, int64_t N, int C, const float* mean, const float* invstd,
                           double* sums, double* zero_next, mde_stream_t stream) {
  if (!x || !dy || !mean || !invstd || !sums) return MDE_ERR_BAD_POINTER;
  if (N <= 0 || C <= 0) return MDE_ERR_BAD_SHAPE;
  if (C % 4 != 0 || !aligned(x, 16) || !aligned(dy, 16) || !aligned(mean, 16) || !aligned(invstd, 16)) return MDE_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  if (!zero_next) cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)C, st);
  bn_bwd_reduce_nhwc_kernel<<<bn_grid(N, C), 256, 0, st>>>(x, dy, N, C, mean, invstd, sums, zero_next, bn_cl_log2(C));
  return check_launch();
}

int mde_bn_bwd_apply_nhwc(const float* x, const float* dy, float* dx, int64_t N, int C, const float* mean,
                          const float* invstd, const float* weight, const double* sums, double count, mde_stream_t stream) {
  if (!x || !dy || !dx || !mean || !invstd || !sums) return MDE_ERR_BAD_POINTER;
  if (N <= 0 || C <= 0 || count <= 0.0) return MDE_ERR_BAD_SHAPE;
  if (C % 4 != 0 || !aligned(x, 16) || !aligned(dy, 16) || !aligned(dx, 16)) return MDE_ERR_UNSUPPORTED;
  bn_bwd_apply_nhwc_kernel<<<bn_grid(N, C), 256, 0, (cudaStream_t)stream>>>(x, dy, dx, N, C, mean, invstd, weight, sums,
                                                                            count, bn_cl_log2(C));
  return check_launch();
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// Fused statistics exchange over NVLink peer memory (no NCCL call in the loop).  Every rank owns a slice of a
// symmetric-memory arena that all peers have mapped (torch symmetric memory; `peers.base[r]` = rank r's arena in THIS
// process's address space).  The statistics kernel's last block publishes the rank's [2C] partial sums into slot
// [my_rank] of EVERY peer's arena with plain stores over NVLink, fences system-wide and raises an epoch flag next to them;
// the consuming kernel (normalise / input gradient) spins on the `world` flags in its own arena, then sums the partials.
// Slots and flags are double-buffered by epoch parity: a rank can run at most one layer call ahead of the slowest peer
// (its next consume needs that peer's publish), so parity e&1 is never rewritten while someone still reads epoch e.
namespace mde {

struct BnPeers {
  unsigned long long base[8];
  int world, rank;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// CUDA-graph replay: kernel arguments are frozen at capture, but the epoch (and with it the parity of the slots / flags) must
// advance by one per step.  With a device-resident replay counter (mde_bn_p2p_set_epoch_counter) the `epoch` argument is a BASE
// to which the counter's current value is added, and slot_off / flag_off address parity 0 of the direction: the kernel derives
// the parity offsets itself (slot rows are world * 2C doubles, flag rows world uint64 -- the layout the host uses).
__device__ __forceinline__ void bn_resolve_epoch(const unsigned long long* ctr, unsigned long long& epoch, long long& slot_off,
                                                 long long& flag_off, int world, int C) {
  if (ctr == nullptr) return;
  epoch += __ldcg(ctr);
  const long long par = (long long)(epoch & 1ULL);
  slot_off += par * (long long)world * 2 * C * 8;
  flag_off += par * (long long)world * 8;
}

// Called by every block after its atomics into `local` [2C] (+ ticket at index 2C): the last block publishes.
__device__ __forceinline__ void bn_publish_if_last(double* local, int C, const BnPeers& peers, long long slot_off,
                                                   long long flag_off, unsigned long long epoch) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long t = atomicAdd(reinterpret_cast<unsigned long long*>(local + 2 * C), 1ULL);
    is_last = (t == (unsigned long long)gridDim.x * gridDim.y - 1ULL);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    const double v = __ldcg(local + i);
    for (int r = 0; r < peers.world; ++r)
      reinterpret_cast<double*>(peers.base[r] + slot_off)[(long long)peers.rank * 2 * C + i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < peers.world)
    st_release_sys(reinterpret_cast<unsigned long long*>(peers.base[threadIdx.x] + flag_off) + peers.rank, epoch);
}

// Wait until every rank's partial sums of this epoch have landed in MY arena.  All ranks must enter every BatchNorm layer
// within the bound below (default 600 s -- the order of NCCL's own collective timeout; the reference would block in NCCL
// for rank skew such as rank-0-only validation, train.py:459-499).  The wait backs off with __nanosleep; when the bound
// expires the rank records the event in g_bn_peer_timeouts (mde_bn_peer_timeouts()) and traps -- the equivalent of the NCCL
// watchdog aborting the process -- rather than normalising with missing statistics.
static __device__ unsigned long long g_bn_timeout_ns = 600ULL * 1000000000ULL;
static __device__ unsigned int g_bn_peer_timeouts = 0;
// time block (0, 0) of the consuming kernels spent waiting for the peers' flags, and the number of such waits (diagnostic:
// mde_bn_wait_stats; the wait is rank skew + NVLink store latency -- the part of a training step SyncBatchNorm adds)
static __device__ unsigned long long g_bn_wait_ns = 0;
static __device__ unsigned long long g_bn_waits = 0;

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void bn_wait_peers(unsigned long long my_base, long long flag_off, int world,
                                              unsigned long long epoch) {
  const bool probe = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0;
  const unsigned long long t_in = probe ? globaltimer_ns() : 0ULL;
  if (threadIdx.x < world) {
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(my_base + flag_off) + threadIdx.x;
    if (ld_acquire_sys(f) < epoch) {
      const unsigned long long t0 = globaltimer_ns();
      unsigned int backoff = 32;
      while (ld_acquire_sys(f) < epoch) {
        __nanosleep(backoff);
        if (backoff < 512) backoff *= 2;  // the expected wait is microseconds: a longer sleep only adds wake-up overshoot
        if (globaltimer_ns() - t0 > g_bn_timeout_ns) {
          atomicAdd(&g_bn_peer_timeouts, 1u);
          __threadfence_system();
          asm volatile("trap;");
        }
      }
    }
  }
  __syncthreads();
  if (probe) {
    atomicAdd(&g_bn_wait_ns, globaltimer_ns() - t_in);
    atomicAdd(&g_bn_waits, 1ULL);
  }
}

__global__ void __launch_bounds__(256) bn_stats_p2p_kernel(const float* __restrict__ x, long long N, int C,
                                                           double* __restrict__ local, double* __restrict__ zero_next,
                                                           int cl_log2, BnPeers peers, long long slot_off, long long flag_off,
                                                           unsigned long long epoch, const unsigned long long* epoch_ctr) {
  bn_resolve_epoch(epoch_ctr, epoch, slot_off, flag_off, peers.world, C);
  if (zero_next != nullptr && blockIdx.x == 0 && blockIdx.y == 0)
    for (int i = threadIdx.x; i < 2 * C + 1; i += blockDim.x) zero_next[i] = 0.0;
  const int c4 = C >> 2;
  const int cl = 1 << cl_log2, npl = 256 >> cl_log2, ci = threadIdx.x & (cl - 1);
  const int cg = blockIdx.x * cl + ci;
  const int pl = threadIdx.x >> cl_log2;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = make_float4(0.f, 0.f, 0.f, 0.f);
  if (cg < c4) {
    for (long long p = (long long)blockIdx.y * npl + pl; p < N; p += (long long)gridDim.y * npl) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + p * C) + cg);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
    }
  }
  __shared__ double red[256][8];
  double* r = red[pl * cl + ci];
  r[0] = s.x; r[1] = s.y; r[2] = s.z; r[3] = s.w; r[4] = q.x; r[5] = q.y; r[6] = q.z; r[7] = q.w;
  __syncthreads();
  if (pl == 0 && cg < c4) {
    double t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      t[i] = 0.0;
      for (int k = 0; k < npl; ++k) t[i] += red[k * cl + ci][i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      atomicAdd(local + cg * 4 + i, t[i]);
      atomicAdd(local + C + cg * 4 + i, t[4 + i]);
    }
  }
  bn_publish_if_last(local, C, peers, slot_off, flag_off, epoch);
}

__global__ void __launch_bounds__(256) bn_bwd_reduce_p2p_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                long long N, int C, const float* __restrict__ mean,
                                                                const float* __restrict__ invstd, double* __restrict__ local,
                                                                double* __restrict__ zero_next, int cl_log2, BnPeers peers,
                                                                long long slot_off, long long flag_off,
                                                                unsigned long long epoch, const unsigned long long* epoch_ctr) {
  bn_resolve_epoch(epoch_ctr, epoch, slot_off, flag_off, peers.world, C);
  if (zero_next != nullptr && blockIdx.x == 0 && blockIdx.y == 0)
    for (int i = threadIdx.x; i < 2 * C + 1; i += blockDim.x) zero_next[i] = 0.0;
  const int c4 = C >> 2;
  const int cl = 1 << cl_log2, npl = 256 >> cl_log2, ci = threadIdx.x & (cl - 1);
  const int cg = blockIdx.x * cl + ci;
  const int pl = threadIdx.x >> cl_log2;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = make_float4(0.f, 0.f, 0.f, 0.f);
  if (cg < c4) {
    const float4 m = reinterpret_cast<const float4*>(mean)[cg], is = reinterpret_cast<const float4*>(invstd)[cg];
    for (long long p = (long long)blockIdx.y * npl + pl; p < N; p += (long long)gridDim.y * npl) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + p * C) + cg);
      const float4 g = __ldg(reinterpret_cast<const float4*>(dy + p * C) + cg);
      s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
      q.x = fmaf(g.x, (v.x - m.x) * is.x, q.x); q.y = fmaf(g.y, (v.y - m.y) * is.y, q.y);
      q.z = fmaf(g.z, (v.z - m.z) * is.z, q.z); q.w = fmaf(g.w, (v.w - m.w) * is.w, q.w);
    }
  }
  __shared__ double red[256][8];
  double* r = red[pl * cl + ci];
  r[0] = s.x; r[1] = s.y; r[2] = s.z; r[3] = s.w; r[4] = q.x; r[5] = q.y; r[6] = q.z; r[7] = q.w;
  __syncthreads();
  if (pl == 0 && cg < c4) {
    double t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      t[i] = 0.0;
      for (int k = 0; k < npl; ++k) t[i] += red[k * cl + ci][i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      atomicAdd(local + cg * 4 + i, t[i]);
      atomicAdd(local + C + cg * 4 + i, t[4 + i]);
    }
  }
  bn_publish_if_last(local, C, peers, slot_off, flag_off, epoch);
}

// gather the world's partials of this epoch from MY arena into gsum [2C] (every block does it for itself: 2C * world
// doubles from L2), then the kernels continue exactly like the NCCL variants with `gsum` as the reduced moments
__device__ __forceinline__ double bn_sum_slots(unsigned long long my_base, long long slot_off, int world, int C, int idx) {
  const double* sl = reinterpret_cast<const double*>(my_base + slot_off);
  double a = 0.0;
  for (int r = 0; r < world; ++r) a += __ldcg(sl + (long long)r * 2 * C + idx);
  return a;
}

__global__ void __launch_bounds__(256) bn_apply_p2p_kernel(const float* __restrict__ x, float* __restrict__ y, long long N,
                                                           int C, unsigned long long my_base, long long slot_off,
                                                           long long flag_off, int world, unsigned long long epoch,
                                                           double count, const float* __restrict__ weight,
                                                           const float* __restrict__ bias, float eps,
                                                           float* __restrict__ save_mean, float* __restrict__ save_invstd,
                                                           float* running_mean, float* running_var, float momentum,
                                                           int cl_log2, const unsigned long long* epoch_ctr) {
  bn_resolve_epoch(epoch_ctr, epoch, slot_off, flag_off, world, C);
  bn_wait_peers(my_base, flag_off, world, epoch);
  const int c4 = C >> 2;
  const int cl = 1 << cl_log2, npl = 256 >> cl_log2, ci = threadIdx.x & (cl - 1);
  const int cg = blockIdx.x * cl + ci;
  const int pl = threadIdx.x >> cl_log2;
  if (cg >= c4) return;
  float sc[4], sh[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = cg * 4 + i;
    const double mean = bn_sum_slots(my_base, slot_off, world, C, c) / count;
    double var = bn_sum_slots(my_base, slot_off, world, C, C + c) / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float w = weight ? weight[c] : 1.f, b = bias ? bias[c] : 0.f;
    sc[i] = invstd * w;
    sh[i] = b - (float)mean * sc[i];
    if (blockIdx.y == 0 && pl == 0) {
      save_mean[c] = (float)mean;
      save_invstd[c] = invstd;
      if (running_mean) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
      }
    }
  }
  for (long long p = (long long)blockIdx.y * npl + pl; p < N; p += (long long)gridDim.y * npl) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + p * C) + cg);
    float4 o;
    o.x = fmaf(v.x, sc[0], sh[0]); o.y = fmaf(v.y, sc[1], sh[1]); o.z = fmaf(v.z, sc[2], sh[2]); o.w = fmaf(v.w, sc[3], sh[3]);
    reinterpret_cast<float4*>(y + p * C)[cg] = o;
  }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_p2p_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                               float* __restrict__ dx, long long N, int C,
                                                               const float* __restrict__ mean, const float* __restrict__ invstd,
                                                               const float* __restrict__ weight, unsigned long long my_base,
                                                               long long slot_off, long long flag_off, int world,
                                                               unsigned long long epoch, double count, int cl_log2,
                                                               const unsigned long long* epoch_ctr) {
  bn_resolve_epoch(epoch_ctr, epoch, slot_off, flag_off, world, C);
  bn_wait_peers(my_base, flag_off, world, epoch);
  const int c4 = C >> 2;
  const int cl = 1 << cl_log2, npl = 256 >> cl_log2, ci = threadIdx.x & (cl - 1);
  const int cg = blockIdx.x * cl + ci;
  const int pl = threadIdx.x >> cl_log2;
  if (cg >= c4) return;
  float m[4], is[4], a[4], k1[4], k2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = cg * 4 + i;
    m[i] = mean[c];
    is[i] = invstd[c];
    a[i] = is[i] * (weight ? weight[c] : 1.f);
    k1[i] = (float)(bn_sum_slots(my_base, slot_off, world, C, c) / count);
    k2[i] = (float)(bn_sum_slots(my_base, slot_off, world, C, C + c) / count);
  }
  for (long long p = (long long)blockIdx.y * npl + pl; p < N; p += (long long)gridDim.y * npl) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + p * C) + cg);
    const float4 g = __ldg(reinterpret_cast<const float4*>(dy + p * C) + cg);
    float4 o;
    o.x = (g.x - k1[0] - (v.x - m[0]) * is[0] * k2[0]) * a[0];
    o.y = (g.y - k1[1] - (v.y - m[1]) * is[1] * k2[1]) * a[1];
    o.z = (g.z - k1[2] - (v.z - m[2]) * is[2] * k2[2]) * a[2];
    o.w = (g.w - k1[3] - (v.w - m[3]) * is[3] * k2[3]) * a[3];
    reinterpret_cast<float4*>(dx + p * C)[cg] = o;
  }
}

// device address of the replay counter (null: eager launches, `epoch` is the epoch).  One process per GPU: process-global.
static const unsigned long long* g_bn_epoch_ctr = nullptr;

static bool fill_peers(BnPeers& p, const uint64_t* bases, int world, int rank) {
  if (!bases || world < 1 || world > 8 || rank < 0 || rank >= world) return false;
  for (int i = 0; i < 8; ++i) p.base[i] = i < world ? bases[i] : 0ULL;
  p.world = world;
  p.rank = rank;
  return true;
}

}  // namespace mde

extern "C" {

int mde_bn_stats_p2p_nhwc(const float* x, int64_t N, int C, double* local, double* zero_next, const uint64_t* peer_bases,
                          int world, int rank, int64_t slot_off, int64_t flag_off, uint64_t epoch, mde_stream_t stream) {
  if (!x || !local || !zero_next) return MDE_ERR_BAD_POINTER;
  if (N <= 0 || C <= 0) return MDE_ERR_BAD_SHAPE;
  if (C % 4 != 0 || !aligned(x, 16)) return MDE_ERR_UNSUPPORTED;
  BnPeers p;
  if (!fill_peers(p, peer_bases, world, rank)) return MDE_ERR_BAD_SHAPE;
  bn_stats_p2p_kernel<<<bn_grid(N, C), 256, 0, (cudaStream_t)stream>>>(x, N, C, local, zero_next, bn_cl_log2(C), p, slot_off,
                                                                       flag_off, epoch, g_bn_epoch_ctr);
  return check_launch();
}

int mde_bn_apply_p2p_nhwc(const float* x, float* y, int64_t N, int C, uint64_t my_base, int64_t slot_off, int64_t flag_off,
                          int world, uint64_t epoch, double count, const float* weight, const float* bias, float eps,
                          float* save_mean, float* save_invstd, float* running_mean, float* running_var, float momentum,
                          mde_stream_t stream) {
  if (!x || !y || !my_base || !save_mean || !save_invstd) return MDE_ERR_BAD_POINTER;
  if (N <= 0 || C <= 0 || count <= 0.0 || world < 1 || world > 8) return MDE_ERR_BAD_SHAPE;
  if (C % 4 != 0 || !aligned(x, 16) || !aligned(y, 16)) return MDE_ERR_UNSUPPORTED;
  bn_apply_p2p_kernel<<<bn_grid(N, C), 256, 0, (cudaStream_t)stream>>>(x, y, N, C, my_base, slot_off, flag_off, world, epoch,
                                                                       count, weight, bias, eps, save_mean, save_invstd,
                                                                       running_mean, running_var, momentum, bn_cl_log2(C),
                                                                       g_bn_epoch_ctr);
  return check_launch();
}

int mde_bn_bwd_reduce_p2p_nhwc(const float* x, const float* dy, int64_t N, int C, const float* mean, const float* invstd,
                               double* local, double* zero_next, const uint64_t* peer_bases, int world, int rank,
                               int64_t slot_off, int64_t flag_off, uint64_t epoch, mde_stream_t stream) {
  if (!x || !dy || !mean || !invstd || !local || !zero_next) return MDE_ERR_BAD_POINTER;
  if (N <= 0 || C <= 0) return MDE_ERR_BAD_SHAPE;
  if (C % 4 != 0 || !aligned(x, 16) || !aligned(dy, 16) || !aligned(mean, 16) || !aligned(invstd, 16)) return MDE_ERR_UNSUPPORTED;
  BnPeers p;
  if (!fill_peers(p, peer_bases, world, rank)) return MDE_ERR_BAD_SHAPE;
  bn_bwd_reduce_p2p_kernel<<<bn_grid(N, C), 256, 0, (cudaStream_t)stream>>>(x, dy, N, C, mean, invstd, local, zero_next,
                                                                            bn_cl_log2(C), p, slot_off, flag_off, epoch,
                                                                            g_bn_epoch_ctr);
  return check_launch();
}

int mde_bn_bwd_apply_p2p_nhwc(const float* x, const float* dy, float* dx, int64_t N, int C, const float* mean,
                              const float* invstd, const float* weight, uint64_t my_base, int64_t slot_off, int64_t flag_off,
                              int world, uint64_t epoch, double count, mde_stream_t stream) {
  if (!x || !dy || !dx || !mean || !invstd || !my_base) return MDE_ERR_BAD_POINTER;
  if (N <= 0 || C <= 0 || count <= 0.0 || world < 1 || world > 8) return MDE_ERR_BAD_SHAPE;
  if (C % 4 != 0 || !aligned(x, 16) || !aligned(dy, 16) || !aligned(dx, 16)) return MDE_ERR_UNSUPPORTED;
  bn_bwd_apply_p2p_kernel<<<bn_grid(N, C), 256, 0, (cudaStream_t)stream>>>(x, dy, dx, N, C, mean, invstd, weight, my_base,
                                                                           slot_off, flag_off, world, epoch, count,
                                                                           bn_cl_log2(C), g_bn_epoch_ctr);
  return check_launch();
}

// CUDA-graph capture of a training step: while `counter` (device address of a uint64 the graph's first node increments once
// per replay) is set, the *_p2p launches above take `epoch` as a base added to the counter's value on the device and slot_off /
// flag_off as the parity-0 offsets of the direction.  NULL restores eager semantics.  Host-side state only (no stream work).
int mde_bn_p2p_set_epoch_counter(const uint64_t* counter) {
  g_bn_epoch_ctr = reinterpret_cast<const unsigned long long*>(counter);
  return MDE_OK;
}

// Bound of the peer-flag wait of the *_p2p kernels (seconds; default 600).  Synchronous (cudaMemcpyToSymbol).
int mde_bn_set_peer_timeout_seconds(double seconds) {
  if (!(seconds > 0.0)) return MDE_ERR_BAD_SHAPE;
  const unsigned long long ns = (unsigned long long)(seconds * 1e9);
  return cudaMemcpyToSymbol(g_bn_timeout_ns, &ns, sizeof(ns)) == cudaSuccess ? MDE_OK : MDE_ERR_LAUNCH;
}

// Diagnostic: nanoseconds block (0, 0) of the consuming SyncBatchNorm kernels has spent waiting for peer flags and the number of
// waits since the last reset (reset != 0 zeroes both after reading).  Synchronises.
int mde_bn_wait_stats(uint64_t* wait_ns, uint64_t* waits, int reset) {
  unsigned long long a = 0, b = 0;
  if (cudaMemcpyFromSymbol(&a, g_bn_wait_ns, sizeof(a)) != cudaSuccess || cudaMemcpyFromSymbol(&b, g_bn_waits, sizeof(b)) != cudaSuccess)
    return MDE_ERR_LAUNCH;
  if (wait_ns) *wait_ns = a;
  if (waits) *waits = b;
  if (reset) {
    const unsigned long long z = 0;
    if (cudaMemcpyToSymbol(g_bn_wait_ns, &z, sizeof(z)) != cudaSuccess || cudaMemcpyToSymbol(g_bn_waits, &z, sizeof(z)) != cudaSuccess)
      return MDE_ERR_LAUNCH;
  }
  return MDE_OK;
}

// Number of peer-flag waits that expired on this device since load (each one trapped its kernel).  Synchronises.
int mde_bn_peer_timeouts(void) {
  unsigned int v = 0;
  if (cudaMemcpyFromSymbol(&v, g_bn_peer_timeouts, sizeof(v)) != cudaSuccess) return -1;
  return (int)v;
}

}  // extern "C"
