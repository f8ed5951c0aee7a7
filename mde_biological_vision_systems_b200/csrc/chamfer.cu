// K5 -- bin-centre chamfer loss: 1-D nearest neighbour between the n_bins centres and the valid target depths.
//
// Reference: BinsChamferLoss.forward (loss.py:33-46) -> pytorch3d.loss.chamfer_distance (v0.6.1 defaults):
//   centres c_k = (e_k + e_{k+1})/2;  targets t >= 1e-3 of image b (T_b of them);
//   loss = 1/B * sum_b [ 1/n * sum_k min_t (c_k - t)^2  +  1/T_b * sum_t min_k (t - c_k)^2 ]
// The reference builds the full n x T distance matrix twice.  Because both clouds are 1-D and the centres are sorted
// (bin widths are positive), this kernel does:
//   * targets -> nearest centre: binary search in the shared-memory-resident centres (8 steps for 256), d^2
//     accumulated in float64 (warp shuffle + one atomic per block);
//   * centres -> nearest target: every target falls in one of n+1 intervals between consecutive centres; per
//     interval the min and max target are kept (shared-memory atomicMin/Max on the float bit pattern -- targets are
//     positive so integer order == float order).  The nearest target of c_k is then either the largest target
//     below it (prefix max) or the smallest above it (suffix min): O(T log n + n) instead of O(n T).
//   The squared distance of the winner is computed as (c - t)^2 in fp32 exactly as the brute force would.
// HBM-bound: reads the target once (4 B/px).  One launch; the last block per image finalises that image, the last
// image finalises the batch mean (ticket counters) -- no host sync (the reference syncs for len() and pad_sequence).
#include "common.cuh"

namespace mde {

struct ChamferWs {  // offsets into the caller's scratch buffer
  unsigned int* imin;       // [B][n+1] float bits, +inf when empty
  unsigned int* imax;       // [B][n+1] float bits, 0 when empty
  double* sum_t;            // [B][n]  sum of targets assigned to centre k
  unsigned int* cnt_t;      // [B][n]  number of targets assigned to centre k
  double* sum_y;            // [B]     sum_t min_k (t-c_k)^2
  unsigned long long* n_y;  // [B]     T_b
  float* nn_t;              // [B][n]  nearest target of centre k (for backward)
  double* cham;             // [B][2]  per-image cham_x, cham_y
  unsigned int* ticket;     // [B]
  unsigned int* done;       // [1]
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__host__ __device__ inline size_t chamfer_layout(void* base, int B, int n, ChamferWs* w) {
  size_t off = 0;
  unsigned char* p = reinterpret_cast<unsigned char*>(base);
  auto take = [&](size_t bytes) {
    unsigned char* r = p ? p + off : nullptr;
    off = align_up(off + bytes, 16);
    return r;
  };
  unsigned char* a0 = take(sizeof(unsigned int) * (size_t)B * (n + 1));
  unsigned char* a1 = take(sizeof(unsigned int) * (size_t)B * (n + 1));
  unsigned char* a2 = take(sizeof(double) * (size_t)B * n);
  unsigned char* a3 = take(sizeof(unsigned int) * (size_t)B * n);
  unsigned char* a4 = take(sizeof(double) * (size_t)B);
  unsigned char* a5 = take(sizeof(unsigned long long) * (size_t)B);
  unsigned char* a6 = take(sizeof(float) * (size_t)B * n);
  unsigned char* a7 = take(sizeof(double) * (size_t)B * 2);
  unsigned char* a8 = take(sizeof(unsigned int) * (size_t)B);
  unsigned char* a9 = take(sizeof(unsigned int));
  if (w) {
    w->imin = reinterpret_cast<unsigned int*>(a0);
    w->imax = reinterpret_cast<unsigned int*>(a1);
    w->sum_t = reinterpret_cast<double*>(a2);
    w->cnt_t = reinterpret_cast<unsigned int*>(a3);
    w->sum_y = reinterpret_cast<double*>(a4);
    w->n_y = reinterpret_cast<unsigned long long*>(a5);
    w->nn_t = reinterpret_cast<float*>(a6);
    w->cham = reinterpret_cast<double*>(a7);
    w->ticket = reinterpret_cast<unsigned int*>(a8);
    w->done = reinterpret_cast<unsigned int*>(a9);
  }
  return off;
}

constexpr unsigned int F_INF = 0x7f800000u;

__global__ void chamfer_init_kernel(unsigned int* imin, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    imin[i] = F_INF;
}

// dynamic smem: centres[n] | imin[n+1] | imax[n+1] | sum_t[n] (float) | cnt_t[n]
__global__ void __launch_bounds__(256) chamfer_fwd_kernel(const float* __restrict__ edges, const float* __restrict__ target,
                                                          int B, int n, long long HW, float min_target, ChamferWs ws,
                                                          float* loss) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sc = reinterpret_cast<float*>(smem_raw);
  unsigned int* smin = reinterpret_cast<unsigned int*>(sc + n);
  unsigned int* smax = smin + (n + 1);
  float* ssum = reinterpret_cast<float*>(smax + (n + 1));
  unsigned int* scnt = reinterpret_cast<unsigned int*>(ssum + n);
  const int b = blockIdx.y;
  const float* e = edges + (long long)b * (n + 1);
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    sc[k] = 0.5f * (e[k + 1] + e[k]);  // loss.py:34 operand order
    ssum[k] = 0.f;
    scnt[k] = 0u;
  }
  for (int k = threadIdx.x; k <= n; k += blockDim.x) {
    smin[k] = F_INF;
    smax[k] = 0u;
  }
  __syncthreads();

  const float* tg = target + (long long)b * HW;
  double acc = 0.0;
  unsigned int cnt = 0;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
    const float t = tg[p];
    if (!(t >= min_target)) continue;  // loss.py:40  mask = target.ge(1e-3)
    // j = number of centres <= t  (upper bound) ; interval j = [c_{j-1}, c_j)
    int lo = 0, hi = n;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (sc[mid] <= t) lo = mid + 1;
      else hi = mid;
    }
    const int j = lo;
    float best = INFINITY;
    int kbest = 0;
    if (j > 0) {
      const float d = t - sc[j - 1];
      best = d * d;
      kbest = j - 1;
    }
    if (j < n) {
      const float d = t - sc[j];
      const float dd = d * d;
      if (dd < best) {
        best = dd;
        kbest = j;
      }
    }
    acc += (double)best;
    ++cnt;
    const unsigned int bits = __float_as_uint(t);
    atomicMin(&smin[j], bits);
    atomicMax(&smax[j], bits);
    atomicAdd(&ssum[kbest], t);
    atomicAdd(&scnt[kbest], 1u);
  }
  // block reduce of the y -> x term
  __shared__ double racc[8];
  __shared__ unsigned int rcnt[8];
  __shared__ bool last_block, last_image;
  acc = warp_sum(acc);
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0) {
    racc[threadIdx.x >> 5] = acc;
    rcnt[threadIdx.x >> 5] = cnt;
  }
  __syncthreads();
  // flush the per-block tables
  unsigned int* gmin = ws.imin + (long long)b * (n + 1);
  unsigned int* gmax = ws.imax + (long long)b * (n + 1);
  for (int k = threadIdx.x; k <= n; k += blockDim.x) {
    if (smin[k] != F_INF) atomicMin(&gmin[k], smin[k]);
    if (smax[k] != 0u) atomicMax(&gmax[k], smax[k]);
  }
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    if (scnt[k]) {
      atomicAdd(&ws.sum_t[(long long)b * n + k], (double)ssum[k]);
      atomicAdd(&ws.cnt_t[(long long)b * n + k], scnt[k]);
    }
  }
  if (threadIdx.x == 0) {
    double a = 0;
    unsigned long long c = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
      a += racc[i];
      c += rcnt[i];
    }
    atomicAdd(&ws.sum_y[b], a);
    atomicAdd(&ws.n_y[b], c);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last_block = atomicAdd(&ws.ticket[b], 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last_block) return;

  // ---- finalise image b: centres -> nearest target ----------------------------------------------------------
  __threadfence();
  for (int k = threadIdx.x; k <= n; k += blockDim.x) {
    smin[k] = atomicMin(&gmin[k], F_INF);  // atomic read (value unchanged)
    smax[k] = atomicMax(&gmax[k], 0u);
  }
  __syncthreads();
  // prefix max of smax (targets below c_k live in intervals 0..k), suffix min of smin (intervals k+1..n).
  // n+1 <= 1025 elements: a single warp walks them in chunks of 32 with shuffles.
  if (threadIdx.x < 32) {
    unsigned int carry = 0u;
    for (int base = 0; base <= n; base += 32) {
      const int k = base + threadIdx.x;
      unsigned int v = k <= n ? smax[k] : 0u;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int u = __shfl_up_sync(0xffffffffu, v, o);
        if ((int)threadIdx.x >= o) v = max(v, u);
      }
      v = max(v, carry);
      if (k <= n) smax[k] = v;
      carry = __shfl_sync(0xffffffffu, v, 31);
    }
  } else if (threadIdx.x < 64) {
    const int lane = threadIdx.x - 32;
    unsigned int carry = F_INF;
    for (int base = n; base >= 0; base -= 32) {
      const int k = base - lane;
      unsigned int v = k >= 0 ? smin[k] : F_INF;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v = min(v, u);
      }
      v = min(v, carry);
      if (k >= 0) smin[k] = v;
      carry = __shfl_sync(0xffffffffu, v, 31);
    }
  }
  __syncthreads();
  double dx = 0.0;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const float c = sc[k];
    const unsigned int lo_bits = smax[k];      // largest target in intervals 0..k  (all < c_k ... <= c_k)
    const unsigned int hi_bits = smin[k + 1];  // smallest target in intervals k+1..n (all >= c_k)
    float best = 0.f, tn = c;                  // no targets at all: pytorch3d leaves the distance at 0
    bool have = false;
    if (lo_bits != 0u) {
      const float t = __uint_as_float(lo_bits);
      const float d = c - t;
      best = d * d;
      tn = t;
      have = true;
    }
    if (hi_bits != F_INF) {
      const float t = __uint_as_float(hi_bits);
      const float d = c - t;
      const float dd = d * d;
      if (!have || dd < best) {
        best = dd;
        tn = t;
      }
    }
    ws.nn_t[(long long)b * n + k] = tn;
    dx += (double)best;
  }
  dx = warp_sum(dx);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) racc[threadIdx.x >> 5] = dx;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) a += racc[i];
    const double sy = atomicAdd(&ws.sum_y[b], 0.0);
    const unsigned long long ny = atomicAdd(&ws.n_y[b], 0ull);
    ws.cham[2 * b + 0] = a / (double)n;
    ws.cham[2 * b + 1] = sy / (double)ny;  // 0/0 -> NaN when the image has no valid target (as the reference)
    __threadfence();
    last_image = atomicAdd(ws.done, 1u) == (unsigned int)B - 1;
    if (last_image) {
      __threadfence();
      double cx = 0, cy = 0;
      for (int i = 0; i < B; ++i) {
        cx += *((volatile double*)&ws.cham[2 * i + 0]);
        cy += *((volatile double*)&ws.cham[2 * i + 1]);
      }
      *loss = (float)(cx / B + cy / B);
    }
  }
}

// grad wrt edges.  d/dc_k = g/B * [ 2 (c_k - nn_t_k)/n + 2 (cnt_k c_k - sum_t_k)/T_b ];  e_j gets half of c_{j-1}, c_j.
__global__ void chamfer_bwd_kernel(const float* __restrict__ edges, int B, int n, ChamferWs ws,
                                   const float* __restrict__ grad_loss, float* __restrict__ grad_edges) {
  const int b = blockIdx.x;
  extern __shared__ float gc[];  // [n]
  const float* e = edges + (long long)b * (n + 1);
  const double g = (double)grad_loss[0] / (double)B;
  const double T = (double)ws.n_y[b];
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const double c = (double)(0.5f * (e[k + 1] + e[k]));
    const double gx = 2.0 * (c - (double)ws.nn_t[(long long)b * n + k]) / (double)n;
    const double gy = 2.0 * ((double)ws.cnt_t[(long long)b * n + k] * c - ws.sum_t[(long long)b * n + k]) / T;
    gc[k] = (float)(g * (gx + gy));
  }
  __syncthreads();
  for (int j = threadIdx.x; j <= n; j += blockDim.x) {
    float v = 0.f;
    if (j > 0) v += 0.5f * gc[j - 1];
    if (j < n) v += 0.5f * gc[j];
    grad_edges[(long long)b * (n + 1) + j] = v;
  }
}

}  // namespace mde

using namespace mde;

extern "C" {

int64_t mde_chamfer_ws_bytes(int B, int n_bins) {
  if (B <= 0 || n_bins <= 0) return 0;
  return (int64_t)chamfer_layout(nullptr, B, n_bins, nullptr);
}

int mde_chamfer_fwd(const float* edges, const float* target, int B, int n_bins, int64_t HW, float min_target, void* ws,
                    float* loss, mde_stream_t stream) {
  if (!edges || !target || !ws || !loss) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || B > 65535 || n_bins <= 0 || n_bins > 2048 || HW <= 0) return MDE_ERR_BAD_SHAPE;
  if (!aligned(ws, 16)) return MDE_ERR_BAD_POINTER;
  cudaStream_t st = (cudaStream_t)stream;
  ChamferWs w;
  const size_t bytes = chamfer_layout(ws, B, n_bins, &w);
  cudaMemsetAsync(ws, 0, bytes, st);
  const long long nmin = (long long)B * (n_bins + 1);
  chamfer_init_kernel<<<(unsigned)((nmin + 255) / 256), 256, 0, st>>>(w.imin, nmin);
  int rc = check_launch();
  if (rc) return rc;
  long long gx = (HW + 2047) / 2048;
  const long long cap = (2 * MDE_NUM_SMS + B - 1) / B;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  const size_t sm = sizeof(float) * (size_t)n_bins * 3 + sizeof(unsigned int) * (size_t)(n_bins + 1) * 2;
  chamfer_fwd_kernel<<<dim3((unsigned)gx, (unsigned)B), 256, sm, st>>>(edges, target, B, n_bins, HW, min_target, w, loss);
  return check_launch();
}

int mde_chamfer_bwd(const float* edges, int B, int n_bins, const void* ws, const float* grad_loss, float* grad_edges,
                    mde_stream_t stream) {
  if (!edges || !ws || !grad_loss || !grad_edges) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || n_bins <= 0 || n_bins > 2048) return MDE_ERR_BAD_SHAPE;
  ChamferWs w;
  chamfer_layout(const_cast<void*>(ws), B, n_bins, &w);
  chamfer_bwd_kernel<<<B, 256, sizeof(float) * n_bins, (cudaStream_t)stream>>>(edges, B, n_bins, w, grad_loss, grad_edges);
  return check_launch();
}

}  // extern "C"
