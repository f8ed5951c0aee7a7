// K4 + K5 -- SILog loss and bin-centre chamfer loss, forward: ONE streaming kernel that reads the target depth once.
//
// Reference: SILogLoss.forward (loss.py:12-25):
//   input = interpolate(input, target.shape[-2:], 'bilinear', align_corners=True); input,target = [mask]
//   g = log(input) - log(target);  Dg = var(g) + 0.15*mean(g)^2  (unbiased var over ALL masked pixels of the batch);  10*sqrt(Dg)
// and BinsChamferLoss.forward (loss.py:33-46) -> pytorch3d.loss.chamfer_distance (v0.6.1 defaults):
//   centres c_k = (e_k + e_{k+1})/2;  targets t >= 1e-3 of image b (T_b of them);
//   loss = 1/B * sum_b [ 1/n * sum_k min_t (c_k - t)^2  +  1/T_b * sum_t min_k (t - c_k)^2 ]
// The reference materialises the up-sampled prediction, runs nonzero() for the mask, and builds the full n x T distance
// matrix twice.  Here (template flags select SILog only / chamfer only / both, so the two drop-in modules and the fused
// call share one kernel):
//   * grid (blocks per image, B): no 64-bit index arithmetic; a thread owns 4 consecutive pixels (one 16-byte load);
//   * SILog: the bilinear sample of the quarter-size prediction (L1/L2 resident) is taken in registers; sum g, sum g^2, n
//     are carried in fp32 per thread (<= a few dozen pixels) and widened to fp64 at the warp / block merge;
//   * chamfer, targets -> nearest centre: the centres are sorted (bin widths are positive; verified on the device, see
//     `unsorted`): a uniform-grid table over [edge_0, edge_n] (built per block by binary search) gives the starting centre
//     and the search walks 0-2 steps forward in the shared-memory-resident centres;
//   * chamfer, centres -> nearest target: every target falls in one of n+1 intervals between consecutive centres; per
//     interval the min and max target are kept in WARP-PRIVATE shared-memory tables (atomicMin/Max on the float bits --
//     targets are positive so integer order == float order), merged per block, written to a per-block row of the scratch
//     buffer with plain stores, and reduced by the last block of the image (ticket counter): prefix max / suffix min give
//     the nearest target below / above every centre: O(T log n + n) instead of O(n T), and no global atomics besides the
//     tickets;
//   * every partial sum goes to its own scratch slot and is added up in a fixed order by the finishing block: results are
//     bit-reproducible run to run.
// One launch, no host synchronisation (the reference syncs for nonzero(), len() and pad_sequence).  HBM traffic: the target
// once (4 B/px) + the optional explicit mask (1 B/px); algorithmic bytes per image F*4 + P*4.
#include "common.cuh"

namespace mde {

constexpr int LS_THREADS = 512;
constexpr int LS_WARPS = LS_THREADS / 32;
constexpr int LS_MAX_BLOCKS = MDE_NUM_SMS * 16;  // bound of blocks per launch (scratch sizing)
constexpr unsigned int F_INF = 0x7f800000u;
constexpr int LS_LUT = 1024;                    // cells of the search accelerator

struct SilogWs {  // read by silog_bwd_kernel
  double sum, sumsq, count;
  unsigned int ticket;
  unsigned int pad;
};

struct ChamferWs {  // offsets into the caller's scratch buffer (nn_t, sum_t, cnt_t, n_y are read by chamfer_bwd_kernel)
  double* sum_t;            // [B][n]  sum of targets assigned to centre k
  unsigned int* cnt_t;      // [B][n]  number of targets assigned to centre k
  unsigned long long* n_y;  // [B]     T_b
  float* nn_t;              // [B][n]  nearest target of centre k
  double* cham;             // [B][2]  per-image cham_x, cham_y
  unsigned int* ticket;     // [B]
  unsigned int* done;       // [1]
  unsigned int* unsorted;   // [1]     raised when the centres of some image are not ascending
  unsigned int* blk_min;    // [B][bx][n+1] per-block interval minima (float bits)
  unsigned int* blk_max;    // [B][bx][n+1]
  float* blk_sum;           // [B][bx][n]   per-block sum of targets per nearest centre   (gradient runs only)
  unsigned int* blk_cnt;    // [B][bx][n]
  double* blk_part;         // [B][bx][2]   per-block {sum_t min_k d^2, T}
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Blocks per image: two 512-thread CTAs per SM over the whole batch.  More, smaller blocks stream no faster (the pass is
// ~10 us of HBM time) but lengthen the finishing block's reduction over the per-block rows, which is the serial tail.
inline int loss_blocks_per_image(int B, long long HW) {
  long long bx = (HW / 4 + LS_THREADS * 2 - 1) / (LS_THREADS * 2);  // >= 8 pixels per thread
  // the whole grid must be resident at once (2 CTAs of 512 threads per SM): rounding the cap UP put 304 blocks on 296 slots at
  // B = 16 and the eight stragglers doubled the kernel's duration (sm__cycles_active = half of elapsed in the ncu capture)
  long long cap = (MDE_NUM_SMS * 2) / B;
  if (cap < 1) cap = 1;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  return (int)bx;
}

__host__ __device__ inline size_t chamfer_layout(void* base, int B, int n, int bx, ChamferWs* w) {
  size_t off = 0;
  unsigned char* p = reinterpret_cast<unsigned char*>(base);
  auto take = [&](size_t bytes) {
    unsigned char* r = p ? p + off : nullptr;
    off = align_up(off + bytes, 16);
    return r;
  };
  unsigned char* a0 = take(sizeof(double) * (size_t)B * n);
  unsigned char* a1 = take(sizeof(unsigned int) * (size_t)B * n);
  unsigned char* a2 = take(sizeof(unsigned long long) * (size_t)B);
  unsigned char* a3 = take(sizeof(float) * (size_t)B * n);
  unsigned char* a4 = take(sizeof(double) * (size_t)B * 2);
  unsigned char* a5 = take(sizeof(unsigned int) * (size_t)B);
  unsigned char* a6 = take(sizeof(unsigned int));
  unsigned char* a7 = take(sizeof(unsigned int));
  const size_t head = off;  // everything up to here is zeroed by the launcher
  unsigned char* b0 = take(sizeof(unsigned int) * (size_t)B * bx * (n + 1));
  unsigned char* b1 = take(sizeof(unsigned int) * (size_t)B * bx * (n + 1));
  unsigned char* b2 = take(sizeof(float) * (size_t)B * bx * n);
  unsigned char* b3 = take(sizeof(unsigned int) * (size_t)B * bx * n);
  unsigned char* b4 = take(sizeof(double) * (size_t)B * bx * 2);
  if (w) {
    w->sum_t = reinterpret_cast<double*>(a0);
    w->cnt_t = reinterpret_cast<unsigned int*>(a1);
    w->n_y = reinterpret_cast<unsigned long long*>(a2);
    w->nn_t = reinterpret_cast<float*>(a3);
    w->cham = reinterpret_cast<double*>(a4);
    w->ticket = reinterpret_cast<unsigned int*>(a5);
    w->done = reinterpret_cast<unsigned int*>(a6);
    w->unsorted = reinterpret_cast<unsigned int*>(a7);
    w->blk_min = reinterpret_cast<unsigned int*>(b0);
    w->blk_max = reinterpret_cast<unsigned int*>(b1);
    w->blk_sum = reinterpret_cast<float*>(b2);
    w->blk_cnt = reinterpret_cast<unsigned int*>(b3);
    w->blk_part = reinterpret_cast<double*>(b4);
  }
  return base ? head : off;  // with a base: bytes to zero; without: total size
}

// torch's area_pixel_compute_source_index for align_corners=True: src = dst * (in-1)/(out-1) in float
__device__ __forceinline__ void src_index(int dst, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
  const float s = scale * (float)dst;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = s - (float)i0;
  l0 = 1.f - l1;
}

struct LossArgs {
  // SILog
  const float* pred;            // [B][h][w]
  const unsigned char* mask;    // [B][H][W] or null
  float mask_thr;               // MASK == 2: valid iff target > mask_thr (train.py:414 mask = depth > args.min_depth)
  int h, w, H, W;
  float sy, sx;
  double* silog_part;           // [B * bx][3] per-block {sum g, sum g^2, n}
  SilogWs* silog_ws;
  float* silog_loss;
  // chamfer
  const float* edges;           // [B][n+1]
  int n;
  int search_step;              // largest power of two <= n
  float min_target;
  ChamferWs cw;
  float* chamfer_loss;
  // common
  const float* target;          // [B][HW]
  int B, HW;
};

// (the centres used to be stored skewed by one word per 32 for the binary search; with the uniform-grid table in front of it
// the search touches 1-3 neighbouring centres per pixel and the extra address arithmetic cost more than the conflicts)
__device__ __forceinline__ int skew(int k) { return k; }

// log2 of a positive NORMAL float on the SFU (depths and predictions are >= 1e-3): the bare instruction, without the
// denormal rescaling __log2f wraps around it (4 extra instructions per call, 8 calls per group)
__device__ __forceinline__ float lg2_fast(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// MASK: 0 none, 1 explicit uint8 mask, 2 derived (target > mask_thr).  dynamic smem (CHAMFER): centres[skew(n)+1] |
// wmin[LS_WARPS][n+1] | wmax[LS_WARPS][n+1] | (GRAD) wsum[LS_WARPS][n] float | wcnt[LS_WARPS][n]
template <bool SILOG, bool INTERP, int MASK, bool CHAMFER, bool GRAD>
__global__ void __launch_bounds__(LS_THREADS) depth_losses_kernel(const LossArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = a.n;
  const int b = blockIdx.y, bx = gridDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sc = reinterpret_cast<float*>(smem_raw);
  const int nsk = CHAMFER ? skew(n) + 1 : 0;
  unsigned int* wmin = reinterpret_cast<unsigned int*>(sc + nsk);
  unsigned int* wmax = wmin + LS_WARPS * (n + 1);
  float* wsum = reinterpret_cast<float*>(wmax + LS_WARPS * (n + 1));
  unsigned int* wcnt = reinterpret_cast<unsigned int*>(wsum + LS_WARPS * n);
  __shared__ double red[LS_WARPS][5];
  __shared__ bool last_block, last_image;
  __shared__ unsigned int bad_order;
  __shared__ unsigned short lut[LS_LUT];
  float lut_lo = 0.f, lut_scale = 0.f;

  if (CHAMFER) {
    const float* e = a.edges + (long long)b * (n + 1);
    if (threadIdx.x == 0) bad_order = 0u;
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += LS_THREADS) {
      const float c = 0.5f * (e[k + 1] + e[k]);  // loss.py:34 operand order
      sc[skew(k)] = c;
      if (k > 0 && !(c >= 0.5f * (e[k] + e[k - 1]))) bad_order = 1u;
    }
    for (int k = threadIdx.x; k < LS_WARPS * (n + 1); k += LS_THREADS) {
      wmin[k] = F_INF;
      wmax[k] = 0u;
    }
    if (GRAD)
      for (int k = threadIdx.x; k < LS_WARPS * n; k += LS_THREADS) {
        wsum[k] = 0.f;
        wcnt[k] = 0u;
      }
    __syncthreads();
    if (threadIdx.x == 0 && bad_order && blockIdx.x == 0) atomicExch(a.cw.unsorted, 1u);
    // search accelerator: lut[c] = number of centres <= the left edge of cell c of a uniform grid over [edge_0, edge_n]
    // (binary search per cell, 2 cells per thread); a target then starts its search at lut[cell(t)] and walks forward --
    // 0-2 steps for any reasonable bin layout, instead of log2(n) dependent shared-memory probes per pixel
    lut_lo = e[0];
    {
      const float span = e[n] - e[0];
      lut_scale = span > 0.f ? (float)LS_LUT / span : 0.f;
    }
    for (int c = threadIdx.x; c < LS_LUT; c += LS_THREADS) {
      const float left = lut_lo + (float)c / lut_scale;
      int j = 0;
      for (int step = a.search_step; step > 0; step >>= 1) {
        const int probe = j + step;
        if (probe <= n && sc[skew(probe - 1)] <= left) j = probe;
      }
      // the walk only moves forward, so the start must not overshoot: step back over centres equal to `left` (rounding of
      // the cell edge) -- one position is enough, the forward walk re-adds it
      lut[c] = (unsigned short)(j > 0 ? j - 1 : 0);
    }
    __syncthreads();
  }

  // ---- stream this block's share of image b: groups of 4 consecutive pixels ------------------------------------
  const float* tg = a.target + (long long)b * a.HW;
  const unsigned char* mk = MASK == 1 ? a.mask + (long long)b * a.HW : nullptr;
  const float* pb = SILOG ? a.pred + (long long)b * a.h * a.w : nullptr;
  float s_g = 0.f, s_gg = 0.f, s_d = 0.f;
  unsigned int n_g = 0, n_t = 0;
  unsigned int* my_min = wmin + warp * (n + 1);
  unsigned int* my_max = wmax + warp * (n + 1);
  const int groups = (a.HW + 3) >> 2;
  const bool vec = (a.HW & 3) == 0;
  // ---- fast path: whole groups, the four pixels of a group share their image row (W % 4 == 0), mask derived from the target.
  // Branch-free per pixel -- an invalid pixel computes with a harmless stand-in and its contribution is selected away, so the
  // compiler interleaves the four pixels and the next group's depth is already in flight while this one is reduced.  (The
  // general path below keeps a validity branch, a row-end test and a search loop per pixel: ~190 warp instructions per pixel,
  // one dependent chain per group -- profiles/r2_ncu_full_summary.txt.)
  const bool fast = vec && MASK != 1 && (!SILOG || !INTERP || (a.W & 3) == 0);
  if (fast) {
    const int stride = bx * LS_THREADS;
    int gidx = blockIdx.x * LS_THREADS + threadIdx.x;
    float4 nxt = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gidx < groups) nxt = ldg_stream(reinterpret_cast<const float4*>(tg) + gidx);
    for (; gidx < groups; gidx += stride) {
      const float4 v4 = nxt;
      if (gidx + stride < groups) nxt = ldg_stream(reinterpret_cast<const float4*>(tg) + gidx + stride);
      const float t4[4] = {v4.x, v4.y, v4.z, v4.w};
      const int p0 = gidx << 2;
      if (SILOG) {
        float pv[4];
        if (INTERP) {
          const int y = p0 / a.W, x = p0 - y * a.W;
          int y0, y1;
          float ly0, ly1;
          src_index(y, a.sy, a.h, y0, y1, ly0, ly1);
          const unsigned int o0 = (unsigned int)(y0 * a.w), o1 = (unsigned int)(y1 * a.w);  // 32-bit offsets: one LEA pair per load
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            int x0, x1;
            float lx0, lx1;
            src_index(x + i, a.sx, a.w, x0, x1, lx0, lx1);
            pv[i] = ly0 * (lx0 * __ldg(pb + (o0 + (unsigned int)x0)) + lx1 * __ldg(pb + (o0 + (unsigned int)x1))) +
                    ly1 * (lx0 * __ldg(pb + (o1 + (unsigned int)x0)) + lx1 * __ldg(pb + (o1 + (unsigned int)x1)));
          }
        } else {
          const float4 q = __ldg(reinterpret_cast<const float4*>(pb) + gidx);
          pv[0] = q.x; pv[1] = q.y; pv[2] = q.z; pv[3] = q.w;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool valid = MASK == 0 || t4[i] > a.mask_thr;
          const float g = (lg2_fast(pv[i]) - lg2_fast(t4[i])) * 0.6931471805599453f;
          const float gs = valid ? g : 0.f;  // a select: the inf / NaN of an invalid pixel never enters the sums
          s_g += gs;
          s_gg = fmaf(gs, gs, s_gg);
          n_g += valid ? 1u : 0u;
        }
      }
      if (CHAMFER) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool valid = t4[i] >= a.min_target;       // loss.py:40  mask = target.ge(1e-3)
          const float t = valid ? t4[i] : lut_lo;         // stand-in inside the table's range
          const float f = (t - lut_lo) * lut_scale;
          const int cell = f > 0.f ? (f < (float)(LS_LUT - 1) ? (int)f : LS_LUT - 1) : 0;
          int j = lut[cell];                               // number of centres <= t is >= j: walk forward
          j += (j < n && sc[skew(j)] <= t) ? 1 : 0;
          j += (j < n && sc[skew(j)] <= t) ? 1 : 0;
          while (j < n && sc[skew(j)] <= t) ++j;           // rare: more than two centres inside one cell of the grid
          // nearest centre: c_{j-1} or c_j; at the ends the clamped index repeats the only candidate
          const int jl = j > 0 ? j - 1 : 0, jr = j < n ? j : n - 1;
          const float dl = t - sc[skew(jl)], dr = t - sc[skew(jr)];
          const float ddl = dl * dl, ddr = dr * dr;
          const bool right = ddr < ddl;                    // the left candidate wins ties, as in the general path
          s_d += valid ? (right ? ddr : ddl) : 0.f;
          n_t += valid ? 1u : 0u;
          const unsigned int bits = __float_as_uint(t4[i]);
          // interval tables: a plain read first -- the entries only move one way, so a stale read can cause a redundant
          // atomic but never a missed one; once the tables have warmed up almost no pixel improves its interval and the
          // shared-memory atomics (the kernel's scarcest resource: two per pixel, 32 scattered addresses per warp) disappear
          const unsigned int lo_new = valid ? bits : F_INF, hi_new = valid ? bits : 0u;  // neutral for an invalid pixel
          if (lo_new < my_min[j]) atomicMin(&my_min[j], lo_new);
          if (hi_new > my_max[j]) atomicMax(&my_max[j], hi_new);
          if (GRAD && valid) {
            const int kbest = right ? jr : jl;
            atomicAdd(&wsum[warp * n + kbest], t);
            atomicAdd(&wcnt[warp * n + kbest], 1u);
          }
        }
      }
    }
  }
  for (int gidx = blockIdx.x * LS_THREADS + threadIdx.x; !fast && gidx < groups; gidx += bx * LS_THREADS) {
    const int p0 = gidx << 2;
    float t4[4];
    if (vec) {
      const float4 v = ldg_stream(reinterpret_cast<const float4*>(tg) + gidx);
      t4[0] = v.x; t4[1] = v.y; t4[2] = v.z; t4[3] = v.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) t4[i] = p0 + i < a.HW ? tg[p0 + i] : 0.f;
    }
    unsigned int m4 = 0xF;
    if (MASK == 1) {
      m4 = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (p0 + i < a.HW && mk[p0 + i]) m4 |= 1u << i;
    }
    if (SILOG) {
      int y = p0 / a.W, x = p0 - y * a.W;
      int y0 = 0, y1 = 0;
      float ly0 = 0.f, ly1 = 0.f;
      bool row_ready = false;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool in = p0 + i < a.HW;
        const bool valid = in && (MASK == 0 || (MASK == 1 ? ((m4 >> i) & 1u) != 0 : t4[i] > a.mask_thr));
        if (valid) {
          float v;
          if (INTERP) {
            if (!row_ready) {  // the four pixels of a group share their row except across a row end
              src_index(y, a.sy, a.h, y0, y1, ly0, ly1);
              row_ready = true;
            }
            int x0, x1;
            float lx0, lx1;
            src_index(x, a.sx, a.w, x0, x1, lx0, lx1);
            v = ly0 * (lx0 * __ldg(pb + y0 * a.w + x0) + lx1 * __ldg(pb + y0 * a.w + x1)) +
                ly1 * (lx0 * __ldg(pb + y1 * a.w + x0) + lx1 * __ldg(pb + y1 * a.w + x1));
          } else {
            v = __ldg(pb + p0 + i);
          }
          // log(v) - log(t) on the SFU (lg2.approx, ~2^-22 absolute in log2 units): the rounding of each pixel's term is far
          // below the 1e-4 the loss is held to and unbiased over the ~10^6 pixels of the sum
          const float g = (__log2f(v) - __log2f(t4[i])) * 0.6931471805599453f;
          s_g += g;
          s_gg = fmaf(g, g, s_gg);
          ++n_g;
        }
        if (++x == a.W) {
          x = 0;
          ++y;
          row_ready = false;
        }
      }
    }
    if (CHAMFER) {
      // j = number of centres <= t (upper bound), interval j = [c_{j-1}, c_j): start from the uniform-grid table and walk
      int j4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float f = (t4[i] - lut_lo) * lut_scale;
        const int cell = f > 0.f ? (f < (float)(LS_LUT - 1) ? (int)f : LS_LUT - 1) : 0;
        j4[i] = lut[cell];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int j = j4[i];
        while (j < n && sc[skew(j)] <= t4[i]) ++j;
        j4[i] = j;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float t = t4[i];
        if (!(p0 + i < a.HW) || !(t >= a.min_target)) continue;  // loss.py:40  mask = target.ge(1e-3)
        const int j = j4[i];
        float best = INFINITY;
        int kbest = 0;
        if (j > 0) {
          const float d = t - sc[skew(j - 1)];
          best = d * d;
          kbest = j - 1;
        }
        if (j < n) {
          const float d = t - sc[skew(j)];
          const float dd = d * d;
          if (dd < best) {
            best = dd;
            kbest = j;
          }
        }
        s_d += best;
        ++n_t;
        const unsigned int bits = __float_as_uint(t);
        if (bits < my_min[j]) atomicMin(&my_min[j], bits);  // see the fast path: read first, atomics only on improvement
        if (bits > my_max[j]) atomicMax(&my_max[j], bits);
        if (GRAD) {
          atomicAdd(&wsum[warp * n + kbest], t);
          atomicAdd(&wcnt[warp * n + kbest], 1u);
        }
      }
    }
  }

  // ---- block merge (fp64 from here on) -------------------------------------------------------------------------
  {
    double v0 = warp_sum((double)s_g), v1 = warp_sum((double)s_gg), v3 = warp_sum((double)s_d);
    const unsigned int c2 = __reduce_add_sync(0xffffffffu, n_g), c4 = __reduce_add_sync(0xffffffffu, n_t);
    if (lane == 0) {
      red[warp][0] = v0; red[warp][1] = v1; red[warp][2] = (double)c2; red[warp][3] = v3; red[warp][4] = (double)c4;
    }
  }
  __syncthreads();
  const int blk = b * bx + blockIdx.x;
  if (threadIdx.x < 5) {
    double acc = 0.0;
    for (int i = 0; i < LS_WARPS; ++i) acc += red[i][threadIdx.x];
    if (SILOG && threadIdx.x < 3) a.silog_part[(long long)blk * 3 + threadIdx.x] = acc;
    if (CHAMFER && threadIdx.x >= 3) a.cw.blk_part[(long long)blk * 2 + (threadIdx.x - 3)] = acc;
  }
  if (CHAMFER) {
    unsigned int* gmin = a.cw.blk_min + (long long)blk * (n + 1);
    unsigned int* gmax = a.cw.blk_max + (long long)blk * (n + 1);
    for (int k = threadIdx.x; k <= n; k += LS_THREADS) {
      unsigned int lo = F_INF, hi = 0u;
#pragma unroll
      for (int wv = 0; wv < LS_WARPS; ++wv) {
        lo = min(lo, wmin[wv * (n + 1) + k]);
        hi = max(hi, wmax[wv * (n + 1) + k]);
      }
      gmin[k] = lo;
      gmax[k] = hi;
    }
    if (GRAD) {
      float* gsum = a.cw.blk_sum + (long long)blk * n;
      unsigned int* gcnt = a.cw.blk_cnt + (long long)blk * n;
      for (int k = threadIdx.x; k < n; k += LS_THREADS) {
        float sacc = 0.f;
        unsigned int cacc = 0u;
#pragma unroll
        for (int wv = 0; wv < LS_WARPS; ++wv) {
          sacc += wsum[wv * n + k];
          cacc += wcnt[wv * n + k];
        }
        gsum[k] = sacc;
        gcnt[k] = cacc;
      }
    }
  }
  __threadfence();
  __syncthreads();
  unsigned int* ticket = CHAMFER ? a.cw.ticket + b : &a.silog_ws->ticket;
  const unsigned int expect = CHAMFER ? (unsigned int)bx : (unsigned int)(bx * a.B);
  if (threadIdx.x == 0) last_block = atomicAdd(ticket, 1u) == expect - 1;
  __syncthreads();
  if (!last_block) return;
  __threadfence();

  if (CHAMFER) {
    // ---- finalise image b: reduce the per-block rows, then centres -> nearest target -------------------------------
    unsigned int* smin = wmin;           // reuse warp 0's tables as the image-level ones
    unsigned int* smax = wmax;
    for (int k = threadIdx.x; k <= n; k += LS_THREADS) {
      unsigned int lo = F_INF, hi = 0u;
      const unsigned int* pmin = a.cw.blk_min + (long long)b * bx * (n + 1) + k;
      const unsigned int* pmax = a.cw.blk_max + (long long)b * bx * (n + 1) + k;
      int i = 0;
      for (; i + 3 < bx; i += 4) {  // eight independent L2 loads per trip (this is the serial tail of the launch)
        const unsigned int l0 = __ldcg(pmin + (long long)i * (n + 1)), l1 = __ldcg(pmin + (long long)(i + 1) * (n + 1));
        const unsigned int l2 = __ldcg(pmin + (long long)(i + 2) * (n + 1)), l3 = __ldcg(pmin + (long long)(i + 3) * (n + 1));
        const unsigned int h0 = __ldcg(pmax + (long long)i * (n + 1)), h1 = __ldcg(pmax + (long long)(i + 1) * (n + 1));
        const unsigned int h2 = __ldcg(pmax + (long long)(i + 2) * (n + 1)), h3 = __ldcg(pmax + (long long)(i + 3) * (n + 1));
        lo = min(min(lo, min(l0, l1)), min(l2, l3));
        hi = max(max(hi, max(h0, h1)), max(h2, h3));
      }
      for (; i < bx; ++i) {
        lo = min(lo, __ldcg(pmin + (long long)i * (n + 1)));
        hi = max(hi, __ldcg(pmax + (long long)i * (n + 1)));
      }
      smin[k] = lo;
      smax[k] = hi;
    }
    if (GRAD)
      for (int k = threadIdx.x; k < n; k += LS_THREADS) {
        double sacc = 0.0;
        unsigned int cacc = 0u;
        for (int i = 0; i < bx; ++i) {
          sacc += (double)__ldcg(a.cw.blk_sum + ((long long)b * bx + i) * n + k);
          cacc += __ldcg(a.cw.blk_cnt + ((long long)b * bx + i) * n + k);
        }
        a.cw.sum_t[(long long)b * n + k] = sacc;
        a.cw.cnt_t[(long long)b * n + k] = cacc;
      }
    __syncthreads();
    // prefix max of smax (targets below c_k live in intervals 0..k), suffix min of smin (intervals k+1..n).
    // n+1 <= 2049 elements: a single warp walks them in chunks of 32 with shuffles.
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
      const int l = threadIdx.x - 32;
      unsigned int carry = F_INF;
      for (int base = n; base >= 0; base -= 32) {
        const int k = base - l;
        unsigned int v = k >= 0 ? smin[k] : F_INF;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned int u = __shfl_up_sync(0xffffffffu, v, o);
          if (l >= o) v = min(v, u);
        }
        v = min(v, carry);
        if (k >= 0) smin[k] = v;
        carry = __shfl_sync(0xffffffffu, v, 31);
      }
    }
    __syncthreads();
    double dx = 0.0;
    for (int k = threadIdx.x; k < n; k += LS_THREADS) {
      const float c = sc[skew(k)];
      const unsigned int lo_bits = smax[k];      // largest target in intervals 0..k  (all <= c_k)
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
      a.cw.nn_t[(long long)b * n + k] = tn;
      dx += (double)best;
    }
    double psy = 0.0, pny = 0.0;
    for (int i = threadIdx.x; i < bx; i += LS_THREADS) {
      psy += __ldcg(a.cw.blk_part + ((long long)b * bx + i) * 2);
      pny += __ldcg(a.cw.blk_part + ((long long)b * bx + i) * 2 + 1);
    }
    dx = warp_sum(dx);
    psy = warp_sum(psy);
    pny = warp_sum(pny);
    if (lane == 0) {
      red[warp][0] = dx; red[warp][1] = psy; red[warp][2] = pny;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double ax = 0.0, sy = 0.0, ny = 0.0;
      for (int i = 0; i < LS_WARPS; ++i) {
        ax += red[i][0]; sy += red[i][1]; ny += red[i][2];
      }
      a.cw.n_y[b] = (unsigned long long)ny;
      a.cw.cham[2 * b + 0] = ax / (double)n;
      a.cw.cham[2 * b + 1] = sy / ny;  // 0/0 -> NaN when the image has no valid target (as the reference)
      __threadfence();
      last_image = atomicAdd(a.cw.done, 1u) == (unsigned int)a.B - 1;
    }
    __syncthreads();
    if (!last_image) return;
    __threadfence();
    if (threadIdx.x < 32) {
      // one warp: lane i reads image i (+32, ...) -- all L2 loads in flight at once (a scalar loop paid one round trip per
      // image at the very end of the launch); the shuffle tree adds in a fixed order, so the result is reproducible
      double cx = 0, cy = 0;
      for (int i = threadIdx.x; i < a.B; i += 32) {
        cx += __ldcg(a.cw.cham + 2 * i);
        cy += __ldcg(a.cw.cham + 2 * i + 1);
      }
      cx = warp_sum(cx);
      cy = warp_sum(cy);
      if (threadIdx.x == 0) {
        // an unsorted centre vector (never produced by the model: widths are positive) would make the search above
        // meaningless: report NaN instead of a silently wrong loss
        const bool bad = __ldcg(a.cw.unsorted) != 0u;
        *a.chamfer_loss = bad ? __int_as_float(0x7fc00000) : (float)(cx / a.B + cy / a.B);
      }
    }
  }
  if (SILOG) {
    // the block that gets here is the last of the whole launch (chamfer: last block of the last image to finish; plain
    // SILog: last ticket): add the per-block partials in index order
    const int nblk = bx * a.B;
    double v0 = 0.0, v1 = 0.0, v2 = 0.0;
    for (int i = threadIdx.x; i < nblk; i += LS_THREADS) {
      v0 += __ldcg(a.silog_part + (long long)i * 3);
      v1 += __ldcg(a.silog_part + (long long)i * 3 + 1);
      v2 += __ldcg(a.silog_part + (long long)i * 3 + 2);
    }
    v0 = warp_sum(v0);
    v1 = warp_sum(v1);
    v2 = warp_sum(v2);
    __syncthreads();
    if (lane == 0) {
      red[warp][0] = v0; red[warp][1] = v1; red[warp][2] = v2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double S = 0, SS = 0, N = 0;
      for (int i = 0; i < LS_WARPS; ++i) {
        S += red[i][0]; SS += red[i][1]; N += red[i][2];
      }
      a.silog_ws->sum = S;
      a.silog_ws->sumsq = SS;
      a.silog_ws->count = N;
      const double mean = S / N;
      const double var = (SS - S * S / N) / (N - 1.0);  // torch.var default: unbiased
      *a.silog_loss = (float)(10.0 * sqrt(var + 0.15 * mean * mean));
    }
  }
}

// MASK as above
template <bool INTERP, int MASK>
__global__ void __launch_bounds__(256) silog_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                        const unsigned char* __restrict__ mask, float mask_thr, int h, int w,
                                                        int H, int W, float sy, float sx, const SilogWs* ws,
                                                        const float* __restrict__ grad_loss, float* grad_pred) {
  const double S = ws->sum, SS = ws->sumsq, N = ws->count;
  const double mean = S / N;
  const double var = (SS - S * S / N) / (N - 1.0);
  const double root = sqrt(var + 0.15 * mean * mean);
  // d(10*sqrt(Dg))/dg_i = 5/sqrt(Dg) * ( 2*(g_i-mean)/(N-1) + 0.3*mean/N )
  const float c0 = (float)(5.0 / root * (double)grad_loss[0]);
  const float ca = (float)(2.0 / (N - 1.0));
  const float bconst = (float)(0.3 * mean / N);
  const float fmean = (float)mean;
  const int b = blockIdx.y;
  const int HW = H * W;
  const float* tg = target + (long long)b * HW;
  const unsigned char* mk = MASK == 1 ? mask + (long long)b * HW : nullptr;
  const float* pb = pred + (long long)b * h * w;
  float* gb = grad_pred + (long long)b * h * w;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < HW; t += gridDim.x * blockDim.x) {
    const float tv = tg[t];
    if (MASK == 1 && !mk[t]) continue;
    if (MASK == 2 && !(tv > mask_thr)) continue;
    if (INTERP) {
      const int y = t / W, x = t - y * W;
      int y0, y1, x0, x1;
      float ly0, ly1, lx0, lx1;
      src_index(y, sy, h, y0, y1, ly0, ly1);
      src_index(x, sx, w, x0, x1, lx0, lx1);
      const float v = ly0 * (lx0 * __ldg(pb + y0 * w + x0) + lx1 * __ldg(pb + y0 * w + x1)) +
                      ly1 * (lx0 * __ldg(pb + y1 * w + x0) + lx1 * __ldg(pb + y1 * w + x1));
      const float g = logf(v) - logf(tv);
      const float gv = c0 * (ca * (g - fmean) + bconst) / v;
      atomicAdd(gb + y0 * w + x0, gv * ly0 * lx0);
      atomicAdd(gb + y0 * w + x1, gv * ly0 * lx1);
      atomicAdd(gb + y1 * w + x0, gv * ly1 * lx0);
      atomicAdd(gb + y1 * w + x1, gv * ly1 * lx1);
    } else {
      const float v = pb[t];
      const float g = logf(v) - logf(tv);
      gb[t] = c0 * (ca * (g - fmean) + bconst) / v;
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

namespace {

inline float scale_of(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }
constexpr size_t SILOG_PART_OFF = 64;  // SilogWs, then the per-block partials

size_t chamfer_smem(int n, bool grad) {
  const size_t nsk = (size_t)n + (n >> 5) + 1;
  return sizeof(float) * nsk + sizeof(unsigned int) * 2 * LS_WARPS * (size_t)(n + 1) +
         (grad ? (sizeof(float) + sizeof(unsigned int)) * LS_WARPS * (size_t)n : 0);
}

// one launcher for the three public forms
int launch_losses(bool silog, bool chamfer, const float* pred, const float* target, const uint8_t* mask, int mask_mode,
                  float mask_thr, int B, int h, int w, int H, int W, int interpolate, void* silog_ws, float* silog_loss,
                  const float* edges, int n_bins, float min_target, int want_grad, void* chamfer_ws, float* chamfer_loss,
                  cudaStream_t st) {
  const long long HW = (long long)H * W;
  if (HW > 0x7fffffffLL || B > 65535) return MDE_ERR_BAD_SHAPE;
  const int bx = loss_blocks_per_image(B, HW);
  LossArgs a = {};
  a.target = target;
  a.B = B;
  a.HW = (int)HW;
  if (silog) {
    a.pred = pred; a.mask = mask; a.mask_thr = mask_thr;
    a.h = h; a.w = w; a.H = H; a.W = W;
    a.sy = scale_of(h, H); a.sx = scale_of(w, W);
    a.silog_ws = reinterpret_cast<SilogWs*>(silog_ws);
    a.silog_part = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(silog_ws) + SILOG_PART_OFF);
    a.silog_loss = silog_loss;
    cudaMemsetAsync(silog_ws, 0, sizeof(SilogWs), st);
  }
  size_t sm = 0;
  if (chamfer) {
    a.edges = edges; a.n = n_bins; a.min_target = min_target; a.chamfer_loss = chamfer_loss;
    a.search_step = 1;
    while (a.search_step * 2 <= n_bins) a.search_step *= 2;
    const size_t zero = chamfer_layout(chamfer_ws, B, n_bins, bx, &a.cw);
    cudaMemsetAsync(chamfer_ws, 0, zero, st);
    sm = chamfer_smem(n_bins, want_grad != 0);
  }
  const dim3 grid((unsigned)bx, (unsigned)B);
#define MDE_LS(S, I, M, C, G)                                                                                          \
  {                                                                                                                    \
    if (sm > 48 * 1024) {                                                                                              \
      if (cudaFuncSetAttribute(depth_losses_kernel<S, I, M, C, G>, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                               (int)sm) != cudaSuccess)                                                                \
        return MDE_ERR_LAUNCH;                                                                                         \
    }                                                                                                                  \
    depth_losses_kernel<S, I, M, C, G><<<grid, LS_THREADS, sm, st>>>(a);                                               \
  }
  const bool g = want_grad != 0, ip = interpolate != 0;
  if (silog && chamfer) {  // fused: derived masks only
    if (ip) { if (g) MDE_LS(true, true, 2, true, true) else MDE_LS(true, true, 2, true, false) }
    else { if (g) MDE_LS(true, false, 2, true, true) else MDE_LS(true, false, 2, true, false) }
  } else if (silog) {
    if (ip) {
      if (mask_mode == 0) MDE_LS(true, true, 0, false, false) else if (mask_mode == 1) MDE_LS(true, true, 1, false, false)
      else MDE_LS(true, true, 2, false, false)
    } else {
      if (mask_mode == 0) MDE_LS(true, false, 0, false, false) else if (mask_mode == 1) MDE_LS(true, false, 1, false, false)
      else MDE_LS(true, false, 2, false, false)
    }
  } else {
    if (g) MDE_LS(false, false, 0, true, true) else MDE_LS(false, false, 0, true, false)
  }
#undef MDE_LS
  return check_launch();
}

int launch_silog_bwd(const float* pred, const float* target, const uint8_t* mask, int mask_mode, float mask_thr, int B, int h,
                     int w, int H, int W, int interpolate, const void* ws, const float* grad_loss, float* grad_pred,
                     cudaStream_t st) {
  if ((long long)H * W > 0x7fffffffLL || B > 65535) return MDE_ERR_BAD_SHAPE;
  cudaMemsetAsync(grad_pred, 0, sizeof(float) * (size_t)B * h * w, st);
  long long gx = ((long long)H * W + 256 * 8 - 1) / (256 * 8);
  const long long cap = (MDE_NUM_SMS * 8 + B - 1) / B;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  const dim3 grid((unsigned)gx, (unsigned)B);
  const float sy = scale_of(h, H), sx = scale_of(w, W);
  const SilogWs* W_ = reinterpret_cast<const SilogWs*>(ws);
#define MDE_SB(I, M) silog_bwd_kernel<I, M><<<grid, 256, 0, st>>>(pred, target, mask, mask_thr, h, w, H, W, sy, sx, W_, grad_loss, grad_pred)
  if (interpolate) {
    if (mask_mode == 0) MDE_SB(true, 0); else if (mask_mode == 1) MDE_SB(true, 1); else MDE_SB(true, 2);
  } else {
    if (mask_mode == 0) MDE_SB(false, 0); else if (mask_mode == 1) MDE_SB(false, 1); else MDE_SB(false, 2);
  }
#undef MDE_SB
  return check_launch();
}

}  // namespace

extern "C" {

int64_t mde_silog_ws_bytes(void) { return (int64_t)(SILOG_PART_OFF + sizeof(double) * 3 * LS_MAX_BLOCKS); }

int mde_silog_fwd(const float* pred, const float* target, const uint8_t* mask, int B, int h, int w, int H, int W,
                  int interpolate, void* ws, float* loss, mde_stream_t stream) {
  if (!pred || !target || !ws || !loss) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return MDE_ERR_BAD_SHAPE;
  if (!interpolate && (h != H || w != W)) return MDE_ERR_BAD_SHAPE;
  if (!aligned(ws, 8)) return MDE_ERR_BAD_POINTER;
  return launch_losses(true, false, pred, target, mask, mask ? 1 : 0, 0.f, B, h, w, H, W, interpolate, ws, loss, nullptr, 0,
                       0.f, 0, nullptr, nullptr, (cudaStream_t)stream);
}

int mde_silog_bwd(const float* pred, const float* target, const uint8_t* mask, int B, int h, int w, int H, int W,
                  int interpolate, const void* ws, const float* grad_loss, float* grad_pred, mde_stream_t stream) {
  if (!pred || !target || !ws || !grad_loss || !grad_pred) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return MDE_ERR_BAD_SHAPE;
  if (!interpolate && (h != H || w != W)) return MDE_ERR_BAD_SHAPE;
  return launch_silog_bwd(pred, target, mask, mask ? 1 : 0, 0.f, B, h, w, H, W, interpolate, ws, grad_loss, grad_pred,
                          (cudaStream_t)stream);
}

int64_t mde_chamfer_ws_bytes(int B, int n_bins) {
  if (B <= 0 || n_bins <= 0) return 0;
  const int bx = (MDE_NUM_SMS * 8 + B - 1) / B;  // upper bound of loss_blocks_per_image
  return (int64_t)chamfer_layout(nullptr, B, n_bins, bx, nullptr);
}

int mde_chamfer_fwd(const float* edges, const float* target, int B, int n_bins, int64_t HW, float min_target, int want_grad,
                    void* ws, float* loss, mde_stream_t stream) {
  if (!edges || !target || !ws || !loss) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || B > 65535 || n_bins <= 0 || n_bins > 2048 || HW <= 0 || HW > 0x7fffffffLL) return MDE_ERR_BAD_SHAPE;
  if (!aligned(ws, 16)) return MDE_ERR_BAD_POINTER;
  return launch_losses(false, true, nullptr, target, nullptr, 0, 0.f, B, 0, 0, 1, (int)HW, 0, nullptr, nullptr, edges, n_bins,
                       min_target, want_grad, ws, loss, (cudaStream_t)stream);
}

int mde_chamfer_bwd(const float* edges, int B, int n_bins, const void* ws, const float* grad_loss, float* grad_edges,
                    mde_stream_t stream) {
  if (!edges || !ws || !grad_loss || !grad_edges) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || n_bins <= 0 || n_bins > 2048) return MDE_ERR_BAD_SHAPE;
  ChamferWs w;
  chamfer_layout(const_cast<void*>(ws), B, n_bins, 1, &w);  // the fields the backward reads precede the per-block rows
  chamfer_bwd_kernel<<<B, 256, sizeof(float) * n_bins, (cudaStream_t)stream>>>(edges, B, n_bins, w, grad_loss, grad_edges);
  return check_launch();
}

// Fused form: SILog(pred, target, mask = target > silog_min_depth, interpolate) AND chamfer(edges, target >= chamfer_min)
// in ONE pass over the target (train.py:414-419: mask = depth > args.min_depth; criterion_ueff; criterion_bins).
int mde_depth_losses_fwd(const float* pred, const float* edges, const float* target, int B, int h, int w, int H, int W,
                         int n_bins, int interpolate, float silog_min_depth, float chamfer_min_target, int want_grad,
                         void* silog_ws, void* chamfer_ws, float* silog_loss, float* chamfer_loss, mde_stream_t stream) {
  if (!pred || !edges || !target || !silog_ws || !chamfer_ws || !silog_loss || !chamfer_loss) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0 || n_bins <= 0 || n_bins > 2048) return MDE_ERR_BAD_SHAPE;
  if (!interpolate && (h != H || w != W)) return MDE_ERR_BAD_SHAPE;
  if (!aligned(silog_ws, 8) || !aligned(chamfer_ws, 16)) return MDE_ERR_BAD_POINTER;
  return launch_losses(true, true, pred, target, nullptr, 2, silog_min_depth, B, h, w, H, W, interpolate, silog_ws, silog_loss,
                       edges, n_bins, chamfer_min_target, want_grad, chamfer_ws, chamfer_loss, (cudaStream_t)stream);
}

// SILog backward for the derived mask of the fused form (target > min_depth)
int mde_silog_bwd_thr(const float* pred, const float* target, float min_depth, int B, int h, int w, int H, int W,
                      int interpolate, const void* ws, const float* grad_loss, float* grad_pred, mde_stream_t stream) {
  if (!pred || !target || !ws || !grad_loss || !grad_pred) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return MDE_ERR_BAD_SHAPE;
  if (!interpolate && (h != H || w != W)) return MDE_ERR_BAD_SHAPE;
  return launch_silog_bwd(pred, target, nullptr, 2, min_depth, B, h, w, H, W, interpolate, ws, grad_loss, grad_pred,
                          (cudaStream_t)stream);
}

}  // extern "C"
