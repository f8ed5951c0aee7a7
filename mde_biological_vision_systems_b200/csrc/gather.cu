// K3 -- label -> embedding gather (shared-memory-staged table, planar NCHW stores), class-area table, casts.
//
// Reference semantics: ExternalInfoLoaders/SemanticsLoader.py:102-145 and
// ExternalInfoLoaders/InstanceSegmentationLoader.py:89-121 (clamp, index_select, permute to [B,D,H,W]).
// HBM-bound: per pixel 8 B of int64 label in, D*sizeof(T) out (108 B/px for the fp32 GloVe-25d case).
// Layout: one thread owns 4 consecutive pixels -> one 32 B label read and one 16 B store per channel plane,
// so a warp writes 512 contiguous bytes per plane (fully coalesced); the table sits in shared memory with an odd
// row pitch (D = 25 or 3) so distinct labels in a warp hit distinct banks.
#include <stdlib.h>

#include "common.cuh"

namespace mde {

unsigned long long g_launch_count = 0;

static int g_pdl = -1;  // mask of kernel classes (common.cuh); -1: not read yet
bool pdl_enabled(int cls) {
  if (g_pdl < 0) {
    const char* e = getenv("MDE_PDL");
    g_pdl = e != nullptr ? (atoi(e) & 7) : 0;
  }
  return (g_pdl & cls) != 0;
}

template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
  static __device__ __forceinline__ void store(float* p, float a, float b, float c, float d) {
    stg_stream(reinterpret_cast<float4*>(p), make_float4(a, b, c, d));
  }
};
template <>
struct Vec4<double> {
  static __device__ __forceinline__ void store(double* p, double a, double b, double c, double d) {
    reinterpret_cast<double2*>(p)[0] = make_double2(a, b);
    reinterpret_cast<double2*>(p)[1] = make_double2(c, d);
  }
};

__device__ __forceinline__ int clamp_label(long long l, int rows, int background, bool& oob) {
  if (l < 0 || l > (long long)(rows - 1)) {
    if (background >= 0) return background;
    oob = true;
    return -1;
  }
  return (int)l;
}

// four consecutive labels of the on-disk / wire label types: int64 (the reference's batch tensors), int32 (the .npz
// instance maps, dataloader.py:136-150) and uint8 (semantic maps after astype(np.ubyte), dataloader.py:121-133)
__device__ __forceinline__ void load4(const long long* p, long long (&v)[4]) {
  const longlong2 a = *reinterpret_cast<const longlong2*>(p), c = *reinterpret_cast<const longlong2*>(p + 2);
  v[0] = a.x; v[1] = a.y; v[2] = c.x; v[3] = c.y;
}
__device__ __forceinline__ void load4(const int* p, long long (&v)[4]) {
  const int4 a = *reinterpret_cast<const int4*>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void load4(const unsigned char* p, long long (&v)[4]) {
  const uchar4 a = *reinterpret_cast<const uchar4*>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}

// VEC = 4: HW % 4 == 0 and 16/32-byte aligned pointers; VEC = 1: generic scalar path.
template <typename T, typename L, int VEC, bool SMEM>
__global__ void __launch_bounds__(256) gather_embed_kernel(const L* __restrict__ labels,
                                                            long long* labels_out, const T* __restrict__ table,
                                                            T* __restrict__ out, long long HW, int rows, int D,
                                                            int background, long long table_image_stride,
                                                            int* oob_flag) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* stab = reinterpret_cast<T*>(smem_raw);
  const int b = blockIdx.y;
  const T* tab = table + (long long)b * table_image_stride;
  if (SMEM) {
    for (int i = threadIdx.x; i < rows * D; i += blockDim.x) stab[i] = tab[i];
    __syncthreads();
    tab = stab;
  }
  const L* lab = labels + (long long)b * HW;
  long long* lab_out = labels_out ? labels_out + (long long)b * HW : nullptr;
  T* o = out + (long long)b * D * HW;
  const long long groups = HW / VEC;
  bool oob = false;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups;
       g += (long long)gridDim.x * blockDim.x) {
    const long long p = g * VEC;
    if (VEC == 4) {
      long long v[4];
      load4(lab + p, v);
      const int l0 = clamp_label(v[0], rows, background, oob), l1 = clamp_label(v[1], rows, background, oob);
      const int l2 = clamp_label(v[2], rows, background, oob), l3 = clamp_label(v[3], rows, background, oob);
      if (lab_out) {
        *reinterpret_cast<longlong2*>(lab_out + p) = make_longlong2(l0 < 0 ? v[0] : l0, l1 < 0 ? v[1] : l1);
        *reinterpret_cast<longlong2*>(lab_out + p + 2) = make_longlong2(l2 < 0 ? v[2] : l2, l3 < 0 ? v[3] : l3);
      }
      const T* r0 = tab + (l0 < 0 ? 0 : l0) * D;
      const T* r1 = tab + (l1 < 0 ? 0 : l1) * D;
      const T* r2 = tab + (l2 < 0 ? 0 : l2) * D;
      const T* r3 = tab + (l3 < 0 ? 0 : l3) * D;
      const T z = T(0);
#pragma unroll 5
      for (int d = 0; d < D; ++d) {
        Vec4<T>::store(o + (long long)d * HW + p, l0 < 0 ? z : r0[d], l1 < 0 ? z : r1[d], l2 < 0 ? z : r2[d],
                       l3 < 0 ? z : r3[d]);
      }
    } else {
      const long long raw = (long long)lab[p];
      const int l = clamp_label(raw, rows, background, oob);
      if (lab_out) lab_out[p] = l < 0 ? raw : l;
      const T* r = tab + (l < 0 ? 0 : l) * D;
      for (int d = 0; d < D; ++d) o[(long long)d * HW + p] = l < 0 ? T(0) : r[d];
    }
  }
  if (oob && oob_flag) atomicExch(oob_flag, 1);
}

template <typename T, typename L>
static int launch_gather(const void* labels, int64_t* labels_out, const void* table, void* out, int B, int64_t HW,
                         int rows, int D, int background, int64_t tis, int32_t* oob_flag, cudaStream_t st) {
  const size_t tab_bytes = (size_t)rows * D * sizeof(T);
  const bool smem = tab_bytes <= 48 * 1024;
  const bool vec = (HW % 4 == 0) && aligned(labels, 4 * sizeof(L) > 16 ? 16 : 4 * sizeof(L)) && aligned(out, 16) &&
                   (!labels_out || aligned(labels_out, 16));
  const long long groups = vec ? HW / 4 : HW;
  long long gx = (groups + 256 * 4 - 1) / (256 * 4);
  if (gx < 1) gx = 1;
  if (gx > 65535) gx = 65535;
  dim3 grid((unsigned)gx, (unsigned)B);
  const size_t sm = smem ? tab_bytes : 0;
  const L* Lp = reinterpret_cast<const L*>(labels);
  long long* LO = reinterpret_cast<long long*>(labels_out);
  const T* Tb = reinterpret_cast<const T*>(table);
  T* O = reinterpret_cast<T*>(out);
#define MDE_GATHER_LAUNCH(V, S) \
  gather_embed_kernel<T, L, V, S><<<grid, 256, sm, st>>>(Lp, LO, Tb, O, HW, rows, D, background, tis, oob_flag)
  if (vec && smem) MDE_GATHER_LAUNCH(4, true);
  else if (vec) MDE_GATHER_LAUNCH(4, false);
  else if (smem) MDE_GATHER_LAUNCH(1, true);
  else MDE_GATHER_LAUNCH(1, false);
#undef MDE_GATHER_LAUNCH
  return check_launch();
}

// The same gather written straight into a channel slice of a (spatially padded) channels_last tensor -- the encoder input
// of the input-insertion configs: out[b, y + pad_top, x + pad_left, c0 + d] = table[clamp(label[b, y, x])][d].  One thread per
// (pixel, d): a warp's 32 stores are consecutive addresses (25 floats of a pixel, then the next pixel's after the c0-wide
// gap); labels are read once per pixel through L1.  Removes the planar [B,D,H,W] intermediate and its transpose
// (2 x 22.6 MB/img at config 2).  fp32 tables in shared memory, clamping mode only.
// VEC: pitch % 4 == 0: a pixel's [c0, c0+D) channels are `head` leading scalars up to the next 16-byte boundary, then
// float4 groups, then `tail` scalars.  Thread layout (8, 32): threadIdx.x = piece (up to 8 pieces per pixel), threadIdx.y =
// pixel of a 32-pixel run of one image row -- a warp covers 4 consecutive pixels, i.e. one contiguous span of the output;
// no integer division per element.  With ``image`` (NCHW fp32 [B][c0][H][W], 1 <= c0 <= 3) the leading channels [0, c0) are
// written by the same kernel (piece 0 becomes the float4 {img..., emb_0...}), so the whole encoder input row comes from
// one pass.
template <typename L, bool VEC>
__global__ void __launch_bounds__(256) gather_embed_nhwc_kernel(const L* __restrict__ labels, long long* labels_out,
                                                                 const float* __restrict__ table,
                                                                 const float* __restrict__ image, float* __restrict__ out,
                                                                 int H, int W, int rows, int D, int background, int pitch,
                                                                 int c0, int Ho, int Wo, int pad_top, int pad_left) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* stab = reinterpret_cast<float*>(smem_raw);
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  // encoder-input case (image planes + embedding, whole float4 pieces): the table is staged as OUTPUT ROWS, 8 float4 pieces
  // per label with the embedding shifted by c0 channels, so that a thread's piece is one 16-byte shared-memory read
  const bool rowform = VEC && image != nullptr && ((D - (4 - c0)) & 3) == 0;
  if (rowform) {
    for (int i = tid; i < rows * 32; i += blockDim.x * blockDim.y) {
      const int e = (i & 31) - c0;
      stab[i] = (e >= 0 && e < D) ? table[(i >> 5) * D + e] : 0.f;
    }
  } else {
    for (int i = tid; i < rows * D; i += blockDim.x * blockDim.y) stab[i] = table[i];
  }
  __syncthreads();
  const int b = blockIdx.z;
  const int HW = H * W;
  const L* lab = labels + (long long)b * HW;
  long long* lab_out = labels_out ? labels_out + (long long)b * HW : nullptr;
  float* o = out + (long long)b * Ho * Wo * pitch;
  bool oob = false;
  if (VEC) {
    const int x = blockIdx.x * 32 + threadIdx.y;
    if (x >= W) return;
    const int q = threadIdx.x;
    const bool fused = image != nullptr;           // then c0 <= 3 leading image channels complete the first float4
    const int head = fused ? 4 - c0 : ((4 - (c0 & 3)) & 3);   // embedding scalars before the first aligned float4
    const int nvec = (D - head) >> 2;
    const int tail = D - head - 4 * nvec;
    const int first = (fused || head) ? 1 : 0;
    constexpr int U = 4;  // rows in flight per thread: the label (and image) loads of U rows are issued before any is used
    // a block walks a strip of rows (the table staging is amortised over H / gridDim.y rows)
    for (int yb = blockIdx.y; yb < H; yb += U * gridDim.y) {
      long long raw[U];
      float img[U][3];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int y = yb + u * gridDim.y;
        raw[u] = y < H ? (long long)lab[y * W + x] : 0;
        img[u][0] = img[u][1] = img[u][2] = 0.f;
        if (fused && q == 0 && y < H) {
          const float* ip = image + (long long)b * c0 * HW + y * W + x;
#pragma unroll
          for (int c = 0; c < 3; ++c) img[u][c] = c < c0 ? __ldg(ip + (long long)c * HW) : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int y = yb + u * gridDim.y;
        if (y >= H) break;
        const int p = y * W + x;
        const int l = clamp_label(raw[u], rows, background, oob);
        if (q == 0 && lab_out) lab_out[p] = l;
        float* dst = o + ((long long)(y + pad_top) * Wo + (x + pad_left)) * pitch;
        const float* row = stab + l * D;
        if (rowform) {
          // uniform (divergence-free) path of the encoder-input case: piece q covers channels [4q, 4q + 4) of the pixel's row;
          // channel c < c0 is an image plane (piece 0 only), channel c >= c0 is embedding element c - c0
          if (q < 1 + nvec) {
            float4 v = *reinterpret_cast<const float4*>(stab + l * 32 + 4 * q);
            if (q == 0) {
              v.x = img[u][0];
              if (c0 > 1) v.y = img[u][1];
              if (c0 > 2) v.z = img[u][2];
            }
            *reinterpret_cast<float4*>(dst + 4 * q) = v;
          }
        } else if (q == 0 && first) {
          if (fused) {
            float v[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) v[c] = c < c0 ? img[u][c < 3 ? c : 2] : row[c - c0];
            *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
          } else {
            for (int d = 0; d < head; ++d) dst[c0 + d] = row[d];
          }
        } else if (q - first < nvec) {
          const int d = head + 4 * (q - first);
          *reinterpret_cast<float4*>(dst + c0 + d) = make_float4(row[d], row[d + 1], row[d + 2], row[d + 3]);
        } else if (q - first == nvec && tail) {
          for (int d = head + 4 * nvec; d < D; ++d) dst[c0 + d] = row[d];
        }
      }
    }
  } else {
    const int total = HW * D;
    for (int i = blockIdx.x * 256 + tid; i < total; i += gridDim.x * 256) {
      const int p = i / D, d = i - p * D;
      const long long raw = (long long)lab[p];
      const int l = clamp_label(raw, rows, background, oob);
      if (d == 0 && lab_out) lab_out[p] = l;
      const int y = p / W, x = p - y * W;
      o[((long long)(y + pad_top) * Wo + (x + pad_left)) * pitch + c0 + d] = stab[l * D + d];
    }
  }
}

// Encoder-input case with a DENSE pixel record (pitch == c0 + D, a whole number NP <= 8 of float4 pieces; config 2: 3 image
// planes + 25 GloVe components = 28 floats): one WARP per run of 32 consecutive pixels of an image row.
//   lane = pixel: label (coalesced 256 B), clamp, labels_out (coalesced), the c0 image planes (coalesced 128 B each), then
//   the pixel's NP pieces go table row -> registers -> the warp's staging tile in shared memory (the 16 * NP byte record
//   stride makes these 16-byte stores bank-conflict free for NP = 7);
//   lane = float4 of the tile: the 32 records are one contiguous 512 * NP byte span of the output, written with NP fully
//   coalesced 16-byte store instructions.
// ~3 warp instructions per pixel instead of ~30 with one thread per (pixel, piece).
template <typename L, int NP>
__global__ void __launch_bounds__(256) gather_embed_dense_kernel(const L* __restrict__ labels, long long* labels_out,
                                                                  const float* __restrict__ table,
                                                                  const float* __restrict__ image, float* __restrict__ out,
                                                                  int H, int W, int rows, int D, int background, int c0,
                                                                  int Ho, int Wo, int pad_top, int pad_left, int segs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* stab = reinterpret_cast<float4*>(smem_raw);                 // [rows][NP] output-row form of the table
  float4* stage_all = stab + rows * NP;                               // [8 warps][32 px][NP]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < rows * NP * 4; i += 256) {
    const int e = (i % (NP * 4)) - c0;
    reinterpret_cast<float*>(stab)[i] = (e >= 0 && e < D) ? table[(i / (NP * 4)) * D + e] : 0.f;
  }
  __syncthreads();
  const int b = blockIdx.y;
  const int HW = H * W;
  const L* lab = labels + (long long)b * HW;
  long long* lab_out = labels_out ? labels_out + (long long)b * HW : nullptr;
  const float* img = image + (long long)b * c0 * HW;
  float* o = out + (long long)b * Ho * Wo * (NP * 4);
  float4* stage = stage_all + warp * 32 * NP;
  bool oob = false;
  const int total = H * segs;  // 32-pixel segments of the image
  for (int sgi = blockIdx.x * 8 + warp; sgi < total; sgi += gridDim.x * 8) {
    const int y = sgi / segs, x0 = (sgi - y * segs) * 32;
    const int npx = min(32, W - x0);
    const int p = y * W + x0 + lane;
    int l = 0;
    float im0 = 0.f, im1 = 0.f, im2 = 0.f;
    if (lane < npx) {
      const long long raw = (long long)lab[p];
      im0 = __ldg(img + p);
      if (c0 > 1) im1 = __ldg(img + HW + p);
      if (c0 > 2) im2 = __ldg(img + 2 * HW + p);
      l = clamp_label(raw, rows, background, oob);
      if (lab_out) lab_out[p] = l;
    }
    const float4* row = stab + l * NP;
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      float4 v = row[q];
      if (q == 0) {
        v.x = im0;
        if (c0 > 1) v.y = im1;
        if (c0 > 2) v.z = im2;
      }
      stage[lane * NP + q] = v;
    }
    __syncwarp();
    float4* dst = reinterpret_cast<float4*>(o + ((long long)(y + pad_top) * Wo + (x0 + pad_left)) * (NP * 4));
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const int i = k * 32 + lane;
      if (i < npx * NP) dst[i] = stage[i];
    }
    __syncwarp();
  }
}

// ---- per-image class histogram -> area fraction table -------------------------------------------------------
__global__ void __launch_bounds__(256) class_hist_kernel(const long long* __restrict__ labels, long long HW, int rows,
                                                          int* __restrict__ counts) {
  extern __shared__ int shist[];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) shist[i] = 0;
  __syncthreads();
  const long long* lab = labels + (long long)b * HW;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
    const long long l = lab[p];
    if (l >= 0 && l < rows) atomicAdd(&shist[(int)l], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < rows; i += blockDim.x)
    if (shist[i]) atomicAdd(&counts[b * rows + i], shist[i]);
}

__global__ void class_frac_kernel(const int* __restrict__ counts, double* __restrict__ frac, int n, long long HW) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  // python: label_area = count / total_area  (int / int true division -> float64)
  if (i < n) frac[i] = (double)counts[i] / (double)HW;
}

__global__ void cast_i64_f32_kernel(const long long* __restrict__ in, float* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = (float)in[i];  // round-to-nearest-even, as Tensor.float()
}

__global__ void relu_eps_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, float eps) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = fmaxf(x[i], 0.f) + eps;
}

}  // namespace mde

using namespace mde;

extern "C" {

int mde_version(void) { return 100; }

int64_t mde_launch_count(void) { return (int64_t)g_launch_count; }

// Programmatic dependent launch of the kernels that support it (common.cuh): a mask of kernel classes (1 chains of small
// kernels, 2 persistent tcgen05 kernels, 4 streaming kernels), 0 off, < 0 query only.  Returns the setting
// in effect (default: the MDE_PDL environment variable, else off).  Host-side state of the process.
int mde_set_pdl(int mask) {
  (void)mde::pdl_enabled(0);
  if (mask >= 0) mde::g_pdl = mask & 7;
  return mde::g_pdl;
}

const char* mde_error_string(int code) {
  switch (code) {
    case MDE_OK: return "ok";
    case MDE_ERR_BAD_SHAPE: return "bad shape / stride / divisibility";
    case MDE_ERR_BAD_POINTER: return "null or misaligned pointer";
    case MDE_ERR_BAD_ARCH: return "device is not sm_100 (B200)";
    case MDE_ERR_LAUNCH: return "kernel launch failed";
    case MDE_ERR_UNSUPPORTED: return "unsupported mode";
    case MDE_ERR_DRIVER: return "CUDA driver entry point unavailable";
  }
  return "unknown error";
}

int mde_check_device(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return MDE_ERR_BAD_ARCH;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return MDE_ERR_BAD_ARCH;
  return major == 10 ? MDE_OK : MDE_ERR_BAD_ARCH;
}

int mde_gather_embed(const int64_t* labels, int64_t* labels_out, const void* table, void* out, int B, int64_t HW,
                     int rows, int D, int background, int out_dtype, int64_t table_image_stride, int32_t* oob_flag,
                     mde_stream_t stream) {
  if (!labels || !table || !out) return MDE_ERR_BAD_POINTER;
  if (B < 0 || HW < 0 || rows <= 0 || D <= 0 || background >= rows || B > 65535) return MDE_ERR_BAD_SHAPE;
  if (B == 0 || HW == 0) return MDE_OK;
  return mde_gather_embed_labels(labels, MDE_I64, labels_out, table, out, B, HW, rows, D, background, out_dtype,
                                 table_image_stride, oob_flag, stream);
}

int mde_gather_embed_labels(const void* labels, int label_dtype, int64_t* labels_out, const void* table, void* out, int B,
                            int64_t HW, int rows, int D, int background, int out_dtype, int64_t table_image_stride,
                            int32_t* oob_flag, mde_stream_t stream) {
  if (!labels || !table || !out) return MDE_ERR_BAD_POINTER;
  if (B < 0 || HW < 0 || rows <= 0 || D <= 0 || background >= rows || B > 65535) return MDE_ERR_BAD_SHAPE;
  if (B == 0 || HW == 0) return MDE_OK;
  cudaStream_t st = (cudaStream_t)stream;
#define MDE_GATHER_DISPATCH(T)                                                                                          \
  switch (label_dtype) {                                                                                                \
    case MDE_I64: return launch_gather<T, long long>(labels, labels_out, table, out, B, HW, rows, D, background,       \
                                                     table_image_stride, oob_flag, st);                                \
    case MDE_I32: return launch_gather<T, int>(labels, labels_out, table, out, B, HW, rows, D, background,             \
                                               table_image_stride, oob_flag, st);                                      \
    case MDE_U8: return launch_gather<T, unsigned char>(labels, labels_out, table, out, B, HW, rows, D, background,    \
                                                        table_image_stride, oob_flag, st);                             \
    default: return MDE_ERR_UNSUPPORTED;                                                                                \
  }
  if (out_dtype == MDE_F32) MDE_GATHER_DISPATCH(float)
  if (out_dtype == MDE_F64) MDE_GATHER_DISPATCH(double)
#undef MDE_GATHER_DISPATCH
  return MDE_ERR_UNSUPPORTED;
}

int mde_class_area_table(const int64_t* labels, int B, int64_t HW, int rows, int32_t* counts, double* frac,
                         mde_stream_t stream) {
  if (!labels || !counts || !frac) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || HW <= 0 || rows <= 0 || rows > 8192 || B > 65535) return MDE_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)B * rows, st);
  long long gx = (HW + 256 * 16 - 1) / (256 * 16);
  if (gx > 1024) gx = 1024;
  class_hist_kernel<<<dim3((unsigned)gx, (unsigned)B), 256, rows * sizeof(int), st>>>(
      reinterpret_cast<const long long*>(labels), HW, rows, counts);
  int rc = check_launch();
  if (rc) return rc;
  class_frac_kernel<<<(B * rows + 255) / 256, 256, 0, st>>>(counts, frac, B * rows, HW);
  return check_launch();
}

int mde_cast_i64_f32(const int64_t* in, float* out, int64_t n, mde_stream_t stream) {
  if (!in || !out) return MDE_ERR_BAD_POINTER;
  if (n <= 0) return n == 0 ? MDE_OK : MDE_ERR_BAD_SHAPE;
  long long gx = (n + 255) / 256;
  if (gx > MDE_NUM_SMS * 16) gx = MDE_NUM_SMS * 16;
  cast_i64_f32_kernel<<<(unsigned)gx, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const long long*>(in), out, n);
  return check_launch();
}

int mde_relu_eps_fwd(const float* x, float* y, int64_t n, float eps, mde_stream_t stream) {
  if (!x || !y) return MDE_ERR_BAD_POINTER;
  if (n <= 0) return n == 0 ? MDE_OK : MDE_ERR_BAD_SHAPE;
  long long gx = (n + 255) / 256;
  if (gx > MDE_NUM_SMS * 16) gx = MDE_NUM_SMS * 16;
  relu_eps_kernel<<<(unsigned)gx, 256, 0, (cudaStream_t)stream>>>(x, y, n, eps);
  return check_launch();
}

int mde_gather_embed_nhwc(const void* labels, int label_dtype, int64_t* labels_out, const float* table, const float* image_nchw,
                          float* out_nhwc, int B, int H, int W, int rows, int D, int background, int pitch, int c0, int Ho,
                          int Wo, int pad_top, int pad_left, mde_stream_t stream) {
  if (!labels || !table || !out_nhwc) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || B > 65535 || H <= 0 || H > 65535 || W <= 0 || rows <= 0 || D <= 0 || c0 < 0 || c0 + D > pitch || pad_top < 0 ||
      pad_left < 0 || Ho < H + pad_top || Wo < W + pad_left || (long long)H * W * D > 0x7fffffffLL)
    return MDE_ERR_BAD_SHAPE;
  if (background < 0 || background >= rows) return MDE_ERR_UNSUPPORTED;  // clamping mode only
  const bool fused = image_nchw != nullptr;
  const size_t sm = (size_t)rows * (fused && D <= 32 - c0 ? 32 : D) * sizeof(float);  // fused: staged as 32-float output rows
  if (sm > 48 * 1024) return MDE_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  if (fused && !(c0 >= 1 && c0 <= 3 && D >= 4 - c0)) return MDE_ERR_UNSUPPORTED;
  const int head = fused ? 4 - c0 : ((4 - (c0 & 3)) & 3);
  const int pieces = ((fused || head) ? 1 : 0) + (D - head) / 4 + (((D - head) % 4) ? 1 : 0);
  const bool vec = (pitch % 4 == 0) && aligned(out_nhwc, 16) && D >= head && pieces <= 8;
  if (fused && !vec) return MDE_ERR_UNSUPPORTED;
  if (fused && pitch == c0 + D && (c0 + D) % 4 == 0 && (pitch == 28 || pitch == 32)) {
    // dense pixel records: the warp-staged kernel
    const int np = pitch / 4, segs = (W + 31) / 32;
    const size_t smd = ((size_t)rows * np + 8 * 32 * np) * sizeof(float4);
    if (smd <= 96 * 1024) {
      long long gx = ((long long)H * segs + 7) / 8;
      // one resident wave: a block holds the table (rows * np float4) + 8 staging tiles, i.e. ~40 KB -> 5 blocks per SM
      const int per_sm = (int)((200 * 1024) / smd) < 8 ? (int)((200 * 1024) / smd) : 8;
      long long cap = ((long long)MDE_NUM_SMS * (per_sm < 1 ? 1 : per_sm)) / B;
      if (cap < 1) cap = 1;
      if (gx > cap) gx = cap;
      const dim3 gd((unsigned)gx, (unsigned)B);
#define MDE_GD(LT, NPV)                                                                                                  \
  {                                                                                                                      \
    static bool attr = false;                                                                                            \
    if (!attr) {                                                                                                         \
      if (cudaFuncSetAttribute(gather_embed_dense_kernel<LT, NPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) != \
          cudaSuccess)                                                                                                   \
        return MDE_ERR_LAUNCH;                                                                                           \
      attr = true;                                                                                                       \
    }                                                                                                                    \
    gather_embed_dense_kernel<LT, NPV><<<gd, 256, smd, st>>>(reinterpret_cast<const LT*>(labels),                        \
        reinterpret_cast<long long*>(labels_out), table, image_nchw, out_nhwc, H, W, rows, D, background, c0, Ho, Wo,     \
        pad_top, pad_left, segs);                                                                                        \
  }
#define MDE_GD2(LT) { if (np == 7) MDE_GD(LT, 7) else MDE_GD(LT, 8) }
      if (label_dtype == MDE_I64) MDE_GD2(long long)
      else if (label_dtype == MDE_I32) MDE_GD2(int)
      else if (label_dtype == MDE_U8) MDE_GD2(unsigned char)
      else return MDE_ERR_UNSUPPORTED;
#undef MDE_GD2
#undef MDE_GD
      return check_launch();
    }
  }
  dim3 grid, block;
  if (vec) {
    block = dim3(8, 32);
    const int gx = (W + 31) / 32;
    int gy = (MDE_NUM_SMS * 16 + gx * B - 1) / (gx * B);  // ~16 blocks per SM over the whole batch, each walking H / gy rows
    if (gy > H) gy = H;
    if (gy < 1) gy = 1;
    grid = dim3((unsigned)gx, (unsigned)gy, (unsigned)B);
  } else {
    long long gx = ((long long)H * W * D + 256 * 8 - 1) / (256 * 8);
    if (gx > MDE_NUM_SMS * 16) gx = MDE_NUM_SMS * 16;
    block = dim3(256, 1);
    grid = dim3((unsigned)gx, 1, (unsigned)B);
  }
#define MDE_GN(LT)                                                                                                     \
  {                                                                                                                    \
    if (vec)                                                                                                           \
      gather_embed_nhwc_kernel<LT, true><<<grid, block, sm, st>>>(reinterpret_cast<const LT*>(labels),                 \
          reinterpret_cast<long long*>(labels_out), table, image_nchw, out_nhwc, H, W, rows, D, background, pitch, c0, Ho, Wo, pad_top, pad_left); \
    else                                                                                                               \
      gather_embed_nhwc_kernel<LT, false><<<grid, block, sm, st>>>(reinterpret_cast<const LT*>(labels),                \
          reinterpret_cast<long long*>(labels_out), table, image_nchw, out_nhwc, H, W, rows, D, background, pitch, c0, Ho, Wo, pad_top, pad_left); \
  }
  if (label_dtype == MDE_I64) MDE_GN(long long)
  else if (label_dtype == MDE_I32) MDE_GN(int)
  else if (label_dtype == MDE_U8) MDE_GN(unsigned char)
  else return MDE_ERR_UNSUPPORTED;
#undef MDE_GN
  return check_launch();
}

}  // extern "C"
