// The stem of the EfficientNet passthrough body in exact fp32: 3x3 / stride 2 convolution on the channels_last encoder input
// (C_in = 3 image planes + external-information channels, e.g. 28 at BASELINE config 2; C_out = 32 / 48), with the folded
// BatchNorm bias and SiLU in the epilogue (geffnet: conv_stem -> bn1 -> act1).
//
// Why a kernel of ours: the 1e-3 contract forbids TF32 here, and the library's exact-fp32 NHWC engine needs 0.83 ms for this one
// layer (14.6 GFLOP, 7 % of the config-2 step).  K = 9 * C_in = 252 is too short and C_out too narrow for the tensor-core
// pipeline to pay; plain fp32 FMAs do it in the time of ~2 passes over the 407 MB input:
//   * a thread computes TWO horizontally adjacent output pixels x all C_out channels (64 / 96 accumulators): every filter word is
//     read once per thread as part of a 16-byte shared-memory broadcast and used for 8 FMAs, which keeps the shared-memory pipe
//     (the co-bottleneck of a direct convolution) at half the FMA pipe's load;
//   * the filter lives in shared memory as [tap][c_in][C_out] (re-laid-out by the host wrapper, cached), the input taps are
//     16-byte loads along the channel axis served by L1 (neighbouring threads share input columns);
//   * zero padding (TensorFlow-SAME, asymmetric) is a clamped address and a zero mask -- no branches in the tap loop.
#include "common.cuh"

namespace mde {

constexpr int STEM_ROWS = 4;

template <int COUT>
__global__ void __launch_bounds__(256) stem_conv3x3s2_kernel(const float* __restrict__ x, const float* __restrict__ w_tcc,
                                                             const float* __restrict__ bias, float* __restrict__ y, int Hi,
                                                             int Wi, int Cin, int pad_t, int pad_l, int Ho, int Wo, int act) {
  extern __shared__ __align__(16) float sw[];  // [9][Cin][COUT]
  pdl_sync();
  for (int i = threadIdx.x; i < 9 * Cin * COUT; i += blockDim.x) sw[i] = w_tcc[i];
  __syncthreads();
  const int b = blockIdx.z;
  const int ox0 = (blockIdx.x * blockDim.x + threadIdx.x) * 2;  // this thread's two output columns: ox0, ox0 + 1
  if (ox0 >= Wo) return;
  const bool two = ox0 + 1 < Wo;
  const float* xb = x + (long long)b * Hi * Wi * Cin;
  // a block walks STEM_ROWS consecutive output rows: staging the 32 KB filter costs about as much as one row of outputs
  for (int oy = blockIdx.y * STEM_ROWS; oy < min(Ho, (int)(blockIdx.y + 1) * STEM_ROWS); ++oy) {
  float acc[2][COUT];
#pragma unroll
  for (int co = 0; co < COUT; ++co) {
    const float bv = bias ? __ldg(bias + co) : 0.f;
    acc[0][co] = bv;
    acc[1][co] = bv;
  }
  const int c4n = Cin >> 2;
#pragma unroll 1
  for (int dy = 0; dy < 3; ++dy) {
    const int iy = oy * 2 + dy - pad_t;
    const int iyc = min(max(iy, 0), Hi - 1);
    const float my = iy == iyc ? 1.f : 0.f;
    const float* row = xb + (long long)iyc * Wi * Cin;
#pragma unroll 1
    for (int dx = 0; dx < 3; ++dx) {
      const int ix0 = ox0 * 2 + dx - pad_l, ix1 = ix0 + 2;
      const int ix0c = min(max(ix0, 0), Wi - 1), ix1c = min(max(ix1, 0), Wi - 1);
      const float m0 = ix0 == ix0c ? my : 0.f, m1 = (ix1 == ix1c && two) ? my : 0.f;
      const float4* p0 = reinterpret_cast<const float4*>(row + (long long)ix0c * Cin);
      const float4* p1 = reinterpret_cast<const float4*>(row + (long long)ix1c * Cin);
      const float* wt = sw + (dy * 3 + dx) * Cin * COUT;
#pragma unroll 1
      for (int c = 0; c < c4n; ++c) {
        const float4 a = __ldg(p0 + c), q = __ldg(p1 + c);
        const float xa[4] = {a.x * m0, a.y * m0, a.z * m0, a.w * m0};
        const float xq[4] = {q.x * m1, q.y * m1, q.z * m1, q.w * m1};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4* wr = reinterpret_cast<const float4*>(wt + (4 * c + k) * COUT);
#pragma unroll
          for (int co = 0; co < COUT; co += 4) {
            const float4 wv = wr[co >> 2];  // all lanes read the same 16 bytes: one broadcast
            acc[0][co] = fmaf(xa[k], wv.x, acc[0][co]); acc[0][co + 1] = fmaf(xa[k], wv.y, acc[0][co + 1]);
            acc[0][co + 2] = fmaf(xa[k], wv.z, acc[0][co + 2]); acc[0][co + 3] = fmaf(xa[k], wv.w, acc[0][co + 3]);
            acc[1][co] = fmaf(xq[k], wv.x, acc[1][co]); acc[1][co + 1] = fmaf(xq[k], wv.y, acc[1][co + 1]);
            acc[1][co + 2] = fmaf(xq[k], wv.z, acc[1][co + 2]); acc[1][co + 3] = fmaf(xq[k], wv.w, acc[1][co + 3]);
          }
        }
      }
    }
  }
  float* yo = y + (((long long)b * Ho + oy) * Wo + ox0) * COUT;
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    if (p == 1 && !two) break;
#pragma unroll
    for (int co = 0; co < COUT; co += 4) {
      float o[4] = {acc[p][co], acc[p][co + 1], acc[p][co + 2], acc[p][co + 3]};
      if (act == 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = __fdividef(o[k], 1.f + __expf(-o[k]));
      }
      *reinterpret_cast<float4*>(yo + p * COUT + co) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
  }  // rows of this block
}

}  // namespace mde

// x [B][Hi][Wi][Cin] fp32 NHWC (Cin % 4 == 0); w_tcc [9][Cin][Cout] (tap = dy*3 + dx major, C_out innermost); bias [Cout] or NULL;
// y [B][Ho][Wo][Cout]; zero padding pad_top / pad_left (and what Ho / Wo imply at the far edges); Cout in {32, 48}; act 0 / 1 (SiLU).
extern "C" int mde_stem_conv3x3s2_nhwc(const float* x, const float* w_tcc, const float* bias, float* y, int B, int Hi, int Wi,
                                       int Cin, int Cout, int pad_top, int pad_left, int Ho, int Wo, int act,
                                       mde_stream_t stream) {
  using namespace mde;
  if (!x || !w_tcc || !y) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || B > 65535 || Hi <= 0 || Wi <= 0 || Cin <= 0 || Ho <= 0 || Ho > 65535 || Wo <= 0 || pad_top < 0 || pad_left < 0 ||
      act < 0 || act > 1)
    return MDE_ERR_BAD_SHAPE;
  if ((Cout != 32 && Cout != 48) || Cin % 4 != 0 || !aligned(x, 16) || !aligned(y, 16) || !aligned(w_tcc, 16))
    return MDE_ERR_UNSUPPORTED;
  if ((Ho - 1) * 2 - pad_top >= Hi || (Wo - 1) * 2 - pad_left >= Wi) return MDE_ERR_BAD_SHAPE;
  const size_t sm = (size_t)9 * Cin * Cout * sizeof(float);
  if (sm > 160 * 1024) return MDE_ERR_UNSUPPORTED;
  // one block per output row when the row's column pairs fit 256 threads (whole warps), else several
  const int pairs = (Wo + 1) / 2;
  const int threads = pairs >= 256 ? 256 : ((pairs + 31) / 32) * 32;
  const dim3 grid((unsigned)((pairs + threads - 1) / threads), (unsigned)((Ho + STEM_ROWS - 1) / STEM_ROWS), (unsigned)B);
  cudaStream_t st = (cudaStream_t)stream;
#define MDE_STEM(CO)                                                                                                       \
  {                                                                                                                        \
    static bool attr = false;                                                                                              \
    if (!attr) {                                                                                                           \
      if (cudaFuncSetAttribute(stem_conv3x3s2_kernel<CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) != cudaSuccess) \
        return MDE_ERR_LAUNCH;                                                                                             \
      attr = true;                                                                                                         \
    }                                                                                                                      \
    launch_pdl(PDL_STREAM, stem_conv3x3s2_kernel<CO>, grid, dim3(threads), sm, st, x, w_tcc, bias, y, Hi, Wi, Cin, pad_top, pad_left, Ho, Wo, \
               act);                                                                                                        \
  }
  if (Cout == 32) MDE_STEM(32) else MDE_STEM(48)
#undef MDE_STEM
  return check_launch();
}
