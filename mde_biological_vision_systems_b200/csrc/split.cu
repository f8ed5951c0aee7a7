// Split-bf16 ("bf16x3") operand format: conversions between fp32 tensors and bf16 (hi, mid) plane pairs.
//
// The tensor-core kernels of the head (conv3x3, patch embedding, fused chain) multiply fp32 values as three bf16 products
// hi*hi + mid*hi + hi*mid (tc_common.cuh).  Their producers write the pairs directly; these streaming kernels serve the
// boundaries where a plain fp32 tensor enters (features from a library kernel, stand-alone calls, per-image weights) or
// leaves (a consumer that wants fp32).  All are single-pass HBM streams: 4 B/element in, 4 B/element out.
//   pair layout: uint16 planes[2][n]  (plane 0 = bf16_rn(v), plane 1 = bf16_rn(v - plane 0)), same element order as the
//   fp32 tensor -- or NHWC element order when the source is NCHW (mde_split_bf16_nchw).
#include "common.cuh"
#include "tc_common.cuh"

namespace mde {

__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ x, uint16_t* __restrict__ planes,
                                                         long long n8, long long n) {
  // 8 elements per thread: two 16-byte loads, one 16-byte store per plane
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = ldg_stream(reinterpret_cast<const float4*>(x) + 2 * i);
    const float4 b = ldg_stream(reinterpret_cast<const float4*>(x) + 2 * i + 1);
    uint4 hi, mid;
    tc::split_bf16x2(a.x, a.y, hi.x, mid.x);
    tc::split_bf16x2(a.z, a.w, hi.y, mid.y);
    tc::split_bf16x2(b.x, b.y, hi.z, mid.z);
    tc::split_bf16x2(b.z, b.w, hi.w, mid.w);
    reinterpret_cast<uint4*>(planes)[i] = hi;
    reinterpret_cast<uint4*>(planes + n)[i] = mid;
  }
}

__global__ void __launch_bounds__(256) merge_bf16_kernel(const uint16_t* __restrict__ planes, float* __restrict__ out,
                                                         long long n8, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const uint4 hi = reinterpret_cast<const uint4*>(planes)[i];
    const uint4 mid = reinterpret_cast<const uint4*>(planes + n)[i];
    const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w}, m[4] = {mid.x, mid.y, mid.z, mid.w};
    float o[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      o[2 * k] = __uint_as_float(h[k] << 16) + __uint_as_float(m[k] << 16);
      o[2 * k + 1] = __uint_as_float(h[k] & 0xFFFF0000u) + __uint_as_float(m[k] & 0xFFFF0000u);
    }
    reinterpret_cast<float4*>(out)[2 * i] = make_float4(o[0], o[1], o[2], o[3]);
    reinterpret_cast<float4*>(out)[2 * i + 1] = make_float4(o[4], o[5], o[6], o[7]);
  }
}

// fp32 NCHW [B][C][P] -> pair NHWC planes[2][B][P][C]: 64 x 64 tile transpose through shared memory
__global__ void __launch_bounds__(256) split_bf16_nchw_kernel(const float* __restrict__ in, uint16_t* __restrict__ planes,
                                                              int C, long long P, long long plane_elems) {
  __shared__ float tile[64][65];
  const int b = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 64;
  const int c0 = blockIdx.y * 64;
  const float* src = in + (long long)b * C * P;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = c0 + ty + i * 4;
    const long long p = p0 + tx;
    tile[ty + i * 4][tx] = (c < C && p < P) ? src[(long long)c * P + p] : 0.f;
  }
  __syncthreads();
  // 32 channel pairs x 8 pixels per pass
  const int cp = threadIdx.x & 31, pr = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long p = p0 + pr + i * 8;
    const int c = c0 + 2 * cp;
    if (p < P && c < C) {  // C is even
      uint32_t hi, mid;
      tc::split_bf16x2(tile[2 * cp][pr + i * 8], tile[2 * cp + 1][pr + i * 8], hi, mid);
      const long long o = ((long long)b * P + p) * C + c;
      *reinterpret_cast<uint32_t*>(planes + o) = hi;
      *reinterpret_cast<uint32_t*>(planes + plane_elems + o) = mid;
    }
  }
}

// fp32 NHWC [B][H][W][C] -> channel-major, zero-PADDED, TF32-rounded out[c][(b*(H+2) + y+1)*Wp + x+1] with row pitch
// ld = B*(H+2)*Wp (the K-major operands of mde_conv3x3_wgrad_tf32; Wp >= W+2): 64 x 64 tile transpose over the padded pixel
// axis, pads written as 0.  SHIFT3: three copies are written, copy s (0..2) shifted by s-1 along k (out3[s][c][k] =
// pad[c][k - (s-1)]), i.e. the horizontal filter tap baked into the data.
template <bool SHIFT3>
__global__ void __launch_bounds__(256) nhwc_to_cpad_tf32_kernel(const float* __restrict__ in, float* __restrict__ out, int C,
                                                                int H, int W, int Wp, long long ld) {
  __shared__ float tile[64][65];
  const long long k0 = (long long)blockIdx.x * 64;
  const int c0 = blockIdx.y * 64;
  const int Hp = H + 2;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
#pragma unroll
  for (int i = 0; i < 16; ++i) {  // rows of the tile = padded pixels, columns = channels (contiguous in the source)
    const long long k = k0 + ty + i * 4;
    const int c = c0 + tx;
    float v = 0.f;
    if (k < ld && c < C) {
      const int xp = (int)(k % Wp);
      const long long r = k / Wp;
      const int yp = (int)(r % Hp);
      const long long b = r / Hp;
      if (xp >= 1 && xp <= W && yp >= 1 && yp <= H) v = in[((b * H + (yp - 1)) * W + (xp - 1)) * C + c];
    }
    tile[ty + i * 4][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = c0 + ty + i * 4;
    const long long k = k0 + tx;
    if (c >= C || k >= ld) continue;
    const float v = tc::tf32_round(tile[tx][ty + i * 4]);
    if (!SHIFT3) {
      out[(long long)c * ld + k] = v;
    } else {
      const long long plane = (long long)C * ld;
      float* row = out + (long long)c * ld;
      if (k >= 1) row[k - 1] = v;                       // copy 0: out3[0][c][k'] = pad[c][k' + 1]
      row[plane + k] = v;                               // copy 1
      if (k + 1 < ld) row[2 * plane + k + 1] = v;       // copy 2: out3[2][c][k'] = pad[c][k' - 1]
      if (k == ld - 1) row[k] = 0.f;                    // the two elements no source position maps to
      if (k == 0) row[2 * plane] = 0.f;
    }
  }
}

}  // namespace mde

using namespace mde;

extern "C" {

int mde_split_bf16(const float* x, uint16_t* planes, int64_t n, mde_stream_t stream) {
  if (!x || !planes) return MDE_ERR_BAD_POINTER;
  if (n <= 0 || n % 8 != 0) return n == 0 ? MDE_OK : MDE_ERR_BAD_SHAPE;
  if (!aligned(x, 16) || !aligned(planes, 16)) return MDE_ERR_BAD_POINTER;
  long long g = (n / 8 + 255) / 256;
  if (g > MDE_NUM_SMS * 16) g = MDE_NUM_SMS * 16;
  split_bf16_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(x, planes, n / 8, n);
  return check_launch();
}

int mde_merge_bf16(const uint16_t* planes, float* out, int64_t n, mde_stream_t stream) {
  if (!planes || !out) return MDE_ERR_BAD_POINTER;
  if (n <= 0 || n % 8 != 0) return n == 0 ? MDE_OK : MDE_ERR_BAD_SHAPE;
  if (!aligned(out, 16) || !aligned(planes, 16)) return MDE_ERR_BAD_POINTER;
  long long g = (n / 8 + 255) / 256;
  if (g > MDE_NUM_SMS * 16) g = MDE_NUM_SMS * 16;
  merge_bf16_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(planes, out, n / 8, n);
  return check_launch();
}

int mde_split_bf16_nchw(const float* x_nchw, uint16_t* planes_nhwc, int B, int C, int64_t P, mde_stream_t stream) {
  if (!x_nchw || !planes_nhwc) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || C <= 0 || P <= 0 || B > 65535 || (C + 63) / 64 > 65535 || C % 2 != 0) return MDE_ERR_BAD_SHAPE;
  dim3 grid((unsigned)((P + 63) / 64), (unsigned)((C + 63) / 64), (unsigned)B);
  split_bf16_nchw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x_nchw, planes_nhwc, C, P, (long long)B * P * C);
  return check_launch();
}

int mde_nhwc_to_cpad_tf32(const float* x_nhwc, float* out, int B, int H, int W, int C, int Wp, int shift3,
                          mde_stream_t stream) {
  if (!x_nhwc || !out) return MDE_ERR_BAD_POINTER;
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || Wp < W + 2 || Wp % 4 != 0 || (C + 63) / 64 > 65535) return MDE_ERR_BAD_SHAPE;
  const long long ld = (long long)B * (H + 2) * Wp;
  dim3 grid((unsigned)((ld + 63) / 64), (unsigned)((C + 63) / 64));
  if (shift3) nhwc_to_cpad_tf32_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(x_nhwc, out, C, H, W, Wp, ld);
  else nhwc_to_cpad_tf32_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(x_nhwc, out, C, H, W, Wp, ld);
  return check_launch();
}

}  // extern "C"
