// Shared helpers for the sm_100a kernels of the AdaBins head / loss / external-info path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mde_b200.h"

#define MDE_NUM_SMS 148  // B200: 2 dies x 74 SMs

namespace mde {

extern unsigned long long g_launch_count;  // bumped by every launch helper (mde_launch_count())

inline int check_launch() {
  g_launch_count++;
  return cudaGetLastError() == cudaSuccess ? MDE_OK : MDE_ERR_LAUNCH;
}

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------
// One inference step is ~195 kernel launches of 5-700 us; between two dependent kernels the GPU pays the launch latency and the
// next kernel's prologue (mbarrier init, TMEM allocation, tensor-map fetch) with nothing running.  Kernels launched through
// launch_pdl() carry cudaLaunchAttributeProgrammaticStreamSerialization: the device may start their CTAs while the previous
// kernel of the stream is still running.  Such a kernel runs pdl_sync() BEFORE its first access to global memory -- it blocks
// until the previous grid has completed and its writes are visible -- and thereby also releases ITS successor (the trigger is
// issued right after the wait: a successor never gets ahead of more than one kernel, and because the trigger only fires once
// every CTA of a grid has issued it, a waiting successor cannot starve CTAs of its predecessor that are not yet resident).
// Everything before pdl_sync() -- the prologue -- overlaps the predecessor.  Without the attribute (MDE_PDL=0, or a kernel that
// follows a memset / a library kernel) both instructions are no-ops and the launch is an ordinary one.
// Kernel classes (a bit each in the MDE_PDL mask / mde_set_pdl): the measured effect differs by class, see DESIGN.md section 7
enum { PDL_CHAIN = 1,    // short dependent chains of small kernels: transformer (GEMM, attention, LayerNorm), regressor, query fold
       PDL_TC = 2,       // persistent one-CTA-per-SM tcgen05 kernels (conv3x3, point-wise GEMM, patch embedding, fused chain)
       PDL_STREAM = 4 }; // streaming SIMT kernels (bias / SiLU / pooling, squeeze-excite gate, resize + concat, stem)
bool pdl_enabled(int cls);  // MDE_PDL environment mask / mde_set_pdl (gather.cu)

__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename... KArgs, typename... Args>
inline void launch_pdl(int cls, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled(cls) ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// streaming (read-once) 128-bit global load / store: keep L1 for the reused data
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

}  // namespace mde
