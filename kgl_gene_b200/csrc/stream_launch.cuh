// stream_launch.cuh -- picks the streaming kernel for a plan and launches it (shared by the C ABI and tools/kbench.cu).
#pragma once
#include "stream_count.cuh"
#include "stream_count_generic.cuh"

namespace kgl {

template <bool WL, bool WG>
inline cudaError_t launch_stream_variant(const StreamParams& P, const StreamPlan& pl, cudaStream_t st) {
  dim3 grid(pl.n_ctas, pl.slices);
  cudaError_t e;
  if (pl.shape == 1) {
    e = cudaFuncSetAttribute(k_stream_count_ct<40, 64, WL, WG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return e;
    k_stream_count_ct<40, 64, WL, WG><<<grid, pl.threads, pl.smem, st>>>(P);
  } else if (pl.shape == 2) {
    e = cudaFuncSetAttribute(k_stream_count_ct<8, 256, WL, WG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return e;
    k_stream_count_ct<8, 256, WL, WG><<<grid, pl.threads, pl.smem, st>>>(P);
  } else {
    e = cudaFuncSetAttribute(k_stream_count_rt<WL, WG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return e;
    k_stream_count_rt<WL, WG><<<grid, pl.threads, pl.smem, st>>>(P);
  }
  return cudaGetLastError();
}

inline cudaError_t launch_stream(const StreamParams& P, const StreamPlan& pl, bool want_locus, bool want_genome, cudaStream_t st) {
  if (want_locus && want_genome) return launch_stream_variant<true, true>(P, pl, st);
  if (want_locus) return launch_stream_variant<true, false>(P, pl, st);
  if (want_genome) return launch_stream_variant<false, true>(P, pl, st);
  return launch_stream_variant<false, false>(P, pl, st);
}

}  // namespace kgl
