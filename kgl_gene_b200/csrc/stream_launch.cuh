// stream_launch.cuh -- picks the streaming kernel for a plan and launches it (shared by the C ABI and tools/kbench.cu).
#pragma once
#include "stream_count.cuh"
#include "stream_count_generic.cuh"

namespace kgl {

// The kernels ask for the largest shared-memory carve-out (228 KB): an SM changes its carve-out only when it is idle, so a
// co-running kernel of a side stream (preparation, tail: kgl_b200_api.cu) must find the configuration it needs already in place
// -- with the default carve-out (the smallest that fits, 196 KB for the 181 KB of the <40,64> shape) a 15 KB block does not fit
// next to a streaming CTA and waits for the whole kernel (measured: tools/kbench --cfg -2).
template <class K>
inline cudaError_t stream_kernel_attributes(K kernel, size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
}

// 2-D tensor map of the matrix for the sliced <40,64> shape (box = 64 rows x 40 units). False: the driver entry point is not
// available or the encode failed -- the kernel then falls back to one bulk copy per row.
inline bool stream_make_tensor_map(StreamParams& P, const StreamPlan& pl, uint64_t padded_rows) {
  P.use_tmap = 0;
  if (pl.shape != 1 || pl.slices <= 1) return false;
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  static bool looked_up = false;
  if (!looked_up) {
    looked_up = true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      encode = reinterpret_cast<EncodeFn>(fn);
    else cudaGetLastError();
  }
  if (!encode) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)P.units * 4, (cuuint64_t)padded_rows};
  const cuuint64_t strides[1] = {(cuuint64_t)P.units * 16};                       // bytes between rows
  const cuuint32_t box[2] = {pl.slice_units * 4, pl.rows_per_stage};
  const cuuint32_t elem[2] = {1, 1};
  const CUresult r = encode(&P.tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint4*>(P.packed), dims, strides, box, elem,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  P.use_tmap = 1;
  return true;
}

template <bool WL, bool WG>
inline cudaError_t launch_stream_variant(const StreamParams& P, const StreamPlan& pl, cudaStream_t st) {
  dim3 grid(pl.n_ctas, pl.slices);
  static size_t configured[3] = {0, 0, 0};      // per template instance: the dynamic shared-memory size the kernel was last set up for
  cudaError_t e = cudaSuccess;
  if (pl.shape == 1) {
    if (configured[1] != pl.smem) { e = stream_kernel_attributes(k_stream_count_ct<40, 64, WL, WG>, pl.smem); if (e != cudaSuccess) return e; configured[1] = pl.smem; }
    k_stream_count_ct<40, 64, WL, WG><<<grid, pl.threads, pl.smem, st>>>(P);
  } else if (pl.shape == 2) {
    if (configured[2] != pl.smem) { e = stream_kernel_attributes(k_stream_count_ct<8, 256, WL, WG>, pl.smem); if (e != cudaSuccess) return e; configured[2] = pl.smem; }
    k_stream_count_ct<8, 256, WL, WG><<<grid, pl.threads, pl.smem, st>>>(P);
  } else {
    if (configured[0] != pl.smem) { e = stream_kernel_attributes(k_stream_count_rt<WL, WG>, pl.smem); if (e != cudaSuccess) return e; configured[0] = pl.smem; }
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
