// gram_launch.cuh -- host-side plan and launch of k_gram_i8.
#pragma once
#include "gram_i8.cuh"

#include <vector>

namespace kgl {

// 256 x 256 tiles (ti, tj), ti <= tj: the upper triangle of an ld x ld matrix (ld a multiple of 256).
inline std::vector<uint2> gram_upper_tiles(uint64_t ld) {
  static_assert(kGramM == kGramN, "square tiles");
  std::vector<uint2> t;
  for (uint32_t ti = 0; ti < ld / kGramM; ++ti)
    for (uint32_t tj = ti; tj < ld / kGramN; ++tj) t.push_back(make_uint2(ti, tj));
  return t;
}

struct GramPlan { uint32_t stages_per_chunk, n_chunks, grid; };

inline GramPlan plan_gram(uint32_t n_tiles, uint32_t k_stages, int sm_count, uint32_t chunk_stages_hint = 0) {
  GramPlan p{};
  uint32_t best = 1;
  if (chunk_stages_hint) {
    best = (k_stages + chunk_stages_hint - 1) / chunk_stages_hint;
  } else {
    const uint32_t max_chunks = k_stages >= 64 ? k_stages / 64 : 1;      // >= 64 stages (8,192 loci) per unit
    double best_eff = -1.0;
    for (uint32_t c = 1; c <= (max_chunks < 2048 ? max_chunks : 2048); ++c) {
      const uint32_t sp = (k_stages + c - 1) / c, nc = (k_stages + sp - 1) / sp;
      const uint64_t units = (uint64_t)n_tiles * nc;
      const uint64_t grid = units < (uint64_t)sm_count ? units : (uint64_t)sm_count;
      const uint64_t rounds = (units + grid - 1) / grid;
      const double t = (double)rounds * ((double)sp + 8.0);                // ~8 stage-times per unit for the epilogue
      const double eff = (double)n_tiles * k_stages / ((double)sm_count * t);
      if (eff > best_eff + 1e-9) { best_eff = eff; best = nc; }
    }
  }
  const uint32_t sp = (k_stages + best - 1) / best;
  p.stages_per_chunk = sp;
  p.n_chunks = (k_stages + sp - 1) / sp;
  const uint64_t units = (uint64_t)n_tiles * p.n_chunks;
  p.grid = (uint32_t)(units < (uint64_t)sm_count ? units : (uint64_t)sm_count);
  if (p.grid == 0) p.grid = 1;
  return p;
}

inline cudaError_t launch_gram(const GramParams& P, const GramPlan& pl, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(k_gram_i8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGramSmem);
  if (e != cudaSuccess) return e;
  k_gram_i8<<<pl.grid, kGramThreads, kGramSmem, stream>>>(P);
  return cudaGetLastError();
}

}  // namespace kgl
