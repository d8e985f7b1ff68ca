// post_kernels.cuh -- everything that follows the streaming kernel of a pass, in ONE launch.
//
// The step used to end with four small, latency-bound launches (k_expand_planes, k_dropped_apply, k_rare_rows,
// k_moment_partials: 8.6 + 21 + 12.7 + 6.8 us on the chr22 shape). The first three are independent of each other, so
// k_post runs them side by side as block roles of one grid; the block that finishes last (ticket counter) then assembles
// the per-genome results, which depend on all three. The streaming kernel leaves no room on the SMs for a concurrent
// kernel (95 registers x 608 threads per SM), so this is the overlap that is available.
#pragma once
#include "misc_kernels.cuh"
#include "sparse_events.cuh"
#include "stream_common.cuh"

namespace kgl {

struct PostParams {
  // role E: counter expansion
  const uint32_t* planes; uint64_t n_vchunks, units, n_genomes_padded; uint32_t* gcounts;
  uint32_t e_bx, e_by;                      // role E grid
  // role D: indexed code-3 cells (d_blocks == 0: none)
  const DroppedKey* keys; const uint64_t* seg; uint32_t d_blocks;
  // role R: rare-major rows (r_blocks == 0: raw mode)
  const uint32_t* rare_rows; uint32_t* n_rare; const uint4* packed; const uint64_t* popmask; uint32_t r_blocks;
  // shared inputs
  uint64_t n_genomes, n_loci; int n_pop;
  const uint16_t* flags16; const uint32_t* all_selected; const uint8_t* superpop; const float* af;
  SparseOut so;
  // tail (last block): 1 = moment partials (+ Simple closed form when results != null), 2 = raw genome counts
  int tail_mode; int unphased;
  const double* totals; double* partials; kgl_b200_locus_results* results; uint64_t* genome_counts;
  unsigned int* ticket;
};

// 6 blocks per SM (40 registers) and 16 counter chunks per expansion block: the 626 + 32 + 190 blocks of the chr22 shape are
// resident together. With 48 registers and 740 expansion blocks the grid needed 1.9 waves and the roles ran one after the other.
__global__ void __launch_bounds__(256, 6)
k_post(const PostParams P) {
  __shared__ int s_last;
  // the code-3 gathers are the long pole (one L2 sector per cell) and go first, in one wave; the rare rows (few blocks, a
  // chain of five dependent loads) and the short expansion blocks fill in behind them
  const uint32_t b = blockIdx.x;
  if (b < P.d_blocks) {
    dropped_apply_block(b, P.keys, P.seg, P.n_genomes, P.flags16, P.all_selected, P.superpop, P.af, P.n_loci, P.so);
  } else if (b < P.d_blocks + P.r_blocks) {
    rare_rows_block(b - P.d_blocks, P.r_blocks, P.rare_rows, P.n_rare, P.packed, (uint32_t)P.units, (uint32_t)P.n_genomes,
                    P.flags16, P.popmask, P.af, P.n_loci, P.n_pop, P.so);
  } else {
    const uint32_t e = b - P.r_blocks - P.d_blocks;
    expand_planes_block(e % P.e_bx, e / P.e_bx, P.planes, P.n_vchunks, P.units, P.n_genomes_padded, P.gcounts);
  }
  if (P.tail_mode == 0) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(P.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (uint64_t g = threadIdx.x; g < P.n_genomes; g += blockDim.x) {
    if (P.tail_mode == 1) {
      moment_partials_one(g, P.gcounts, P.so.n3, P.totals, P.so.ecorr, P.so.nz_rare, P.superpop, P.unphased, P.partials, P.results);
    } else {
      const uint64_t n3 = __ldcg(&P.so.n3[g]), n1 = __ldcg(&P.gcounts[g * 2]) - n3, n2 = __ldcg(&P.gcounts[g * 2 + 1]) - n3;
      uint64_t* o = P.genome_counts + g * 4;
      o[0] = P.n_loci - n1 - n2 - n3; o[1] = n1; o[2] = n2; o[3] = n3;
    }
  }
  if (threadIdx.x == 0) *P.ticket = 0;     // ready for the next pass
}

}  // namespace kgl
