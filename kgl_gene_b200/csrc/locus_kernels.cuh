// locus_kernels.cuh -- per-locus preparation: selection flags and the genome-independent ("dense") totals.
//
// generateFrequencies (kga_analysis_inbreed_freq.cpp:425-583) rebuilds, for every genome, quantities that depend only
// on (locus, super-population): the AlleleFreqVector, its validity, q > 0.01, and alleleClassFrequencies(0.0). Here
// they are computed once per (locus, population); the per-genome pass then only has to correct for the genotypes that
// deviate from hom-ref ("dense minus sparse").
#pragma once
#include "common.cuh"

namespace kgl {

// Per-population totals over the selected, valid loci of this shard.
enum { TOT_T = 0,       // number of selected valid loci
       TOT_TQ,          // ... with q > 0.01 (a hom-ref genome is MAJOR_HOMOZYGOUS there, else dropped; freq.cpp:532-539)
       TOT_EMAJHOM, TOT_EMAJHET, TOT_EMINHOM,   // sums of alleleClassFrequencies(0.0)
       TOT_W0,          // sum over q > 0.01 loci of (1/q - 1): the Ritland term of a hom-ref genome (calc.cpp:397-401)
       TOT_COUNT };

constexpr int kPrepThreads = 256;
constexpr int kPrepIters = 8;                                   // 64-locus groups per warp
constexpr int kPrepLociPerBlock = (kPrepThreads / 32) * 64 * kPrepIters;   // 4096

// flags16[l] bit k: locus selected for population k AND its AF vector is valid; bit 8+k: ... AND q_k <= 0.01 ("rare-q")
// sum64[l/64]  low byte: AND over the group's rows of the low flag byte, high byte: OR (rows >= n_loci count as 0)
// selw[k][w]   bit i: flags bit k of locus 32w+i (sample-major kernels)
// rare_rows / n_rare: list of the rows with a rare-q bit (any order)
// block_totals[block][k][TOT_COUNT]
// A warp owns 64 consecutive loci per iteration (lane -> l, l+32), so the group summary and the selection words need no
// shared memory; the totals stay in registers until the end of the block.
__global__ void __launch_bounds__(kPrepThreads)
k_locus_prepare(const float* __restrict__ af, const uint8_t* __restrict__ sel, uint64_t n_loci, uint64_t padded_rows, int n_pop,
                int select_all, uint16_t* __restrict__ flags16, uint16_t* __restrict__ sum64,
                uint32_t* __restrict__ selw, uint64_t n_words, uint32_t* __restrict__ rare_rows, uint32_t* __restrict__ n_rare,
                double* __restrict__ block_totals) {
  __shared__ double s_tot[kPrepThreads / 32][kMaxPop][TOT_COUNT];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc[kMaxPop][TOT_COUNT];
#pragma unroll
  for (int k = 0; k < kMaxPop; ++k)
#pragma unroll
    for (int j = 0; j < TOT_COUNT; ++j) acc[k][j] = 0.0;

  for (int it = 0; it < kPrepIters; ++it) {
    const uint64_t l0 = (uint64_t)blockIdx.x * kPrepLociPerBlock + ((uint64_t)it * (kPrepThreads / 32) + warp) * 64 + lane;
    if (l0 - lane >= padded_rows && l0 - lane >= n_words * 32) break;       // warp-uniform
    uint32_t fls[2];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const uint64_t l = l0 + 32 * half;
      const uint8_t s = (l < n_loci) ? (select_all ? 0x3f : sel[l]) : 0;
      uint32_t fl = 0, rq = 0;
#pragma unroll
      for (int k = 0; k < kMaxPop; ++k) {
        bool on = false;
        if (k < n_pop && ((s >> k) & 1)) {
          const LocusFreq f = locus_freq(af[(uint64_t)k * n_loci + l]);
          if (f.valid) {
            on = true;
            fl |= 1u << k;
            double a, b, c;
            class_freqs(f.p, a, b, c);
            acc[k][TOT_T] += 1.0; acc[k][TOT_EMAJHOM] += a; acc[k][TOT_EMAJHET] += b; acc[k][TOT_EMINHOM] += c;
            if (f.q > kMinMajorFreq) { acc[k][TOT_TQ] += 1.0; acc[k][TOT_W0] += __dsub_rn(__ddiv_rn(1.0, f.q), 1.0); }
            else rq |= 1u << k;
          }
        }
        const uint32_t word = __ballot_sync(kFull, on);
        if (lane == 0 && selw != nullptr && k < n_pop) {
          const uint64_t w = l >> 5;
          if (w < n_words) selw[(uint64_t)k * n_words + w] = word;
        }
      }
      if (l < padded_rows) flags16[l] = (uint16_t)(fl | (rq << 8));
      if (rq != 0 && rare_rows != nullptr) rare_rows[atomicAdd(n_rare, 1u)] = (uint32_t)l;
      fls[half] = fl;
    }
    const uint32_t g_and = __reduce_and_sync(kFull, fls[0] & fls[1]);
    const uint32_t g_or = __reduce_or_sync(kFull, fls[0] | fls[1]);
    if (lane == 0 && l0 < padded_rows) sum64[l0 >> 6] = (uint16_t)(g_and | (g_or << 8));
  }

#pragma unroll
  for (int k = 0; k < kMaxPop; ++k)
#pragma unroll
    for (int j = 0; j < TOT_COUNT; ++j) {
      const double v = warp_sum(acc[k][j]);
      if (lane == 0) s_tot[warp][k][j] = v;
    }
  __syncthreads();
  if (threadIdx.x < kMaxPop * TOT_COUNT) {
    const int k = threadIdx.x / TOT_COUNT, j = threadIdx.x % TOT_COUNT;
    double v = 0.0;
    for (int w = 0; w < kPrepThreads / 32; ++w) v += s_tot[w][k][j];
    block_totals[((uint64_t)blockIdx.x * kMaxPop + k) * TOT_COUNT + j] = v;
  }
}

// Fixed-order reduction of the block totals: totals[item], one block per item, deterministic.
__global__ void __launch_bounds__(256)
k_reduce_totals(const double* __restrict__ block_totals, uint64_t n_blocks, double* __restrict__ totals) {
  __shared__ double s[256];
  const int item = blockIdx.x;
  double v = 0.0;
  for (uint64_t b = threadIdx.x; b < n_blocks; b += 256) v += block_totals[b * kMaxPop * TOT_COUNT + item];
  s[threadIdx.x] = v;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[item] = s[0];
}

}  // namespace kgl
