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

// flags16[l] bit k: locus selected for population k AND its AF vector is valid; bit 8+k: ... AND q_k <= 0.01 ("rare-q")
// selw[k][w] bit i: flags bit k of locus 32w+i (sample-major kernels)
// block_totals[block][k][TOT_COUNT]
__global__ void __launch_bounds__(kPrepThreads)
k_locus_prepare(const float* __restrict__ af, const uint8_t* __restrict__ sel, uint64_t n_loci, int n_pop,
                int select_all, uint16_t* __restrict__ flags16,
                uint32_t* __restrict__ selw, uint64_t n_words, double* __restrict__ block_totals) {
  __shared__ double s_tot[kPrepThreads / 32][kMaxPop][TOT_COUNT];
  const uint64_t l = (uint64_t)blockIdx.x * kPrepThreads + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint8_t s = (l < n_loci) ? (select_all ? 0x3f : sel[l]) : 0;
  uint8_t fl = 0, rq = 0;
  for (int k = 0; k < n_pop; ++k) {
    double t[TOT_COUNT] = {0, 0, 0, 0, 0, 0};
    bool on = false;
    if (l < n_loci && ((s >> k) & 1)) {
      const LocusFreq f = locus_freq(af[(uint64_t)k * n_loci + l]);
      if (f.valid) {
        on = true;
        fl |= (uint8_t)(1u << k);
        double a, b, c;
        class_freqs(f.p, a, b, c);
        t[TOT_T] = 1.0; t[TOT_EMAJHOM] = a; t[TOT_EMAJHET] = b; t[TOT_EMINHOM] = c;
        if (f.q > kMinMajorFreq) { t[TOT_TQ] = 1.0; t[TOT_W0] = __dsub_rn(__ddiv_rn(1.0, f.q), 1.0); } else rq |= (uint8_t)(1u << k);
      }
    }
    const uint32_t word = __ballot_sync(kFull, on);
    if (lane == 0 && selw != nullptr) {
      const uint64_t w = l >> 5;   // kPrepThreads is a multiple of 32, so a warp covers exactly one word
      if (w < n_words) selw[(uint64_t)k * n_words + w] = word;
    }
#pragma unroll
    for (int j = 0; j < TOT_COUNT; ++j) {
      const double v = warp_sum(t[j]);
      if (lane == 0) s_tot[warp][k][j] = v;
    }
  }
  if (l < n_loci) flags16[l] = (uint16_t)(fl | ((uint16_t)rq << 8));
  __syncthreads();
  if (threadIdx.x < n_pop * TOT_COUNT) {
    const int k = threadIdx.x / TOT_COUNT, j = threadIdx.x % TOT_COUNT;
    double v = 0.0;
    for (int w = 0; w < kPrepThreads / 32; ++w) v += s_tot[w][k][j];
    block_totals[((uint64_t)blockIdx.x * kMaxPop + k) * TOT_COUNT + j] = v;
  }
}

// Fixed-order reduction of the block totals: totals[k][j]. One block, deterministic.
__global__ void __launch_bounds__(256)
k_reduce_totals(const double* __restrict__ block_totals, uint64_t n_blocks, double* __restrict__ totals) {
  __shared__ double s[256];
  for (int item = 0; item < kMaxPop * TOT_COUNT; ++item) {
    double v = 0.0;
    for (uint64_t b = threadIdx.x; b < n_blocks; b += 256) v += block_totals[b * kMaxPop * TOT_COUNT + item];
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) totals[item] = s[0];
    __syncthreads();
  }
}

}  // namespace kgl
