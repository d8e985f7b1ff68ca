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

// The totals are accumulated by the preparation blocks with 64-bit integer atomics -- counts as they are, sums in fixed point
// (terms in [0, 1]; TOT_W0 terms are below 128) -- so they are exact sums of the rounded terms, independent of the schedule,
// and need no "last block" pass.
struct DenseTotals {
  const unsigned long long* fx;   // [kMaxPop][TOT_COUNT]
  double inv, inv_w0;             // 1 / scale of the class-frequency sums, of TOT_W0
  __device__ __forceinline__ double get(int k, int j) const {
    const long long v = (long long)__ldcg(&fx[k * TOT_COUNT + j]);
    return (j == TOT_T || j == TOT_TQ) ? (double)v : (double)v * (j == TOT_W0 ? inv_w0 : inv);
  }
};

constexpr int kPrepThreads = 256;
constexpr int kPrepIters = 4;                                   // 64-locus groups per warp
constexpr int kPrepLociPerBlock = (kPrepThreads / 32) * 64 * kPrepIters;   // 2048

// flags16[l] bit k: locus selected for population k AND its AF vector is valid; bit 8+k: ... AND q_k <= 0.01 ("rare-q")
// sum64[l/64]  low byte: AND over the group's rows of the low flag byte, high byte: OR (rows >= n_loci count as 0)
// selw[k][w]   bit i: flags bit k of locus 32w+i (sample-major kernels; only with WANT_SELW -- the Simple step has no use
//              for it and the 48 ballots per thread were a fifth of the kernel)
// totals_fx[k][TOT_COUNT] (zeroed by the caller; TOT_W0 only when WANT_W0: it costs a double-precision divide per locus and
//              only the Ritland estimator reads it)
// rare-major rows (some population has q <= 0.01 there: a hom-ref genome of that population is dropped, freq.cpp:532-539) are
// settled by the block that finds them, right behind its loci: nz_rare[g] += non-reference cells of the population's genomes in
// the row, ecorr_fx[g][2] += class frequencies {majHom, minHom} of the hom-ref ones (64-bit fixed point: the atomic adds commute).
//
// The expected class frequencies alleleClassFrequencies(0.0) of a locus are {q^2, 2qp, p^2} / (q^2 + 2qp + p^2) (freq.cpp:127-217,
// freq.h:54-63). q = fl(1 - p) is exact for every float p >= 2^-29, so the divisor is (q + p)^2 = 1 up to the rounding of the
// three products, |divisor - 1| < 3e-16: the dense totals accumulate the products themselves (three DFMA per locus and
// population instead of nine multiplies, two adds and a subtract); a total differs from the reference's by < 1e-15 relative.
//
// A warp owns 64 consecutive loci per iteration (lane -> l, l+32), so the group summary and the selection words need no
// shared memory. The population loop is the OUTER loop: only one population's totals are live at a time.
// Persistent grid (one or two blocks per SM): a block walks tiles of kPrepLociPerBlock loci with a grid stride, in groups of 512
// loci (a warp owns 64 consecutive loci of a group: lane -> l, l + 32, so the 64-row summary and the selection words need no
// shared memory). The kernel is built to run NEXT to a CTA of the streaming kernel on the same SM (<= 80 registers x 256
// threads, 15 KB of shared memory: kgl_b200_api.cu hides the preparation of pass i+1 behind the streaming kernel of pass i), i.e.
// with eight warps per SM and HBM saturated by its neighbour: the twelve frequencies of a thread's two loci (six populations) are
// loaded one group AHEAD of their use, and the per-population sums stay in registers until the block ends (one warp reduction and
// 36 integer atomics per block instead of one per tile and population).
constexpr int kPrepRareMax = kPrepLociPerBlock;
constexpr int kPrepGroup = (kPrepThreads / 32) * 64;             // 512 loci per group
template <bool WANT_W0, bool WANT_SELW>
__global__ void __maxnreg__(96)
k_locus_prepare(const float* __restrict__ af, const uint8_t* __restrict__ sel, uint64_t n_loci, uint64_t padded_rows, uint64_t n_tiles,
                int n_pop, uint16_t* __restrict__ flags16, uint16_t* __restrict__ sum64, uint32_t* __restrict__ selw, uint64_t n_words,
                const uint4* __restrict__ packed, uint32_t units, const uint64_t* __restrict__ popmask,
                uint32_t* __restrict__ nz_rare, unsigned long long* __restrict__ ecorr_fx, double fx,
                unsigned long long* __restrict__ totals_fx, uint32_t* __restrict__ unselected_blocks) {
  __shared__ double s_tot[kPrepThreads / 32][kMaxPop][TOT_COUNT];
  __shared__ uint32_t s_rare[kPrepRareMax];
  __shared__ uint16_t s_rare_fl[kPrepRareMax];
  __shared__ uint32_t s_n_rare;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t all_pops = (1u << n_pop) - 1u;
  bool all_ok = true;
  // per population: sum q^2, sum p^2 (the heterozygous class is what is left of the count: the three classes of a locus sum to one)
  double e_majhom[kMaxPop], e_minhom[kMaxPop], w0[WANT_W0 ? kMaxPop : 1];
  uint32_t n_t[kMaxPop], n_tq[kMaxPop];
#pragma unroll
  for (int k = 0; k < kMaxPop; ++k) { e_majhom[k] = 0.0; e_minhom[k] = 0.0; n_t[k] = 0; n_tq[k] = 0; if (WANT_W0) w0[k] = 0.0; }

  // groups of this block, in order: tile blockIdx.x groups 0..3, tile blockIdx.x + gridDim.x groups 0..3, ...
  const uint64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const uint64_t n_groups = my_tiles * kPrepIters;
  auto group_base = [&](uint64_t q) -> uint64_t {       // first locus of this thread in group q
    const uint64_t tile = blockIdx.x + (q / kPrepIters) * gridDim.x;
    return tile * kPrepLociPerBlock + (q % kPrepIters) * kPrepGroup + (uint64_t)warp * 64 + lane;
  };
  // The frequencies of a group are LOADED one group ahead of their use (registers: a prefetch hint can be dropped by a memory
  // system that the neighbouring streaming kernel keeps saturated, a load cannot), and pulled into L2 two groups ahead of that.
  float nxt[2][kMaxPop];
  uint32_t nsel[2];
  auto fetch = [&](uint64_t q) {
    const uint64_t base0 = group_base(q);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const uint64_t l = base0 + 32 * half;
      const bool in = l < n_loci;
      nsel[half] = in ? (uint32_t)sel[l] : 0u;
#pragma unroll
      for (int k = 0; k < kMaxPop; ++k) nxt[half][k] = (in && k < n_pop) ? __ldg(af + (uint64_t)k * n_loci + l) : 0.0f;
    }
  };
  auto prefetch = [&](uint64_t q) {
    if (lane < 2 * n_pop) {
      const uint64_t l = group_base(q) - lane + 32 * (lane & 1);
      if (l < n_loci) asm volatile("prefetch.global.L2 [%0];" ::"l"(af + (uint64_t)(lane >> 1) * n_loci + l));
    }
  };
  if (n_groups > 0) fetch(0);
  if (n_groups > 1) prefetch(1);
  if (n_groups > 2) prefetch(2);

  for (uint64_t q = 0; q < n_groups; ++q) {
    if (q % kPrepIters == 0) {
      if (threadIdx.x == 0) s_n_rare = 0;
      __syncthreads();
    }
    float cur[2][kMaxPop];
    uint32_t csel[2];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      csel[half] = nsel[half];
#pragma unroll
      for (int k = 0; k < kMaxPop; ++k) cur[half][k] = nxt[half][k];
    }
    if (q + 1 < n_groups) fetch(q + 1);
    if (q + 3 < n_groups) prefetch(q + 3);
    const uint64_t base = group_base(q);
    uint32_t fl[2] = {0, 0};
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll
      for (int k = 0; k < kMaxPop; ++k) {
        const float a = cur[half][k];
        const bool on = ((csel[half] >> k) & 1u) && !(a != a) && k < n_pop;
        if (on) {
          // clamp(AF, 0, 1) (freq.cpp:47) on the float: the widening conversion is exact and monotone, so the clamp commutes
          // with it; q = 1 - p then lies in [0, 1] by itself (majorAlleleFrequency(), freq.cpp:119-123)
          const double p = (double)fminf(fmaxf(a, 0.0f), 1.0f);
          const double qf = __dsub_rn(1.0, p);
          fl[half] |= 1u << k;
          ++n_t[k];
          e_majhom[k] = fma(qf, qf, e_majhom[k]);
          e_minhom[k] = fma(p, p, e_minhom[k]);
          if (qf > kMinMajorFreq) { ++n_tq[k]; if (WANT_W0) w0[k] += __dsub_rn(__ddiv_rn(1.0, qf), 1.0); }
          else fl[half] |= 0x100u << k;
        }
        if (WANT_SELW) {
          const uint32_t word = __ballot_sync(kFull, on);
          if (lane == 0 && k < n_pop) {
            const uint64_t w = (base + 32 * half) >> 5;
            if (w < n_words) selw[(uint64_t)k * n_words + w] = word;
          }
        }
      }
      const uint64_t l = base + 32 * half;
      if (l < padded_rows) flags16[l] = (uint16_t)fl[half];
      if ((fl[half] >> 8) != 0 && nz_rare != nullptr) {
        const uint32_t slot = atomicAdd(&s_n_rare, 1u);
        s_rare[slot] = (uint32_t)l; s_rare_fl[slot] = (uint16_t)(fl[half] >> 8);
      }
      if (l < n_loci && (fl[half] & all_pops) != all_pops) all_ok = false;
    }
    {
      const uint64_t l0 = base - lane;
      const uint32_t g_and = __reduce_and_sync(kFull, fl[0] & fl[1]) & 0xFFu;
      const uint32_t g_or = __reduce_or_sync(kFull, fl[0] | fl[1]) & 0xFFu;
      if (lane == 0 && l0 < padded_rows) sum64[l0 >> 6] = (uint16_t)(g_and | (g_or << 8));
    }
    if (q % kPrepIters != kPrepIters - 1) continue;

    // ---- end of a tile: its rare-major rows, one thread per (row, 128-bit unit); the loads of an item do not depend on each other ----
    __syncthreads();
    const uint32_t n_rare_tile = s_n_rare;
    const uint64_t total = (uint64_t)n_rare_tile * units;
    for (uint64_t t = threadIdx.x; t < total; t += kPrepThreads) {
      const uint32_t slot = (uint32_t)(t / units), u = (uint32_t)(t % units);
      const uint32_t row = s_rare[slot];
      const uint32_t rq = s_rare_fl[slot];                 // populations with q <= 0.01 at this row (selected & valid)
      const uint4 v = packed[(uint64_t)row * units + u];
      const uint64_t lo = (uint64_t)v.x | ((uint64_t)v.y << 32), hi = (uint64_t)v.z | ((uint64_t)v.w << 32);
      for (int k = 0; k < n_pop; ++k) {
        if (!((rq >> k) & 1u)) continue;
        const uint64_t mask = popmask[(uint64_t)k * units + u];
        if (mask == 0) continue;
        double ca, ch, cm;
        class_freqs(locus_freq(af[(uint64_t)k * n_loci + row]).p, ca, ch, cm);
        const unsigned long long qa = (unsigned long long)__double2ll_rn(ca * fx), qm = (unsigned long long)__double2ll_rn(cm * fx);
        uint64_t homref = ~(lo | hi) & mask;
        uint64_t nonref = (lo | hi) & mask;
        while (homref) {
          const int b = __ffsll((long long)homref) - 1;
          homref &= homref - 1;
          const uint64_t g = (uint64_t)u * 64 + b;
          atomicAdd(&ecorr_fx[g * 2 + 0], qa);
          atomicAdd(&ecorr_fx[g * 2 + 1], qm);
        }
        while (nonref) {
          const int b = __ffsll((long long)nonref) - 1;
          nonref &= nonref - 1;
          atomicAdd(&nz_rare[(uint64_t)u * 64 + b], 1u);
        }
      }
    }
  }

  // the block's sums: one warp reduction per (population, total), then one 64-bit integer atomic each
#pragma unroll
  for (int k = 0; k < kMaxPop; ++k) {
    double acc[TOT_COUNT];
    acc[TOT_T] = (double)n_t[k]; acc[TOT_TQ] = (double)n_tq[k];
    acc[TOT_EMAJHOM] = e_majhom[k]; acc[TOT_EMAJHET] = 0.0; acc[TOT_EMINHOM] = e_minhom[k]; acc[TOT_W0] = WANT_W0 ? w0[WANT_W0 ? k : 0] : 0.0;
#pragma unroll
    for (int j = 0; j < TOT_COUNT; ++j) {
      const double v = warp_sum(acc[j]);
      if (lane == 0) s_tot[warp][k][j] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x < kMaxPop * TOT_COUNT) {
    const int k = threadIdx.x / TOT_COUNT, j = threadIdx.x % TOT_COUNT;
    if (k < n_pop) {
      double v = 0.0;
      if (j == TOT_EMAJHET) {      // 2qp summed = count - sum q^2 - sum p^2: q + p = 1, so the three products of a locus add up to one
        double t = 0.0, a = 0.0, m = 0.0;
        for (int w = 0; w < kPrepThreads / 32; ++w) { t += s_tot[w][k][TOT_T]; a += s_tot[w][k][TOT_EMAJHOM]; m += s_tot[w][k][TOT_EMINHOM]; }
        v = (t - a) - m;
      } else {
        for (int w = 0; w < kPrepThreads / 32; ++w) v += s_tot[w][k][j];
      }
      const double scale = (j == TOT_T || j == TOT_TQ) ? 1.0 : (j == TOT_W0 ? fx * (1.0 / 128.0) : fx);
      const long long qv = __double2ll_rn(v * scale);
      if (qv != 0) atomicAdd(&totals_fx[k * TOT_COUNT + j], (unsigned long long)qv);
    }
  }
  // unselected_blocks (zeroed by the caller) counts the blocks that hold a row < n_loci which is not selected and valid for
  // every population: 0 at the end = the sparse kernels need not look at the flags.
  if (!__syncthreads_and(all_ok) && threadIdx.x == 0) atomicAdd(unselected_blocks, 1u);
}

// RetrieveLociiVector::getAllelesFromTo (kga_analysis_inbreed_locus.cpp:21-72) with lociiSpacing == 0: a locus is taken for
// population k iff lower <= offset <= upper, its AF entry exists, and p = clamp(AF, 0, 1) satisfies p != 0 and
// min_af <= p <= max_af (:53-54). sel[l] bit k; counts[k] += selected loci.
// keep (nullable): per-locus verdict of the variant-level filters (kgl_b200_set_locus_filter: SNP and PASS of the frequency
// source, kga_analysis_inbreed.cpp:79; the Pf7 INFO-field filters, kga_analysis_lib_PfFilter.cpp:62-92) -- 0 = the locus is no candidate.
__global__ void __launch_bounds__(256)
k_select_dense(const float* __restrict__ af, const uint32_t* __restrict__ offsets, uint64_t n_loci, int n_pop, uint64_t lower,
               uint64_t upper, double min_af, double max_af, const uint8_t* __restrict__ keep, uint8_t* __restrict__ sel,
               unsigned long long* __restrict__ counts) {
  const uint64_t l = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t bits = 0;
  if (l < n_loci) {
    const uint64_t offset = offsets[l];
    if (offset >= lower && offset <= upper && (keep == nullptr || keep[l] != 0)) {
      for (int k = 0; k < n_pop; ++k) {
        const float a = af[(uint64_t)k * n_loci + l];
        if (a != a) continue;
        double p = (double)a;
        p = p < 0.0 ? 0.0 : (p > 1.0 ? 1.0 : p);
        if (p == 0.0 || p < min_af || p > max_af) continue;
        bits |= 1u << k;
      }
    }
    sel[l] = (uint8_t)bits;
  }
  for (int k = 0; k < n_pop; ++k) {
    const uint32_t n = __popc(__ballot_sync(kFull, (bits >> k) & 1u));
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(&counts[k], (unsigned long long)n);
  }
}

// RetrieveLociiVector::getAllelesCount (kga_analysis_inbreed_locus.cpp:105-156) on an accepted-locus mask: the first `count`
// loci of population k at or after row `begin` (the walk stops as soon as `count` loci have been taken, :117). out[0] = loci
// found (<= count), out[1] = row of the last one. One block; the rows are walked 1,024 at a time.
__global__ void __launch_bounds__(1024)
k_rank_find(const uint8_t* __restrict__ sel, uint64_t begin, uint64_t end, int k, uint64_t count, unsigned long long* __restrict__ out) {
  __shared__ uint32_t s_warp[32];
  __shared__ unsigned long long s_found, s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { s_found = 0; s_last = ~0ull; }
  __syncthreads();
  for (uint64_t l0 = begin; l0 < end; l0 += 1024) {
    const uint64_t l = l0 + threadIdx.x;
    const bool on = l < end && ((sel[l] >> k) & 1u);
    const uint32_t bal = __ballot_sync(kFull, on);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    unsigned long long before = s_found;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    const unsigned long long rank = before + __popc(bal & ((1u << lane) - 1u));     // loci found before this one
    uint32_t total = 0;
    for (int w = 0; w < 32; ++w) total += s_warp[w];
    // the last locus that is still taken: rank count - 1, or the last one found when there are fewer
    if (on && (rank + 1 == count || (rank + 1 < count && rank + 1 == s_found + total))) s_last = l;
    __syncthreads();
    if (threadIdx.x == 0) s_found = min((unsigned long long)count, s_found + total);
    __syncthreads();
    if (s_found >= count) break;
  }
  if (threadIdx.x == 0) { out[0] = s_found; out[1] = s_last; }
}

// ---- spaced selection on the device (getAllelesFromTo with SamplingDistance > 0, kga_analysis_inbreed_locus.cpp:21-72) ------
// The reference walks the window in offset order and accepts a locus when it is a candidate (valid frequency vector inside
// [min_af, max_af], the bits k_select_dense leaves in sel) and `offset >= previous_offset + spacing || previous_offset == 0`
// (:38), previous_offset being the offset of the last ACCEPTED locus. That is a chain over the candidates: it starts at the
// first candidate of the window, and the successor of an accepted locus i is the first candidate j > i with
// offset[j] >= offset[i] + spacing (the next candidate if offset[i] == 0). The chain is marked in parallel by pointer doubling:
// round r marks the successors 2^r steps ahead of every marked locus and squares the jump table, so ceil(log2 L) + 1 rounds
// mark a chain of any length (marking a locus "early" is harmless: whatever a marked locus jumps to is on the chain).
constexpr uint32_t kChainEnd = 0xFFFFFFFFu;
constexpr int kChainBlock = 256;

// next_valid[k][l]: first candidate of population k at or after l INSIDE l's 256-locus block (kChainEnd: none);
// block_first[k][b]: first candidate of block b.
__global__ void __launch_bounds__(kChainBlock)
k_chain_next_valid(const uint8_t* __restrict__ sel, uint64_t n_loci, uint64_t n_blocks, uint32_t* __restrict__ next_valid,
                   uint32_t* __restrict__ block_first) {
  __shared__ uint32_t s_first[kChainBlock / 32];
  const int k = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t l = (uint64_t)blockIdx.x * kChainBlock + threadIdx.x;
  const bool valid = l < n_loci && ((sel[l] >> k) & 1u);
  const uint32_t bal = __ballot_sync(kFull, valid);
  if (lane == 0) s_first[warp] = bal ? (uint32_t)(l + __ffs(bal) - 1) : kChainEnd;
  __syncthreads();
  uint32_t nv = kChainEnd;
  const uint32_t rest = bal >> lane;
  if (rest) nv = (uint32_t)l + (uint32_t)__ffs(rest) - 1;
  else for (int w = warp + 1; w < kChainBlock / 32; ++w) if (s_first[w] != kChainEnd) { nv = s_first[w]; break; }
  if (l < n_loci) next_valid[(uint64_t)k * n_loci + l] = nv;
  if (threadIdx.x == 0) {
    uint32_t f = kChainEnd;
    for (int w = 0; w < kChainBlock / 32; ++w) if (s_first[w] != kChainEnd) { f = s_first[w]; break; }
    block_first[(uint64_t)k * n_blocks + blockIdx.x] = f;
  }
}

// block_after[k][b]: first candidate in any block > b. One warp per population walks the blocks backwards, 32 at a time.
__global__ void __launch_bounds__(32)
k_chain_block_suffix(const uint32_t* __restrict__ block_first, uint64_t n_blocks, uint32_t* __restrict__ block_after) {
  const int k = blockIdx.x, lane = threadIdx.x;
  uint32_t carry = kChainEnd;                                         // first candidate beyond the current group of 32 blocks
  for (int64_t g0 = (int64_t)((n_blocks + 31) / 32 * 32) - 32; g0 >= 0; g0 -= 32) {
    const uint64_t b = (uint64_t)g0 + lane;
    const uint32_t f = b < n_blocks ? block_first[(uint64_t)k * n_blocks + b] : kChainEnd;
    // suffix minimum over the lanes above this one (candidates are indices, blocks are ordered: the minimum is the first)
    uint32_t after = carry, group_first = carry;
    for (int o = 31; o >= 0; --o) {                                   // every lane takes part in every shuffle
      const uint32_t v = __shfl_sync(kFull, f, o);
      if (v != kChainEnd) { group_first = v; if (o > lane) after = v; }
    }
    if (b < n_blocks) block_after[(uint64_t)k * n_blocks + b] = after;
    carry = group_first;
  }
}

__device__ __forceinline__ uint32_t chain_next_valid(const uint32_t* next_valid, const uint32_t* block_after, uint64_t n_loci,
                                                     uint64_t n_blocks, int k, uint64_t l) {
  if (l >= n_loci) return kChainEnd;
  const uint32_t nv = next_valid[(uint64_t)k * n_loci + l];
  return nv != kChainEnd ? nv : block_after[(uint64_t)k * n_blocks + l / kChainBlock];
}

// jump[k][i] = successor of candidate i; mark[k][.] = 1 at the first candidate of the window.
__global__ void __launch_bounds__(256)
k_chain_successor(const uint8_t* __restrict__ sel, const uint32_t* __restrict__ offsets, uint64_t n_loci, uint64_t n_blocks,
                  uint64_t spacing, const uint32_t* __restrict__ next_valid, const uint32_t* __restrict__ block_after,
                  uint32_t* __restrict__ jump, uint8_t* __restrict__ mark) {
  const int k = blockIdx.y;
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_loci) return;
  if (i == 0) {
    const uint32_t start = chain_next_valid(next_valid, block_after, n_loci, n_blocks, k, 0);
    if (start != kChainEnd) mark[(uint64_t)k * n_loci + start] = 1;
  }
  if (!((sel[i] >> k) & 1u)) return;
  const uint64_t o = offsets[i];
  uint64_t t = i + 1;
  if (o != 0) {                                        // lower_bound(offsets, o + spacing) in (i, n_loci)
    const uint64_t want = o + spacing;
    uint64_t lo = i + 1, hi = n_loci;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if ((uint64_t)offsets[mid] < want) lo = mid + 1; else hi = mid; }
    t = lo;
  }
  jump[(uint64_t)k * n_loci + i] = chain_next_valid(next_valid, block_after, n_loci, n_blocks, k, t);
}

__global__ void __launch_bounds__(256)
k_chain_round(const uint8_t* __restrict__ sel, uint64_t n_loci, const uint32_t* __restrict__ jump_in, uint32_t* __restrict__ jump_out,
              uint8_t* __restrict__ mark) {
  const int k = blockIdx.y;
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_loci || !((sel[i] >> k) & 1u)) return;
  const uint64_t base = (uint64_t)k * n_loci;
  const uint32_t j = jump_in[base + i];
  if (j == kChainEnd) { jump_out[base + i] = kChainEnd; return; }
  if (mark[base + i]) mark[base + j] = 1;
  jump_out[base + i] = jump_in[base + j];
}

// sel[l] = accepted bits; counts[k] += accepted loci.
__global__ void __launch_bounds__(256)
k_chain_finish(const uint8_t* __restrict__ mark, uint64_t n_loci, int n_pop, uint8_t* __restrict__ sel, unsigned long long* __restrict__ counts) {
  const uint64_t l = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t bits = 0;
  if (l < n_loci) {
    const uint32_t cand = sel[l];
    for (int k = 0; k < n_pop; ++k) if (((cand >> k) & 1u) && mark[(uint64_t)k * n_loci + l]) bits |= 1u << k;
    sel[l] = (uint8_t)bits;
  }
  for (int k = 0; k < n_pop; ++k) {
    const uint32_t n = __popc(__ballot_sync(kFull, (bits >> k) & 1u));
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(&counts[k], (unsigned long long)n);
  }
}

}  // namespace kgl
