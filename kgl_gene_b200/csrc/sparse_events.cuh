// sparse_events.cuh -- the sparse side of the fused pass: code-3 ("dropped") cells and rare-major rows.
//
// The streaming kernel (stream_count.cuh) only counts set lo / hi bits. Two kinds of cells need more than a count, and
// both are sparse in real data:
//   * code-3 cells (first variant not in the AF list, >2 variants, missing ...; SURVEY flattener contract): they are not
//     classified, so the locus' class frequencies leave the genome's expected sums (kga_analysis_inbreed_freq.cpp:462-543)
//   * rows whose major allele is rare for a population (q <= 0.01): a hom-ref genome is dropped there (freq.cpp:532-539)
// Dropped cells are indexed ONCE per uploaded matrix (k_dropped_count / k_dropped_index + a key sort, part of the upload:
// the index is the "side list" of the flattener contract in device form, genome-major and row-sorted); every pass then
// visits only the indexed cells, one warp per genome, without atomics and in a fixed summation order. Populations with
// too many code-3 cells for an index fall back to k_dropped_scan, which re-reads the matrix.
#pragma once
#include "common.cuh"

namespace kgl {

typedef unsigned long long DroppedKey;     // (genome << 32) | row

// ---- index construction (upload time) ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_dropped_count(const uint4* __restrict__ packed, uint64_t n_cells128, unsigned long long* __restrict__ total) {
  uint32_t c = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cells128; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint4 v = ld_stream_u4(packed + i);
    c += __popc(v.x & v.z) + __popc(v.y & v.w);
  }
  c = __reduce_add_sync(kFull, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(total, (unsigned long long)c);
}

// Appends one key per code-3 cell of a range of units (any order; sorted afterwards). The cursor keeps counting past the
// capacity, so the caller learns the true total and can retry with an exact allocation.
__global__ void __launch_bounds__(256)
k_dropped_index(const uint4* __restrict__ packed /* first unit of the range */, uint64_t n_cells128, uint64_t cell0 /* its global unit index */,
                uint32_t units, unsigned long long* __restrict__ cursor, DroppedKey* __restrict__ cells, uint64_t capacity) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (uint64_t base = i0 - lane; base < n_cells128; base += stride) {   // whole warps iterate together
    const uint64_t i = base + lane;
    uint64_t both = 0;
    if (i < n_cells128) {
      const uint4 v = ld_stream_u4(packed + i);
      both = (uint64_t)(v.x & v.z) | ((uint64_t)(v.y & v.w) << 32);
    }
    const uint32_t n = (uint32_t)__popcll(both);
    if (__any_sync(kFull, n != 0)) {
      // warp-level exclusive scan of n, one atomic per warp
      uint32_t incl = n;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, incl, o);
        if ((int)lane >= o) incl += t;
      }
      const uint32_t warp_total = __shfl_sync(kFull, incl, 31);
      unsigned long long start = 0;
      if (lane == 0) start = atomicAdd(cursor, (unsigned long long)warp_total);
      start = __shfl_sync(kFull, start, 0);
      uint64_t o = start + (incl - n);
      const uint32_t row = (uint32_t)((cell0 + i) / units), unit = (uint32_t)((cell0 + i) % units);
      while (both) {
        const int b = __ffsll((long long)both) - 1;
        both &= both - 1;
        if (o < capacity) cells[o] = ((DroppedKey)(unit * 64 + (uint32_t)b) << 32) | row;
        ++o;
      }
    }
  }
}

// ---- per pass ---------------------------------------------------------------------------------------------------------------
// Accumulators per genome: n3[g] (code-3 cells in rows selected for the genome), nz_rare[g] (non-reference cells in
// rare-major rows), ecorr[g][2] (sum over the genome's dropped loci of e_majHom, e_minHom).
struct SparseOut {
  uint32_t* n3;
  uint32_t* nz_rare;
  double* ecorr;
};

// seg[g] = first key of genome g in the sorted key array (g = 0 .. n_genomes_padded inclusive).
__global__ void __launch_bounds__(256)
k_dropped_segments(const DroppedKey* __restrict__ keys, uint64_t n_keys, uint64_t n_genomes_padded, uint64_t* __restrict__ seg) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g > n_genomes_padded) return;
  const DroppedKey want = (DroppedKey)g << 32;
  uint64_t lo = 0, hi = n_keys;
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if (keys[mid] < want) lo = mid + 1; else hi = mid;
  }
  seg[g] = lo;
}

// Four genomes per block, two warps per genome over the genome's row-sorted dropped cells; the partial sums are combined
// in a fixed order, so the result does not depend on scheduling. flags16 == null: raw mode (allele_count) -- every
// code-3 cell counts, no frequency corrections. Writes n3[g]; adds the class frequencies of the selected dropped loci
// to ecorr[g].
constexpr int kDropGenomesPerBlock = 4;          // 64 threads (two warps) per genome: N / 4 blocks fit the machine in one wave
__device__ __forceinline__ void
dropped_apply_block(uint32_t block, const DroppedKey* __restrict__ keys, const uint64_t* __restrict__ seg, uint64_t n_genomes,
                    const uint16_t* __restrict__ flags16, const uint32_t* __restrict__ all_selected,
                    const uint8_t* __restrict__ superpop, const float* __restrict__ af, uint64_t n_loci, SparseOut out) {
  constexpr int GPB = kDropGenomesPerBlock, TPG = 256 / GPB, WPG = TPG / 32;
  __shared__ double s_sum[GPB][WPG][2];
  __shared__ uint32_t s_n[GPB][WPG];
  const uint32_t tid = threadIdx.x, lane = tid & 31, gi = tid / TPG, wi = (tid % TPG) >> 5, tl = tid % TPG;
  const uint64_t g = (uint64_t)block * GPB + gi;
  const bool raw = flags16 == nullptr;
  uint32_t n = 0;
  double sa = 0.0, sm = 0.0;
  if (g < n_genomes) {
    const uint64_t i0 = seg[g], i1 = seg[g + 1];
    if (raw) {
      n = (tl == 0) ? (uint32_t)(i1 - i0) : 0u;
    } else {
      const int k = superpop[g];
      const float* afk = af + (uint64_t)k * n_loci;
      const bool all_sel = all_selected[0] != 0;       // every row selected & valid for every population: skip the flag gather
      // four cells per thread and trip: the key, flag and frequency gathers of a trip are independent and overlap
      for (uint64_t i = i0 + tl; i < i1; i += 4 * TPG) {
        uint32_t row[4], fl[4];
        float fa[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) row[j] = (i + TPG * j < i1) ? (uint32_t)keys[i + TPG * j] : 0xFFFFFFFFu;
#pragma unroll
        for (int j = 0; j < 4; ++j) fl[j] = (row[j] == 0xFFFFFFFFu) ? 0u : (all_sel ? 0xFFu : (uint32_t)flags16[row[j]]);
#pragma unroll
        for (int j = 0; j < 4; ++j) fa[j] = ((fl[j] >> k) & 1u) ? afk[row[j]] : 0.0f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if ((fl[j] >> k) & 1u) {
            double a, h, m;
            class_freqs(locus_freq(fa[j]).p, a, h, m);
            ++n; sa += a; sm += m;
          }
        }
      }
    }
  }
  n = __reduce_add_sync(kFull, n);
  sa = warp_sum(sa); sm = warp_sum(sm);
  if (lane == 0) { s_n[gi][wi] = n; s_sum[gi][wi][0] = sa; s_sum[gi][wi][1] = sm; }
  __syncthreads();
  if (tl == 0 && g < n_genomes) {
    uint32_t nn = 0;
    double ta = 0.0, tm = 0.0;
#pragma unroll
    for (int w = 0; w < WPG; ++w) { nn += s_n[gi][w]; ta += s_sum[gi][w][0]; tm += s_sum[gi][w][1]; }   // fixed order
    out.n3[g] = nn;
    if (nn && !raw) {
      atomicAdd(&out.ecorr[g * 2 + 0], ta);
      atomicAdd(&out.ecorr[g * 2 + 1], tm);
    }
  }
}

__global__ void __launch_bounds__(256)
k_dropped_apply(const DroppedKey* __restrict__ keys, const uint64_t* __restrict__ seg, uint64_t n_genomes,
                const uint16_t* __restrict__ flags16, const uint32_t* __restrict__ all_selected,
                const uint8_t* __restrict__ superpop, const float* __restrict__ af, uint64_t n_loci, SparseOut out) {
  dropped_apply_block(blockIdx.x, keys, seg, n_genomes, flags16, all_selected, superpop, af, n_loci, out);
}

// Fallback without an index: one thread per 128-bit unit-row.
__global__ void __launch_bounds__(256)
k_dropped_scan(const uint4* __restrict__ packed, uint64_t n_cells128, uint32_t units, const uint16_t* __restrict__ flags16,
               const uint8_t* __restrict__ superpop, const float* __restrict__ af, uint64_t n_loci, SparseOut out) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cells128; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint4 v = ld_stream_u4(packed + i);
    uint64_t both = (uint64_t)(v.x & v.z) | ((uint64_t)(v.y & v.w) << 32);
    if (both == 0) continue;
    const uint32_t row = (uint32_t)(i / units), unit = (uint32_t)(i % units);
    const uint32_t fl = flags16 ? flags16[row] : 0u;
    while (both) {
      const int b = __ffsll((long long)both) - 1;
      both &= both - 1;
      const uint32_t g = unit * 64 + (uint32_t)b;
      if (flags16 == nullptr) { atomicAdd(&out.n3[g], 1u); continue; }
      const int k = superpop[g];
      if (!((fl >> k) & 1u)) continue;
      atomicAdd(&out.n3[g], 1u);
      double a, h, m;
      class_freqs(locus_freq(af[(uint64_t)k * n_loci + row]).p, a, h, m);
      atomicAdd(&out.ecorr[(uint64_t)g * 2 + 0], a);
      atomicAdd(&out.ecorr[(uint64_t)g * 2 + 1], m);
    }
  }
}

// Rare-major rows (flags16 high byte != 0; listed by k_locus_prepare). One thread per (listed row, unit): for every genome
// whose population has q <= 0.01 at this row: hom-ref -> the locus is dropped for it; else nz_rare++.
__device__ __forceinline__ void
rare_rows_block(uint32_t block, uint32_t n_blocks, const uint32_t* __restrict__ rare_rows, const uint32_t* __restrict__ n_rare,
                const uint4* __restrict__ packed, uint32_t units, uint32_t n_genomes, const uint16_t* __restrict__ flags16,
                const uint64_t* __restrict__ popmask, const float* __restrict__ af, uint64_t n_loci, int n_pop, SparseOut out) {
  const uint64_t total = (uint64_t)(*n_rare) * units;
  for (uint64_t t = (uint64_t)block * blockDim.x + threadIdx.x; t < total; t += (uint64_t)n_blocks * blockDim.x) {
    const uint32_t row = rare_rows[t / units], u = (uint32_t)(t % units);
    const uint32_t rq = (uint32_t)flags16[row] >> 8;
    const uint4 v = packed[(uint64_t)row * units + u];
    const uint64_t lo = (uint64_t)v.x | ((uint64_t)v.y << 32), hi = (uint64_t)v.z | ((uint64_t)v.w << 32);
    for (int k = 0; k < n_pop; ++k) {
      if (!((rq >> k) & 1u)) continue;
      const uint64_t mask = popmask[(uint64_t)k * units + u];
      if (mask == 0) continue;
      double a, h, m;
      class_freqs(locus_freq(af[(uint64_t)k * n_loci + row]).p, a, h, m);
      uint64_t homref = ~(lo | hi) & mask;
      uint64_t nonref = (lo | hi) & mask;
      while (homref) {
        const int b = __ffsll((long long)homref) - 1;
        homref &= homref - 1;
        const uint64_t g = (uint64_t)u * 64 + b;
        atomicAdd(&out.ecorr[g * 2 + 0], a);
        atomicAdd(&out.ecorr[g * 2 + 1], m);
      }
      while (nonref) {
        const int b = __ffsll((long long)nonref) - 1;
        nonref &= nonref - 1;
        atomicAdd(&out.nz_rare[(uint64_t)u * 64 + b], 1u);
      }
    }
  }
}

__global__ void __launch_bounds__(256)
k_rare_rows(const uint32_t* __restrict__ rare_rows, const uint32_t* __restrict__ n_rare, const uint4* __restrict__ packed,
            uint32_t units, uint32_t n_genomes, const uint16_t* __restrict__ flags16, const uint64_t* __restrict__ popmask,
            const float* __restrict__ af, uint64_t n_loci, int n_pop, SparseOut out) {
  rare_rows_block(blockIdx.x, gridDim.x, rare_rows, n_rare, packed, units, n_genomes, flags16, popmask, af, n_loci, n_pop, out);
}

}  // namespace kgl
