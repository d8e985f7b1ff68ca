// count_moments.cuh -- K1+K2: the fused streaming pass over the loci-major genotype matrix.
//
// One read of the matrix (16 B per 64 genomes per locus, coalesced 128-bit loads) produces
//   * per-locus allele counts  {n0,n1,n2,n3}            -> VariantDBVariant::summaryByVariant (kgl_variant_db_variant.cpp:126)
//   * per-genome class counts over the SELECTED loci      -> generateFrequencies' count block (kga_analysis_inbreed_freq.cpp:559-577)
//     (or over all loci in `raw` mode                     -> VariantDBVariant::summaryByGenome, :180)
//   * the sparse corrections to the expected class-frequency sums for dropped loci (freq.cpp:532-539 and the :462 no-match case)
//
// Mapping: a thread owns one 128-bit unit column (64 genomes) and walks down the rows of its CTA's row range, eight rows
// per step (eight independent LDG.128 in flight). Per-locus counts are horizontal popcounts reduced with REDUX over the
// warp segment that shares a row. Per-genome counts are kept as bit-sliced vertical counters fed by a Harley-Seal
// carry-save tree (two LOP3 per full adder), so no per-genotype instruction is ever issued: ~80 thread-instructions per
// 64 genotypes, under the ~90 the HBM roofline allows ("instruction budget").
#pragma once
#include "common.cuh"

namespace kgl {

constexpr int kCountThreads = 256;
constexpr int kCountUnroll = 8;
constexpr int kCountMaxTY = 32;
constexpr int kLevels = 10;            // vertical counters hold up to 1023 rows per thread
constexpr int kMaxRowsPerThread = (1 << kLevels) - 8;

struct CountParams {
  const uint4* packed;       // [n_loci][units]
  uint64_t units;            // 128-bit units per row
  uint64_t n_loci;
  uint64_t n_genomes;
  uint32_t rows_per_cta;     // multiple of 8*ty
  int sw, ty;                // slice width (units per CTA row) and rows in flight
  const uint16_t* flags16;   // [n_loci] low byte: selected&valid per pop, high byte: ... and q <= 0.01 (rare-q rows)
  const uint8_t* unit_pop;   // [units] super-population of the unit's genomes, 0xFF = mixed
  const uint64_t* popmask;   // [n_pop][units] genomes of population k inside the unit
  const uint8_t* superpop;   // [n_genomes]
  const float* af;           // [n_pop][n_loci]
  int n_pop;
  int raw;                   // 1: count over all loci, no selection / corrections (allele_count)
  int multi_slice;           // gridDim.y > 1: per-locus counts are combined with global atomics
  uint32_t* locus_counts;    // [n_loci][4] or null
  uint32_t* planes;          // [gridDim.x*ty][units][3][2][kLevels] bit-sliced per-genome counters, or null
  double* ecorr;             // [n_genomes_padded][2]: sum over dropped loci of e_majHom, e_minHom
  uint32_t* nz_rare;         // [n_genomes_padded]: non-reference cells in rare-q rows
};

// Eight bit-plane words into a bit-sliced counter (levels 0,1,2 via carry-save, then a ripple into the high levels).
__device__ __forceinline__ void hs_add8(uint32_t (&c)[kLevels], const uint32_t (&x)[kCountUnroll]) {
  uint32_t t2a, t2b, t4a, t4b, t8;
  csa(t2a, c[0], c[0], x[0], x[1]);
  csa(t2b, c[0], c[0], x[2], x[3]);
  csa(t4a, c[1], c[1], t2a, t2b);
  csa(t2a, c[0], c[0], x[4], x[5]);
  csa(t2b, c[0], c[0], x[6], x[7]);
  csa(t4b, c[1], c[1], t2a, t2b);
  csa(t8, c[2], c[2], t4a, t4b);
  uint32_t carry = t8;
#pragma unroll
  for (int lv = 3; lv < kLevels; ++lv) {
    const uint32_t t = c[lv] & carry;
    c[lv] ^= carry;
    carry = t;
  }
}

// Rare events of one unit-row (kept out of line): dropped cells need their class frequencies subtracted from the dense
// totals, and rows whose major allele is rare (q <= 0.01) turn every hom-ref genome into a dropped locus.
__device__ __noinline__ void count_events(const CountParams& P, uint64_t row, uint64_t unit, uint64_t lo, uint64_t hi,
                                          uint64_t sel_mask, uint64_t rare_mask) {
  uint64_t dropped = (lo & hi & sel_mask) | (~(lo | hi) & rare_mask);
  while (dropped) {
    const int b = __ffsll((long long)dropped) - 1;
    dropped &= dropped - 1;
    const uint64_t g = unit * 64 + b;
    const int k = P.superpop[g];
    const LocusFreq f = locus_freq(P.af[(uint64_t)k * P.n_loci + row]);
    double a, h, c;
    class_freqs(f.p, a, h, c);
    atomicAdd(&P.ecorr[g * 2 + 0], a);
    atomicAdd(&P.ecorr[g * 2 + 1], c);
  }
  uint64_t nz = (lo | hi) & rare_mask;
  while (nz) {
    const int b = __ffsll((long long)nz) - 1;
    nz &= nz - 1;
    atomicAdd(&P.nz_rare[unit * 64 + b], 1u);
  }
}

template <bool MIXED>
__global__ void __launch_bounds__(kCountThreads, 2)
k_count_moments(const CountParams P) {
  __shared__ uint32_t s_rowcnt[2][kCountUnroll * kCountMaxTY][2];

  const int tid = threadIdx.x, lane = tid & 31;
  const int tx = tid % P.sw, ty = tid / P.sw;
  const uint64_t unit = (uint64_t)blockIdx.y * P.sw + tx;
  const bool active = (ty < P.ty) && (unit < P.units);
  const unsigned seg_mask = __match_any_sync(kFull, active ? ty : 0xFFFF);
  const bool leader = active && (lane == __ffs(seg_mask) - 1);

  for (int i = tid; i < 2 * kCountUnroll * kCountMaxTY * 2; i += kCountThreads) (&s_rowcnt[0][0][0])[i] = 0;
  __syncthreads();

  const int upop = active ? P.unit_pop[unit] : 0;
  uint64_t pm[kMaxPop];
  if (MIXED) {
#pragma unroll
    for (int k = 0; k < kMaxPop; ++k) pm[k] = (active && k < P.n_pop) ? P.popmask[(uint64_t)k * P.units + unit] : 0ull;
  }

  uint32_t cnt[3][2][kLevels];
#pragma unroll
  for (int p = 0; p < 3; ++p)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int lv = 0; lv < kLevels; ++lv) cnt[p][h][lv] = 0;

  const uint64_t row0 = (uint64_t)blockIdx.x * P.rows_per_cta;
  const uint64_t row_end = min(row0 + (uint64_t)P.rows_per_cta, P.n_loci);
  const int group_rows = kCountUnroll * P.ty;
  int buf = 0;

  for (uint64_t base = row0; base < row_end; base += group_rows, buf ^= 1) {
    uint4 v[kCountUnroll];
    uint32_t fl[kCountUnroll];
#pragma unroll
    for (int i = 0; i < kCountUnroll; ++i) {
      const uint64_t r = base + (uint64_t)i * P.ty + ty;
      const bool ok = active && r < row_end;
      v[i] = ok ? ld_stream_u4(P.packed + r * P.units + unit) : make_uint4(0, 0, 0, 0);
      fl[i] = (ok && !P.raw) ? (uint32_t)P.flags16[r] : 0u;
    }

    // ---- per-locus counts: horizontal popcounts, REDUX over the lanes that share the row ----
    if (P.locus_counts != nullptr) {
#pragma unroll
      for (int i = 0; i < kCountUnroll; ++i) {
        const uint32_t c1 = __popc(v[i].x & ~v[i].z) + __popc(v[i].y & ~v[i].w);
        const uint32_t c2 = __popc(v[i].z & ~v[i].x) + __popc(v[i].w & ~v[i].y);
        const uint32_t c3 = __popc(v[i].x & v[i].z) + __popc(v[i].y & v[i].w);
        const uint32_t s12 = __reduce_add_sync(seg_mask, c1 | (c2 << 16));   // a slice holds <= 16384 genomes: no carry
        const uint32_t s3 = __reduce_add_sync(seg_mask, c3);
        if (leader) {
          atomicAdd(&s_rowcnt[buf][i * P.ty + ty][0], s12);
          atomicAdd(&s_rowcnt[buf][i * P.ty + ty][1], s3);
        }
      }
    }

    // ---- per-genome counters over the selected rows ----
    if (P.planes != nullptr) {
      uint32_t mlo[kCountUnroll], mhi[kCountUnroll];
#pragma unroll
      for (int i = 0; i < kCountUnroll; ++i) {
        uint64_t m;
        if (P.raw) m = ~0ull;
        else if (!MIXED) m = ((fl[i] >> upop) & 1u) ? ~0ull : 0ull;
        else {
          m = 0;
#pragma unroll
          for (int k = 0; k < kMaxPop; ++k) m |= ((fl[i] >> k) & 1u) ? pm[k] : 0ull;
        }
        mlo[i] = (uint32_t)m; mhi[i] = (uint32_t)(m >> 32);
      }
      uint32_t x[kCountUnroll];
      // het plane (code 1): lo & ~hi
#pragma unroll
      for (int i = 0; i < kCountUnroll; ++i) x[i] = v[i].x & ~v[i].z & mlo[i];
      hs_add8(cnt[0][0], x);
#pragma unroll
      for (int i = 0; i < kCountUnroll; ++i) x[i] = v[i].y & ~v[i].w & mhi[i];
      hs_add8(cnt[0][1], x);
      // hom-alt plane (code 2): hi & ~lo
#pragma unroll
      for (int i = 0; i < kCountUnroll; ++i) x[i] = v[i].z & ~v[i].x & mlo[i];
      hs_add8(cnt[1][0], x);
#pragma unroll
      for (int i = 0; i < kCountUnroll; ++i) x[i] = v[i].w & ~v[i].y & mhi[i];
      hs_add8(cnt[1][1], x);
      // dropped plane (code 3): lo & hi
#pragma unroll
      for (int i = 0; i < kCountUnroll; ++i) x[i] = v[i].x & v[i].z & mlo[i];
      hs_add8(cnt[2][0], x);
#pragma unroll
      for (int i = 0; i < kCountUnroll; ++i) x[i] = v[i].y & v[i].w & mhi[i];
      hs_add8(cnt[2][1], x);

      // ---- rare events: dropped cells inside selected rows, and rare-major rows ----
      if (!P.raw) {
#pragma unroll
        for (int i = 0; i < kCountUnroll; ++i) {
          const uint32_t rq = fl[i] >> 8;
          bool rare;
          if (!MIXED) rare = (rq >> upop) & 1u; else rare = rq != 0;
          const uint32_t mis = (v[i].x & v[i].z & mlo[i]) | (v[i].y & v[i].w & mhi[i]);
          if (mis != 0 || rare) {
            uint64_t rare_mask = 0;
            if (rare) {
              if (!MIXED) rare_mask = P.popmask[(uint64_t)upop * P.units + unit];
              else {
#pragma unroll
                for (int k = 0; k < kMaxPop; ++k) rare_mask |= ((rq >> k) & 1u) ? pm[k] : 0ull;
              }
            }
            const uint64_t lo = (uint64_t)v[i].x | ((uint64_t)v[i].y << 32);
            const uint64_t hi = (uint64_t)v[i].z | ((uint64_t)v[i].w << 32);
            count_events(P, base + (uint64_t)i * P.ty + ty, unit, lo, hi,
                         (uint64_t)mlo[i] | ((uint64_t)mhi[i] << 32), rare_mask);
          }
        }
      }
    }

    // ---- publish the per-locus counts of this row group ----
    if (P.locus_counts != nullptr) {
      __syncthreads();
      if (tid < group_rows) {
        const uint64_t r = base + tid;
        const uint32_t s12 = s_rowcnt[buf][tid][0], s3 = s_rowcnt[buf][tid][1];
        s_rowcnt[buf][tid][0] = 0;
        s_rowcnt[buf][tid][1] = 0;
        if (r < row_end) {
          const uint32_t c1 = s12 & 0xFFFFu, c2 = s12 >> 16;
          uint32_t* out = P.locus_counts + r * 4;
          if (!P.multi_slice) {
            *reinterpret_cast<uint4*>(out) = make_uint4((uint32_t)P.n_genomes - c1 - c2 - s3, c1, c2, s3);
          } else {
            atomicAdd(out + 1, c1); atomicAdd(out + 2, c2); atomicAdd(out + 3, s3);
          }
        }
      }
    }
  }

  // ---- epilogue: bit-sliced counters of this thread's (row-range, unit) ----
  if (P.planes != nullptr && active) {
    uint32_t* out = P.planes + (((uint64_t)blockIdx.x * P.ty + ty) * P.units + unit) * (3 * 2 * kLevels);
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int lv = 0; lv < kLevels; ++lv) out[(p * 2 + h) * kLevels + lv] = cnt[p][h][lv];
  }
}

// multi-slice only: n0 = N - n1 - n2 - n3
__global__ void k_fix_locus_n0(uint32_t* locus_counts, uint64_t n_loci, uint32_t n_genomes) {
  const uint64_t l = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (l < n_loci) {
    uint32_t* c = locus_counts + l * 4;
    c[0] = n_genomes - c[1] - c[2] - c[3];
  }
}

// Expand the bit-sliced counters: gcounts[g][p] += sum over virtual chunks. Thread = (genome, group of virtual chunks).
constexpr int kExpandChunkGroup = 16;
__global__ void __launch_bounds__(256)
k_expand_counts(const uint32_t* __restrict__ planes, uint64_t n_vchunks, uint64_t units, uint64_t n_genomes_padded,
                uint32_t* __restrict__ gcounts /* [n_genomes_padded][4] (col 3 unused pad) */) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_genomes_padded) return;
  const uint64_t unit = g >> 6;
  const int h = (int)((g >> 5) & 1), bit = (int)(g & 31);
  const uint64_t vc0 = (uint64_t)blockIdx.y * kExpandChunkGroup;
  const uint64_t vc1 = min(vc0 + (uint64_t)kExpandChunkGroup, n_vchunks);
  uint32_t acc[3] = {0, 0, 0};
  for (uint64_t vc = vc0; vc < vc1; ++vc) {
    const uint32_t* base = planes + (vc * units + unit) * (3 * 2 * kLevels);
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      uint32_t c = 0;
#pragma unroll
      for (int lv = 0; lv < kLevels; ++lv) c |= ((base[(p * 2 + h) * kLevels + lv] >> bit) & 1u) << lv;
      acc[p] += c;
    }
  }
#pragma unroll
  for (int p = 0; p < 3; ++p) if (acc[p]) atomicAdd(&gcounts[g * 4 + p], acc[p]);
}

}  // namespace kgl
