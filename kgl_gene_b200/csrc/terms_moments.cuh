// terms_moments.cuh -- the iterative estimators (HallME, log-likelihood root search) without a pass over the genotype matrix per
// sweep: per-genome moment tables of the homozygous terms, built ONCE per selection.
//
// Every homozygous term of processHallME (calc.cpp:257-285) and of d/df logLikelihood (calc.cpp:94-129) is a function of
//   t(f) = 1/(f + r),  r = a/(1 - a),  a = the frequency of the homozygous allele (q for a hom-ref cell, p for a hom-alt cell):
//   HallME      f/(f + (1-f) a) = f (1 + (1-f) t)          ->  f (n_hom + (1-f) S1)
//   Newton      (1-a)/(a + f (1-a)) = t,  its square        ->  S1 = sum t,  S2 = sum t^2
// and f differs from genome to genome, so the sums cannot be shared between genomes -- but they can be shared between SWEEPS. r is
// binned (32 bins per octave: bin = exponent and five mantissa bits of the double, centre rc, half width w = 2^e/64); inside a
// bin  1/(f + rc + u) = t_c sum_j (-u t_c)^j,  t_c = 1/(f + rc),  |u t_c| <= w/(f + rc) <= 1/64 for f >= 0 (1/48 at the left end
// of the domain below), so the six moments  m_j = sum over the genome's homozygous cells in the bin of (u/w)^j  give the bin's
// S1 and S2 to < 1e-10 relative (six terms; typically 1e-12). A sweep is then bins x 6 FMAs per genome instead of L reciprocals.
//
// Building the moments is dense-minus-sparse, as the rest of the path: at every locus one homozygous class is COMMON (the
// allele with frequency >= 1/2) and one RARE. The population's moments over all selected loci (k_mom_dense) stand for "every
// genome is common-homozygous everywhere"; one pass over the matrix subtracts, per genome, the cells that are not, and adds the
// rare homozygous cells to their own bins -- on the tensor cores (k_mom_mma, terms_moments_mma.cuh) or, as the fallback and the
// cross-check, on the CUDA cores (k_mom_build below: lane = genome, predicated 64-bit adds). Loci are walked in the order of the population's frequency (a radix
// sort of the float bits), so both bins change monotonically and the accumulators live in registers between bin changes. All
// moments are 64-bit FIXED-POINT integers (normalised powers in [-1,1] x 2^s), added with integer atomics: dense minus sparse
// cancels exactly, a bin the genome has no cell in is exactly zero, and the result does not depend on the schedule.
//
// Left of zero the poles of the rare cells come close (the feasible region of the likelihood ends at f = -r of the genome's
// rarest homozygous allele), so for f < 0 the bins of the octaves below 4|f| are not used: the rare cells themselves are kept as
// a per-genome list of r (k_mom_unit_fill / k_mom_fill; ascending hom-alt part, descending hom-ref part) and the few below the
// threshold are evaluated exactly. The domain of the tables is f >= kMomValidMin; a genome left of it takes the exact cell-by-cell kernel
// (k_genome_terms), as do genomes within rounding of their feasible end (k_newton_reduce, unchanged).
#pragma once
#include "common.cuh"
#include "terms_fast.cuh"
#include "misc_kernels.cuh"

namespace kgl {

constexpr int kMomJ = 6;                               // moments per bin (j = 0 is the cell count)
constexpr int kMomSubBits = 5, kMomSub = 1 << kMomSubBits;
constexpr int kMomExpLo = -40, kMomExpHi = 40;         // octaves [2^-40, 2^40): AF down to 1e-12
constexpr int kMomBinsMax = (kMomExpHi - kMomExpLo) * kMomSub;     // global bin kMomBinsMax = the a == 1 class (r = +inf, t = 0)
constexpr double kMomValidMin = -0.2;                  // f below this: exact fallback (common bins need f >= -2^e / 4, e >= 0)
constexpr int kMomTile = 256;                          // genomes per CTA
constexpr int kMomStep = 128;                          // loci per table step
constexpr uint32_t kMomChunk = 4096;                   // loci per CTA

// Global bin of r: [0, kMomBinsMax) in range, kMomBinsMax for +inf, -1 out of range (the run then uses the exact kernels).
__device__ __forceinline__ int mom_bin(double r) {
  const long long bits = __double_as_longlong(r);
  const int e = (int)((bits >> 52) & 0x7FF) - 1023;
  if (e == 1024) return ((bits & 0xFFFFFFFFFFFFFll) == 0 && bits > 0) ? kMomBinsMax : -1;
  if (e < kMomExpLo || e >= kMomExpHi) return -1;
  return (e - kMomExpLo) * kMomSub + (int)((bits >> (52 - kMomSubBits)) & (kMomSub - 1));
}
__device__ __forceinline__ void mom_geometry(int b, double& rc, double& w) {
  const int e = b / kMomSub + kMomExpLo, sub = b % kMomSub;
  w = ldexp(1.0, e - kMomSubBits - 1);
  rc = ldexp(1.0 + (sub + 0.5) / kMomSub, e);
}

// The two homozygous classes of one (population, locus). Same constants as fast_constants<FAST_NEWTON> / the LIMITS table.
struct MomClass {
  double r_common, r_rare;     // a/(1-a); +inf when a == 1 (t = 0: the cell only counts in n_hom)
  int common_code, rare_code;  // genotype code of the class (0 hom-ref, 2 hom-alt); -1: its cells are not terms
  double c_het;                // 2 p q of a heterozygous cell (calc.cpp:117)
};
__device__ __forceinline__ MomClass mom_classify(float af, bool unphased) {
  const LocusFreq lf = locus_freq(af);
  const double p = lf.p, q = lf.q;
  const bool ref_in = q > kMinMajorFreq;                  // freq.cpp:532
  const bool alt_in = !unphased;                          // Q6: an unphased hom-alt pair is a heterozygous term
  const double uq = __dsub_rn(1.0, q), up = __dsub_rn(1.0, p);
  const double inf = __longlong_as_double(0x7FF0000000000000ll);
  const double r0 = uq > 0.0 ? __ddiv_rn(q, uq) : inf, r2 = up > 0.0 ? __ddiv_rn(p, up) : inf;
  const bool ref_common = q >= p;
  MomClass m;
  m.common_code = ref_common ? (ref_in ? 0 : -1) : (alt_in ? 2 : -1);
  m.rare_code = ref_common ? (alt_in ? 2 : -1) : (ref_in ? 0 : -1);
  m.r_common = ref_common ? r0 : r2;
  m.r_rare = ref_common ? r2 : r0;
  m.c_het = __dmul_rn(__dmul_rn(2.0, p), q);
  return m;
}

// Fixed-point normalised powers ((r - rc)/w)^j 2^s, j = 1..5, of a class in bin b.
__device__ __forceinline__ void mom_powers(double r, int b, double scale, long long (&u)[kMomJ - 1]) {
  if (b < 0 || b >= kMomBinsMax) {
#pragma unroll
    for (int j = 0; j < kMomJ - 1; ++j) u[j] = 0;
    return;
  }
  double rc, w;
  mom_geometry(b, rc, w);
  const double v = (r - rc) / w;
  double pw = v;
#pragma unroll
  for (int j = 0; j < kMomJ - 1; ++j) { u[j] = __double2ll_rn(pw * scale); pw *= v; }
}

// ---- sort keys ---------------------------------------------------------------------------------------------------------------
// One key per (population, row of the selection window): population << 32 | float bits of p (selected rows), all ones otherwise.
// stats = {smallest bin, largest bin, classes out of the bin range}.
__global__ void __launch_bounds__(256)
k_mom_keys(const uint32_t* __restrict__ selw, const float* __restrict__ af, uint64_t n_loci, uint64_t n_words, uint64_t row_lo,
           uint64_t n_rows, int unphased, uint64_t* __restrict__ keys, uint32_t* __restrict__ rows, int* __restrict__ stats,
           unsigned long long* __restrict__ pop_cmin /* [kMaxPop] smallest 2 p q of a selected locus, as the bits of the double */) {
  __shared__ int s_lo, s_hi, s_bad;
  __shared__ unsigned long long s_cmin;
  if (threadIdx.x == 0) { s_lo = kMomBinsMax; s_hi = -1; s_bad = 0; s_cmin = ~0ull; }
  __syncthreads();
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int kk = blockIdx.y;
  int lo = kMomBinsMax, hi = -1, bad = 0;
  unsigned long long cmin = ~0ull;
  if (i < n_rows) {
    const uint64_t l = row_lo + i;
    const bool sel = l < n_loci && ((selw[(uint64_t)kk * n_words + (l >> 5)] >> (l & 31)) & 1u);
    uint64_t key = ~0ull;
    if (sel) {
      const float a = af[(uint64_t)kk * n_loci + l];
      const LocusFreq lf = locus_freq(a);
      key = ((uint64_t)kk << 32) | __float_as_uint((float)lf.p);
      const MomClass m = mom_classify(a, unphased != 0);
      cmin = (unsigned long long)__double_as_longlong(m.c_het);          // non-negative doubles order like their bits
      if (m.common_code >= 0) { const int b = mom_bin(m.r_common); if (b < 0) ++bad; else if (b < kMomBinsMax) { lo = min(lo, b); hi = max(hi, b); } }
      if (m.rare_code >= 0) { const int b = mom_bin(m.r_rare); if (b < 0 || b >= kMomBinsMax) ++bad; else { lo = min(lo, b); hi = max(hi, b); } }
    }
    keys[(uint64_t)kk * n_rows + i] = key;
    rows[(uint64_t)kk * n_rows + i] = (uint32_t)l;
  }
  lo = __reduce_min_sync(kFull, lo); hi = __reduce_max_sync(kFull, hi); bad = __reduce_add_sync(kFull, bad);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(kFull, cmin, o); cmin = t < cmin ? t : cmin; }
  if ((threadIdx.x & 31) == 0) { atomicMin(&s_lo, lo); atomicMax(&s_hi, hi); if (bad) atomicAdd(&s_bad, bad); atomicMin(&s_cmin, cmin); }
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicMin(&stats[0], s_lo); atomicMax(&stats[1], s_hi); if (s_bad) atomicAdd(&stats[2], s_bad);
    atomicMin(&pop_cmin[kk], s_cmin);
  }
}

// pop_begin[k] = first sorted item of population k (k = 0 .. n_pop; the unselected items sort behind the last population).
__global__ void k_mom_ranges(const uint64_t* __restrict__ keys, uint64_t n, int n_pop, uint32_t* __restrict__ pop_begin) {
  const int k = threadIdx.x;
  if (k > n_pop) return;
  const uint64_t mask = (1ull << 35) - 1, want = (uint64_t)k << 32;
  uint64_t lo = 0, hi = n;
  while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if ((keys[mid] & mask) < want) lo = mid + 1; else hi = mid; }
  pop_begin[k] = (uint32_t)lo;
}

// ---- the population's moments: every selected locus, common class -----------------------------------------------------------
__global__ void __launch_bounds__(256)
k_mom_dense(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ rows, const uint32_t* __restrict__ pop_begin, int n_pop,
            const float* __restrict__ af, uint64_t n_loci, int unphased, int b_lo, int nbt, double scale, long long* __restrict__ pm) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t total = pop_begin[n_pop];
  int slot = -1;
  long long u[kMomJ - 1];
#pragma unroll
  for (int j = 0; j < kMomJ - 1; ++j) u[j] = 0;
  if (i < total) {
    const int kk = (int)(keys[i] >> 32);
    const MomClass m = mom_classify(af[(uint64_t)kk * n_loci + rows[i]], unphased != 0);
    if (m.common_code >= 0) {
      const int b = mom_bin(m.r_common);
      mom_powers(m.r_common, b, scale, u);
      slot = kk * nbt + (b >= kMomBinsMax ? nbt - 1 : b - b_lo);
    }
  }
  // a warp whose items share the bin adds once
  const int first = __shfl_sync(kFull, slot, 0);
  if (__all_sync(kFull, slot == first)) {
    if (first < 0) return;
#pragma unroll
    for (int j = 0; j < kMomJ - 1; ++j)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) u[j] += __shfl_xor_sync(kFull, u[j], o);
    if ((threadIdx.x & 31) == 0) {
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(pm) + (size_t)first * kMomJ;
      atomicAdd(dst, 32ull);
#pragma unroll
      for (int j = 0; j < kMomJ - 1; ++j) atomicAdd(dst + 1 + j, (unsigned long long)u[j]);
    }
    return;
  }
  if (slot < 0) return;
  unsigned long long* dst = reinterpret_cast<unsigned long long*>(pm) + (size_t)slot * kMomJ;
  atomicAdd(dst, 1ull);
#pragma unroll
  for (int j = 0; j < kMomJ - 1; ++j) atomicAdd(dst + 1 + j, (unsigned long long)u[j]);
}

// ---- the pass over the matrix ------------------------------------------------------------------------------------------------
struct MomParams {
  const uint4* packed; uint64_t units;               // loci-major matrix: row = locus, unit = {u64 lo, u64 hi} of 64 genomes
  const float* af; uint64_t n_loci; int n_pop, unphased;
  const uint8_t* superpop; uint64_t n_genomes, n_genomes_padded;
  const uint32_t* rows; const uint32_t* pop_begin;   // the populations' loci in frequency order
  uint32_t chunks_per_pop;                           // grid.x = n_pop * chunks_per_pop
  int b_lo, nbt; double scale;
  long long* mi;                                     // [n_genomes_padded][nbt][kMomJ] (zeroed): - not-common cells + rare cells
  uint32_t* cnt;                                     // [chunks_per_pop][n_genomes_padded]: rare hom-alt cells | rare hom-ref cells << 16
  // k_mom_fill
  const uint64_t* base; const uint2* offs; double* list;   // list[base[g] + ...] = r of the genome's rare homozygous cells
};

__device__ __forceinline__ void mom_flush(long long (&a)[kMomJ], long long* __restrict__ dst, int bin) {
  if (bin >= 0 && a[0] != 0) {
    unsigned long long* d = reinterpret_cast<unsigned long long*>(dst) + (size_t)bin * kMomJ;
#pragma unroll
    for (int j = 0; j < kMomJ; ++j) atomicAdd(d + j, (unsigned long long)a[j]);
  }
#pragma unroll
  for (int j = 0; j < kMomJ; ++j) a[j] = 0;
}

// Masks of the genotype classes over the 32 genomes of a warp: code 0, 1, 2 (3 = both planes set).
__device__ __forceinline__ uint32_t mom_code_mask(uint2 w, int code) {
  return code == 0 ? ~(w.x | w.y) : code == 2 ? (w.y & ~w.x) : code == 1 ? (w.x & ~w.y) : 0u;
}

// Masks over the 32 genomes of warp slice `k` of a tile's four units: cells that are not common-homozygous, rare-homozygous, heterozygous.
struct MomMasks { uint32_t nc, rare, het; };
__device__ __forceinline__ MomMasks mom_masks(uint32_t lo, uint32_t hi, int common_code, int rare_code) {
  const uint2 w = make_uint2(lo, hi);
  MomMasks m;
  m.nc = common_code >= 0 ? ~mom_code_mask(w, common_code) : 0u;
  m.rare = rare_code >= 0 ? mom_code_mask(w, rare_code) : 0u;
  m.het = mom_code_mask(w, 1);
  return m;
}

__global__ void __launch_bounds__(kMomTile)
k_mom_build(const MomParams P) {
  constexpr int kWarps = kMomTile / 32;
  __shared__ int2 s_hdr[kMomStep];                                  // {common bin, rare bin} (-1: the class has no terms)
  __shared__ uint2 s_m[kMomStep][kWarps];                           // per locus and warp: {not-common mask, rare mask}
  __shared__ __align__(16) long long s_cu[kMomStep][kMomJ], s_ru[kMomStep][kMomJ];   // {-1, -u1..-u5} and {+1, +u1..+u5}
  const int pop = blockIdx.x / P.chunks_per_pop, chunk = blockIdx.x % P.chunks_per_pop;
  const uint32_t begin = P.pop_begin[pop] + chunk * kMomChunk, end = min(P.pop_begin[pop + 1], begin + kMomChunk);
  if (begin >= end) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t g = (uint64_t)blockIdx.y * kMomTile + threadIdx.x;
  const bool mine = g < P.n_genomes && P.superpop[g] == pop;
  if (!__syncthreads_or(mine)) return;
  const uint32_t mine_mask = __ballot_sync(kFull, mine);
  const uint32_t bit = mine ? (1u << lane) : 0u;
  long long ca[kMomJ], ra[kMomJ];
#pragma unroll
  for (int j = 0; j < kMomJ; ++j) { ca[j] = 0; ra[j] = 0; }
  int cur_cb = -1, cur_rb = -1;
  uint32_t n_alt = 0, n_ref = 0;
  long long* mi_g = P.mi + g * (uint64_t)P.nbt * kMomJ;
  // table phase: two threads per locus, each with two of the tile's four units (four warp slices)
  const int tj = threadIdx.x >> 1, th = threadIdx.x & 1;
  static_assert(kMomStep * 2 == kMomTile && kMomTile == 256, "two table threads per locus, four units per tile");

  for (uint32_t s = begin; s < end; s += kMomStep) {
    __syncthreads();
    {
      const uint32_t i = s + tj;
      int cb = -1, rb = -1;
      MomMasks mk[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) mk[k] = MomMasks{0u, 0u, 0u};
      if (i < end) {
        const uint32_t l = P.rows[i];
        const MomClass m = mom_classify(P.af[(uint64_t)pop * P.n_loci + l], P.unphased != 0);
        long long u[kMomJ - 1];
        if (m.common_code >= 0) {
          const int b = mom_bin(m.r_common);
          cb = b >= kMomBinsMax ? P.nbt - 1 : b - P.b_lo;
          if (th == 0) {
            mom_powers(m.r_common, b, P.scale, u);
            s_cu[tj][0] = -1;
#pragma unroll
            for (int j = 0; j < kMomJ - 1; ++j) s_cu[tj][j + 1] = -u[j];
          }
        }
        if (m.rare_code >= 0) {
          const int b = mom_bin(m.r_rare);
          rb = (b - P.b_lo) | (m.rare_code == 0 ? (1 << 20) : 0);      // bit 20: the rare class is hom-ref (second part of the list)
          if (th == 1) {
            mom_powers(m.r_rare, b, P.scale, u);
            s_ru[tj][0] = 1;
#pragma unroll
            for (int j = 0; j < kMomJ - 1; ++j) s_ru[tj][j + 1] = u[j];
          }
        }
        const uint64_t unit0 = (uint64_t)blockIdx.y * (kMomTile / 64) + 2 * th;
        const uint4* row = P.packed + (uint64_t)l * P.units + unit0;
#pragma unroll
        for (int k = 0; k < 2; ++k)
          if (unit0 + k < P.units) {
            const uint4 v = __ldg(row + k);                       // {lo.lo32, lo.hi32, hi.lo32, hi.hi32}
            mk[2 * k] = mom_masks(v.x, v.z, m.common_code, m.rare_code);
            mk[2 * k + 1] = mom_masks(v.y, v.w, m.common_code, m.rare_code);
          }
      }
      if (th == 0) s_hdr[tj] = make_int2(cb, rb);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        s_m[tj][4 * th + k] = make_uint2(mk[k].nc, mk[k].rare);
      }
    }
    __syncthreads();
    const int n_here = (int)min((uint32_t)kMomStep, end - s);
    // loci at which one of this warp's genomes has something to add
    uint32_t actw[kMomStep / 32];
#pragma unroll
    for (int k = 0; k < kMomStep / 32; ++k) {
      const int j = k * 32 + lane;
      const uint2 m = s_m[j][warp];
      uint32_t act = m.x | m.y;
      actw[k] = __ballot_sync(kFull, j < n_here && (act & mine_mask) != 0);
    }
#pragma unroll
    for (int k = 0; k < kMomStep / 32; ++k) {
      uint32_t aw = actw[k];
      while (aw) {
        const int j = k * 32 + __ffs(aw) - 1;
        aw &= aw - 1;
        const int2 hdr = s_hdr[j];
        const uint2 m = s_m[j][warp];
        if (m.x & mine_mask) {
          if (hdr.x != cur_cb) { mom_flush(ca, mi_g, mine ? cur_cb : -1); cur_cb = hdr.x; }
          if (m.x & bit) {
            const longlong2* u = reinterpret_cast<const longlong2*>(s_cu[j]);
            const longlong2 a = u[0], b = u[1], c = u[2];
            ca[0] += a.x; ca[1] += a.y; ca[2] += b.x; ca[3] += b.y; ca[4] += c.x; ca[5] += c.y;
          }
        }
        if (m.y & mine_mask) {
          if (hdr.y != cur_rb) { mom_flush(ra, mi_g, mine ? (cur_rb & 0xFFFFF) : -1); cur_rb = hdr.y; }
          if (m.y & bit) {
            const longlong2* u = reinterpret_cast<const longlong2*>(s_ru[j]);
            const longlong2 a = u[0], b = u[1], c = u[2];
            ra[0] += a.x; ra[1] += a.y; ra[2] += b.x; ra[3] += b.y; ra[4] += c.x; ra[5] += c.y;
            if (hdr.y >> 20) ++n_ref; else ++n_alt;
          }
        }
      }
    }
  }
  if (!mine) return;
  mom_flush(ca, mi_g, cur_cb);
  mom_flush(ra, mi_g, cur_rb < 0 ? -1 : (cur_rb & 0xFFFFF));
  P.cnt[(uint64_t)chunk * P.n_genomes_padded + g] = n_alt | (n_ref << 16);
}

// Per genome: prefix of the chunk counts (offs: cells of the chunks before, hom-alt and hom-ref part; the hom-ref part of the list
// follows the whole hom-alt part), totals = {rare cells, rare hom-alt cells}.
__global__ void __launch_bounds__(256)
k_mom_scan(const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ pop_begin, const uint8_t* __restrict__ superpop,
           uint64_t n_genomes, uint64_t n_genomes_padded, uint32_t* __restrict__ totals /* [Npad][2] */,
           uint2* __restrict__ offs /* [chunks][Npad] {hom-alt, hom-ref} before the chunk; may be null */) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_genomes) return;
  const int pop = superpop[g];
  const uint32_t n = pop_begin[pop + 1] - pop_begin[pop], n_chunks = (n + kMomChunk - 1) / kMomChunk;
  uint32_t n_alt = 0, n_ref = 0;
  for (uint32_t ch = 0; ch < n_chunks; ++ch) {
    const uint32_t c = cnt[(uint64_t)ch * n_genomes_padded + g];
    if (offs) offs[(uint64_t)ch * n_genomes_padded + g] = make_uint2(n_alt, n_ref);
    n_alt += c & 0xFFFFu; n_ref += c >> 16;
  }
  totals[g * 2 + 0] = n_alt + n_ref; totals[g * 2 + 1] = n_alt;
}

// base[g] = sum of the totals of the genomes before g; base[n_genomes] = the length of the list (one block).
__global__ void __launch_bounds__(1024)
k_mom_base(const uint32_t* __restrict__ totals, uint64_t n_genomes, uint64_t* __restrict__ base) {
  __shared__ uint64_t s_warp[32];
  __shared__ uint64_t s_run;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_run = 0;
  __syncthreads();
  for (uint64_t g0 = 0; g0 < n_genomes; g0 += 1024) {
    const uint64_t g = g0 + threadIdx.x;
    const uint64_t v = g < n_genomes ? totals[g * 2] : 0;
    uint64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint64_t t = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint64_t before = s_run;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    if (g < n_genomes) base[g] = before + incl - v;
    __syncthreads();
    if (threadIdx.x == 0) { uint64_t t = 0; for (int w = 0; w < 32; ++w) t += s_warp[w]; s_run += t; }
    __syncthreads();
  }
  if (threadIdx.x == 0) base[n_genomes] = s_run;
}

// Second walk, rare homozygous cells only: the list of their r, per genome the hom-alt cells (r ascending) and then the hom-ref
// cells (r descending). Positions: base[g] + the counts of the chunks before (k_mom_scan) + the cells met so far.
__global__ void __launch_bounds__(kMomTile)
k_mom_fill(const MomParams P, const uint32_t* __restrict__ totals) {
  constexpr int kWarps = kMomTile / 32;
  __shared__ uint32_t s_rare[kMomStep][kWarps];
  __shared__ double s_r[kMomStep];
  __shared__ int s_ref[kMomStep];
  const int pop = blockIdx.x / P.chunks_per_pop, chunk = blockIdx.x % P.chunks_per_pop;
  const uint32_t begin = P.pop_begin[pop] + chunk * kMomChunk, end = min(P.pop_begin[pop + 1], begin + kMomChunk);
  if (begin >= end) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t g = (uint64_t)blockIdx.y * kMomTile + threadIdx.x;
  const bool mine = g < P.n_genomes && P.superpop[g] == pop;
  if (!__syncthreads_or(mine)) return;
  const uint32_t mine_mask = __ballot_sync(kFull, mine);
  const uint32_t bit = mine ? (1u << lane) : 0u;
  uint64_t pos_alt = 0, pos_ref = 0;
  if (mine) {
    const uint2 o = P.offs[(uint64_t)chunk * P.n_genomes_padded + g];
    pos_alt = P.base[g] + o.x;
    pos_ref = P.base[g] + totals[g * 2 + 1] + o.y;
  }
  const int tj = threadIdx.x >> 1, th = threadIdx.x & 1;
  for (uint32_t s = begin; s < end; s += kMomStep) {
    __syncthreads();
    {
      const uint32_t i = s + tj;
      uint32_t mk[4] = {0u, 0u, 0u, 0u};
      if (i < end) {
        const uint32_t l = P.rows[i];
        const MomClass m = mom_classify(P.af[(uint64_t)pop * P.n_loci + l], P.unphased != 0);
        if (th == 0) { s_r[tj] = m.r_rare; s_ref[tj] = m.rare_code == 0; }
        if (m.rare_code >= 0) {
          const uint64_t unit0 = (uint64_t)blockIdx.y * (kMomTile / 64) + 2 * th;
          const uint4* row = P.packed + (uint64_t)l * P.units + unit0;
#pragma unroll
          for (int k = 0; k < 2; ++k)
            if (unit0 + k < P.units) {
              const uint4 v = __ldg(row + k);
              mk[2 * k] = mom_code_mask(make_uint2(v.x, v.z), m.rare_code);
              mk[2 * k + 1] = mom_code_mask(make_uint2(v.y, v.w), m.rare_code);
            }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) s_rare[tj][4 * th + k] = mk[k];
    }
    __syncthreads();
    const int n_here = (int)min((uint32_t)kMomStep, end - s);
#pragma unroll
    for (int k = 0; k < kMomStep / 32; ++k) {
      const int jj = k * 32 + lane;
      uint32_t aw = __ballot_sync(kFull, jj < n_here && (s_rare[jj][warp] & mine_mask) != 0);
      while (aw) {
        const int j = k * 32 + __ffs(aw) - 1;
        aw &= aw - 1;
        if (s_rare[j][warp] & bit) {
          if (s_ref[j]) P.list[pos_ref++] = s_r[j]; else P.list[pos_alt++] = s_r[j];
        }
      }
    }
  }
}

// mi (fixed point, sparse part) + the population's dense moments -> doubles, in place.
__global__ void __launch_bounds__(256)
k_mom_finalize(long long* __restrict__ mi, const long long* __restrict__ pm, const uint8_t* __restrict__ superpop, uint64_t n_genomes,
               int nbt, double inv_scale) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_genomes * (uint64_t)nbt) return;
  const uint64_t g = i / nbt; const int b = (int)(i % nbt);
  const long long* src = pm + ((size_t)superpop[g] * nbt + b) * kMomJ;
  long long* dst = mi + i * kMomJ;
#pragma unroll
  for (int j = 0; j < kMomJ; ++j) {
    const long long v = dst[j] + src[j];
    dst[j] = __double_as_longlong(j == 0 ? (double)v : (double)v * inv_scale);
  }
}

// 1/d for the sweeps: MUFU.RCP64H and the second-order correction x0 (1 + e + e^2), e = 1 - d x0: 2.2e-16 relative (measured over
// 2^26 arguments, tools/kbench), four instructions instead of the IEEE divide's dependent chain.
__device__ __forceinline__ double mom_rcp(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  const double e = fma(-d, x, 1.0);
  return fma(x, fma(e, e, e), x);
}

// ---- one sweep -----------------------------------------------------------------------------------------------------------------
// Block = genome. NEWTON: out[pos] = {S1, S2} in the layout of one chunk of k_terms_fast<FAST_NEWTON> (k_newton_reduce follows);
// HALL: iter[g][0] = f (n_hom + (1 - f) S1).
template <int MODE>
__global__ void __launch_bounds__(128)
k_mom_eval(const double* __restrict__ mom, int b_lo, int nbt, const double* __restrict__ f, const uint32_t* __restrict__ list,
           uint64_t n_list, const uint32_t* __restrict__ n_list_dev, uint64_t n_genomes, const double* __restrict__ rare,
           const uint64_t* __restrict__ base, const uint32_t* __restrict__ totals, double* __restrict__ out, double* __restrict__ iter) {
  __shared__ double s_red[3][4];
  const uint64_t pos = blockIdx.x;
  if (list && n_list_dev) n_list = min(n_list, (uint64_t)*n_list_dev);
  if (pos >= (list ? n_list : n_genomes)) return;
  const uint64_t g = list ? list[pos] : pos;
  const double x = f[g];
  // f < 0: the octaves below 4|f| are evaluated from the list (their bins hold rare cells only: 4|f| <= 0.8 < 1)
  double edge = 0.0; int b_edge = 0;
  if (x < 0.0) {
    int e; const double m = frexp(-4.0 * x, &e);          // 4|x| = m 2^e, m in [0.5, 1)
    if (m == 0.5) --e;                                      // an exact power of two is its own edge
    edge = ldexp(1.0, e);
    b_edge = (e - kMomExpLo) * kMomSub;
  }
  double s1 = 0.0, s2 = 0.0, n0 = 0.0;
  const double* M = mom + g * (uint64_t)nbt * kMomJ;
  for (int b = threadIdx.x; b < nbt - 1; b += blockDim.x) {
    const int gb = b + b_lo;
    if (gb < b_edge) continue;
    const double m0 = M[b * kMomJ];
    if (m0 == 0.0) continue;
    double rc, w;
    mom_geometry(gb, rc, w);
    const double t = mom_rcp(x + rc), z = -w * t;
    const double m1 = M[b * kMomJ + 1], m2 = M[b * kMomJ + 2], m3 = M[b * kMomJ + 3], m4 = M[b * kMomJ + 4], m5 = M[b * kMomJ + 5];
    const double p1 = fma(z, fma(z, fma(z, fma(z, fma(z, m5, m4), m3), m2), m1), m0);
    s1 = fma(t, p1, s1);
    if (MODE == FAST_NEWTON) {
      const double p2 = fma(z, fma(z, fma(z, fma(z, fma(z, 6.0 * m5, 5.0 * m4), 4.0 * m3), 3.0 * m2), 2.0 * m1), m0);
      s2 = fma(t * t, p2, s2);
    }
    n0 += m0;
  }
  if (threadIdx.x == 0) n0 += M[(nbt - 1) * kMomJ];           // a == 1 cells: t = 0
  if (x < 0.0 && rare) {
    const uint64_t n = totals[g * 2], n_alt = totals[g * 2 + 1];
    const double* L = rare + base[g];
    for (uint64_t i = threadIdx.x; i < n_alt; i += blockDim.x) {           // hom-alt part: r ascending
      const double r = L[i];
      if (!(r < edge)) break;
      const double t = mom_rcp(x + r);
      s1 += t; s2 = fma(t, t, s2); n0 += 1.0;
    }
    for (uint64_t i = threadIdx.x; i < n - n_alt; i += blockDim.x) {       // hom-ref part: r descending, from its end
      const double r = L[n - 1 - i];
      if (!(r < edge)) break;
      const double t = mom_rcp(x + r);
      s1 += t; s2 = fma(t, t, s2); n0 += 1.0;
    }
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2); n0 = warp_sum(n0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_red[0][warp] = s1; s_red[1][warp] = s2; s_red[2][warp] = n0; }
  __syncthreads();
  if (threadIdx.x != 0) return;
  s1 = (s_red[0][0] + s_red[0][1]) + (s_red[0][2] + s_red[0][3]);
  s2 = (s_red[1][0] + s_red[1][1]) + (s_red[1][2] + s_red[1][3]);
  n0 = (s_red[2][0] + s_red[2][1]) + (s_red[2][2] + s_red[2][3]);
  if (MODE == FAST_NEWTON) { out[pos * 2 + 0] = s1; out[pos * 2 + 1] = s2; }
  else iter[g * 4] = x * (n0 + (1.0 - x) * s1);
}

// ---- whole runs in one launch (one context holds every locus: no all-reduce between sweeps) --------------------------------------
// Block = genome; its table row sits in shared memory ({rc, w} and the six moments of every bin) for all sweeps.
struct MomRunParams {
  const double* mom; int b_lo, nbt;
  uint64_t n_genomes;
  const double* partials;            // moments of the counting pass: n = terms of the genome (calc.cpp:285)
  double* f;                         // in: start, out: result
  // HallME
  int hall_sweeps;                   // > 0: that many sweeps; < 0: to the fixed point (|delta| < 1e-15, at most 100000 sweeps)
  // root search
  const double* rare; const uint64_t* base; const uint32_t* totals; const double* limits;
  double* bracket; uint32_t* done; double tol; int max_iterations;
  unsigned long long* remaining;     // genomes the kernel left unfinished (they need the exact cell-by-cell kernel)
};

// COUNT: also the number of homozygous cells n0 (the same in every sweep: HallME asks once).
template <int MODE, bool COUNT>
__device__ __forceinline__ void mom_block_sums(const double* s_m, const double2* s_geo, int b_lo, int nbt, double x, const double* L,
                                               uint64_t n, uint64_t n_alt, double* s_red, double& s1, double& s2, double& n0) {
  double edge = 0.0; int b_edge = 0;
  if (x < 0.0) {
    int e; const double m = frexp(-4.0 * x, &e);
    if (m == 0.5) --e;
    edge = ldexp(1.0, e);
    b_edge = (e - kMomExpLo) * kMomSub;
  }
  s1 = 0.0; s2 = 0.0; n0 = 0.0;
  for (int b = threadIdx.x; b < nbt - 1; b += blockDim.x) {
    if (b + b_lo < b_edge) continue;
    const double m0 = s_m[b * kMomJ];
    if (m0 == 0.0) continue;
    const double2 geo = s_geo[b];
    const double t = mom_rcp(x + geo.x), z = -geo.y * t;
    const double m1 = s_m[b * kMomJ + 1], m2 = s_m[b * kMomJ + 2], m3 = s_m[b * kMomJ + 3], m4 = s_m[b * kMomJ + 4], m5 = s_m[b * kMomJ + 5];
    s1 = fma(t, fma(z, fma(z, fma(z, fma(z, fma(z, m5, m4), m3), m2), m1), m0), s1);
    if (MODE == FAST_NEWTON) s2 = fma(t * t, fma(z, fma(z, fma(z, fma(z, fma(z, 6.0 * m5, 5.0 * m4), 4.0 * m3), 3.0 * m2), 2.0 * m1), m0), s2);
    if (COUNT) n0 += m0;
  }
  if (COUNT && threadIdx.x == 0) n0 += s_m[(nbt - 1) * kMomJ];
  if (MODE == FAST_NEWTON && x < 0.0 && L) {
    for (uint64_t i = threadIdx.x; i < n_alt; i += blockDim.x) {
      const double r = L[i];
      if (!(r < edge)) break;
      const double t = mom_rcp(x + r);
      s1 += t; s2 = fma(t, t, s2); if (COUNT) n0 += 1.0;
    }
    for (uint64_t i = threadIdx.x; i < n - n_alt; i += blockDim.x) {
      const double r = L[n - 1 - i];
      if (!(r < edge)) break;
      const double t = mom_rcp(x + r);
      s1 += t; s2 = fma(t, t, s2); if (COUNT) n0 += 1.0;
    }
  }
  s1 = warp_sum(s1);
  if (MODE == FAST_NEWTON) s2 = warp_sum(s2);
  if (COUNT) n0 = warp_sum(n0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();                                   // the sums of the sweep before have been read
  if (lane == 0) { s_red[warp] = s1; if (MODE == FAST_NEWTON) s_red[8 + warp] = s2; if (COUNT) s_red[16 + warp] = n0; }
  __syncthreads();
  s1 = ((s_red[0] + s_red[1]) + (s_red[2] + s_red[3])) + ((s_red[4] + s_red[5]) + (s_red[6] + s_red[7]));
  if (MODE == FAST_NEWTON) s2 = ((s_red[8] + s_red[9]) + (s_red[10] + s_red[11])) + ((s_red[12] + s_red[13]) + (s_red[14] + s_red[15]));
  if (COUNT) n0 = ((s_red[16] + s_red[17]) + (s_red[18] + s_red[19])) + ((s_red[20] + s_red[21]) + (s_red[22] + s_red[23]));
}

constexpr int kMomRunThreads = 256;                    // eight warps per genome: mom_block_sums adds eight partial sums
template <int MODE>
__global__ void __launch_bounds__(kMomRunThreads)
k_mom_run(const MomRunParams P) {
  extern __shared__ __align__(16) unsigned char mom_run_smem[];
  double* s_m = reinterpret_cast<double*>(mom_run_smem);                       // [nbt][kMomJ]
  double2* s_geo = reinterpret_cast<double2*>(s_m + (size_t)P.nbt * kMomJ);    // [nbt] {rc, w}
  __shared__ double s_red[24];
  const uint64_t g = blockIdx.x;
  const double* M = P.mom + g * (uint64_t)P.nbt * kMomJ;
  for (int i = threadIdx.x; i < P.nbt * kMomJ; i += blockDim.x) s_m[i] = M[i];
  for (int b = threadIdx.x; b < P.nbt; b += blockDim.x) {
    double rc = 0.0, w = 0.0;
    if (b < P.nbt - 1) mom_geometry(b + P.b_lo, rc, w);
    s_geo[b] = make_double2(rc, w);
  }
  const double* Pg = P.partials + g * PART_COUNT;
  const double n_terms = Pg[PART_NMAJHOM] + Pg[PART_NMAJHET] + Pg[PART_NMINHOM] + Pg[PART_NMINHET];
  double x = P.f[g];
  __syncthreads();
  if (MODE == FAST_HALL) {
    const int limit = P.hall_sweeps > 0 ? P.hall_sweeps : 100000;
    double n0 = 0.0;                                   // the homozygous cells: the same in every sweep
    for (int it = 0; it < limit; ++it) {
      double s1, s2, unused;
      if (it == 0) mom_block_sums<FAST_HALL, true>(s_m, s_geo, P.b_lo, P.nbt, x, nullptr, 0, 0, s_red, s1, s2, n0);
      else mom_block_sums<FAST_HALL, false>(s_m, s_geo, P.b_lo, P.nbt, x, nullptr, 0, 0, s_red, s1, s2, unused);
      const double sum = x * (n0 + (1.0 - x) * s1);
      const double nx = (x == 0.0 && n_terms > 0.0) ? 0.0 : __ddiv_rn(sum, n_terms);       // as k_hall_update
      const bool stop = P.hall_sweeps < 0 && fabs(nx - x) < 1e-15;
      x = nx;
      if (stop) break;                               // every thread computed the same nx
    }
    if (threadIdx.x == 0) P.f[g] = x;
    return;
  }
  // root search: k_newton_reduce + k_ll_step per iteration, until the search ends or the genome needs the exact kernel
  if (P.done[g]) return;
  const double fmin_ = P.limits[g * 3 + 0], cmin = P.limits[g * 3 + 1], nhet = P.limits[g * 3 + 2];
  const double band = kLimitMargin * fmax(1.0, fabs(fmin_));
  const uint64_t n = P.totals[g * 2], n_alt = P.totals[g * 2 + 1];
  const double* L = P.rare + P.base[g];
  double a = P.bracket[g * 2 + 0], b = P.bracket[g * 2 + 1];
  bool finished = false;
  for (int it = 0; it < P.max_iterations; ++it) {
    int st = 0;
    if (x < fmin_ - band) st = 1;
    else if (x < fmin_ + band || x < kMomValidMin) st = 2;
    else if (nhet > 0.0 && !((1.0 - x) * cmin > kSmallProb * (1.0 + kLimitMargin))) st = 2;
    if (st == 2) break;                              // uniform: every thread holds the same x
    double g1 = 0.0, g2 = 0.0;
    bool hom_clamped = st == 1;
    if (st == 0) {
      double s1, s2, n0;
      mom_block_sums<FAST_NEWTON, false>(s_m, s_geo, P.b_lo, P.nbt, x, L, n, n_alt, s_red, s1, s2, n0);
      const double t = 1.0 / (1.0 - x);
      g1 = s1 - nhet * t;
      g2 = -s2 - nhet * t * t;
    }
    if (hom_clamped || g1 > 0.0) a = x; else b = x;              // k_ll_step
    double nx = 0.5 * (a + b);
    bool newton_converged = false;
    if (!hom_clamped && g2 < 0.0) {
      const double cand = x - g1 / g2;
      if (fabs(cand - x) < P.tol) { nx = (cand > a && cand < b) ? cand : x; newton_converged = true; }
      else if (cand > a && cand < b) nx = cand;
    }
    const bool stop = newton_converged || fabs(nx - x) < P.tol || (b - a) < P.tol;
    x = nx;
    if (stop) { finished = true; break; }
  }
  if (threadIdx.x == 0) {
    P.f[g] = x; P.bracket[g * 2 + 0] = a; P.bracket[g * 2 + 1] = b;
    if (finished) P.done[g] = 1; else atomicAdd(P.remaining, 1ull);
  }
}

}  // namespace kgl
