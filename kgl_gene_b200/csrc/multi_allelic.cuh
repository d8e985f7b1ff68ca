// multi_allelic.cuh -- loci with several alternate alleles (the "3 + side list" rows of the flattener contract).
//
// At such a locus the reference's classification and class frequencies take their general form
// (kga_analysis_inbreed_freq.cpp:18-57 AlleleFreqVector over several alleles, :127-217 alleleClassFrequencies, :452-543
// generateFrequencies): the major allele frequency is the complement of the SUM of the alternate frequencies, a genome with two
// different alternate alleles is MINOR_HETEROZYGOUS with both allele frequencies, and an allele without a frequency for the
// genome's super-population is not in the list (the locus is dropped for that genome). These loci are few (about 1 % of a
// 1000 Genomes contig), so they stay out of the dense machinery altogether: their rows of the frequency table hold NaN (never
// selected, no dense totals, neutral in every table-driven sweep), their cells are coded 0 / 3 in the matrix, and this file
// evaluates them cell by cell from a byte matrix cells[m][genome] -- lane = genome, chunks of loci, fixed-order reduction --
// adding their share to whatever the dense path produced: moments, Ritland / HallME sums, likelihood derivatives and grid values.
#pragma once
#include "common.cuh"
#include "misc_kernels.cuh"
#include "sample_major.cuh"

namespace kgl {

constexpr int kMultiSlots = 3;            // a SNP has at most three alternate alleles
constexpr double kMultiEpsilon = 1.0e-05; // epsilon_class_, checkValidAlleleVector (freq.cpp:61-75)

// Per (population, multi-allelic locus), rebuilt for every selection: what generateFrequencies derives from the AlleleFreqVector.
struct MultiLocus {
  double q;                  // majorAlleleFrequency(): clamp(1 - clamp(sum p, 0, 1), 0, 1)    (freq.cpp:113-123)
  double cf[4];              // alleleClassFrequencies(0.0): majHom, majHet, minHom, minHet     (freq.cpp:127-217, freq.h:54-63)
  double p[kMultiSlots];     // clamped frequency of every allele slot
  uint32_t in_list;          // bit a: slot a has a frequency for this population
  uint32_t active;           // selected for the population and the vector is valid
};

// The allele vector of one (population, locus): frequencies in slot order (= the order of the locus' variant array).
struct MultiVector { int n; int slot[kMultiSlots]; double p[kMultiSlots]; double sum; bool valid; };
__device__ __forceinline__ MultiVector multi_vector(const float* __restrict__ af3) {
  MultiVector v;
  v.n = 0; v.sum = 0.0;
#pragma unroll
  for (int a = 0; a < kMultiSlots; ++a) {
    const float f = af3[a];
    if (f != f) continue;
    double p = (double)f;
    p = p < 0.0 ? 0.0 : (p > 1.0 ? 1.0 : p);                       // freq.cpp:47
    v.slot[v.n] = a; v.p[v.n] = p; ++v.n;
    v.sum = __dadd_rn(v.sum, p);                                    // sumAlleleFrequencies(), slot order (freq.cpp:97-105)
  }
  v.valid = v.n > 0 && !(__dsub_rn(v.sum, 1.0) > kMultiEpsilon);
  return v;
}

// RetrieveLociiVector::getAllelesFromTo candidates among the multi-allelic loci (kga_analysis_inbreed_locus.cpp:21-72): k_select_dense
// left their bits clear (NaN in the frequency table). Runs before the spaced accept chain, which then treats them like any locus.
__global__ void __launch_bounds__(128)
k_multi_select(const uint32_t* __restrict__ rows, const float* __restrict__ af, const uint32_t* __restrict__ offsets, uint64_t n_multi,
               int n_pop, uint64_t lower, uint64_t upper, double min_af, double max_af, const uint8_t* __restrict__ keep,
               uint8_t* __restrict__ sel, unsigned long long* __restrict__ counts) {
  const uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_multi) return;
  const uint32_t row = rows[m];
  const uint64_t offset = offsets[row];
  uint32_t bits = 0;
  if (offset >= lower && offset <= upper && (keep == nullptr || keep[row] != 0)) {
    for (int k = 0; k < n_pop; ++k) {
      const MultiVector v = multi_vector(af + ((uint64_t)k * n_multi + m) * kMultiSlots);
      if (!v.valid) continue;
      const double s = v.sum < 0.0 ? 0.0 : (v.sum > 1.0 ? 1.0 : v.sum);     // minorAlleleFrequencies() (freq.cpp:107-111)
      if (s == 0.0 || s < min_af || s > max_af) continue;                   // locus.cpp:53-54
      bits |= 1u << k;
      atomicAdd(&counts[k], 1ull);
    }
  }
  sel[row] = (uint8_t)bits;
}

__global__ void __launch_bounds__(128)
k_multi_prepare(const uint32_t* __restrict__ rows, const float* __restrict__ af, const uint8_t* __restrict__ sel, uint64_t n_multi,
                int n_pop, MultiLocus* __restrict__ tab /* [n_pop][n_multi] */) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_multi * (uint64_t)n_pop) return;
  const uint64_t k = i / n_multi, m = i % n_multi;
  const MultiVector v = multi_vector(af + (k * n_multi + m) * kMultiSlots);
  MultiLocus t;
  t.active = (v.valid && ((sel[rows[m]] >> k) & 1u)) ? 1u : 0u;
  t.in_list = 0;
#pragma unroll
  for (int a = 0; a < kMultiSlots; ++a) t.p[a] = 0.0;
  for (int j = 0; j < v.n; ++j) { t.in_list |= 1u << v.slot[j]; t.p[v.slot[j]] = v.p[j]; }
  double s = v.sum < 0.0 ? 0.0 : (v.sum > 1.0 ? 1.0 : v.sum);
  double q = __dsub_rn(1.0, s);
  t.q = q < 0.0 ? 0.0 : (q > 1.0 ? 1.0 : q);
  // unadjustedAlleleClassFrequencies(0.0) + normalize(), the reference's operation order (freq.cpp:131-190, freq.h:54-63)
  const double major = fmax(0.0, __dsub_rn(1.0, v.sum));
  double f[kMultiSlots];
  for (int j = 0; j < v.n; ++j) f[j] = (v.sum > 1.0) ? __ddiv_rn(v.p[j], v.sum) : v.p[j];
  double min_hom = 0.0, min_het = 0.0, maj_het = 0.0;
  for (int j = 0; j < v.n; ++j) min_hom = __dadd_rn(min_hom, __dmul_rn(f[j], f[j]));
  for (int j = 0; j < v.n; ++j)
    for (int j2 = j + 1; j2 < v.n; ++j2) min_het = __dadd_rn(min_het, __dmul_rn(__dmul_rn(2.0, f[j]), f[j2]));
  const double maj_hom = __dmul_rn(major, major);
  for (int j = 0; j < v.n; ++j) maj_het = __dadd_rn(maj_het, __dmul_rn(__dmul_rn(2.0, major), f[j]));
  const double sum = __dadd_rn(__dadd_rn(__dadd_rn(maj_hom, maj_het), min_hom), min_het);
  t.cf[0] = __ddiv_rn(maj_hom, sum); t.cf[1] = __ddiv_rn(maj_het, sum); t.cf[2] = __ddiv_rn(min_hom, sum); t.cf[3] = __ddiv_rn(min_het, sum);
  tab[i] = t;
}

// Classification of one cell (freq.cpp:452-543). cls: 0 majHom, 1 majHet, 2 minHom, 3 minHet; returns false = dropped.
__device__ __forceinline__ bool multi_classify(uint32_t cell, const MultiLocus& t, bool unphased, int& cls, double& a1, double& a2) {
  if (!t.active) return false;
  if (cell == 0) {                                              // no variant at the offset (:521-541)
    if (!(t.q > kMinMajorFreq)) return false;
    cls = 0; a1 = t.q; a2 = t.q;
    return true;
  }
  if (cell == 0xFFu) return false;                              // more than two variants
  const uint32_t first = (cell & 15u) - 1u, second = cell >> 4;
  if (first >= (uint32_t)kMultiSlots || !((t.in_list >> first) & 1u)) return false;     // front variant not in the list (:462)
  if (second == 0) { cls = 1; a1 = t.p[first]; a2 = t.q; return true; }                 // one variant (:464-472)
  const uint32_t sec = second - 1u;
  if (sec == first && !unphased) { cls = 2; a1 = t.p[first]; a2 = a1; return true; }    // homozygous(): same allele, phases differ (:476)
  if (sec >= (uint32_t)kMultiSlots || !((t.in_list >> sec) & 1u)) return false;         // second minor not found (:500)
  cls = 3; a1 = t.p[first]; a2 = t.p[sec];                                              // :482-511 (also the unphased pair, Q6)
  return true;
}

enum { MULTI_MOMENTS = 0, MULTI_HALL = 1, MULTI_NEWTON = 2, MULTI_GRID = 3 };
// outputs per genome: MOMENTS {4 class counts, 4 expected sums, Ritland sum, Ritland count}; HALL {sum}; NEWTON {dLL, d2LL, clamped
// homozygous, clamped heterozygous}; GRID {kGridMax values}
__host__ __device__ constexpr int multi_n_out(int mode) { return mode == MULTI_MOMENTS ? 10 : mode == MULTI_HALL ? 1 : mode == MULTI_NEWTON ? 4 : kGridMax; }
constexpr int kMultiChunk = 256;          // loci per chunk (gridDim.y chunks)

struct MultiParams {
  const uint8_t* cells; uint64_t n_multi, n_genomes, n_genomes_padded;
  const MultiLocus* tab; const uint8_t* superpop; int unphased;
  const double* f; const double* grid; int n_grid;
  double* out;                            // [n_chunks][n_genomes_padded][multi_n_out]
};

template <int MODE>
__global__ void __launch_bounds__(128)
k_multi_terms(const MultiParams P) {
  constexpr int NOUT = multi_n_out(MODE);
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= P.n_genomes) return;
  const int k = P.superpop[g];
  const bool unphased = P.unphased != 0;
  const double f = (MODE == MULTI_HALL || MODE == MULTI_NEWTON) ? P.f[g] : 0.0;
  double gridv[kGridMax];
  if (MODE == MULTI_GRID) {
#pragma unroll
    for (int j = 0; j < kGridMax; ++j) gridv[j] = (j < P.n_grid) ? P.grid[j] : 0.0;
  }
  double acc[NOUT];
#pragma unroll
  for (int j = 0; j < NOUT; ++j) acc[j] = 0.0;
  const uint64_t m0 = (uint64_t)blockIdx.y * kMultiChunk, m1 = min(m0 + (uint64_t)kMultiChunk, P.n_multi);
  const MultiLocus* tab = P.tab + (uint64_t)k * P.n_multi;
  for (uint64_t m = m0; m < m1; ++m) {
    const uint32_t cell = P.cells[m * P.n_genomes + g];
    const MultiLocus t = tab[m];
    int cls; double a1, a2;
    if (!multi_classify(cell, t, unphased, cls, a1, a2)) continue;
    const bool hom = cls == 0 || cls == 2;
    if (MODE == MULTI_MOMENTS) {
      acc[0] += cls == 0 ? 1.0 : 0.0; acc[1] += cls == 1 ? 1.0 : 0.0; acc[2] += cls == 2 ? 1.0 : 0.0; acc[3] += cls == 3 ? 1.0 : 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[4 + j] += t.cf[j];                        // statistics block (freq.cpp:549-556)
      if (hom) {                                                                 // processRitlandLocus (calc.cpp:390-415)
        if (a1 > kRitlandMinFreq) { acc[8] += __dsub_rn(__ddiv_rn(1.0, a1), 1.0); acc[9] += 1.0; }
      } else { acc[8] -= 1.0; acc[9] += 1.0; }
    } else if (MODE == MULTI_HALL) {                                            // processHallME (calc.cpp:261-283)
      if (hom) {
        const double denominator = __dadd_rn(f, __dmul_rn(__dsub_rn(1.0, f), a1));
        if (denominator != 0) acc[0] = __dadd_rn(acc[0], __ddiv_rn(f, denominator));
      }
    } else if (MODE == MULTI_NEWTON) {                                          // d/df logLikelihood (calc.cpp:94-129), as k_genome_terms
      if (hom) {
        const double prob = __dadd_rn(__dmul_rn(f, a1), __dmul_rn(__dsub_rn(1.0, f), __dmul_rn(a1, a1)));
        if (prob >= kSmallProb) {
          if (prob <= 1.0) { const double t1 = (1.0 - a1) / (a1 + f * (1.0 - a1)); acc[0] += t1; acc[1] -= t1 * t1; }
        } else acc[2] += 1.0;
      } else {
        const double prob = 2 * (1.0 - f) * a1 * a2;
        if (prob >= kSmallProb && prob <= 1.0) { const double t1 = 1.0 / (1.0 - f); acc[0] -= t1; acc[1] -= t1 * t1; }
        else acc[3] += 1.0;
      }
    } else {                                                                    // logLikelihood on the grid (calc.cpp:100-125)
#pragma unroll
      for (int j = 0; j < kGridMax; ++j) {
        const double fj = gridv[j];
        double prob;
        if (hom) prob = __dadd_rn(__dmul_rn(fj, a1), __dmul_rn(__dsub_rn(1.0, fj), __dmul_rn(a1, a1)));
        else prob = __dmul_rn(__dmul_rn(__dmul_rn(2.0, __dsub_rn(1.0, fj)), a1), a2);
        prob = prob < kSmallProb ? kSmallProb : (prob > 1.0 ? 1.0 : prob);
        acc[j] += log(prob);
      }
    }
  }
  double* out = P.out + ((uint64_t)blockIdx.y * P.n_genomes_padded + g) * NOUT;
#pragma unroll
  for (int j = 0; j < NOUT; ++j) out[j] = acc[j];
}

// Adds the chunk outputs, in chunk order, to what the dense path left: MOMENTS -> partials (+ the Simple closed form again when
// results != null); HALL / NEWTON -> iter[g][0..]; GRID is reduced by the host like the dense grid values.
template <int MODE>
__global__ void __launch_bounds__(128)
k_multi_add(const double* __restrict__ chunk_out, uint64_t n_chunks, uint64_t n_genomes_padded, uint64_t n_genomes,
            double* __restrict__ target, kgl_b200_locus_results* __restrict__ results) {
  constexpr int NOUT = multi_n_out(MODE);
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_genomes) return;
  double s[NOUT];
#pragma unroll
  for (int j = 0; j < NOUT; ++j) s[j] = 0.0;
  for (uint64_t c = 0; c < n_chunks; ++c) {
    const double* o = chunk_out + (c * n_genomes_padded + g) * NOUT;
#pragma unroll
    for (int j = 0; j < NOUT; ++j) s[j] += o[j];
  }
  if (MODE == MULTI_MOMENTS) {
    double* P = target + g * PART_COUNT;
    P[PART_NMAJHOM] += s[0]; P[PART_NMAJHET] += s[1]; P[PART_NMINHOM] += s[2]; P[PART_NMINHET] += s[3];
    P[PART_EMAJHOM] += s[4]; P[PART_EMAJHET] += s[5]; P[PART_EMINHOM] += s[6]; P[PART_EMINHET] += s[7];
    P[PART_RSUM] += s[8]; P[PART_RCOUNT] += s[9];
    if (results) results[g] = closed_form(P, KGL_B200_ALGO_SIMPLE);
  } else {
    double* I = target + g * ITER_COUNT;
#pragma unroll
    for (int j = 0; j < NOUT && j < ITER_COUNT; ++j) I[j] += s[j];
  }
}

// Per-allele summaries of the multi-allelic loci (VariantDBVariant has one column per variant, kgl_variant_db_variant.cpp:14-30):
// counts[m][a][c] = genomes with c = 0, 1, 2 copies of allele slot a. One block per locus.
__global__ void __launch_bounds__(256)
k_multi_allele_count(const uint8_t* __restrict__ cells, uint64_t n_genomes, uint32_t* __restrict__ counts) {
  __shared__ uint32_t s_c[kMultiSlots][3];
  if (threadIdx.x < kMultiSlots * 3) (&s_c[0][0])[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t m = blockIdx.x;
  uint32_t c[kMultiSlots][3];
#pragma unroll
  for (int a = 0; a < kMultiSlots; ++a) { c[a][0] = 0; c[a][1] = 0; c[a][2] = 0; }
  for (uint64_t g = threadIdx.x; g < n_genomes; g += blockDim.x) {
    const uint32_t cell = cells[m * n_genomes + g];
#pragma unroll
    for (int a = 0; a < kMultiSlots; ++a) {
      uint32_t copies = ((cell & 15u) == (uint32_t)(a + 1)) + ((cell >> 4) == (uint32_t)(a + 1));
      if (cell == 0xFFu) copies = 0;                        // more than two variants: which ones is not recorded
      ++c[a][copies];
    }
  }
#pragma unroll
  for (int a = 0; a < kMultiSlots; ++a)
#pragma unroll
    for (int j = 0; j < 3; ++j) atomicAdd(&s_c[a][j], c[a][j]);
  __syncthreads();
  if (threadIdx.x < kMultiSlots * 3) counts[m * kMultiSlots * 3 + threadIdx.x] = (&s_c[0][0])[threadIdx.x];
}

// CalcFWS::updateGenomeFWSMap (kga_PfEMP/kga_analysis_PfEMP_FWS.cpp:72-101) over the alleles of the multi-allelic loci, one AF bin
// per launch, added to what the masked pass over the ordinary rows left in out[g] = {refHom, het, minorHom, code 3}: every listed
// allele is a variant of its own with its own AF (P7FrequencyFilter reads the variant's element of the Number=A vector,
// kgl_variant_filter_Pf7.cpp:22-48); it is in the bin when lower <= AF < upper and, with allele_counts, when some genome carries
// it. A genome has 0, 1 or 2 copies of the allele (kgl_variant_db_variant.cpp:73-103). Cells with more than two variants (0xFF)
// do not record which: they count as no copy, as in k_multi_allele_count. One thread per genome; the membership test is
// recomputed by every thread (3 M float reads from L1).
__global__ void __launch_bounds__(128)
k_multi_bin_counts(const uint8_t* __restrict__ cells, const float* __restrict__ af_pop /* [M][3] of the population */,
                   const uint32_t* __restrict__ allele_counts /* nullable [M][3][3] */, const uint32_t* __restrict__ rows,
                   const uint8_t* __restrict__ keep /* nullable, per row */, uint64_t n_multi, uint64_t n_genomes, double lower, double upper,
                   uint64_t* __restrict__ out, uint64_t* __restrict__ rows_out) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t members = 0, het = 0, hom = 0;
  for (uint64_t m = 0; m < n_multi; ++m) {
    if (keep && keep[rows[m]] == 0) continue;
    const uint32_t cell = g < n_genomes ? cells[m * n_genomes + g] : 0u;
#pragma unroll
    for (int a = 0; a < kMultiSlots; ++a) {
      const float af = af_pop[m * kMultiSlots + a];
      if (!(af == af)) continue;
      const double v = (double)af;
      if (!(v >= lower && !(v >= upper))) continue;
      if (allele_counts && allele_counts[(m * kMultiSlots + a) * 3 + 1] + allele_counts[(m * kMultiSlots + a) * 3 + 2] == 0) continue;
      ++members;
      if (cell == 0u || cell == 0xFFu) continue;
      const uint32_t copies = ((cell & 15u) == (uint32_t)(a + 1)) + ((cell >> 4) == (uint32_t)(a + 1));
      het += copies == 1u; hom += copies == 2u;
    }
  }
  if (g == 0) *rows_out += members;
  if (g >= n_genomes) return;
  out[g * 4 + 0] += members - het - hom; out[g * 4 + 1] += het; out[g * 4 + 2] += hom;
}

// HeteroHomoZygous::updateVariantAnalysisType (kga_PfEMP/kga_analysis_PfEMP_heterozygous.cpp:61-105) for every genome, from the raw
// per-genome code counts of a counting pass and the multi-allelic side cells. Per offset of a genome: every variant entry counts
// (total_variants_, snp_count_: the matrix path holds SNPs only); one entry -> heterozygous_reference_minor_alleles_; otherwise
// homozygous_minor_alleles_ += distinct alleles (UniqueUnphasedFilter) and heterozygous_minor_alleles_ += alleles that occur
// once (HeterozygousFilter): a hom-alt pair gives {1, 0}, two different alternate alleles {2, 2}.
// out[g] = {total, snp, indel, homMinor, hetMinor, hetRefMinor, homRef}. other_entries: variant entries a code-3 cell of an
// ordinary row stands for (1: one other allele, the flattener's usual case).
__global__ void __launch_bounds__(128)
k_hetero_homo(const uint32_t* __restrict__ gcounts, const uint32_t* __restrict__ n3s, const uint8_t* __restrict__ cells /* nullable */,
              uint64_t n_multi, uint64_t n_genomes, uint32_t other_entries, uint64_t* __restrict__ out) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_genomes) return;
  uint64_t single = 0, same = 0, diff = 0, many = 0;
  for (uint64_t m = 0; m < n_multi; ++m) {
    const uint32_t cell = cells[m * n_genomes + g];
    if (cell == 0) continue;
    if (cell == 0xFFu) ++many;
    else if ((cell >> 4) == 0) ++single;
    else if ((cell >> 4) == (cell & 15u)) ++same;
    else ++diff;
  }
  const uint64_t n3_all = n3s[g], n1 = gcounts[g * 2] - n3_all, n2 = gcounts[g * 2 + 1] - n3_all;
  const uint64_t n3 = (n3_all - (single + same + diff + many)) * other_entries;      // code-3 cells of ordinary rows
  const uint64_t total = n1 + 2 * n2 + n3 + single + 2 * same + 2 * diff + 3 * many;
  uint64_t* o = out + g * 7;
  o[0] = total; o[1] = total; o[2] = 0;
  o[3] = n2 + same + 2 * diff + 2 * many;
  o[4] = 2 * diff + many;
  o[5] = n1 + n3 + single;
  o[6] = 0;
}

// The frequency-table entries of the multi-allelic rows are not used: they are set to "no value" so that the dense path never
// selects such a row, whatever the caller left there.
__global__ void __launch_bounds__(128)
k_multi_mask_af(const uint32_t* __restrict__ rows, uint64_t n_multi, int n_pop, uint64_t n_loci, float* __restrict__ af) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_multi * (uint64_t)n_pop) return;
  af[(i / n_multi) * n_loci + rows[i % n_multi]] = __int_as_float(0x7fc00000);
}

}  // namespace kgl
