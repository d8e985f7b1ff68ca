// common.cuh -- shared device helpers for the kgl_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace kgl {

constexpr int kMaxPop = 6;
constexpr unsigned kFull = 0xffffffffu;

// Streaming 128-bit load that does not allocate in L1: every genotype byte is touched exactly once per pass.
__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// Carry-save adder on bit-planes: (h,l) = a + b + c per bit lane. Two LOP3 (0x96 = xor3, 0xE8 = majority).
__device__ __forceinline__ void csa(uint32_t& h, uint32_t& l, uint32_t a, uint32_t b, uint32_t c) {
  uint32_t lo, hi;
  asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(lo) : "r"(a), "r"(b), "r"(c));
  asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(hi) : "r"(a), "r"(b), "r"(c));
  h = hi; l = lo;
}

// ---- per-locus quantities, rounded exactly like the reference (no FMA contraction) --------------------------------
// AlleleFreqVector for ONE alt allele (kga_analysis_inbreed_freq.cpp:18-57): p = clamp((double)af, 0, 1); valid iff AF present.
struct LocusFreq {
  double p, q;     // minor (alt) and major allele frequency (majorAlleleFrequency(), freq.cpp:119-123)
  bool valid;
};

__device__ __forceinline__ LocusFreq locus_freq(float af) {
  LocusFreq f;
  f.valid = !(af != af);
  double p = (double)af;
  p = p < 0.0 ? 0.0 : (p > 1.0 ? 1.0 : p);
  if (!f.valid) p = 0.0;
  double q = __dsub_rn(1.0, p);
  q = q < 0.0 ? 0.0 : (q > 1.0 ? 1.0 : q);
  f.p = p; f.q = q;
  return f;
}

// alleleClassFrequencies(0.0) (freq.cpp:127-217 + freq.h:54-63) for one alt allele: normalised {majHom, majHet, minHom}.
// (minHet is identically 0 with a single alt allele.) The products follow the reference's operation order. The reference
// then divides each class by sum = q^2 + 2qp + p^2, which is 1 + e with |e| < 1e-15 because q = fl(1 - p); here the
// division is the multiplication by 2 - sum (= 1/sum to within e^2 < 1e-30, and exact in floating point), so a term can
// differ from the reference's by at most one unit in the last place -- and the kernels need no double-precision divide.
__device__ __forceinline__ void class_freqs(double p, double& maj_hom, double& maj_het, double& min_hom) {
  const double major = fmax(0.0, __dsub_rn(1.0, p));                      // :140
  const double minor = p;                                                  // sum_minor_freq <= 1 after the clamp (:143-151)
  double mh = __dmul_rn(minor, minor);                                     // :158 with inbreeding = 0
  double Mh = __dmul_rn(major, major);                                     // :176
  double Mt = __dmul_rn(__dmul_rn(2.0, major), minor);                     // :181
  const double sum = __dadd_rn(__dadd_rn(Mh, Mt), mh);                     // sumFrequencies() order: MH + Mt + mh + mt(=0)
  const double r = __dsub_rn(2.0, sum);
  maj_hom = __dmul_rn(Mh, r);
  maj_het = __dmul_rn(Mt, r);
  min_hom = __dmul_rn(mh, r);
}

constexpr double kMinMajorFreq = 0.01;     // minimum_major_frequency, freq.cpp:532
constexpr double kRitlandMinFreq = 0.001;  // minimum_frequency, calc.cpp:381
constexpr double kSmallProb = 1e-10;       // small_prob, calc.cpp:98

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

}  // namespace kgl
