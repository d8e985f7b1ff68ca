// stream_common.cuh -- shared pieces of the fused streaming pass: parameters, mbarrier/TMA helpers, the bit-sliced vertical
// counter, the launch plan and the small follow-up kernels. The streaming kernels themselves are in stream_count.cuh
// (compile-time shapes: the 1000 Genomes / Pf7 / biobank widths) and stream_count_generic.cuh (any width).
#pragma once
#include "common.cuh"

#include <cuda.h>       // CUtensorMap (the type only: cuTensorMapEncodeTiled is looked up through the runtime, no libcuda link)

namespace kgl {

constexpr int kScConsumerWarps = 10;
constexpr int kScConsumerThreads = kScConsumerWarps * 32;
constexpr int kScHThreads = 256;               // threads that take the H role (rows_per_stage * h_parts)
constexpr int kScThreads = kScConsumerThreads + 32;   // + producer warp
constexpr int kScLevels = 12;                  // bit-sliced counter depth: up to 4095 rows per thread between flushes
constexpr int kScMaxCalls = 511;               // vc_add8 calls between flushes (511 * 8 = 4088 rows)
constexpr int kScMaxStages = 8;
constexpr int kScMaxSliceUnits = 56;           // 64 rows * 56 units * 16 B = 56 KB per stage

struct StreamParams {
  // Wide populations (row pitch of several 40-unit slices): 2-D tensor map over the matrix as uint32[padded_rows][units * 4], box
  // = one stage of one slice (64 rows x 160 words): ONE cp.async.bulk.tensor per stage instead of 64 row copies of 640 bytes.
  alignas(64) CUtensorMap tmap;
  uint32_t use_tmap;
  const uint4* packed;        // [padded_rows][units]; rows >= n_loci are zero
  uint32_t units;             // 128-bit units per row
  uint32_t n_loci;
  uint32_t n_genomes;
  uint32_t slice_units;       // units per slice (blockIdx.y); the last slice may own fewer
  uint32_t rows_per_stage;    // R: 64, 128 or 256
  uint32_t h_parts;           // kScHThreads / R: lanes that share a row in the H role
  uint32_t h_units;           // ceil(slice_units / h_parts): units per H thread
  uint32_t v_row_lanes;       // RL: V threads per word column; each owns R / RL consecutive rows of a stage
  uint32_t n_stages;          // ring depth
  uint32_t stages_per_cta;    // consecutive stages owned by one CTA
  uint32_t total_stages;      // ceil(n_loci / R)
  uint32_t flush_stages;      // stages between counter flushes
  uint32_t chunks_per_cta;    // ceil(stages_per_cta / flush_stages)
  const uint16_t* flags16;    // [padded_rows] bit k: row selected for population k; null = raw mode (every row counts)
  const uint16_t* sum64;      // [padded_rows/64] per 64-row group, low byte: AND of the rows' selection bits, high byte: OR
  const uint32_t* popmask32;  // [n_pop][units*2] genomes of population k inside each 32-genome group
  const uint8_t* need32;      // [units*2] populations present in each 32-genome group
  uint32_t n_pop;
  uint32_t multi_slice;       // gridDim.y > 1: per-locus counts are combined with global atomics
  uint32_t* locus_counts;     // [n_loci][4] or null
  uint32_t* cta_counts;       // [gridDim.x][2][n_genomes_padded] or null: set lo / hi bits per genome over the CTA's selected rows
  uint32_t n_genomes_padded;  // units * 64
};

// ---- mbarrier / TMA bulk-copy helpers ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_tensor2d_g2s(uint32_t dst, const CUtensorMap* tmap, uint32_t c0, uint32_t c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ---- bit-sliced vertical counter ---------------------------------------------------------------------------------------
struct VCount {
  uint32_t c[kScLevels];
  uint32_t p8, p16;           // pending carries of weight 8 / 16, folded in on every second / fourth call
};
__device__ __forceinline__ void vc_clear(VCount& v) {
#pragma unroll
  for (int i = 0; i < kScLevels; ++i) v.c[i] = 0;
  v.p8 = 0; v.p16 = 0;
}
template <int FROM>
__device__ __forceinline__ void vc_ripple(VCount& v, uint32_t carry) {
#pragma unroll
  for (int lv = FROM; lv < kScLevels; ++lv) {
    const uint32_t t = v.c[lv] & carry;
    v.c[lv] ^= carry;
    carry = t;
  }
}
// Adds eight plane words. j = call index since the last flush (warp-uniform).
__device__ __forceinline__ void vc_add8(VCount& v, const uint32_t (&x)[8], uint32_t j) {
  uint32_t a, b, q0, q1, t8;
  csa(a, v.c[0], v.c[0], x[0], x[1]);
  csa(b, v.c[0], v.c[0], x[2], x[3]);
  csa(q0, v.c[1], v.c[1], a, b);
  csa(a, v.c[0], v.c[0], x[4], x[5]);
  csa(b, v.c[0], v.c[0], x[6], x[7]);
  csa(q1, v.c[1], v.c[1], a, b);
  csa(t8, v.c[2], v.c[2], q0, q1);
  if (j & 1) {
    uint32_t t16;
    csa(t16, v.c[3], v.c[3], v.p8, t8);
    v.p8 = 0;
    if (j & 2) {
      uint32_t t32;
      csa(t32, v.c[4], v.c[4], v.p16, t16);
      v.p16 = 0;
      vc_ripple<5>(v, t32);
    } else {
      v.p16 = t16;
    }
  } else {
    v.p8 = t8;
  }
}
__device__ __forceinline__ void vc_finish(VCount& v) {
  vc_ripple<3>(v, v.p8);
  vc_ripple<4>(v, v.p16);
  v.p8 = 0; v.p16 = 0;
}

// Dynamic shared memory: uint4 stage[n_stages][R][slice_units] ; uint16_t flags[n_stages][R] ; uint64_t full[8], empty[8] ;
// uint32_t cnt[slice_units * 4][32] (per-genome counts of the CTA, filled when the bit-sliced counters are flushed)
__host__ __device__ inline size_t stream_smem_bytes(uint32_t slice_units, uint32_t rows, uint32_t n_stages) {
  return (size_t)n_stages * rows * slice_units * 16 + (size_t)n_stages * rows * 2 + 2 * kScMaxStages * 8 + 64 + (size_t)slice_units * 512;
}
__device__ __forceinline__ uint32_t* stream_smem_counts(uint64_t* s_bar) { return reinterpret_cast<uint32_t*>(s_bar + 2 * kScMaxStages); }

// Flush of one thread's bit-sliced counter (32 genomes of one plane-word column, kScLevels = 12 bits each) into the CTA's
// count array: levels are gathered four at a time into nibbles (genome 4n + k sits in nibble n of pass k), so a genome's
// count is three nibbles -- ~350 instructions per flush instead of 3 x 384 single-bit moves.
// The count array is swizzled: counter `bit` of word column `col` lives at col * 32 + ((bit + col) & 31), so the 32 lanes of a
// warp (consecutive columns, same bit) hit 32 different banks instead of one.
__device__ __forceinline__ void vc_flush_counts(const VCount& v, uint32_t* __restrict__ s_cnt, uint32_t col) {
  static_assert(kScLevels == 12, "three nibbles per counter");
  uint32_t* cnt32 = s_cnt + col * 32;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint32_t A = 0, B = 0, Cg = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      A |= ((v.c[j] >> k) & 0x11111111u) << j;
      B |= ((v.c[4 + j] >> k) & 0x11111111u) << j;
      Cg |= ((v.c[8 + j] >> k) & 0x11111111u) << j;
    }
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const uint32_t c = ((A >> (4 * n)) & 15u) | (((B >> (4 * n)) & 15u) << 4) | (((Cg >> (4 * n)) & 15u) << 8);
      if (c) atomicAdd(&cnt32[(4 * n + k + col) & 31], c);
    }
  }
}

// End of a CTA: its count array goes to cta_counts[blockIdx.x][plane][genome]. Word column j of unit u holds plane j >> 1 of
// the genomes 64 u + 32 (j & 1) .. + 31. Called by `n_threads` threads (index t) after they have synchronised.
__device__ __forceinline__ void stream_store_counts(const StreamParams& P, const uint32_t* __restrict__ s_cnt, uint32_t unit0,
                                                    uint32_t n_cols, uint32_t t, uint32_t n_threads) {
  uint32_t* out = P.cta_counts + (size_t)blockIdx.x * 2 * P.n_genomes_padded;
  for (uint32_t i = t; i < n_cols * 32; i += n_threads) {
    const uint32_t col = i >> 5, bit = i & 31;
    const uint32_t unit = unit0 + (col >> 2), j = col & 3;
    out[(size_t)(j >> 1) * P.n_genomes_padded + unit * 64 + (j & 1) * 32 + bit] = s_cnt[col * 32 + ((bit + col) & 31)];
  }
}

// gcounts[g] = {set lo bits, set hi bits} summed over the CTAs (the tail kernel does this itself; used by the fallback path of
// populations whose code-3 cells are too many to index, and by tools/kbench).
__global__ void __launch_bounds__(256)
k_sum_cta_counts(const uint32_t* __restrict__ cta_counts, uint32_t n_ctas, uint64_t n_genomes_padded, uint32_t* __restrict__ gcounts) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_genomes_padded) return;
  uint32_t a = 0, b = 0;
  for (uint32_t c = 0; c < n_ctas; ++c) {
    a += cta_counts[((size_t)c * 2 + 0) * n_genomes_padded + g];
    b += cta_counts[((size_t)c * 2 + 1) * n_genomes_padded + g];
  }
  gcounts[g * 2 + 0] = a; gcounts[g * 2 + 1] = b;
}

// multi-slice only: n0 = N - n1 - n2 - n3
__global__ void k_fix_locus_n0(uint32_t* locus_counts, uint64_t n_loci, uint32_t n_genomes) {
  const uint64_t l = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (l < n_loci) {
    uint32_t* c = locus_counts + l * 4;
    c[0] = n_genomes - c[1] - c[2] - c[3];
  }
}

// ---- host-side launch plan ----------------------------------------------------------------------------------------------
// shape: 0 = runtime-shape kernel (stream_count_generic.cuh); 1 = compile-time <SU 40, R 64>; 2 = compile-time <SU 8, R 256>
struct StreamPlan {
  int shape;
  uint32_t units_padded;      // device row pitch in units (>= units; a multiple of slice_units for the compile-time shapes)
  uint32_t slice_units, slices, rows_per_stage, h_parts, h_units, v_row_lanes, n_stages, threads;
  uint32_t total_stages, stages_per_cta, n_ctas, flush_stages, chunks_per_cta;
  uint64_t padded_rows;
  size_t smem;
};

// Device row pitch for a host width of `units`: wide populations are cut into 40-unit slices, so their pitch is padded.
inline uint32_t stream_units_padded(uint64_t units) {
  return (uint32_t)(units > (uint64_t)kScMaxSliceUnits ? (units + 39) / 40 * 40 : units);
}

inline StreamPlan plan_stream(uint64_t units_padded, uint64_t n_loci, int sm_count, int rows_hint = 0, int stages_hint = 0,
                              bool allow_ct = true) {
  StreamPlan p{};
  p.units_padded = (uint32_t)units_padded;
  const uint64_t units = units_padded;
  uint32_t R = 64;
  if (allow_ct && rows_hint == 0 && units == 8) { p.shape = 2; p.slice_units = 8; p.slices = 1; R = 256; }
  else if (allow_ct && rows_hint == 0 && units % 40 == 0) { p.shape = 1; p.slice_units = 40; p.slices = (uint32_t)(units / 40); R = 64; }
  else {
    p.shape = 0;
    p.slices = (uint32_t)((units + kScMaxSliceUnits - 1) / kScMaxSliceUnits);     // as few as possible, balanced
    p.slice_units = (uint32_t)((units + p.slices - 1) / p.slices);
    p.slices = (uint32_t)((units + p.slice_units - 1) / p.slice_units);
    if (rows_hint > 0) R = (uint32_t)rows_hint;
    else if (p.slice_units <= 10) R = 256;
    else if (p.slice_units <= 20) R = 128;
  }
  p.rows_per_stage = R;
  p.h_parts = kScHThreads / R;
  p.h_units = (p.slice_units + p.h_parts - 1) / p.h_parts;
  // V role: word columns x row lanes <= consumer threads, rows per thread a multiple of 8
  const uint32_t W = p.slice_units * 4;
  uint32_t rl = 8;
  if (p.shape != 0) rl = R / 32;
  else while (rl > 1 && (W * rl > (uint32_t)kScConsumerThreads || (R / rl) % 8 != 0)) rl >>= 1;
  p.v_row_lanes = rl;
  p.threads = p.shape == 0 ? (uint32_t)kScThreads : (8 + p.slice_units * R / 256 + 1) * 32;
  const size_t stage_bytes = (size_t)R * p.slice_units * 16 + R * 2;
  int S = stages_hint > 0 ? stages_hint : (int)((200 * 1024 - (size_t)p.slice_units * 512) / stage_bytes);
  if (S > kScMaxStages) S = kScMaxStages;
  if (S < 2) S = 2;
  p.n_stages = (uint32_t)S;
  p.total_stages = (uint32_t)((n_loci + R - 1) / R);
  uint32_t ctas_x = (uint32_t)sm_count / p.slices;
  if (ctas_x < 1) ctas_x = 1;
  if (p.total_stages > 0 && ctas_x > p.total_stages) ctas_x = p.total_stages;
  p.stages_per_cta = (p.total_stages + ctas_x - 1) / ctas_x;
  if (p.stages_per_cta == 0) p.stages_per_cta = 1;
  p.n_ctas = (p.total_stages + p.stages_per_cta - 1) / p.stages_per_cta;
  if (p.n_ctas == 0) p.n_ctas = 1;
  p.flush_stages = (uint32_t)kScMaxCalls / ((R / rl) / 8);
  p.chunks_per_cta = (p.stages_per_cta + p.flush_stages - 1) / p.flush_stages;
  p.padded_rows = (uint64_t)p.total_stages * R;
  p.smem = stream_smem_bytes(p.slice_units, R, p.n_stages);
  return p;
}

inline void fill_stream_params(StreamParams& P, const StreamPlan& pl) {
  P.slice_units = pl.slice_units; P.rows_per_stage = pl.rows_per_stage; P.h_parts = pl.h_parts; P.h_units = pl.h_units;
  P.v_row_lanes = pl.v_row_lanes; P.n_stages = pl.n_stages; P.stages_per_cta = pl.stages_per_cta; P.total_stages = pl.total_stages;
  P.flush_stages = pl.flush_stages; P.chunks_per_cta = pl.chunks_per_cta; P.multi_slice = pl.slices > 1 ? 1 : 0;
}

}  // namespace kgl
