// stream_count.cuh -- K1+K2: the fused streaming pass over the loci-major 2-bit genotype matrix (sm_100a).
//
// One read of the matrix (16 B per 64 genomes per locus) yields
//   * per-locus allele counts {n0,n1,n2,n3}             -> VariantDBVariant::summaryByVariant (kgl_variant_db_variant.cpp:126)
//   * per-genome counts of set lo / hi bits over the rows selected for the genome's super-population, as bit-sliced
//     vertical counters                                   -> generateFrequencies' class counts (kga_analysis_inbreed_freq.cpp:559-577)
//     (raw mode: over all rows                            -> VariantDBVariant::summaryByGenome, :180)
// Everything that is sparse -- code-3 cells and rows whose major allele is rare -- is handled by the small kernels in
// sparse_events.cuh from a pre-built index, so this kernel issues no per-genotype instruction at all.
//
// Structure: persistent CTAs (one per SM), each owning a contiguous range of rows. A producer warp streams the range
// through a ring of shared-memory stages with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx); consumer
// thread (tx, ty) owns unit column tx of its slice and rows {i*TY + ty} of every stage:
//   horizontal: 6 POPC per unit-row, REDUX.SUM over the lanes that share the row, shared-memory row accumulators
//   vertical  : Harley-Seal carry-save tree (2.25 LOP3 per 32-bit plane word) into 10-level bit-sliced counters
// Instruction budget per unit-row (64 genotypes): ~26 issue slots, ~15 on the ALU pipe, 6 POPC -- under what the HBM
// roofline allows (1.45 unit-rows/clk/SM at 6.5 TB/s leaves 44 ALU-pipe slots and 11 POPC slots per unit-row).
#pragma once
#include "common.cuh"

namespace kgl {

constexpr int kStreamU = 8;                 // rows per consumer thread per stage
constexpr int kStreamLevels = 10;           // bit-sliced counter depth: up to 1023 rows per thread per chunk
constexpr int kStreamItersPerChunk = 127;   // 127 * 8 = 1016 rows per thread between counter flushes
constexpr int kStreamMaxStages = 8;
constexpr int kStreamPlaneWords = 2 * 2 * kStreamLevels;   // per (vchunk, unit): [plane lo|hi][half][level]

struct StreamParams {
  const uint4* packed;        // [n_rows_padded][units]; rows >= n_loci are zero
  uint64_t units;             // 128-bit units per row
  uint64_t n_loci;
  uint32_t n_genomes;
  // Column decomposition of a slice (blockIdx.y): `blocks_per_slice` full column blocks of 32 units, and -- in the last
  // slice -- a remainder block of rem_units (< 32) units handled by warps whose lanes are (row lane, unit) pairs with
  // rem_w (a power of two >= rem_units) units per row lane.
  int blocks_per_slice;       // full 32-unit blocks per slice (the last slice may own fewer)
  int n_full_blocks;          // units / 32
  int rem_units, rem_w;
  int tyw;                    // row lanes: a stage is 8*tyw rows; thread rows are i*tyw + rl, i = 0..7
  int n_stages;               // ring depth
  uint32_t stages_per_cta;    // consecutive stages owned by one CTA
  uint32_t total_stages;      // ceil(n_loci / (8*tyw))
  const uint16_t* flags16;    // [n_rows_padded] bit k: row selected for population k; null = raw mode (every row counts)
  const uint8_t* unit_need;   // [units] populations present in the unit (bit set), raw mode: unused
  const uint64_t* popmask;    // [n_pop][units] genomes of population k inside the unit
  int n_pop;
  int multi_slice;            // gridDim.y > 1: per-locus counts are combined with global atomics
  uint32_t* locus_counts;     // [n_loci][4] or null
  uint32_t* planes;           // [vchunks][units][kStreamPlaneWords] or null; vchunk = (cta*chunks_per_cta + chunk)*tyw + row lane
  uint32_t chunks_per_cta;
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
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int n_threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

// ---- bit-sliced vertical counter ---------------------------------------------------------------------------------------
struct VCount {
  uint32_t c[kStreamLevels];
  uint32_t p8, p16;           // pending carries of weight 8 / 16, folded in on every second / fourth call
};
__device__ __forceinline__ void vc_clear(VCount& v) {
#pragma unroll
  for (int i = 0; i < kStreamLevels; ++i) v.c[i] = 0;
  v.p8 = 0; v.p16 = 0;
}
template <int FROM>
__device__ __forceinline__ void vc_ripple(VCount& v, uint32_t carry) {
#pragma unroll
  for (int lv = FROM; lv < kStreamLevels; ++lv) {
    const uint32_t t = v.c[lv] & carry;
    v.c[lv] ^= carry;
    carry = t;
  }
}
// Adds eight plane words. j = call index inside the chunk (warp-uniform).
__device__ __forceinline__ void vc_add8(VCount& v, const uint32_t (&x)[kStreamU], int j) {
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

// Dynamic shared memory layout (16-byte aligned pieces):
//   uint4    stage[n_stages][R][slice_units]
//   uint16_t flags[n_stages][R]
//   uint2    part[n_stages][n_contrib][R]   per-row partial counts of every column block, one slot per ring stage
//   uint64_t full[kStreamMaxStages], done[kStreamMaxStages], empty[kStreamMaxStages]
__host__ __device__ inline size_t stream_smem_bytes(int slice_units, int tyw, int n_stages, int n_contrib) {
  const size_t R = (size_t)kStreamU * tyw;
  return (size_t)n_stages * R * slice_units * 16 + (size_t)n_stages * R * 2 + (size_t)n_stages * n_contrib * R * 8 +
         3 * kStreamMaxStages * 8 + 64;
}

struct StreamCounters { VCount lo0, lo1, hi0, hi1; };

// One stage of one consumer thread: horizontal popcounts (reduced over the lanes that share a row, stored to part[] by
// the row leaders) and the vertical carry-save adds. REM: remainder warp (rem_w lanes per row lane). Kept as one
// straight-line block so that ptxas interleaves the POPC/REDUX chain with the LOP3 tree.
template <bool WANT_LOCUS, bool WANT_GENOME, bool REM, bool MASKED>
__device__ __forceinline__ void stream_consume(const uint4 (&v)[kStreamU], const uint32_t (&mlo)[kStreamU], const uint32_t (&mhi)[kStreamU],
                                               StreamCounters& C, int j, int rem_w, bool row_leader, uint2* part, int tyw) {
  if (WANT_LOCUS) {
    uint32_t sab[kStreamU], sc[kStreamU];
#pragma unroll
    for (int i = 0; i < kStreamU; ++i) {
      const uint32_t a = __popc(v[i].x) + __popc(v[i].y);                       // set lo bits: n1 + n3
      const uint32_t b = __popc(v[i].z) + __popc(v[i].w);                       // set hi bits: n2 + n3
      sc[i] = __popc(v[i].x & v[i].z) + __popc(v[i].y & v[i].w);                // both: n3
      sab[i] = a | (b << 16);                                                   // 32 lanes * 64 genomes: no carry between fields
    }
    if (!REM) {
#pragma unroll
      for (int i = 0; i < kStreamU; ++i) {
        sab[i] = __reduce_add_sync(kFull, sab[i]);
        sc[i] = __reduce_add_sync(kFull, sc[i]);
      }
    } else {
      for (int o = rem_w >> 1; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < kStreamU; ++i) {
          sab[i] += __shfl_xor_sync(kFull, sab[i], o);
          sc[i] += __shfl_xor_sync(kFull, sc[i], o);
        }
      }
    }
    if (row_leader) {
#pragma unroll
      for (int i = 0; i < kStreamU; ++i) part[i * tyw] = make_uint2(sab[i], sc[i]);
    }
  }
  if (WANT_GENOME) {
    uint32_t x[kStreamU];
#pragma unroll
    for (int i = 0; i < kStreamU; ++i) x[i] = MASKED ? (v[i].x & mlo[i]) : v[i].x;
    vc_add8(C.lo0, x, j);
#pragma unroll
    for (int i = 0; i < kStreamU; ++i) x[i] = MASKED ? (v[i].y & mhi[i]) : v[i].y;
    vc_add8(C.lo1, x, j);
#pragma unroll
    for (int i = 0; i < kStreamU; ++i) x[i] = MASKED ? (v[i].z & mlo[i]) : v[i].z;
    vc_add8(C.hi0, x, j);
#pragma unroll
    for (int i = 0; i < kStreamU; ++i) x[i] = MASKED ? (v[i].w & mhi[i]) : v[i].w;
    vc_add8(C.hi1, x, j);
  }
}

// WANT_LOCUS requires P.locus_counts, WANT_GENOME requires P.planes.
template <bool WANT_LOCUS, bool WANT_GENOME, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
k_stream_count(const StreamParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int TYW = P.tyw, S = P.n_stages;
  const int R = kStreamU * TYW;
  // this slice's columns
  const int blk0 = (int)blockIdx.y * P.blocks_per_slice;
  const int n_full = max(0, min(P.blocks_per_slice, P.n_full_blocks - blk0));
  const bool has_rem = (P.rem_units > 0) && (blockIdx.y == gridDim.y - 1);
  const int slice_units = n_full * 32 + (has_rem ? P.rem_units : 0);
  const int slice_units_max = P.blocks_per_slice * 32 + P.rem_units;   // smem row pitch (same for every slice)
  const int n_rem_warps = has_rem ? (TYW * P.rem_w) / 32 : 0;
  const int n_cons_warps = n_full * TYW + n_rem_warps;
  const int n_contrib_max = P.blocks_per_slice + (P.rem_units > 0 ? 1 : 0);
  const int n_contrib = n_full + (has_rem ? 1 : 0);
  const uint64_t slice_unit0 = (uint64_t)blk0 * 32;

  uint4* s_stage = reinterpret_cast<uint4*>(smem_raw);
  uint16_t* s_flags = reinterpret_cast<uint16_t*>(s_stage + (size_t)S * R * slice_units_max);
  uint2* s_part = reinterpret_cast<uint2*>(s_flags + (size_t)S * R);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(s_part + (size_t)S * n_contrib_max * R) + 15) & ~(uintptr_t)15);
  const uint32_t bar_full = smem_u32(s_bar), bar_done = smem_u32(s_bar + kStreamMaxStages), bar_empty = smem_u32(s_bar + 2 * kStreamMaxStages);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool raw = P.flags16 == nullptr;
  const uint32_t stage0 = blockIdx.x * P.stages_per_cta;
  const uint32_t stage_end = min(stage0 + P.stages_per_cta, P.total_stages);
  const uint32_t n_iters = stage_end > stage0 ? stage_end - stage0 : 0;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_done + 8 * s, n_cons_warps);
      mbar_init(bar_empty + 8 * s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == n_cons_warps) {
    // ===== producer warp: 1-D TMA bulk copies into the stage ring =====
    const bool contiguous = (slice_units == (int)P.units);
    const uint32_t row_bytes_slice = (uint32_t)slice_units * 16;
    for (uint32_t it = 0; it < n_iters; ++it) {
      const int s = it % S;
      if (it >= (uint32_t)S) mbar_wait(bar_empty + 8 * s, ((it / S) & 1) ^ 1);
      const uint64_t r0 = (uint64_t)(stage0 + it) * R;
      const uint32_t fbytes = raw ? 0u : (uint32_t)R * 2;
      const uint32_t dst = smem_u32(s_stage + (size_t)s * R * slice_units_max);
      if (lane == 0) {
        mbar_expect_tx(bar_full + 8 * s, (uint32_t)R * row_bytes_slice + fbytes);
        if (!raw) tma_bulk_g2s(smem_u32(s_flags + (size_t)s * R), P.flags16 + r0, fbytes, bar_full + 8 * s);
        if (contiguous) tma_bulk_g2s(dst, P.packed + r0 * P.units, (uint32_t)R * row_bytes_slice, bar_full + 8 * s);
      }
      __syncwarp();
      if (!contiguous) {
        for (int r = lane; r < R; r += 32)
          tma_bulk_g2s(dst + (uint32_t)r * slice_units_max * 16, P.packed + (r0 + r) * P.units + slice_unit0, row_bytes_slice,
                       bar_full + 8 * s);
      }
    }
    return;
  }
  if (warp == n_cons_warps + 1) {
    // ===== epilogue warp: waits until every consumer warp is done with a stage, publishes the per-locus counts of its
    // rows and hands the stage slot back to the producer =====
    for (uint32_t it = 0; it < n_iters; ++it) {
      const int s = it % S;
      mbar_wait(bar_done + 8 * s, (it / S) & 1);
      if (WANT_LOCUS) {
        const uint64_t r0 = (uint64_t)(stage0 + it) * R;
        for (int rr = lane; rr < R; rr += 32) {
          uint32_t sab = 0, sc = 0;
          for (int cb = 0; cb < n_contrib; ++cb) {
            const uint2 pc = s_part[((size_t)s * n_contrib_max + cb) * R + rr];
            sab += pc.x; sc += pc.y;
          }
          const uint64_t r = r0 + rr;
          if (r < P.n_loci) {
            const uint32_t a = sab & 0xFFFFu, b = sab >> 16;
            uint32_t* out = P.locus_counts + r * 4;
            if (!P.multi_slice) {
              *reinterpret_cast<uint4*>(out) = make_uint4(P.n_genomes - a - b + sc, a - sc, b - sc, sc);
            } else {
              atomicAdd(out + 1, a - sc); atomicAdd(out + 2, b - sc); atomicAdd(out + 3, sc);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty + 8 * s);
    }
    return;
  }
  if (warp > n_cons_warps + 1) return;

  // ===== consumers =====
  // full-block warp: lane = unit inside the block, one row lane per warp; remainder warp: lane = (row lane, unit)
  const bool rem_warp = warp >= n_full * TYW;
  int col, rl;
  bool active = true;
  if (!rem_warp) {
    col = (warp / TYW) * 32 + lane;
    rl = warp % TYW;
  } else {
    const int per = 32 / P.rem_w;                       // row lanes per remainder warp
    col = n_full * 32 + (lane % P.rem_w);
    rl = (warp - n_full * TYW) * per + lane / P.rem_w;
    active = (lane % P.rem_w) < P.rem_units;
  }
  const int contrib = rem_warp ? n_full : warp / TYW;    // which column block of the slice this warp works on
  const uint64_t unit = slice_unit0 + col;
  const bool row_leader = active && (rem_warp ? (lane % P.rem_w) == 0 : lane == 0);
  const uint32_t need = (!raw && active) ? P.unit_need[unit] : 0u;

  StreamCounters C;
  vc_clear(C.lo0); vc_clear(C.lo1); vc_clear(C.hi0); vc_clear(C.hi1);

  for (uint32_t chunk = 0; chunk < P.chunks_per_cta; ++chunk) {
    const uint32_t it_end = min(n_iters, (chunk + 1) * (uint32_t)kStreamItersPerChunk);
    int j = 0;           // iteration inside the current counter chunk
    for (uint32_t it = chunk * kStreamItersPerChunk; it < it_end; ++it, ++j) {
      const int s = it % S;
      mbar_wait(bar_full + 8 * s, (it / S) & 1);

      uint4 v[kStreamU];
      uint32_t f[kStreamU];
      const uint4* st = s_stage + (size_t)s * R * slice_units_max + col;
      const uint16_t* fl = s_flags + (size_t)s * R;
      bool allfull = true;
#pragma unroll
      for (int i = 0; i < kStreamU; ++i) {
        const int r = i * TYW + rl;
        v[i] = active ? st[r * slice_units_max] : make_uint4(0, 0, 0, 0);
        f[i] = (raw || !active) ? 0u : ((uint32_t)fl[r] & need);
        allfull = allfull && (f[i] == need);
      }
      uint2* part = s_part + ((size_t)s * n_contrib_max + contrib) * R + rl;
      uint32_t mlo[kStreamU], mhi[kStreamU];
      if (!WANT_GENOME || allfull) {
        if (!rem_warp) stream_consume<WANT_LOCUS, WANT_GENOME, false, false>(v, mlo, mhi, C, j, 32, row_leader, part, TYW);
        else stream_consume<WANT_LOCUS, WANT_GENOME, true, false>(v, mlo, mhi, C, j, P.rem_w, row_leader, part, TYW);
      } else {
#pragma unroll
        for (int i = 0; i < kStreamU; ++i) {
          if (f[i] == need) { mlo[i] = 0xFFFFFFFFu; mhi[i] = 0xFFFFFFFFu; }
          else if (f[i] == 0) { mlo[i] = 0; mhi[i] = 0; }
          else {
            uint64_t m = 0;
            for (int k = 0; k < P.n_pop; ++k)
              if ((f[i] >> k) & 1u) m |= P.popmask[(uint64_t)k * P.units + unit];
            mlo[i] = (uint32_t)m; mhi[i] = (uint32_t)(m >> 32);
          }
        }
        if (!rem_warp) stream_consume<WANT_LOCUS, WANT_GENOME, false, true>(v, mlo, mhi, C, j, 32, row_leader, part, TYW);
        else stream_consume<WANT_LOCUS, WANT_GENOME, true, true>(v, mlo, mhi, C, j, P.rem_w, row_leader, part, TYW);
      }

      // stage consumed, partial counts written: tell the epilogue warp
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_done + 8 * s);
    }
    // every virtual chunk is written, also the ones that saw no rows
    if (WANT_GENOME) {
      vc_finish(C.lo0); vc_finish(C.lo1); vc_finish(C.hi0); vc_finish(C.hi1);
      if (active) {
        const uint64_t vchunk = ((uint64_t)blockIdx.x * P.chunks_per_cta + chunk) * TYW + rl;
        uint32_t* out = P.planes + (vchunk * P.units + unit) * kStreamPlaneWords;
#pragma unroll
        for (int lv = 0; lv < kStreamLevels; ++lv) {
          out[0 * kStreamLevels + lv] = C.lo0.c[lv];
          out[1 * kStreamLevels + lv] = C.lo1.c[lv];
          out[2 * kStreamLevels + lv] = C.hi0.c[lv];
          out[3 * kStreamLevels + lv] = C.hi1.c[lv];
        }
      }
      vc_clear(C.lo0); vc_clear(C.lo1); vc_clear(C.hi0); vc_clear(C.hi1);
    }
  }
}

// Expand the bit-sliced counters: gcounts[g] = {set lo bits, set hi bits} summed over the virtual chunks.
constexpr int kExpandGroup = 16;
__global__ void __launch_bounds__(256)
k_expand_planes(const uint32_t* __restrict__ planes, uint64_t n_vchunks, uint64_t units, uint64_t n_genomes_padded,
                uint32_t* __restrict__ gcounts /* [n_genomes_padded][2], zeroed */) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_genomes_padded) return;
  const uint64_t unit = g >> 6;
  const int h = (int)((g >> 5) & 1), bit = (int)(g & 31);
  const uint64_t vc0 = (uint64_t)blockIdx.y * kExpandGroup;
  const uint64_t vc1 = min(vc0 + (uint64_t)kExpandGroup, n_vchunks);
  uint32_t acc[2] = {0, 0};
  for (uint64_t vc = vc0; vc < vc1; ++vc) {
    const uint32_t* base = planes + (vc * units + unit) * kStreamPlaneWords;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      uint32_t c = 0;
#pragma unroll
      for (int lv = 0; lv < kStreamLevels; ++lv) c |= ((base[(p * 2 + h) * kStreamLevels + lv] >> bit) & 1u) << lv;
      acc[p] += c;
    }
  }
#pragma unroll
  for (int p = 0; p < 2; ++p) if (acc[p]) atomicAdd(&gcounts[g * 2 + p], acc[p]);
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
struct StreamPlan {
  int blocks_per_slice, n_full_blocks, rem_units, rem_w, tyw, n_stages, threads, slice_units_max;
  unsigned slices, n_ctas;
  uint32_t rows_per_stage, total_stages, stages_per_cta, chunks_per_cta;
  uint64_t n_vchunks, padded_rows;
  size_t smem;
};

inline int pow2_ceil_int(int v) { int p = 1; while (p < v) p <<= 1; return p; }

inline StreamPlan plan_stream(uint64_t units, uint64_t n_loci, int sm_count, int tyw_hint = 0, int stages_hint = 0) {
  StreamPlan p{};
  p.n_full_blocks = (int)(units / 32);
  p.rem_units = (int)(units % 32);
  p.blocks_per_slice = p.n_full_blocks < 8 ? p.n_full_blocks : 8;
  if (p.n_full_blocks == 0) {
    p.slices = 1;
  } else {
    p.slices = (unsigned)((p.n_full_blocks + p.blocks_per_slice - 1) / p.blocks_per_slice);
    // balance the full blocks over the slices
    p.blocks_per_slice = (int)((p.n_full_blocks + p.slices - 1) / p.slices);
  }
  int tyw;
  if (tyw_hint > 0) tyw = tyw_hint;
  else if (p.n_full_blocks == 0) tyw = 256 / pow2_ceil_int(p.rem_units);       // 8 remainder warps
  else tyw = 8 / p.blocks_per_slice;
  if (tyw < 1) tyw = 1;
  tyw = pow2_ceil_int(tyw);
  if (tyw > 32) tyw = 32;
  p.slice_units_max = p.blocks_per_slice * 32 + p.rem_units;
  while (tyw > 1 && (size_t)kStreamU * tyw * p.slice_units_max * 16 > 56 * 1024) tyw >>= 1;
  p.tyw = tyw;
  // remainder lanes: a power of two >= rem_units, and wide enough that tyw row lanes fill whole warps
  p.rem_w = 0;
  if (p.rem_units > 0) {
    p.rem_w = pow2_ceil_int(p.rem_units);
    if (p.rem_w * tyw < 32) p.rem_w = 32 / tyw;
  }
  p.rows_per_stage = (uint32_t)(kStreamU * tyw);
  const size_t stage_bytes = (size_t)p.rows_per_stage * p.slice_units_max * 16;
  int S = stages_hint > 0 ? stages_hint : (int)(190 * 1024 / stage_bytes);
  if (S > kStreamMaxStages) S = kStreamMaxStages;
  if (S < 2) S = 2;
  p.n_stages = S;
  const int cons_warps = p.blocks_per_slice * tyw + (p.rem_units > 0 ? (tyw * p.rem_w) / 32 : 0);
  p.threads = (cons_warps + 2) * 32;     // + producer warp + epilogue warp
  p.total_stages = (uint32_t)((n_loci + p.rows_per_stage - 1) / p.rows_per_stage);
  unsigned ctas_x = (unsigned)(sm_count / (int)p.slices);
  if (ctas_x < 1) ctas_x = 1;
  if (p.total_stages > 0 && ctas_x > p.total_stages) ctas_x = p.total_stages;
  p.stages_per_cta = (p.total_stages + ctas_x - 1) / ctas_x;
  if (p.stages_per_cta == 0) p.stages_per_cta = 1;
  p.n_ctas = (p.total_stages + p.stages_per_cta - 1) / p.stages_per_cta;
  if (p.n_ctas == 0) p.n_ctas = 1;
  p.chunks_per_cta = (p.stages_per_cta + kStreamItersPerChunk - 1) / kStreamItersPerChunk;
  p.n_vchunks = (uint64_t)p.n_ctas * p.chunks_per_cta * tyw;
  p.padded_rows = (uint64_t)p.total_stages * p.rows_per_stage;
  p.smem = stream_smem_bytes(p.slice_units_max, tyw, S, p.blocks_per_slice + (p.rem_units > 0 ? 1 : 0));
  return p;
}

inline void fill_stream_params(StreamParams& P, const StreamPlan& pl) {
  P.blocks_per_slice = pl.blocks_per_slice; P.n_full_blocks = pl.n_full_blocks; P.rem_units = pl.rem_units; P.rem_w = pl.rem_w;
  P.tyw = pl.tyw; P.n_stages = pl.n_stages; P.stages_per_cta = pl.stages_per_cta; P.total_stages = pl.total_stages;
  P.multi_slice = pl.slices > 1 ? 1 : 0; P.chunks_per_cta = pl.chunks_per_cta;
}

}  // namespace kgl
