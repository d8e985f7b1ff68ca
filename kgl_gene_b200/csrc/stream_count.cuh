// stream_count.cuh -- K1+K2: the fused streaming pass over the loci-major 2-bit genotype matrix (sm_100a), for the
// row widths fixed at compile time: SU = 40 units (2,504 genomes: 1000 Genomes, BASELINE configs 2-4; also the slice
// width for very wide populations such as the 100k-genome config 5) and SU = 8 (500 genomes, the Pf7-shaped config 1).
//
// One read of the matrix (16 B per 64 genomes per locus) yields
//   * per-locus allele counts {n0,n1,n2,n3}             -> VariantDBVariant::summaryByVariant (kgl_variant_db_variant.cpp:126)
//   * per-genome counts of set lo / hi bits over the rows selected for the genome's super-population, as bit-sliced
//     vertical counters                                   -> generateFrequencies' class counts (kga_analysis_inbreed_freq.cpp:559-577)
//     (raw mode: over all rows                            -> VariantDBVariant::summaryByGenome, :180)
// Everything that is sparse -- code-3 cells and rows whose major allele is rare -- is handled by the small kernels in
// sparse_events.cuh, so this kernel issues no per-genotype instruction at all.
//
// Structure: persistent CTAs (one per SM), each owning a contiguous range of stages (R rows x SU units, 32-40 KB). One
// producer warp streams the range through a ring of shared-memory stages with TMA bulk copies (cp.async.bulk + mbarrier
// complete_tx). Role-specialised consumer warps read every stage from shared memory, fully unrolled:
//   H warps (8)   : thread = (row, part of the row's units). The 2*HU plane words of the thread go through a complete
//                   carry-save tree before any POPC: 20 words -> 14 CSA (2 LOP3 each) + 6 POPC. Parts of a row sit in
//                   adjacent lanes (shuffle combine); the row's {n0,n1,n2,n3} goes straight to HBM as one uint4.
//   V warps (SU*R/256): thread = (32-bit plane word column, row lane), 32 rows per stage: 32 words + 12-level bit-sliced
//                   counter -> 31 CSA + 7 half adders, branch free (2.4 LOP3 per word).
// Per stage of 2,560 unit-rows (SU 40): ~3,000 warp instructions, i.e. ~40% of the issue slots and ~47% of the ALU pipe
// at the HBM roofline (1.4 unit-rows/clk/SM). Measured pipe rates (tools/kbench.cu): LOP3 62.5, POPC 15.9 per clk per SM.
#pragma once
#include "stream_common.cuh"

namespace kgl {

// ---- compile-time carry-save trees ---------------------------------------------------------------------------------------
// One level: passes of independent triples until fewer than three words are left. w[0..N) in/out, carries appended at
// carry[C0..). All indices are compile-time constants after inlining, so the arrays live in registers.
template <int N, int C0>
__device__ __forceinline__ void csa_level(uint32_t* w, uint32_t* carry) {
  if constexpr (N >= 3) {
    constexpr int T = N / 3, r = N % 3;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      uint32_t h, l;
      csa(h, l, w[3 * t], w[3 * t + 1], w[3 * t + 2]);
      carry[C0 + t] = h;
      w[t] = l;
    }
#pragma unroll
    for (int i = 0; i < r; ++i) w[T + i] = w[3 * T + i];
    csa_level<T + r, C0 + T>(w, carry);
  }
}
__host__ __device__ constexpr int csa_carries(int n) { return n < 3 ? 0 : ((n % 2) ? (n - 1) / 2 : (n - 2) / 2); }
__host__ __device__ constexpr int csa_left(int n) { return n < 3 ? n : ((n % 2) ? 1 : 2); }

// Sum of the popcounts of N words: full carry-save reduction, then POPC of the <= 2 words left per level.
template <int N, int SHIFT>
__device__ __forceinline__ uint32_t popc_tree(uint32_t* w) {
  constexpr int NC = csa_carries(N), NL = csa_left(N);
  uint32_t carry[NC > 0 ? NC : 1];
  csa_level<N, 0>(w, carry);
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < NL; ++i) s += __popc(w[i]);
  s <<= SHIFT;
  if constexpr (NC > 0) s += popc_tree<NC, SHIFT + 1>(carry);
  return s;
}

// Adds N words of weight 2^LV into the bit-sliced counter: the counter word joins the level's tree, carries go up.
template <int N, int LV>
__device__ __forceinline__ void vc_add_level(uint32_t* w, VCount& v) {
  if constexpr (LV < kScLevels && N > 0) {
    constexpr int M = N + 1, NC = csa_carries(M), NL = csa_left(M);
    uint32_t arr[M], carry[NC + 1];
#pragma unroll
    for (int i = 0; i < N; ++i) arr[i] = w[i];
    arr[N] = v.c[LV];
    csa_level<M, 0>(arr, carry);
    if constexpr (NL == 2) {
      carry[NC] = arr[0] & arr[1];
      v.c[LV] = arr[0] ^ arr[1];
      vc_add_level<NC + 1, LV + 1>(carry, v);
    } else {
      v.c[LV] = arr[0];
      vc_add_level<NC, LV + 1>(carry, v);
    }
  }
}

constexpr int kCtHWarps = 8;
constexpr int kCtVRows = 32;                    // rows per V thread per stage
__host__ __device__ constexpr int ct_v_warps(int su, int r) { return su * r / 256; }
__host__ __device__ constexpr int ct_threads(int su, int r) { return (kCtHWarps + ct_v_warps(su, r) + 1) * 32; }

// P.slice_units == SU, P.rows_per_stage == R, P.units a multiple of SU (the device row pitch is padded by the upload).
template <int SU, int R, bool WANT_LOCUS, bool WANT_GENOME>
__global__ void __maxnreg__(64)
k_stream_count_ct(const __grid_constant__ StreamParams P) {
  constexpr int PARTS = kScHThreads / R;        // lanes that share a row in the H role
  constexpr int HU = SU / PARTS;                // units per H thread
  static_assert(SU % PARTS == 0 && R % 64 == 0 && (SU * R) % 256 == 0, "shape");
  constexpr int W = SU * 4;                     // 32-bit word columns of a stage
  constexpr int RL = R / kCtVRows;              // V row lanes
  constexpr int V_WARPS = ct_v_warps(SU, R);
  static_assert(W * RL == V_WARPS * 32, "V mapping");
  constexpr int N_CONSUMER_WARPS = kCtHWarps + V_WARPS;
  constexpr uint32_t STAGE_BYTES = R * SU * 16;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t S = P.n_stages;
  const uint32_t unit0 = blockIdx.y * SU;
  uint4* s_stage = reinterpret_cast<uint4*>(smem_raw);
  uint16_t* s_flags = reinterpret_cast<uint16_t*>(s_stage + (size_t)S * R * SU);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(s_flags + (size_t)S * R) + 15) & ~(uintptr_t)15);
  const uint32_t bar_full = smem_u32(s_bar), bar_empty = smem_u32(s_bar + kScMaxStages);
  uint32_t* s_cnt = stream_smem_counts(s_bar);                          // [W][32] per-genome counts of this CTA

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool raw = P.flags16 == nullptr;
  const uint32_t stage_begin = blockIdx.x * P.stages_per_cta;
  const uint32_t stage_end = min(stage_begin + P.stages_per_cta, P.total_stages);
  const uint32_t n_iters = stage_end > stage_begin ? stage_end - stage_begin : 0;

  if (tid == 0) {
    for (uint32_t s = 0; s < S; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, N_CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (WANT_GENOME)
    for (uint32_t i = tid; i < (uint32_t)W * 32; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();

  if (warp == N_CONSUMER_WARPS) {
    // ===== producer warp =====
    const bool contiguous = (P.units == (uint32_t)SU);
    const uint32_t fbytes = (raw || !WANT_GENOME) ? 0u : R * 2;
    uint32_t s = 0, ph = 0;
    for (uint32_t it = 0; it < n_iters; ++it) {
      if (it >= S) mbar_wait(bar_empty + 8 * s, ph ^ 1);
      const uint64_t r0 = (uint64_t)(stage_begin + it) * R;
      const uint32_t dst = smem_u32(s_stage + (size_t)s * R * SU);
      if (lane == 0) {
        mbar_expect_tx(bar_full + 8 * s, STAGE_BYTES + fbytes);
        if (fbytes) tma_bulk_g2s(smem_u32(s_flags + (size_t)s * R), P.flags16 + r0, fbytes, bar_full + 8 * s);
        if (contiguous) tma_bulk_g2s(dst, P.packed + r0 * P.units, STAGE_BYTES, bar_full + 8 * s);
        else if (P.use_tmap) tma_tensor2d_g2s(dst, &P.tmap, unit0 * 4, (uint32_t)r0, bar_full + 8 * s);
      }
      __syncwarp();
      if (!contiguous && !P.use_tmap) {
        for (uint32_t r = lane; r < (uint32_t)R; r += 32)
          tma_bulk_g2s(dst + r * SU * 16, P.packed + (r0 + r) * P.units + unit0, SU * 16, bar_full + 8 * s);
      }
      if (++s == S) { s = 0; ph ^= 1; }
    }
    return;
  }

  uint32_t s = 0, ph = 0;
  if (warp < kCtHWarps) {
    // ===== H warps: per-locus counts =====
    const uint32_t h_row = tid / PARTS, h_part = tid % PARTS;
    // Bank-conflict-free order: the eight lanes of a quarter warp must touch eight different 16-byte bank groups.
    constexpr bool POW2 = (HU & (HU - 1)) == 0 && PARTS == 1;
    const uint32_t rot = POW2 ? (h_row & (HU - 1)) : (h_row & 1u);
    for (uint32_t it = 0; it < n_iters; ++it) {
      mbar_wait(bar_full + 8 * s, ph);
      if (WANT_LOCUS) {
        const uint4* sr = s_stage + (size_t)s * R * SU + (size_t)h_row * SU + h_part * HU;
        uint32_t lo[2 * HU], hi[2 * HU], bo[2 * HU];
#pragma unroll
        for (int i = 0; i < HU; ++i) {
          uint4 a;
          if constexpr (POW2) a = sr[(i + rot) & (HU - 1)];
          else if (i < HU - 1) a = (sr + rot)[i];
          else a = sr[rot ? 0 : HU - 1];
          lo[2 * i] = a.x; lo[2 * i + 1] = a.y;
          hi[2 * i] = a.z; hi[2 * i + 1] = a.w;
          bo[2 * i] = a.x & a.z; bo[2 * i + 1] = a.y & a.w;
        }
        const uint32_t A = popc_tree<2 * HU, 0>(lo);
        const uint32_t B = popc_tree<2 * HU, 0>(hi);
        uint32_t Cc = popc_tree<2 * HU, 0>(bo);
        uint32_t ab = A | (B << 16);               // <= 40 * 64 genomes per slice: no carry between the fields
#pragma unroll
        for (int o = PARTS >> 1; o > 0; o >>= 1) {
          ab += __shfl_xor_sync(kFull, ab, o);
          Cc += __shfl_xor_sync(kFull, Cc, o);
        }
        const uint64_t r = (uint64_t)(stage_begin + it) * R + h_row;
        if (h_part == 0 && r < P.n_loci) {
          const uint32_t a = ab & 0xFFFFu, b = ab >> 16;
          uint32_t* out = P.locus_counts + r * 4;
          if (!P.multi_slice) {
            *reinterpret_cast<uint4*>(out) = make_uint4(P.n_genomes - a - b + Cc, a - Cc, b - Cc, Cc);
          } else {
            if (a - Cc) atomicAdd(out + 1, a - Cc);
            if (b - Cc) atomicAdd(out + 2, b - Cc);
            if (Cc) atomicAdd(out + 3, Cc);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty + 8 * s);
      if (++s == S) { s = 0; ph ^= 1; }
    }
    return;
  }

  // ===== V warps: per-genome bit-sliced counters =====
  const uint32_t vt = tid - kCtHWarps * 32;
  const uint32_t v_wcol = vt % W, v_rl = vt / W;
  const uint32_t g32 = (unit0 + (v_wcol >> 2)) * 2 + (v_wcol & 1);       // this word's 32-genome group
  uint32_t need = 0, pm[kMaxPop];
#pragma unroll
  for (int k = 0; k < kMaxPop; ++k) pm[k] = 0;
  if (WANT_GENOME && !raw) {
    need = P.need32[g32];
#pragma unroll
    for (int k = 0; k < kMaxPop; ++k)
      if (k < (int)P.n_pop) pm[k] = P.popmask32[(size_t)k * P.units * 2 + g32];
  }
  const uint32_t all_pops = (1u << P.n_pop) - 1u;
  VCount C;
  vc_clear(C);
  uint32_t stages_in_chunk = 0;

  // the counters hold up to 4,095 rows: every flush_stages stages (and at the end) they are added to the CTA's count array
  auto flush = [&]() {
    vc_flush_counts(C, s_cnt, v_wcol);
    vc_clear(C);
    stages_in_chunk = 0;
  };

  for (uint32_t it = 0; it < n_iters; ++it) {
    uint32_t s_and = all_pops, s_or = 0xFFu;
    if (WANT_GENOME && !raw) {
      s_and = 0xFFu; s_or = 0;
#pragma unroll
      for (int q = 0; q < R / 64; ++q) {
        const uint32_t v = P.sum64[(size_t)(stage_begin + it) * (R / 64) + q];
        s_and &= v & 0xFFu; s_or |= v >> 8;
      }
    }
    mbar_wait(bar_full + 8 * s, ph);
    if (WANT_GENOME) {
      if (stages_in_chunk == P.flush_stages) flush();
      ++stages_in_chunk;
      if (s_or != 0) {                                                   // some row of the stage is selected for somebody
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(s_stage + (size_t)s * R * SU) + (size_t)v_rl * kCtVRows * W + v_wcol;
        uint32_t x[kCtVRows];
#pragma unroll
        for (int i = 0; i < kCtVRows; ++i) x[i] = sw[i * W];
        if (!raw && (s_and & all_pops) != all_pops) {                    // stage-uniform: some rows are not selected for everybody
          const uint16_t* fl = s_flags + (size_t)s * R + v_rl * kCtVRows;
#pragma unroll
          for (int i = 0; i < kCtVRows; ++i) {
            const uint32_t fn = (uint32_t)fl[i] & need;
            uint32_t m = (fn == need) ? 0xFFFFFFFFu : 0u;
            if (fn != 0 && fn != need) {
              m = 0;
#pragma unroll
              for (int k = 0; k < kMaxPop; ++k) m |= ((fn >> k) & 1u) ? pm[k] : 0u;
            }
            x[i] &= m;
          }
        }
        vc_add_level<kCtVRows, 0>(x, C);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + 8 * s);
    if (++s == S) { s = 0; ph ^= 1; }
  }
  if (WANT_GENOME) {
    flush();
    asm volatile("bar.sync 1, %0;" ::"n"(V_WARPS * 32) : "memory");    // the V warps only: H warps and the producer have left
    stream_store_counts(P, s_cnt, unit0, W, vt, V_WARPS * 32);
  }
}

}  // namespace kgl
