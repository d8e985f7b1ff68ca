// stream_count_generic.cuh -- K1+K2 streaming pass for ARBITRARY row widths (runtime slice width / rows per stage).
// Same algorithm and outputs as stream_count.cuh, whose compile-time shapes cover the widths BASELINE.json names; this
// kernel is the fallback for every other width. Ten consumer warps take both roles in turn:
//   H (per locus) : thread = (row, part of the row's units); three units' plane words share one carry-save adder before POPC
//   V (per genome): thread = (32-bit plane word column, row lane); Harley-Seal tree into 12-level bit-sliced counters
#pragma once
#include "stream_common.cuh"

namespace kgl {

// popcount of three words with one carry-save adder in front: popc(a)+popc(b)+popc(c) = popc(l) + 2*popc(h)
__device__ __forceinline__ uint32_t popc3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t h, l;
  csa(h, l, a, b, c);
  return __popc(l) + 2u * __popc(h);
}

// WANT_LOCUS requires P.locus_counts, WANT_GENOME requires P.cta_counts.
template <bool WANT_LOCUS, bool WANT_GENOME>
__global__ void __launch_bounds__(kScThreads, 1)
k_stream_count_rt(const StreamParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t R = P.rows_per_stage, S = P.n_stages, SU = P.slice_units;
  const uint32_t unit0 = blockIdx.y * SU;                              // first unit of this slice
  const uint32_t su = min(SU, P.units - unit0);                        // units this slice owns
  uint4* s_stage = reinterpret_cast<uint4*>(smem_raw);
  uint16_t* s_flags = reinterpret_cast<uint16_t*>(s_stage + (size_t)S * R * SU);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(s_flags + (size_t)S * R) + 15) & ~(uintptr_t)15);
  const uint32_t bar_full = smem_u32(s_bar), bar_empty = smem_u32(s_bar + kScMaxStages);
  uint32_t* s_cnt = stream_smem_counts(s_bar);                          // [SU * 4][32] per-genome counts of this CTA

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool raw = P.flags16 == nullptr;
  const uint32_t stage_begin = blockIdx.x * P.stages_per_cta;
  const uint32_t stage_end = min(stage_begin + P.stages_per_cta, P.total_stages);
  const uint32_t n_iters = stage_end > stage_begin ? stage_end - stage_begin : 0;

  if (tid == 0) {
    for (uint32_t s = 0; s < S; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, kScConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (WANT_GENOME)
    for (uint32_t i = tid; i < SU * 4 * 32; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();

  if (warp == kScConsumerWarps) {
    // ===== producer warp: 1-D TMA bulk copies into the stage ring =====
    const bool contiguous = (su == P.units);
    const uint32_t row_bytes_slice = su * 16;
    const uint32_t fbytes = raw ? 0u : R * 2;
    uint32_t s = 0, ph = 0;
    for (uint32_t it = 0; it < n_iters; ++it) {
      if (it >= S) mbar_wait(bar_empty + 8 * s, ph ^ 1);
      const uint64_t r0 = (uint64_t)(stage_begin + it) * R;
      const uint32_t dst = smem_u32(s_stage + (size_t)s * R * SU);
      if (lane == 0) {
        mbar_expect_tx(bar_full + 8 * s, R * row_bytes_slice + fbytes);
        if (!raw) tma_bulk_g2s(smem_u32(s_flags + (size_t)s * R), P.flags16 + r0, fbytes, bar_full + 8 * s);
        if (contiguous) tma_bulk_g2s(dst, P.packed + r0 * P.units, R * row_bytes_slice, bar_full + 8 * s);
      }
      __syncwarp();
      if (!contiguous) {
        for (uint32_t r = lane; r < R; r += 32)
          tma_bulk_g2s(dst + r * SU * 16, P.packed + (r0 + r) * P.units + unit0, row_bytes_slice, bar_full + 8 * s);
      }
      if (++s == S) { s = 0; ph ^= 1; }
    }
    return;
  }

  // ===== consumers =====
  // H role: thread = (row, part)
  const bool h_active = WANT_LOCUS && tid < kScHThreads;
  const uint32_t h_row = tid / P.h_parts, h_part = tid % P.h_parts;
  const uint32_t h_u0 = min(h_part * P.h_units, su);
  const uint32_t h_cnt = min(h_u0 + P.h_units, su) - h_u0;
  const uint32_t h_rot = (h_cnt > 1) ? (h_row & 1u) : 0u;
  // V role: thread = (word column, row lane)
  const uint32_t W = su * 4;
  const uint32_t v_wcol = tid % (SU * 4), v_rl = tid / (SU * 4);
  const bool v_active = WANT_GENOME && v_wcol < W && v_rl < P.v_row_lanes;
  const uint32_t v_rows = R / P.v_row_lanes;                            // rows per V thread per stage, multiple of 8
  const uint32_t v_row0 = v_active ? v_rl * v_rows : 0;                // inactive threads still read (and discard) flags
  const uint32_t g32 = (unit0 + (v_wcol >> 2)) * 2 + (v_wcol & 1);       // this word's 32-genome group
  uint32_t need = 0, pm[kMaxPop];
#pragma unroll
  for (int k = 0; k < kMaxPop; ++k) pm[k] = 0;
  if (v_active && !raw) {
    need = P.need32[g32];
#pragma unroll
    for (int k = 0; k < kMaxPop; ++k)
      if (k < (int)P.n_pop) pm[k] = P.popmask32[(size_t)k * P.units * 2 + g32];
  }
  const uint32_t all_pops = (1u << P.n_pop) - 1u;

  VCount C;
  vc_clear(C);
  uint32_t calls = 0, stages_in_chunk = 0;
  uint32_t s = 0, ph = 0;

  // the counters hold up to 4,095 rows: every flush_stages stages (and at the end) they are added to the CTA's count array
  auto flush = [&]() {
    vc_finish(C);
    if (v_active) vc_flush_counts(C, s_cnt, v_wcol);
    vc_clear(C);
    calls = 0; stages_in_chunk = 0;
  };

  for (uint32_t it = 0; it < n_iters; ++it) {
    const uint32_t stage = stage_begin + it;
    uint32_t ssum = 0xFF00u | all_pops;
    if (WANT_GENOME && !raw) {
      uint32_t a = 0xFFu, o = 0;
      for (uint32_t q = 0; q < (R >> 6); ++q) {
        const uint32_t v = P.sum64[(size_t)stage * (R >> 6) + q];
        a &= v & 0xFFu; o |= v >> 8;
      }
      ssum = a | (o << 8);
    }
    mbar_wait(bar_full + 8 * s, ph);
    const uint4* st = s_stage + (size_t)s * R * SU;

    if (WANT_LOCUS) {
      // ---- H role: per-locus counts of set lo bits (A), set hi bits (B) and both (C) over this slice's units ----
      uint32_t A = 0, B = 0, Cc = 0;
      if (h_active && h_cnt > 0) {
        const uint4* sr = st + (size_t)h_row * SU + h_u0;
        uint32_t j = h_rot;
        uint32_t n = 0;
        for (; n + 3 <= h_cnt; n += 3) {
          const uint32_t j0 = j;  j = (j + 1 == h_cnt) ? 0 : j + 1;
          const uint32_t j1 = j;  j = (j + 1 == h_cnt) ? 0 : j + 1;
          const uint32_t j2 = j;  j = (j + 1 == h_cnt) ? 0 : j + 1;
          const uint4 a = sr[j0], b = sr[j1], c = sr[j2];
          A += popc3(a.x, b.x, c.x) + popc3(a.y, b.y, c.y);
          B += popc3(a.z, b.z, c.z) + popc3(a.w, b.w, c.w);
          Cc += popc3(a.x & a.z, b.x & b.z, c.x & c.z) + popc3(a.y & a.w, b.y & b.w, c.y & c.w);
        }
        for (; n < h_cnt; ++n) {
          const uint4 a = sr[j];
          j = (j + 1 == h_cnt) ? 0 : j + 1;
          A += __popc(a.x) + __popc(a.y);
          B += __popc(a.z) + __popc(a.w);
          Cc += __popc(a.x & a.z) + __popc(a.y & a.w);
        }
      }
      if (tid < kScHThreads) {           // whole warps: warps 0..7
        uint32_t ab = A | (B << 16);     // a slice holds <= 56 * 64 genomes: no carry between the fields
        for (uint32_t o = P.h_parts >> 1; o > 0; o >>= 1) {
          ab += __shfl_xor_sync(kFull, ab, o);
          Cc += __shfl_xor_sync(kFull, Cc, o);
        }
        const uint64_t r = (uint64_t)stage * R + h_row;
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
    }

    if (WANT_GENOME) {
      // ---- V role: per-genome bit-sliced counters over the selected rows ----
      if (stages_in_chunk == P.flush_stages) flush();
      ++stages_in_chunk;
      const uint32_t s_and = ssum & 0xFFu, s_or = ssum >> 8;
      if (s_or != 0) {                                                  // some row of the stage is selected for somebody
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(st) + v_wcol;
        const uint16_t* fl = s_flags + (size_t)s * R + v_row0;
        const bool unmasked = raw || ((s_and & all_pops) == all_pops);  // stage-uniform
        for (uint32_t g = 0; g < v_rows; g += 8) {
          uint32_t x[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] = v_active ? sw[(size_t)(v_row0 + g + i) * SU * 4] : 0u;
          if (!unmasked) {
            uint32_t fn[8];
            bool partial = false;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              fn[i] = (uint32_t)fl[g + i] & need;
              partial = partial || (fn[i] != 0 && fn[i] != need);
            }
            if (__any_sync(kFull, partial)) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                uint32_t m = 0;
#pragma unroll
                for (int k = 0; k < kMaxPop; ++k) m |= ((fn[i] >> k) & 1u) ? pm[k] : 0u;
                x[i] &= m;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) x[i] = (fn[i] == need && need != 0) ? x[i] : 0u;
            }
          }
          vc_add8(C, x, calls);
          ++calls;
        }
      }
    }

    // this warp is done reading the stage: hand the slot back to the producer
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + 8 * s);
    if (++s == S) { s = 0; ph ^= 1; }
  }

  if (WANT_GENOME) {
    flush();
    asm volatile("bar.sync 1, %0;" ::"n"(kScConsumerThreads) : "memory");    // the consumer warps only: the producer has left
    stream_store_counts(P, s_cnt, unit0, W, tid, kScConsumerThreads);
  }
}

}  // namespace kgl
