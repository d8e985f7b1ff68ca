// terms_moments_mma.cuh -- the sparse half of the moment tables (terms_moments.cuh) on the tensor cores.
//
// The per-genome fixed-point moments are a matrix product: M[g][j] = sum over the loci l of a bin of X[g][l] U[l][j], X the
// 0/1 indicator "genome g is not common-homozygous at l" (and, for the rare class, "is rare-homozygous"), U the locus'
// normalised powers. With U + 2^s split into six unsigned 8-bit limbs the product is exact on tcgen05.mma kind::i8 (int32
// accumulators in TMEM): 128 genomes x 32 payload columns {count, 5 moments x 6 limbs} x 32 loci per instruction. The payload
// operand is precomputed once per selection in the shared-memory layout the MMA reads (k_mom_btiles; one TMA bulk copy per
// stage); the indicator operand is expanded from the genotype bit planes stage by stage (one byte per cell, never in HBM).
//
// Work unit = (unit, tile of 128 genomes). A UNIT is a run of the population's frequency-sorted loci with one common bin, one
// rare bin and one pair of class codes, at most kMomChunk long (k_mom_bounds / k_mom_units): its accumulators live in TMEM from
// its first stage to its last and are flushed once -- limbs recombined to 64-bit integers, the 2^s offsets removed with the
// count column, integer atomics into the same table the CUDA-core builder (k_mom_build) fills. Same integers, same result.
#pragma once
#include "gram_i8.cuh"
#include "terms_moments.cuh"

namespace kgl {

constexpr int kMmaM = 128;                      // genomes per tile = TMEM lanes
constexpr int kMmaK = 128;                      // loci per stage = bytes per operand row (one 128-byte swizzle atom)
constexpr int kMmaN = 32;                       // payload columns: count, 5 moments x 6 limbs, one unused
constexpr int kMmaLimbs = 6;
constexpr uint32_t kMmaATile = kMmaM * kMmaK;   // 16 KB per class
constexpr uint32_t kMmaBTile = kMmaN * kMmaK;   // 4 KB per class; a stage's two payload tiles are 8 KB contiguous in global memory
constexpr uint32_t kMmaIdesc = (2u << 4) | ((uint32_t)(kMmaN >> 3) << 17) | ((uint32_t)(kMmaM >> 4) << 24);   // as kGramIdesc, N = 32
constexpr uint32_t kMomMaxUnits = 1u << 17;
static_assert(1 + (kMomJ - 1) * kMmaLimbs <= kMmaN, "payload columns");

struct MomUnit {
  uint32_t begin, end;       // sorted items [begin, end)
  uint32_t tile_base;        // index of the unit's first payload tile pair
  int16_t cb, rb;            // GLOBAL bins of the common / rare class (kMomBinsMax: a == 1), -1: the class has no terms
  int8_t pop, common_code, rare_code, pad;
};

__device__ __forceinline__ uint64_t mom_seg_key(const MomClass& m) {
  const uint64_t cb = m.common_code >= 0 ? (uint64_t)mom_bin(m.r_common) : 0xFFFFull;
  const uint64_t rb = m.rare_code >= 0 ? (uint64_t)mom_bin(m.r_rare) : 0xFFFFull;
  return cb | (rb << 16) | ((uint64_t)(m.common_code & 0xFF) << 32) | ((uint64_t)(m.rare_code & 0xFF) << 40);
}

// Unit boundaries: the first item of a population, every kMomChunk-th item, every change of (bins, codes).
__global__ void __launch_bounds__(256)
k_mom_bounds(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ rows, const uint32_t* __restrict__ pop_begin, int n_pop,
             const float* __restrict__ af, uint64_t n_loci, int unphased, uint32_t* __restrict__ bounds, uint32_t* __restrict__ n_bounds) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pop_begin[n_pop]) return;
  const int kk = (int)(keys[i] >> 32);
  const uint32_t first = pop_begin[kk];
  bool b = i == first || ((uint32_t)i - first) % kMomChunk == 0;
  if (!b) {
    const float* a = af + (uint64_t)kk * n_loci;
    b = mom_seg_key(mom_classify(a[rows[i]], unphased != 0)) != mom_seg_key(mom_classify(a[rows[i - 1]], unphased != 0));
  }
  if (b) { const uint32_t pos = atomicAdd(n_bounds, 1u); if (pos < kMomMaxUnits) bounds[pos] = (uint32_t)i; }
}

// Units from the sorted boundaries (one block). out = {units, payload tile pairs}; unit_range[k] = {first unit, past-the-last unit} of population k.
__global__ void __launch_bounds__(1024)
k_mom_units(const uint32_t* __restrict__ bounds, const uint32_t* __restrict__ n_bounds, const uint64_t* __restrict__ keys,
            const uint32_t* __restrict__ rows, const uint32_t* __restrict__ pop_begin, int n_pop, const float* __restrict__ af,
            uint64_t n_loci, int unphased, MomUnit* __restrict__ units, uint2* __restrict__ unit_range, uint32_t* __restrict__ out) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_run;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t n = min(*n_bounds, kMomMaxUnits), total = pop_begin[n_pop];
  if (threadIdx.x == 0) s_run = 0;
  if ((int)threadIdx.x < kMaxPop) unit_range[threadIdx.x] = make_uint2(0u, 0u);
  __syncthreads();
  for (uint32_t u0 = 0; u0 < n; u0 += 1024) {
    const uint32_t u = u0 + threadIdx.x;
    uint32_t begin = 0, end = 0, stages = 0;
    if (u < n) { begin = bounds[u]; end = u + 1 < n ? bounds[u + 1] : total; stages = (end - begin + kMmaK - 1) / kMmaK; }
    uint32_t incl = stages;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t before = s_run;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    if (u < n) {
      const int kk = (int)(keys[begin] >> 32);
      const MomClass m = mom_classify(af[(uint64_t)kk * n_loci + rows[begin]], unphased != 0);
      MomUnit r;
      r.begin = begin; r.end = end; r.tile_base = before + incl - stages;
      r.cb = m.common_code >= 0 ? (int16_t)mom_bin(m.r_common) : (int16_t)-1;
      r.rb = m.rare_code >= 0 ? (int16_t)mom_bin(m.r_rare) : (int16_t)-1;
      r.pop = (int8_t)kk; r.common_code = (int8_t)m.common_code; r.rare_code = (int8_t)m.rare_code; r.pad = 0;
      units[u] = r;
      if (begin == pop_begin[kk]) unit_range[kk].x = u;
      if (end == pop_begin[kk + 1]) unit_range[kk].y = u + 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < 32; ++w) t += s_warp[w]; s_run += t; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = n; out[1] = s_run; out[2] = *n_bounds; }
}

// Byte offset of (row n, byte k) inside a K-major operand tile with the 128-byte swizzle (rows of 128 B, 8-row groups of 1 KB).
__device__ __forceinline__ uint32_t mma_tile_offset(uint32_t n, uint32_t k) {
  return (n >> 3) * 1024u + (n & 7u) * 128u + ((((k >> 4) ^ (n & 7u)) & 7u) << 4) + (k & 15u);
}

// Payload tiles of a unit, stage by stage: [common 4 KB | rare 4 KB], column t of a tile = the limbs of item begin + 128 stage + t.
__global__ void __launch_bounds__(kMmaK)
k_mom_btiles(const MomUnit* __restrict__ units, const uint32_t* __restrict__ rows, const float* __restrict__ af, uint64_t n_loci,
             int unphased, double scale, unsigned char* __restrict__ btiles, double* __restrict__ rr /* r of the rare class per sorted item */,
             int b_lo, int nbt, long long* __restrict__ pm /* [n_pop][nbt][kMomJ]: the population's moments (what k_mom_dense computes) */) {
  __shared__ long long s_pm[kMmaK / 32][kMomJ];
  const MomUnit U = units[blockIdx.x];
  const uint32_t n_stages = (U.end - U.begin + kMmaK - 1) / kMmaK;
  const long long offset = (long long)scale;
  long long dense[kMomJ];
#pragma unroll
  for (int j = 0; j < kMomJ; ++j) dense[j] = 0;
  for (uint32_t s = 0; s < n_stages; ++s) {
    const uint32_t i = U.begin + s * kMmaK + threadIdx.x;
    unsigned char* tile = btiles + (size_t)(U.tile_base + s) * (2 * kMmaBTile);
    unsigned char col[2][kMmaN];
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
      for (int n = 0; n < kMmaN; ++n) col[x][n] = 0;
    if (i < U.end) {
      const MomClass m = mom_classify(af[(uint64_t)U.pop * n_loci + rows[i]], unphased != 0);
      rr[i] = m.r_rare;
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        const int code = x == 0 ? m.common_code : m.rare_code;
        if (code < 0) continue;
        const double r = x == 0 ? m.r_common : m.r_rare;
        long long u[kMomJ - 1];
        mom_powers(r, mom_bin(r), scale, u);
        if (x == 0) {
          dense[0] += 1;
#pragma unroll
          for (int j = 0; j < kMomJ - 1; ++j) dense[j + 1] += u[j];
        }
        col[x][0] = 1;
#pragma unroll
        for (int j = 0; j < kMomJ - 1; ++j) {
          const unsigned long long v = (unsigned long long)(u[j] + offset);          // in [0, 2^(s+1)]
#pragma unroll
          for (int b = 0; b < kMmaLimbs; ++b) col[x][1 + j * kMmaLimbs + b] = (unsigned char)(v >> (8 * b));
        }
      }
    }
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
      for (int n = 0; n < kMmaN; ++n) tile[x * kMmaBTile + mma_tile_offset(n, threadIdx.x)] = col[x][n];
  }
  // the unit's share of the population's moments: every item of a unit has the unit's common bin
  if (U.cb < 0) return;
#pragma unroll
  for (int j = 0; j < kMomJ; ++j)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dense[j] += __shfl_xor_sync(kFull, dense[j], o);
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int j = 0; j < kMomJ; ++j) s_pm[threadIdx.x >> 5][j] = dense[j];
  __syncthreads();
  if (threadIdx.x < kMomJ) {
    long long t = 0;
    for (int w = 0; w < kMmaK / 32; ++w) t += s_pm[w][threadIdx.x];
    const int bin = U.cb >= kMomBinsMax ? nbt - 1 : U.cb - b_lo;
    atomicAdd(reinterpret_cast<unsigned long long*>(pm) + ((size_t)U.pop * nbt + bin) * kMomJ + threadIdx.x, (unsigned long long)t);
  }
}

struct MomMmaParams {
  const uint4* packed; uint64_t units;               // loci-major matrix (row = locus, unit = 64 genomes)
  const uint8_t* superpop; uint64_t n_genomes, n_genomes_padded;
  const uint32_t* rows;                              // sorted items -> rows
  const MomUnit* unit_table;
  const unsigned char* btiles;
  int b_lo, nbt; double scale;
  long long* mi;                                     // [n_genomes_padded][nbt][kMomJ]
  uint32_t* cnt;                                     // [n_units][n_genomes_padded] rare homozygous cells (null: not wanted)
  uint32_t tile_lo[kMaxPop], tile_hi[kMaxPop];       // 128-genome tiles that hold a genome of the population
  uint32_t tiles_per_unit;                           // the widest of those ranges: grid = n_units * tiles_per_unit blocks
  uint32_t* rare_bits;                               // [payload tile pairs][tiles_per_unit][4]: loci of the stage at which a genome of the
                                                     // tile has a rare homozygous cell (k_mom_unit_fill loads only those rows); may be null
};

__device__ __forceinline__ void mma_i8_n32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kMmaIdesc), "r"(accumulate), "r"(0u) : "memory");
}

// 32 x 32 bit transpose across a warp: lane i holds row i, afterwards lane i holds column i (bit k = bit i of row k).
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, uint32_t lane) {
#pragma unroll
  for (uint32_t j = 16; j > 0; j >>= 1) {
    const uint32_t m = j == 16 ? 0x0000FFFFu : j == 8 ? 0x00FF00FFu : j == 4 ? 0x0F0F0F0Fu : j == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t y = __shfl_xor_sync(kFull, x, j);
    x = (lane & j) ? ((x & ~m) | ((y >> j) & m)) : ((x & m) | ((y << j) & ~m));
  }
  return x;
}

// One operand row (genome = thread) of an indicator tile from the stage's masks: byte k = bit `lane` of masks[k]. Per 32 loci the
// warp transposes the 32 mask words (lane k loads locus k) and every lane spreads its 32 bits to 32 bytes (4 bits -> 4 bytes:
// (nibble * 0x00204081) & 0x01010101).
__device__ __forceinline__ void mma_build_row(unsigned char* tile, const uint32_t* masks /* [128] of this warp's slice */, uint32_t m,
                                              uint32_t lane, uint32_t keep /* all ones, or 0 for a genome of another population */) {
  unsigned char* row = tile + (m >> 3) * 1024u + (m & 7u) * 128u;
  const uint32_t sw = m & 7u;
#pragma unroll
  for (uint32_t q = 0; q < 4; ++q) {                              // 32 loci = two 16-byte chunks
    const uint32_t t = warp_transpose32(masks[q * 32 + lane], lane) & keep;
    uint32_t o[8];
#pragma unroll
    for (uint32_t n = 0; n < 8; ++n) o[n] = (((t >> (4 * n)) & 0xFu) * 0x00204081u) & 0x01010101u;
    *reinterpret_cast<uint4*>(row + (((2 * q) ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<uint4*>(row + (((2 * q + 1) ^ sw) << 4)) = make_uint4(o[4], o[5], o[6], o[7]);
  }
}
__device__ __forceinline__ void mma_zero_row(unsigned char* tile, uint32_t m) {
  uint4* row = reinterpret_cast<uint4*>(tile + (m >> 3) * 1024u + (m & 7u) * 128u);
#pragma unroll
  for (uint32_t c = 0; c < 8; ++c) row[c] = make_uint4(0u, 0u, 0u, 0u);
}

__global__ void __launch_bounds__(kMmaM)
k_mom_mma(const MomMmaParams P) {
  __shared__ __align__(1024) unsigned char s_a[2][kMmaATile];      // indicator tiles: not-common, rare
  __shared__ __align__(1024) unsigned char s_b[2 * kMmaBTile];     // payload tiles: common | rare
  __shared__ __align__(16) uint32_t s_mask[2][4][kMmaK];           // [class][warp slice][locus]
  __shared__ uint64_t s_bar[2];
  __shared__ uint32_t s_tmem, s_mine[4];
  // 1-D grid, block = unit * tiles_per_unit + tile: the tiles of one unit run together, so its rows and payload tiles come from L2
  // (a 2-D grid would cap the units at 65,535)
  const uint32_t unit_index = blockIdx.x / P.tiles_per_unit;
  const MomUnit U = P.unit_table[unit_index];
  const uint32_t tile = P.tile_lo[U.pop] + blockIdx.x % P.tiles_per_unit;
  if (tile >= P.tile_hi[U.pop]) return;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t g = (uint64_t)tile * kMmaM + tid;
  const bool mine = g < P.n_genomes && P.superpop[g] == U.pop;
  if (!__syncthreads_or(mine)) return;
  const uint32_t mine_mask = __ballot_sync(kFull, mine);
  if (lane == 0) s_mine[warp] = mine_mask;
  const uint32_t bar_b = smem_u32(&s_bar[0]), bar_mma = smem_u32(&s_bar[1]);
  if (tid == 0) {
    mbar_init(bar_b, 1); mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  const uint32_t mm0 = s_mine[0], mm1 = s_mine[1], mm2 = s_mine[2], mm3 = s_mine[3];
  const uint32_t n_stages = (U.end - U.begin + kMmaK - 1) / kMmaK;
  const uint64_t unit0 = (uint64_t)tile * (kMmaM / 64);
  uint32_t b_phase = 0, mma_phase = 0;
  bool pending = false, acc[2] = {false, false};
  // the two units of this thread's locus, requested one stage ahead
  auto fetch = [&](uint32_t s, uint4 (&v)[2]) {
    v[0] = make_uint4(0u, 0u, 0u, 0u); v[1] = v[0];
    const uint32_t i = U.begin + s * kMmaK + tid;
    if (s < n_stages && i < U.end) {
      const uint4* row = P.packed + (uint64_t)P.rows[i] * P.units + unit0;
      if (unit0 < P.units) v[0] = __ldg(row);
      if (unit0 + 1 < P.units) v[1] = __ldg(row + 1);
    }
  };
  uint4 nxt[2];
  fetch(0, nxt);

  for (uint32_t s = 0; s < n_stages; ++s) {
    // masks of this thread's locus over the tile's four warp slices (the MMAs of the stage before may still be running)
    uint32_t nc[4] = {0u, 0u, 0u, 0u}, ra[4] = {0u, 0u, 0u, 0u};
    if (U.begin + s * kMmaK + tid < U.end) {
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (unit0 + k < P.units) {
          const uint4 v = nxt[k];                               // {lo.lo32, lo.hi32, hi.lo32, hi.hi32}
          const MomMasks a = mom_masks(v.x, v.z, U.common_code, U.rare_code), b = mom_masks(v.y, v.w, U.common_code, U.rare_code);
          nc[2 * k] = a.nc; nc[2 * k + 1] = b.nc; ra[2 * k] = a.rare; ra[2 * k + 1] = b.rare;
        }
    }
    fetch(s + 1, nxt);
#pragma unroll
    for (int k = 0; k < 4; ++k) { s_mask[0][k][tid] = nc[k]; s_mask[1][k][tid] = ra[k]; }
    const bool has_c = __syncthreads_or(((nc[0] & mm0) | (nc[1] & mm1) | (nc[2] & mm2) | (nc[3] & mm3)) != 0u);
    const bool rare_here = ((ra[0] & mm0) | (ra[1] & mm1) | (ra[2] & mm2) | (ra[3] & mm3)) != 0u;
    if (P.rare_bits) {
      const uint32_t bal = __ballot_sync(kFull, rare_here);
      if (lane == 0) P.rare_bits[((size_t)(U.tile_base + s) * P.tiles_per_unit + (tile - P.tile_lo[U.pop])) * 4 + warp] = bal;
    }
    const bool has_r = __syncthreads_or(rare_here);
    if (pending) { mbar_wait(bar_mma, mma_phase); mma_phase ^= 1; pending = false; }      // the MMAs of the stage before have read s_a / s_b
    if (has_c || has_r) {
      if (tid == 0) {
        mbar_expect_tx(bar_b, 2 * kMmaBTile);
        tma_bulk_g2s(smem_u32(s_b), P.btiles + (size_t)(U.tile_base + s) * (2 * kMmaBTile), 2 * kMmaBTile, bar_b);
      }
      const uint32_t keep = mine ? ~0u : 0u;
      if (has_c) mma_build_row(s_a[0], s_mask[0][warp], tid, lane, keep);
      if (has_r) {            // rare homozygous cells are few: most warps have none in a stage
        const uint32_t* mr = s_mask[1][warp];
        if (__any_sync(kFull, ((mr[lane] | mr[lane + 32] | mr[lane + 64] | mr[lane + 96]) & mine_mask) != 0u)) mma_build_row(s_a[1], mr, tid, lane, keep);
        else mma_zero_row(s_a[1], tid);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();                                   // s_a complete; s_mask free for the next stage
    if (has_c || has_r) {
      if (tid == 0) {
        mbar_wait(bar_b, b_phase);
        tc_fence_after();
#pragma unroll
        for (uint32_t x = 0; x < 2; ++x) {
          if (!(x == 0 ? has_c : has_r)) continue;
          const uint32_t sa = smem_u32(s_a[x]), sb = smem_u32(s_b) + x * kMmaBTile;
#pragma unroll
          for (uint32_t kk = 0; kk < kMmaK / 32; ++kk)
            mma_i8_n32(tmem_base + x * kMmaN, tc_smem_desc(sa + kk * 32), tc_smem_desc(sb + kk * 32), (acc[x] || kk) ? 1u : 0u);
        }
        tc_commit(bar_mma);
      }
      b_phase ^= 1;
      pending = true;
      acc[0] = acc[0] || has_c; acc[1] = acc[1] || has_r;
    }
  }
  if (pending) mbar_wait(bar_mma, mma_phase);
  tc_fence_after();

  // flush: limbs -> 64-bit integers, offsets out, into the genome's bins
  uint32_t n_rare = 0;
#pragma unroll
  for (uint32_t x = 0; x < 2; ++x) {
    if (!acc[x]) continue;                                           // CTA-uniform
    uint32_t v[32];
    const uint32_t taddr = tmem_base + x * kMmaN + ((warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int bin_g = x == 0 ? U.cb : U.rb;
    if (!mine || v[0] == 0u || bin_g < 0) continue;
    const int bin = bin_g >= kMomBinsMax ? P.nbt - 1 : bin_g - P.b_lo;
    const long long count = (long long)v[0], offset = (long long)P.scale;
    unsigned long long* d = reinterpret_cast<unsigned long long*>(P.mi) + (g * (uint64_t)P.nbt + bin) * kMomJ;
    const long long sign = x == 0 ? -1 : 1;
    atomicAdd(d, (unsigned long long)(sign * count));
#pragma unroll
    for (int j = 0; j < kMomJ - 1; ++j) {
      long long sum = 0;
#pragma unroll
      for (int b = 0; b < kMmaLimbs; ++b) sum += (long long)v[1 + j * kMmaLimbs + b] << (8 * b);
      sum -= count * offset;
      atomicAdd(d + 1 + j, (unsigned long long)(sign * sum));
    }
    if (x == 1) n_rare = v[0];
  }
  if (P.cnt && mine) P.cnt[(uint64_t)unit_index * P.n_genomes_padded + g] = n_rare;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(64) : "memory");
}

// ---- lists of the rare homozygous cells (root search) -------------------------------------------------------------------------
// Per genome over the units of its population, in order: offs[unit][g] = rare cells of the units before within the same part
// (hom-alt units come first: they are the loci with p <= 1/2), totals = {rare cells, rare hom-alt cells}. 32 genomes x 8 stripes of
// units per block: every stripe sums its units, the stripes are scanned, then every stripe writes its offsets.
__global__ void __launch_bounds__(256)
k_mom_unit_scan(const uint32_t* __restrict__ cnt, const MomUnit* __restrict__ units, const uint2* __restrict__ unit_range,
                const uint8_t* __restrict__ superpop, uint64_t n_genomes, uint64_t n_genomes_padded, uint32_t* __restrict__ offs,
                uint32_t* __restrict__ totals) {
  __shared__ uint32_t s_alt[8][32], s_ref[8][32];
  const int gl = threadIdx.x & 31, stripe = threadIdx.x >> 5;
  const uint64_t g = (uint64_t)blockIdx.x * 32 + gl;
  uint2 r = make_uint2(0u, 0u);
  if (g < n_genomes) r = unit_range[superpop[g]];
  const uint32_t n_u = r.y - r.x, per = (n_u + 7) / 8;
  const uint32_t u0 = r.x + min(n_u, stripe * per), u1 = r.x + min(n_u, (stripe + 1) * per);
  uint32_t n_alt = 0, n_ref = 0;
  for (uint32_t u = u0; u < u1; ++u) {
    const uint32_t c = cnt[(uint64_t)u * n_genomes_padded + g];
    if (units[u].rare_code == 0) n_ref += c; else n_alt += c;
  }
  s_alt[stripe][gl] = n_alt; s_ref[stripe][gl] = n_ref;
  __syncthreads();
  uint32_t a = 0, b = 0, ta = 0, tb = 0;
  for (int k = 0; k < 8; ++k) { if (k < stripe) { a += s_alt[k][gl]; b += s_ref[k][gl]; } ta += s_alt[k][gl]; tb += s_ref[k][gl]; }
  if (g >= n_genomes) return;
  for (uint32_t u = u0; u < u1; ++u) {
    const uint32_t c = cnt[(uint64_t)u * n_genomes_padded + g];
    if (units[u].rare_code == 0) { offs[(uint64_t)u * n_genomes_padded + g] = b; b += c; }
    else { offs[(uint64_t)u * n_genomes_padded + g] = a; a += c; }
  }
  if (stripe == 0) { totals[g * 2 + 0] = ta + tb; totals[g * 2 + 1] = ta; }
}

struct MomFillParams {
  const uint4* packed; uint64_t units;
  const double* rr;
  const uint8_t* superpop; uint64_t n_genomes, n_genomes_padded;
  const uint32_t* rows; const MomUnit* unit_table;
  const uint32_t* cnt; const uint32_t* offs; const uint32_t* totals; const uint64_t* base;
  double* list;
  uint32_t tile_lo[kMaxPop], tile_hi[kMaxPop];       // kMomTile-genome tiles of the population
  uint32_t tiles_per_unit;
  const uint32_t* rare_bits;                         // k_mom_mma's bitmap and the tile ranges it is indexed with
  uint32_t mma_tile_lo[kMaxPop], mma_tile_hi[kMaxPop], mma_tiles_per_unit;
};

// list[base[g] + ...] = r of the genome's rare homozygous cells: hom-alt part (r ascending), then hom-ref part (r descending).
__global__ void __launch_bounds__(kMomTile)
k_mom_unit_fill(const MomFillParams P) {
  constexpr int kWarps = kMomTile / 32;
  __shared__ uint32_t s_rare[2][kMomStep][kWarps];          // two steps: one barrier per step (a warp may write step i + 1 while
  __shared__ double s_r[2][kMomStep];                       // another still walks step i)
  const uint32_t unit_index = blockIdx.x / P.tiles_per_unit;
  const MomUnit U = P.unit_table[unit_index];
  const uint32_t tile = P.tile_lo[U.pop] + blockIdx.x % P.tiles_per_unit;
  if (U.rare_code < 0 || tile >= P.tile_hi[U.pop]) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t g = (uint64_t)tile * kMomTile + threadIdx.x;
  const bool mine = g < P.n_genomes && P.superpop[g] == U.pop && P.cnt[(uint64_t)unit_index * P.n_genomes_padded + g] != 0u;
  if (!__syncthreads_or(mine)) return;                      // no genome of the tile has a rare homozygous cell in this unit
  const uint32_t mine_mask = __ballot_sync(kFull, mine);
  uint64_t pos = 0;
  if (mine) pos = P.base[g] + (U.rare_code == 0 ? P.totals[g * 2 + 1] : 0u) + P.offs[(uint64_t)unit_index * P.n_genomes_padded + g];
  const int tj = threadIdx.x >> 1, th = threadIdx.x & 1;
  const uint64_t unit0 = (uint64_t)tile * (kMomTile / 64) + 2 * th;
  static_assert(kMomStep == kMmaK && kMomTile == 2 * kMmaM, "a step is a stage of k_mom_mma, a tile two of its tiles");
  // this thread's half of the tile is one 128-genome tile of k_mom_mma: its bitmap says which loci of a stage have a rare
  // homozygous cell there -- only those rows are loaded (rare cells are a few per cent of the cells)
  const uint32_t t128 = tile * 2 + (uint32_t)th;
  const bool half_live = t128 >= P.mma_tile_lo[U.pop] && t128 < P.mma_tile_hi[U.pop];
  const uint32_t* bits = P.rare_bits + ((size_t)U.tile_base * P.mma_tiles_per_unit + (half_live ? t128 - P.mma_tile_lo[U.pop] : 0u)) * 4 + (tj >> 5);
  // this thread's half row of the step after the current one, requested a step ahead
  auto fetch = [&](uint32_t s, uint4 (&v)[2], double& r, bool& flag) {
    v[0] = make_uint4(0u, 0u, 0u, 0u); v[1] = v[0]; r = 0.0; flag = false;
    const uint32_t i = s + tj;
    if (s < U.end && i < U.end && half_live) {
      flag = (bits[(size_t)((s - U.begin) / kMomStep) * P.mma_tiles_per_unit * 4] >> (tj & 31)) & 1u;
      if (flag) {
        const uint4* row = P.packed + (uint64_t)P.rows[i] * P.units + unit0;
        if (unit0 < P.units) v[0] = __ldg(row);
        if (unit0 + 1 < P.units) v[1] = __ldg(row + 1);
        r = P.rr[i];
      }
    }
  };
  uint4 nxt[2]; double nxt_r; bool nxt_flag;
  fetch(U.begin, nxt, nxt_r, nxt_flag);
  uint32_t buf = 0;
  for (uint32_t s = U.begin; s < U.end; s += kMomStep, buf ^= 1u) {
    {
      uint32_t mk[4] = {0u, 0u, 0u, 0u};
      if (nxt_flag) {
        if (unit0 < P.units) { mk[0] = mom_code_mask(make_uint2(nxt[0].x, nxt[0].z), U.rare_code); mk[1] = mom_code_mask(make_uint2(nxt[0].y, nxt[0].w), U.rare_code); }
        if (unit0 + 1 < P.units) { mk[2] = mom_code_mask(make_uint2(nxt[1].x, nxt[1].z), U.rare_code); mk[3] = mom_code_mask(make_uint2(nxt[1].y, nxt[1].w), U.rare_code); }
        s_r[buf][tj] = nxt_r;                                   // both halves may write it: the same value
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) s_rare[buf][tj][4 * th + k] = mk[k];
    }
    const bool step_live = __syncthreads_or(nxt_flag);
    fetch(s + kMomStep, nxt, nxt_r, nxt_flag);
    if (!step_live) continue;
    const int n_here = (int)min((uint32_t)kMomStep, U.end - s);
    // 32 loci at a time: lane = locus holds the mask of its 32 genomes; after the transpose lane = genome holds the loci at which
    // it has a cell, in order -- every lane then walks only its own cells
#pragma unroll
    for (int k = 0; k < kMomStep / 32; ++k) {
      const int jj = k * 32 + lane;
      const uint32_t w = jj < n_here ? (s_rare[buf][jj][warp] & mine_mask) : 0u;
      if (!__any_sync(kFull, w != 0u)) continue;
      uint32_t cells = warp_transpose32(w, lane);
      while (cells) {
        const int j = __ffs(cells) - 1;
        cells &= cells - 1;
        P.list[pos++] = s_r[buf][k * 32 + j];
      }
    }
  }
}

// limits[g] = {left end of the genome's feasible region = max over its rare homozygous cells of e(a) = (1e-10 - a^2)/(a (1 - a)),
// a = r/(1 + r); the population's smallest 2 p q}. e = 1e-10 (1 + r)^2 / r - r falls with r on (0, 1), so the maximum sits at the
// genome's smallest r: the first entry of the hom-alt part or the last of the hom-ref part. The second limit is a lower bound of
// the genome's own smallest 2 p q: the clamp test it feeds (k_newton_reduce) sends a genome to the exact kernel when it fails.
__global__ void __launch_bounds__(256)
k_mom_list_limits(const double* __restrict__ list, const uint64_t* __restrict__ base, const uint32_t* __restrict__ totals,
                  const uint8_t* __restrict__ superpop, const unsigned long long* __restrict__ pop_cmin, uint64_t n_genomes,
                  double* __restrict__ limits) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_genomes) return;
  const uint64_t n = totals[g * 2], n_alt = totals[g * 2 + 1];
  const double* L = list + base[g];
  double r = kHuge;
  if (n_alt > 0) r = fmin(r, L[0]);
  if (n > n_alt) r = fmin(r, L[n - 1]);
  double fmin_ = -kHuge;
  if (r < kHuge) {
    const double a = r / (1.0 + r), d = a * (1.0 - a);
    fmin_ = d > 0.0 ? (kSmallProb - a * a) / d : (a == 0.0 ? kHuge : -kHuge);
  }
  limits[g * 3 + 0] = fmin_;
  limits[g * 3 + 1] = __longlong_as_double((long long)pop_cmin[superpop[g]]);
}

}  // namespace kgl
