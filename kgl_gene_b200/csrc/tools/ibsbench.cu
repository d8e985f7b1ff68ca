// ibsbench.cu -- standalone timing + verification of the pairwise IBS tile kernel (development tool; not part of the library).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o ibsbench ibsbench.cu
#include "../ibs_tile.cuh"
#include "../ibs_launch.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace kgl;

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); std::exit(2); } } while (0)

// planes [n_gblocks][n_words][32]: random codes (het 14%, hom-alt 2.5%), code 3 with probability miss/65536, padding loci coded 3
__global__ void k_fill_planes(uint32_t* lo, uint32_t* hi, uint64_t n_gblocks, uint64_t n_words, uint64_t n_loci, uint32_t miss, uint64_t seed) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_gblocks * n_words * 32) return;
  const uint64_t w = (i / 32) % n_words;
  uint32_t l = 0, h = 0;
  for (int b = 0; b < 32; ++b) {
    unsigned code = 3;
    if (w * 32 + b < n_loci) {
      const uint64_t r = mix64(seed ^ (i * 32 + b));
      const uint32_t u = (uint32_t)(r & 1023), m = (uint32_t)((r >> 20) & 65535);
      code = u < 143 ? 1u : (u < 169 ? 2u : 0u);
      if (m < miss) code = 3;
    }
    l |= (code & 1u) << b; h |= (code >> 1) << b;
  }
  lo[i] = l; hi[i] = h;
}

__global__ void k_ref_tile(const uint32_t* lo, const uint32_t* hi, uint64_t n_words, uint32_t words_used, uint2 tc, uint32_t* out /* [3][4096] */) {
  const uint32_t cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= kIbsTileCells) return;
  const uint64_t a = (uint64_t)tc.x * 64 + cell / 64, b = (uint64_t)tc.y * 64 + cell % 64;
  uint32_t c0 = 0, c1 = 0, cv = 0;
  for (uint32_t w = 0; w < words_used; ++w) {
    const uint64_t oa = ((a >> 5) * n_words + w) * 32 + (a & 31), ob = ((b >> 5) * n_words + w) * 32 + (b & 31);
    const uint32_t la = lo[oa], ha = hi[oa], lb = lo[ob], hb = hi[ob];
    for (int i = 0; i < 32; ++i) {
      const int ca = ((la >> i) & 1) | (((ha >> i) & 1) << 1), cb = ((lb >> i) & 1) | (((hb >> i) & 1) << 1);
      if (ca == 3 || cb == 3) continue;
      ++cv;
      const int d = ca > cb ? ca - cb : cb - ca;
      c0 += d == 2; c1 += d == 1;
    }
  }
  out[cell] = c0; out[kIbsTileCells + cell] = c1; out[2 * kIbsTileCells + cell] = cv;
}

int main(int argc, char** argv) {
  uint64_t n_genomes = 2504, n_loci = 1100000; int reps = 3; uint32_t miss = 66; uint32_t chunk_hint = 0; int verify = 1; int tj = 2;
  for (int i = 1; i < argc; ++i) {
    if (!std::strcmp(argv[i], "--genomes")) n_genomes = std::strtoull(argv[++i], nullptr, 10);
    else if (!std::strcmp(argv[i], "--loci")) n_loci = std::strtoull(argv[++i], nullptr, 10);
    else if (!std::strcmp(argv[i], "--reps")) reps = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--miss")) miss = (uint32_t)std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--chunk")) chunk_hint = (uint32_t)std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--no-verify")) verify = 0;
    else if (!std::strcmp(argv[i], "--tj")) tj = std::atoi(argv[++i]);
  }
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  const double clk = prop.clockRate * 1e-6;
  const uint64_t side = (n_genomes + 63) / 64, n_gblocks = side * 2;
  const uint64_t n_words = ((n_loci + 31) / 32 + 3) / 4 * 4;
  const bool missing = miss > 0;
  std::printf("device %s, %d SMs, %.3f GHz; %llu genomes x %llu loci, %llu words, missing %s\n", prop.name, sms, clk,
              (unsigned long long)n_genomes, (unsigned long long)n_loci, (unsigned long long)n_words, missing ? "yes" : "no");
  const size_t plane_words = (size_t)n_gblocks * n_words * 32;
  uint32_t *d_lo, *d_hi, *d_v;
  CK(cudaMalloc(&d_lo, plane_words * 4)); CK(cudaMalloc(&d_hi, plane_words * 4)); CK(cudaMalloc(&d_v, plane_words * 4));
  k_fill_planes<<<(unsigned)((plane_words + 255) / 256), 256>>>(d_lo, d_hi, n_gblocks, n_words, n_loci, miss, 20261018ull);
  k_valid_plane<<<(unsigned)((plane_words / 4 + 255) / 256), 256>>>((const uint4*)d_lo, (const uint4*)d_hi, plane_words / 4, (uint4*)d_v);
  CK(cudaDeviceSynchronize());

  std::vector<uint2> tiles;
  for (uint32_t a = 0; a < side; ++a) for (uint32_t b = a; b < side; ++b) tiles.push_back(make_uint2(a, b));
  uint2* d_tiles; CK(cudaMalloc(&d_tiles, tiles.size() * sizeof(uint2)));
  CK(cudaMemcpy(d_tiles, tiles.data(), tiles.size() * sizeof(uint2), cudaMemcpyHostToDevice));
  uint32_t* d_acc; CK(cudaMalloc(&d_acc, tiles.size() * 3 * kIbsTileCells * 4));

  IbsPlan pl = plan_ibs((uint32_t)tiles.size(), (uint32_t)n_words, sms, chunk_hint);
  IbsParams P{};
  P.plane[0] = d_lo; P.plane[1] = d_hi; P.plane[2] = d_v; P.n_words = n_words; P.words_used = (uint32_t)n_words;
  P.tiles = d_tiles; P.n_tiles = (uint32_t)tiles.size(); P.words_per_chunk = pl.words_per_chunk; P.n_chunks = pl.n_chunks; P.acc = d_acc;
  std::printf("tiles %zu, chunks %u x %u words, grid %u, smem %zu\n", tiles.size(), pl.n_chunks, pl.words_per_chunk, pl.grid, ibs_smem_bytes(missing));

  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto go = [&]() {
    if (pl.n_chunks > 1) CK(cudaMemsetAsync(d_acc, 0, tiles.size() * 3 * kIbsTileCells * 4));
    CK(launch_ibs(P, pl, missing, 0, tj));
  };
  go(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; ++r) go();
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
  const double pair_loci_tri = (double)n_genomes * (n_genomes + 1) / 2 * n_loci;
  const double pair_words_exec = (double)tiles.size() * kIbsTileCells * n_words;
  std::printf("k_ibs_tiles %9.3f ms  %.3e upper-triangle pair-loci/s  (executed %.3e pair-words/s = %.4f clk per pair-word per SM)\n", ms,
              pair_loci_tri / (ms * 1e-3), pair_words_exec / (ms * 1e-3), (ms * 1e-3) * clk * 1e9 * sms / pair_words_exec);

  if (verify) {
    uint32_t* d_ref; CK(cudaMalloc(&d_ref, 3 * kIbsTileCells * 4));
    std::vector<uint32_t> a(3 * kIbsTileCells), b(3 * kIbsTileCells);
    int bad = 0;
    const size_t picks[] = {0, 1, tiles.size() / 2, tiles.size() - 1};
    for (size_t t : picks) {
      if (t >= tiles.size()) continue;
      k_ref_tile<<<16, 256>>>(d_lo, d_hi, n_words, (uint32_t)n_words, tiles[t], d_ref);
      CK(cudaMemcpy(a.data(), d_acc + t * 3 * kIbsTileCells, a.size() * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(b.data(), d_ref, b.size() * 4, cudaMemcpyDeviceToHost));
      size_t diff = 0;
      const size_t n_cmp = missing ? a.size() : 2 * kIbsTileCells;
      for (size_t i = 0; i < n_cmp; ++i) diff += a[i] != b[i];
      std::printf("verify tile %zu (%u,%u): %zu mismatches  [c0 %u c1 %u cv %u | ref %u %u %u]\n", t, tiles[t].x, tiles[t].y, diff, a[5], a[4096 + 5],
                  a[8192 + 5], b[5], b[4096 + 5], b[8192 + 5]);
      bad += diff != 0;
    }
    std::printf(bad ? "VERIFY FAILED\n" : "VERIFY OK\n");
    if (bad) return 1;
  }
  return 0;
}
