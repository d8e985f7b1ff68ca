// grambench.cu -- standalone timing + verification of the tcgen05 int8 Gram kernel (development tool).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o grambench grambench.cu
#include "../gram_i8.cuh"
#include "../gram_launch.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace kgl;
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); std::exit(2); } } while (0)

__global__ void k_fill_codes(uint32_t* codes, uint64_t n_rows, uint64_t pitch, uint64_t n_genomes, uint64_t n_loci, uint64_t seed) {
  const uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 >= n_rows * pitch) return;
  const uint64_t g = i0 / pitch, w = i0 % pitch;
  const uint64_t i = gram_code_index(g, w, pitch / 8);
  uint32_t x = 0;
  if (g < n_genomes)
    for (int b = 0; b < 16; ++b) {
      if (w * 16 + b >= n_loci) break;
      const uint64_t r = mix64(seed ^ (i0 * 16 + b));
      const uint32_t u = (uint32_t)(r & 1023);
      const uint32_t code = u < 143 ? 1u : (u < 169 ? 2u : 0u);
      x |= code << (2 * b);
    }
  codes[i] = x;
}

__global__ void k_ref_pairs(const uint32_t* codes, uint64_t pitch, const uint2* pairs, uint32_t n_pairs, long long* out) {
  const uint32_t p = blockIdx.x;
  if (p >= n_pairs) return;
  long long s = 0;
  for (uint64_t w = threadIdx.x; w < pitch; w += blockDim.x) {
    const uint32_t x = codes[gram_code_index(pairs[p].x, w, pitch / 8)], y = codes[gram_code_index(pairs[p].y, w, pitch / 8)];
    for (int k = 0; k < 16; ++k) s += (long long)((x >> (2 * k)) & 3) * ((y >> (2 * k)) & 3);
  }
  __shared__ long long sh[256];
  sh[threadIdx.x] = s; __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) out[p] = sh[0];
}

int main(int argc, char** argv) {
  uint64_t n_genomes = 2504, n_loci = 1100000; int reps = 3; uint32_t chunk_hint = 0; int skip = 0;
  for (int i = 1; i < argc; ++i) {
    if (!std::strcmp(argv[i], "--genomes")) n_genomes = std::strtoull(argv[++i], nullptr, 10);
    else if (!std::strcmp(argv[i], "--loci")) n_loci = std::strtoull(argv[++i], nullptr, 10);
    else if (!std::strcmp(argv[i], "--reps")) reps = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--skip")) skip = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--chunk")) chunk_hint = (uint32_t)std::atoi(argv[++i]);
  }
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  const uint64_t ld = (n_genomes + 255) / 256 * 256;
  const uint32_t k_stages = (uint32_t)((n_loci + 127) / 128);
  const uint64_t pitch = (uint64_t)k_stages * 8;
  std::printf("device %s, %d SMs; %llu genomes (ld %llu) x %llu loci, %u K-stages\n", prop.name, sms, (unsigned long long)n_genomes,
              (unsigned long long)ld, (unsigned long long)n_loci, k_stages);
  uint32_t* d_codes; CK(cudaMalloc(&d_codes, ld * pitch * 4));
  k_fill_codes<<<(unsigned)((ld * pitch + 255) / 256), 256>>>(d_codes, ld, pitch, n_genomes, n_loci, 20261018ull);
  std::vector<uint2> tiles = gram_upper_tiles(ld);
  uint2* d_tiles; CK(cudaMalloc(&d_tiles, tiles.size() * sizeof(uint2)));
  CK(cudaMemcpy(d_tiles, tiles.data(), tiles.size() * sizeof(uint2), cudaMemcpyHostToDevice));
  int32_t* d_out; CK(cudaMalloc(&d_out, ld * ld * 4));
  GramPlan pl = plan_gram((uint32_t)tiles.size(), k_stages, sms, chunk_hint);
  GramParams P{};
  P.codes = d_codes; P.k_stages = k_stages; P.tiles = d_tiles; P.n_tiles = (uint32_t)tiles.size();
  P.stages_per_chunk = pl.stages_per_chunk; P.n_chunks = pl.n_chunks; P.out = d_out; P.ld = ld; (void)skip;
  P.table_a = P.table_b = 0x03020100u;
  std::printf("tiles %zu, chunks %u x %u stages, grid %u, smem %zu\n", tiles.size(), pl.n_chunks, pl.stages_per_chunk, pl.grid, kGramSmem);
  auto go = [&]() {
    if (pl.n_chunks > 1) CK(cudaMemsetAsync(d_out, 0, ld * ld * 4));
    CK(launch_gram(P, pl, 0));
  };
  go();
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; ++r) go();
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
  const double macs_exec = (double)tiles.size() * kGramM * kGramN * (double)k_stages * kGramK;
  const double pair_loci_tri = (double)n_genomes * (n_genomes + 1) / 2 * n_loci;
  std::printf("k_gram_i8 %9.3f ms  %.3e upper-triangle pair-loci/s  executed %.1f int8 TOP/s (2 ops per MAC)\n", ms, pair_loci_tri / (ms * 1e-3),
              2.0 * macs_exec / (ms * 1e-3) / 1e12);
  // verification on sampled pairs (upper triangle: row <= col within computed tiles)
  std::vector<uint2> pairs;
  for (int i = 0; i < 64; ++i) {
    uint64_t a = ((uint64_t)(i * 2 + 1) * 0x9E3779B97F4A7C15ULL >> 20) % n_genomes, b = ((uint64_t)(i * 2 + 2) * 0xBF58476D1CE4E5B9ULL >> 20) % n_genomes;
    if (a > b) std::swap(a, b);
    pairs.push_back(make_uint2((uint32_t)a, (uint32_t)b));
  }
  pairs.push_back(make_uint2(0, 0)); pairs.push_back(make_uint2(0, (uint32_t)n_genomes - 1));
  pairs.push_back(make_uint2((uint32_t)n_genomes - 1, (uint32_t)n_genomes - 1)); pairs.push_back(make_uint2(127, 128)); pairs.push_back(make_uint2(128, 255));
  uint2* d_pairs; long long* d_ref;
  CK(cudaMalloc(&d_pairs, pairs.size() * sizeof(uint2))); CK(cudaMalloc(&d_ref, pairs.size() * 8));
  CK(cudaMemcpy(d_pairs, pairs.data(), pairs.size() * sizeof(uint2), cudaMemcpyHostToDevice));
  k_ref_pairs<<<(unsigned)pairs.size(), 256>>>(d_codes, pitch, d_pairs, (uint32_t)pairs.size(), d_ref);
  std::vector<long long> ref(pairs.size());
  CK(cudaMemcpy(ref.data(), d_ref, pairs.size() * 8, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (size_t i = 0; i < pairs.size(); ++i) {
    int32_t got;
    CK(cudaMemcpy(&got, d_out + (uint64_t)pairs[i].x * ld + pairs[i].y, 4, cudaMemcpyDeviceToHost));
    if ((long long)got != ref[i]) { if (bad < 8) std::printf("MISMATCH (%u,%u): got %d want %lld\n", pairs[i].x, pairs[i].y, got, ref[i]); ++bad; }
  }
  std::printf(bad ? "VERIFY FAILED (%d)\n" : "VERIFY OK (%d mismatches)\n", bad);
  return bad ? 1 : 0;
}
