// kbench.cu -- standalone micro-benchmarks for the kgl_b200 kernels (development tool; not part of the library).
//   1. issue rates of the integer pipes the kernels lean on (POPC, LOP3, IADD3, REDUX, SHFL) -> the INT-pipe roofline
//      denominators of DESIGN.md
//   2. the streaming pass k_stream_count in its configurations, timed with CUDA events and checked against naive kernels
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o kbench kbench.cu
#include "../stream_launch.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace kgl;

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); std::exit(2); } } while (0)

// ------------------------------------------------------------------------------------------------ pipe rates --------
template <int OP>
__global__ void __launch_bounds__(256) k_pipe(uint32_t* out, int iters, uint32_t seed) {
  uint32_t a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B9u;
  uint32_t b = seed ^ 0x5bd1e995u, c = seed * 7u + 1u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) a[i] = __popc(a[i]) + b;                                   // POPC + IADD
      else if (OP == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
      else if (OP == 2) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
      else if (OP == 3) a[i] = __reduce_add_sync(0xffffffffu, a[i]) + b;     // REDUX + IADD
      else if (OP == 4) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1) + b;    // SHFL + IADD
      else if (OP == 5) { a[i] = __popc(a[i]); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c)); }  // POPC + LOP3
      else if (OP == 6) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c)); a[i] = a[i] * 3u + b; }  // LOP3 + IMAD
      else if (OP == 7) a[i] = __popc(a[i]);                                  // POPC alone (dependent chain per register)
      else if (OP == 8) a[i] = __reduce_add_sync(0xffffffffu, __popc(a[i]) + b);   // POPC + IADD + REDUX
      else if (OP == 9) { a[i] += b; asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c)); }  // IADD + LOP3
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) r ^= a[i];
  if (r == 0x12345678u) out[0] = r;
}

// FP64 pipe: the estimator passes behind the streaming kernel (terms_fast.cuh) are bound by it.
//   0 DFMA   1 DADD   2 MUFU.RCP64H (rcp.approx.ftz.f64)   3 fast reciprocal (MUFU + 3 DFMA)   4 IEEE divide (__ddiv_rn)
template <int OP>
__global__ void __launch_bounds__(256) k_pipe_f64(double* out, int iters, double seed) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + 1e-3 * (threadIdx.x + 1) + 0.125 * i;
  const double b = 1.0 + 1e-9 * seed, c = 1e-7 * seed;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) a[i] = fma(a[i], b, c);
      else if (OP == 1) a[i] = __dadd_rn(a[i], c);
      else if (OP == 2) { double x; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a[i])); a[i] = x; }
      else if (OP == 3) { double x; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a[i]));
                          const double e = fma(-a[i], x, 1.0); a[i] = fma(x, fma(e, e, e), x) + b; }
      else if (OP == 4) a[i] = __ddiv_rn(b, a[i]) + b;
    }
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += a[i];
  if (r == 0.12345) out[0] = r;
}

// Accuracy of MUFU.RCP64H and of the refinements used by terms_fast.cuh, against the IEEE divide, over a log-uniform sweep.
__global__ void k_rcp_accuracy(double* out /* [3] max relative errors: x0, x0(1+e), x0(1+e+e^2) */, int n) {
  double m0 = 0, m1 = 0, m2 = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double d = exp2(-40.0 + 80.0 * (double)i / n) * (1.0 + 0.37 * (double)(i % 977) / 977.0);
    double x; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    const double ref = 1.0 / d, e = fma(-d, x, 1.0);
    const double x1 = fma(x, e, x), x2 = fma(x, fma(e, e, e), x);
    m0 = fmax(m0, fabs(x - ref) / ref); m1 = fmax(m1, fabs(x1 - ref) / ref); m2 = fmax(m2, fabs(x2 - ref) / ref);
  }
  for (int o = 16; o > 0; o >>= 1) {
    m0 = fmax(m0, __shfl_xor_sync(0xffffffffu, m0, o)); m1 = fmax(m1, __shfl_xor_sync(0xffffffffu, m1, o)); m2 = fmax(m2, __shfl_xor_sync(0xffffffffu, m2, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax((unsigned long long*)&out[0], (unsigned long long)__double_as_longlong(m0));
    atomicMax((unsigned long long*)&out[1], (unsigned long long)__double_as_longlong(m1));
    atomicMax((unsigned long long*)&out[2], (unsigned long long)__double_as_longlong(m2));
  }
}

template <int OP>
static void pipe_rate_f64(const char* name, double* d_out, int sms, double clk_ghz) {
  const int iters = 1024, blocks = sms * 8;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_pipe_f64<OP><<<blocks, 256>>>(d_out, 16, 1.0);
  CK(cudaEventRecord(e0));
  k_pipe_f64<OP><<<blocks, 256>>>(d_out, iters, 1.0);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  const double inner = (double)blocks * 256 * iters * 8;
  const double per_s = inner / (ms * 1e-3);
  std::printf("pipe %-22s %8.3f ms  %.3e inner-ops/s  = %.2f inner/clk/SM at %.3f GHz\n", name, ms, per_s, per_s / sms / (clk_ghz * 1e9), clk_ghz);
}

template <int OP>
static void pipe_rate(const char* name, int ops_per_inner, uint32_t* d_out, int sms, double clk_ghz) {
  const int iters = 4096, blocks = sms * 8;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_pipe<OP><<<blocks, 256>>>(d_out, 64, 1u);
  CK(cudaEventRecord(e0));
  k_pipe<OP><<<blocks, 256>>>(d_out, iters, 1u);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  const double inner = (double)blocks * 256 * iters * 8;
  const double per_s = inner / (ms * 1e-3);
  std::printf("pipe %-12s %8.3f ms  %.3e inner-ops/s  = %.2f inner/clk/SM at %.3f GHz (%d instr per inner)\n", name, ms, per_s,
              per_s / sms / (clk_ghz * 1e9), clk_ghz, ops_per_inner);
}

// ------------------------------------------------------------------------------------------------ data --------------
__global__ void k_fill(uint4* packed, uint64_t n_loci, uint64_t units, uint32_t n_genomes, uint32_t p_het_1024, uint32_t p_hom_1024,
                       uint32_t p_miss_65536, uint64_t seed) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_loci * units) return;
  const uint64_t unit = idx % units;
  uint64_t lo = 0, hi = 0;
  for (int b = 0; b < 64; ++b) {
    if (unit * 64 + b >= n_genomes) break;
    const uint64_t h = mix64(seed ^ (idx * 64 + b));
    const uint32_t u = (uint32_t)(h & 1023), m = (uint32_t)((h >> 20) & 65535);
    unsigned code = u < p_het_1024 ? 1u : (u < p_het_1024 + p_hom_1024 ? 2u : 0u);
    if (m < p_miss_65536) code = 3;
    lo |= (uint64_t)(code & 1) << b;
    hi |= (uint64_t)(code >> 1) << b;
  }
  packed[idx] = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
}

__global__ void k_fill_flags(uint16_t* flags, uint64_t n, int mode, uint64_t seed) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint16_t f = 0x3f;
  if (mode == 1) { const uint64_t h = mix64(seed ^ i); if ((h & 63) == 0) f = (uint16_t)((h >> 8) & 0x3f); }   // ~1.6% partial rows
  if (mode == 2) { const uint64_t h = mix64(seed ^ (i >> 9)); f = (h & 1) ? 0x3f : 0; }                        // windows of 512 rows on/off
  flags[i] = f;
}
__global__ void k_fill_sum64(const uint16_t* flags, uint64_t n_groups, uint16_t* sum64) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  uint32_t a = 0xFF, o = 0;
  for (int i = 0; i < 64; ++i) { const uint32_t f = flags[g * 64 + i] & 0xFF; a &= f; o |= f; }
  sum64[g] = (uint16_t)(a | (o << 8));
}

// naive references ------------------------------------------------------------------------------------------------------
__global__ void k_ref_locus(const uint4* packed, uint64_t n_loci, uint64_t units, uint32_t* out) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_loci * units) return;
  const uint4 v = packed[idx];
  const uint64_t row = idx / units;
  atomicAdd(out + row * 4 + 1, __popc(v.x & ~v.z) + __popc(v.y & ~v.w));
  atomicAdd(out + row * 4 + 2, __popc(v.z & ~v.x) + __popc(v.w & ~v.y));
  atomicAdd(out + row * 4 + 3, __popc(v.x & v.z) + __popc(v.y & v.w));
}
__global__ void k_ref_genome(const uint4* packed, uint64_t n_loci, uint64_t units, const uint16_t* flags, const uint8_t* superpop,
                             uint32_t n_genomes, uint32_t* out /* [N][2] */) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_genomes) return;
  const uint64_t unit = g >> 6;
  const int b = g & 63;
  const int k = superpop ? superpop[g] : 0;
  uint32_t nlo = 0, nhi = 0;
  for (uint64_t l = blockIdx.y; l < n_loci; l += gridDim.y) {
    if (flags && !((flags[l] >> k) & 1)) continue;
    const uint4 v = packed[l * units + unit];
    const uint64_t lo = (uint64_t)v.x | ((uint64_t)v.y << 32), hi = (uint64_t)v.z | ((uint64_t)v.w << 32);
    nlo += (lo >> b) & 1; nhi += (hi >> b) & 1;
  }
  atomicAdd(out + g * 2, nlo); atomicAdd(out + g * 2 + 1, nhi);
}

__global__ void __launch_bounds__(256) k_read_only(const uint4* packed, uint64_t n, uint32_t* out) {
  uint32_t acc = 0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    const uint4 a = ld_stream_u4(packed + i), b = ld_stream_u4(packed + i + stride), c = ld_stream_u4(packed + i + 2 * stride),
                d = ld_stream_u4(packed + i + 3 * stride);
    acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
  }
  for (; i < n; i += stride) { const uint4 a = ld_stream_u4(packed + i); acc ^= a.x ^ a.y ^ a.z ^ a.w; }
  if (acc == 0x12345678u) out[0] = acc;
}

// Co-residency probe: a block that spins for `ns` nanoseconds, with or without a static shared-memory footprint.
template <int SMEM_WORDS>
__global__ void __launch_bounds__(256) k_spin(unsigned long long ns, unsigned* sink) {
  __shared__ unsigned s_pad[SMEM_WORDS];
  s_pad[threadIdx.x % SMEM_WORDS] = threadIdx.x;
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  unsigned acc = 0;
  do { acc += s_pad[(threadIdx.x * 7 + acc) % SMEM_WORDS]; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < ns);
  if (acc == 0xFFFFFFFFu) *sink = acc;
}

int main(int argc, char** argv) {
  uint64_t n_loci = 1100000; uint32_t n_genomes = 2504; int reps = 10; int do_pipes = 1; uint64_t verify_loci = 65536; int only_cfg = -1; int do_verify = 1;
  for (int i = 1; i < argc; ++i) {
    if (!std::strcmp(argv[i], "--loci")) n_loci = std::strtoull(argv[++i], nullptr, 10);
    else if (!std::strcmp(argv[i], "--genomes")) n_genomes = (uint32_t)std::strtoul(argv[++i], nullptr, 10);
    else if (!std::strcmp(argv[i], "--reps")) reps = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--no-pipes")) do_pipes = 0;
    else if (!std::strcmp(argv[i], "--verify-loci")) verify_loci = std::strtoull(argv[++i], nullptr, 10);
    else if (!std::strcmp(argv[i], "--cfg")) only_cfg = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--no-verify")) do_verify = 0;
  }
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  const double clk = prop.clockRate * 1e-6;
  std::printf("device %s, %d SMs, %.3f GHz max, L2 %d MB\n", prop.name, sms, clk, prop.l2CacheSize >> 20);
  uint32_t* d_out; CK(cudaMalloc(&d_out, 64));

  if (do_pipes) {
    pipe_rate<7>("POPC", 1, d_out, sms, clk);
    pipe_rate<0>("POPC+IADD", 2, d_out, sms, clk);
    pipe_rate<1>("LOP3", 1, d_out, sms, clk);
    pipe_rate<2>("IADD", 1, d_out, sms, clk);
    pipe_rate<3>("REDUX+IADD", 2, d_out, sms, clk);
    pipe_rate<4>("SHFL+IADD", 2, d_out, sms, clk);
    pipe_rate<5>("POPC+LOP3", 2, d_out, sms, clk);
    pipe_rate<6>("LOP3+IMAD", 2, d_out, sms, clk);
    pipe_rate<8>("POPC+IADD+REDUX", 3, d_out, sms, clk);
    pipe_rate<9>("IADD+LOP3", 2, d_out, sms, clk);
    pipe_rate_f64<0>("DFMA", (double*)d_out, sms, clk);
    pipe_rate_f64<1>("DADD", (double*)d_out, sms, clk);
    pipe_rate_f64<2>("MUFU.RCP64H", (double*)d_out, sms, clk);
    pipe_rate_f64<3>("rcp (MUFU+3DFMA)+DADD", (double*)d_out, sms, clk);
    pipe_rate_f64<4>("IEEE div + DADD", (double*)d_out, sms, clk);
    {
      double* d_acc; CK(cudaMalloc(&d_acc, 24)); CK(cudaMemset(d_acc, 0, 24));
      k_rcp_accuracy<<<sms * 4, 256>>>(d_acc, 1 << 26);
      double h[3]; CK(cudaMemcpy(h, d_acc, 24, cudaMemcpyDeviceToHost));
      std::printf("rcp accuracy (max relative error over 2^26 arguments in [2^-40, 2^40]): MUFU.RCP64H %.3e (2^%.1f), one step x0(1+e) %.3e, x0(1+e+e^2) %.3e\n",
                  h[0], std::log2(h[0]), h[1], h[2]);
      CK(cudaFree(d_acc));
    }
  }

  const uint64_t units = stream_units_padded((n_genomes + 63) / 64);
  std::vector<uint8_t> h_superpop(n_genomes), h_need(units * 2, 0);
  std::vector<uint32_t> h_popmask(6 * units * 2, 0);
  for (uint32_t g = 0; g < n_genomes; ++g) {
    const int k = (int)((uint64_t)g * 5 / n_genomes);
    h_superpop[g] = (uint8_t)k;
    h_need[g >> 5] |= (uint8_t)(1u << k);
    h_popmask[(size_t)k * units * 2 + (g >> 5)] |= 1u << (g & 31);
  }
  uint8_t *d_superpop, *d_need; uint32_t* d_popmask;
  CK(cudaMalloc(&d_superpop, n_genomes)); CK(cudaMalloc(&d_need, units * 2)); CK(cudaMalloc(&d_popmask, 6 * units * 8));
  CK(cudaMemcpy(d_superpop, h_superpop.data(), n_genomes, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_need, h_need.data(), units * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_popmask, h_popmask.data(), 6 * units * 8, cudaMemcpyHostToDevice));

  const uint64_t pad_rows = (n_loci + 255) / 256 * 256;
  uint4* d_packed; CK(cudaMalloc(&d_packed, pad_rows * units * 16));
  CK(cudaMemset(d_packed, 0, pad_rows * units * 16));
  k_fill<<<(unsigned)((n_loci * units + 255) / 256), 256>>>(d_packed, n_loci, units, n_genomes, 140, 25, 66, 20261018ull);
  uint16_t* d_flags[3]; uint16_t* d_sum64[3];
  for (int m = 0; m < 3; ++m) {
    CK(cudaMalloc(&d_flags[m], pad_rows * 2)); CK(cudaMemset(d_flags[m], 0, pad_rows * 2));
    CK(cudaMalloc(&d_sum64[m], pad_rows / 64 * 2));
    k_fill_flags<<<(unsigned)((n_loci + 255) / 256), 256>>>(d_flags[m], n_loci, m, 99);
    k_fill_sum64<<<(unsigned)((pad_rows / 64 + 255) / 256), 256>>>(d_flags[m], pad_rows / 64, d_sum64[m]);
  }
  uint32_t *d_lc, *d_lc_ref, *d_gc, *d_gc_ref, *d_planes;
  CK(cudaMalloc(&d_lc, n_loci * 16)); CK(cudaMalloc(&d_lc_ref, n_loci * 16));
  CK(cudaMalloc(&d_gc, units * 64 * 8)); CK(cudaMalloc(&d_gc_ref, units * 64 * 8));
  CK(cudaDeviceSynchronize());

  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double matrix_bytes = (double)n_loci * units * 16;
  {
    for (int w = 0; w < 2; ++w) k_read_only<<<sms * 8, 256>>>(d_packed, n_loci * units, d_out);
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) k_read_only<<<sms * 8, 256>>>(d_packed, n_loci * units, d_out);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    std::printf("read-only LDG.128 x4        %8.4f ms  %7.1f GB/s\n", ms / reps, matrix_bytes / (ms / reps * 1e-3) / 1e9);
  }

  auto make_params = [&](const StreamPlan& pl, uint64_t loci, int flags_mode, bool wl, bool wg, uint32_t* planes) {
    StreamParams P{};
    fill_stream_params(P, pl);
    P.packed = d_packed; P.units = (uint32_t)units; P.n_loci = (uint32_t)loci; P.n_genomes = n_genomes;
    P.flags16 = flags_mode < 0 ? nullptr : d_flags[flags_mode];
    P.sum64 = flags_mode < 0 ? nullptr : d_sum64[flags_mode];
    P.need32 = d_need; P.popmask32 = d_popmask; P.n_pop = 6; P.locus_counts = wl ? d_lc : nullptr;
    P.cta_counts = wg ? planes : nullptr; P.n_genomes_padded = (uint32_t)(units * 64);
    if (!getenv("KBENCH_NO_TMAP")) stream_make_tensor_map(P, pl, pad_rows);
    return P;
  };

  struct Cfg { const char* name; bool wl, wg; int flags_mode; int rows, stages; bool generic; };
  const Cfg cfgs[] = {
      {"locus+genome full  auto   ", true, true, 0, 0, 0},    {"locus+genome full  s3     ", true, true, 0, 0, 3},
      {"locus+genome full  s2     ", true, true, 0, 0, 2},    {"locus+genome full  generic", true, true, 0, 0, 0, true},
      {"locus only         auto   ", true, false, 0, 0, 0},   {"genome only full   auto   ", false, true, 0, 0, 0},
      {"neither (stream)   auto   ", false, false, 0, 0, 0},  {"locus+genome part. auto   ", true, true, 1, 0, 0},
      {"locus+genome windows auto ", true, true, 2, 0, 0},    {"locus+genome raw   auto   ", true, true, -1, 0, 0},
  };
  int cfg_index = -1;
  for (const Cfg& c : cfgs) {
    ++cfg_index;
    if (only_cfg >= 0 && cfg_index != only_cfg) continue;
    StreamPlan pl = plan_stream(units, n_loci, sms, c.rows, c.stages, !c.generic);
    if (pl.smem > 227 * 1024 || pl.rows_per_stage * pl.h_parts != (uint32_t)kScHThreads) { std::printf("%-28s skipped (smem %zu)\n", c.name, pl.smem); continue; }
    CK(cudaMalloc(&d_planes, (size_t)pl.n_ctas * 2 * units * 64 * 4));
    StreamParams P = make_params(pl, n_loci, c.flags_mode, c.wl, c.wg, d_planes);
    auto go = [&]() { CK(launch_stream(P, pl, c.wl, c.wg, 0)); };
    for (int w = 0; w < 3; ++w) go();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) go();
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    std::printf("%-28s %8.4f ms  %7.1f GB/s  (shape %d grid %ux%u, R %u, SU %u, parts %u, RL %u, %u stages, smem %zu KB)\n", c.name, ms / reps,
                matrix_bytes / (ms / reps * 1e-3) / 1e9, pl.shape, pl.n_ctas, pl.slices, pl.rows_per_stage, pl.slice_units, pl.h_parts, pl.v_row_lanes,
                pl.n_stages, pl.smem >> 10);
    CK(cudaFree(d_planes));
  }

  // ---- co-residency: does a small kernel on a second stream run NEXT to the streaming kernel? ----
  if (only_cfg == -2) {
    int least, greatest;
    CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    unsigned* d_sink; CK(cudaMalloc(&d_sink, 4));
    for (int stages : {0, 2, 1}) {                 // 0: auto (4 stages, 180 KB); 2: 100 KB; 1: 2 stages and the spin kernel launched FIRST
      const bool spin_first = stages == 1;
      StreamPlan pl = plan_stream(units, n_loci, sms, 0, stages == 0 ? 0 : 2);
      CK(cudaMalloc(&d_planes, (size_t)pl.n_ctas * 2 * units * 64 * 4));
      StreamParams P = make_params(pl, n_loci, 0, true, true, d_planes);
      for (int equal_prio = 0; equal_prio < 2; ++equal_prio) {
        cudaStream_t sa, sb;
        CK(cudaStreamCreateWithPriority(&sa, cudaStreamNonBlocking, greatest));
        CK(cudaStreamCreateWithPriority(&sb, cudaStreamNonBlocking, equal_prio ? greatest : least));
        for (int smem : {0, 1}) {
          float t[3];
          for (int mode = 0; mode < 3; ++mode) {      // 0: streaming kernel alone; 1: spin alone (148 blocks x 50 us); 2: both
            CK(cudaDeviceSynchronize());
            cudaEvent_t a0, a1, b1; CK(cudaEventCreate(&a0)); CK(cudaEventCreate(&a1)); CK(cudaEventCreate(&b1));
            CK(cudaEventRecord(a0, sa));
            CK(cudaStreamWaitEvent(sb, a0, 0));
            auto spin = [&]() { if (smem) k_spin<3600><<<sms, 256, 0, sb>>>(50000ull, d_sink); else k_spin<32><<<sms, 256, 0, sb>>>(50000ull, d_sink); };
            if (mode != 0 && spin_first) spin();
            if (mode != 1) CK(launch_stream(P, pl, true, true, sa));
            if (mode != 0 && !spin_first) spin();
            CK(cudaEventRecord(b1, sb));
            CK(cudaStreamWaitEvent(sa, b1, 0));
            CK(cudaEventRecord(a1, sa));
            CK(cudaEventSynchronize(a1));
            CK(cudaEventElapsedTime(&t[mode], a0, a1));
          }
          std::printf("corun stages %u smem %zu KB, spin %s, %s priority, spin smem %s: stream %.4f  spin %.4f  both %.4f ms\n", pl.n_stages, pl.smem >> 10,
                      spin_first ? "first" : "second", equal_prio ? "equal" : "lower", smem ? "14 KB" : "128 B", t[0], t[1], t[2]);
        }
      }
      CK(cudaFree(d_planes));
    }
    return 0;
  }

  // ---- correctness on a prefix of the matrix ----
  if (do_verify) {
    const uint64_t vl = verify_loci < n_loci ? verify_loci : n_loci;
    int bad = 0;
    for (int mode = -1; mode <= 2; ++mode) {
      for (int rows : {0, 1, 64, 128, 256}) {
        StreamPlan pl = plan_stream(units, vl, sms, rows == 1 ? 0 : rows, 3, rows != 1);
        if (pl.smem > 227 * 1024) continue;
        CK(cudaMalloc(&d_planes, (size_t)pl.n_ctas * 2 * units * 64 * 4));
        CK(cudaMemset(d_planes, 0xAB, (size_t)pl.n_ctas * 2 * units * 64 * 4));
        StreamParams P = make_params(pl, vl, mode, true, true, d_planes);
        CK(cudaMemset(d_lc, 0, vl * 16)); CK(cudaMemset(d_lc_ref, 0, vl * 16));
        CK(cudaMemset(d_gc, 0, units * 64 * 8)); CK(cudaMemset(d_gc_ref, 0, units * 64 * 8));
        CK(launch_stream(P, pl, true, true, 0));
        if (pl.slices > 1) k_fix_locus_n0<<<(unsigned)((vl + 255) / 256), 256>>>(d_lc, vl, n_genomes);
        k_sum_cta_counts<<<(unsigned)((units * 64 + 255) / 256), 256>>>(d_planes, pl.n_ctas, units * 64, d_gc);
        k_ref_locus<<<(unsigned)((vl * units + 255) / 256), 256>>>(d_packed, vl, units, d_lc_ref);
        k_fix_locus_n0<<<(unsigned)((vl + 255) / 256), 256>>>(d_lc_ref, vl, n_genomes);
        dim3 rg((n_genomes + 127) / 128, 64);
        k_ref_genome<<<rg, 128>>>(d_packed, vl, units, mode < 0 ? nullptr : d_flags[mode], d_superpop, n_genomes, d_gc_ref);
        CK(cudaDeviceSynchronize());
        std::vector<uint32_t> a(vl * 4), b(vl * 4), ga(units * 128), gb(units * 128);
        CK(cudaMemcpy(a.data(), d_lc, vl * 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(b.data(), d_lc_ref, vl * 16, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(ga.data(), d_gc, units * 512, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(gb.data(), d_gc_ref, units * 512, cudaMemcpyDeviceToHost));
        uint64_t dl = 0, dg = 0;
        for (size_t i = 0; i < a.size(); ++i) dl += a[i] != b[i];
        for (size_t i = 0; i < (size_t)n_genomes * 2; ++i) dg += ga[i] != gb[i];
        std::printf("verify mode %2d shape %d R %3u (grid %ux%u, chunks %u, RL %u): locus mismatches %llu, genome mismatches %llu  [g0 lo %u hi %u]\n", mode, pl.shape,
                    pl.rows_per_stage, pl.n_ctas, pl.slices, pl.chunks_per_cta, pl.v_row_lanes, (unsigned long long)dl, (unsigned long long)dg, ga[0], ga[1]);
        bad += (dl != 0) + (dg != 0);
        CK(cudaFree(d_planes));
      }
    }
    std::printf(bad ? "VERIFY FAILED\n" : "VERIFY OK\n");
    if (bad) return 1;
  }
  return 0;
}
