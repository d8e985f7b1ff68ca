// ref_harness.cpp -- TEST INFRASTRUCTURE (oracle/_ref build only). Not part of the product.
//
// Drives the UNMODIFIED reference implementation of the population-genotype hot path on a flattened
// population (oracle/flat_io.h, KGLFLAT1) and dumps what it computes (KGLTENS1):
//
//   * locus selection per super-population   InbreedSampling::getPopulationLocusMap
//                                            (kga_analytic/kga_inbreed/kga_analysis_inbreed_locus.cpp:192)
//   * the four inbreeding estimators         InbreedingCalculation::namedAlgorithm -> process{Simple,RitlandLocus,
//                                            HallME,LogLikelihood} (kga_analysis_inbreed_calc.cpp:71,319,375,226,154),
//                                            fanned out over the reference's own WorkflowThreads pool exactly as
//                                            InbreedingAnalysis::processResults does (kga_analysis_inbreed_diploid.cpp:117-160)
//   * logLikelihood(f) on a fixed grid       (kga_analysis_inbreed_calc.cpp:94, captured inside the Optimize shim)
//   * allele summaries                       VariantDBVariant::{summaryByVariant,summaryByGenome,populationSummary}
//                                            (kgl_genomics/kgl_variant_db/kgl_variant_db_variant.cpp:126,180,234)
//
// The population is rebuilt through the reference's own containers (EvidenceFactory -> VariantEvidence ->
// Variant -> PopulationDB::addVariant), so everything downstream is reference code. It is used to
// (a) pin the C restatement (oracle/kgl_oracle.c) and generate tests/golden/*, (b) time the reference CPU
// path for bench.py's cpu_baseline / --impl reference.
#include "kel_exec_env_app.h"
#include "kel_workflow_threads.h"
#include "kgl_variant_db_population.h"
#include "kgl_variant_db_variant.h"
#include "kgl_variant_db_freq.h"
#include "kgl_variant_factory_vcf_evidence.h"
#include "kga_analysis_inbreed_calc.h"
#include "kga_analysis_inbreed_locus.h"
#include "kga_analysis_PfEMP_FWS.h"

#include "flat_io.h"
#include "ref_population.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <map>
#include <random>
#include <sstream>

namespace kel = kellerberrin;
namespace kgl = kellerberrin::genome;
namespace kga = kellerberrin::genome::analysis;

namespace kglref {
extern thread_local std::vector<double> tl_start_points;
extern thread_local std::vector<double> tl_end_points;
extern thread_local std::vector<double> tl_grid_values;
extern thread_local std::vector<size_t> tl_evals;
extern std::vector<double> g_grid;
extern std::atomic<bool> g_fixed_seed_enabled;
extern unsigned g_fixed_seed;
}  // namespace kglref

namespace {

struct Options {
  std::string in_path, out_path;
  std::vector<std::string> algos{"Simple", "RitlandLocus", "HallME", "Loglikelihood"};
  double min_af{0.0}, max_af{1.0};
  size_t spacing{0}, count{1000000000}, lower{0}, upper{1000000000};
  size_t threads{0};  // 0 = reference default (hardware_concurrency - 1)
  size_t grid{0};
  long seed{-1};      // >= 0: std::random_device is pinned to this value (see ref_stubs.cpp)
  size_t repeat{1};   // time the per-genome fan-out this many times (bench.py --impl reference)
  bool variantdb{true};
  bool fws{false};
  bool quiet{true};
};
Options g_opt;

using kglref::kSuperPops;
using kglref::kContig;

struct GenomeOut {
  kga::LocusResults results;
  std::vector<double> starts, ends, grid;
  std::vector<size_t> evals;
};

void run() {
  using Clock = std::chrono::steady_clock;
  const kglflat::Flat flat = kglflat::readFlat(g_opt.in_path);
  const uint32_t N = flat.N(), L = flat.L();
  const bool unphased = (flat.hdr.flags & kglflat::FLAG_UNPHASED) != 0;
  kglflat::TensorWriter out;

  // ---- the populations, through the reference's own containers (ref_population.h) ---------------------
  kglref::BuiltPopulations built = kglref::buildPopulations(flat, kgl::DataSourceEnum::Genome1000, kgl::DataSourceEnum::Genome1000);
  auto& af_population = built.af_population;
  auto& diploid = built.diploid;
  auto& genome_ids = built.genome_ids;
  auto& locus_variant = built.locus_variant;
  std::vector<uint32_t> present(N, 0);
  for (uint32_t g = 0; g < N; ++g) present[g] = diploid->getMap().contains(genome_ids[g]) ? 1 : 0;
  out.addU32("genome_present", {N}, present);

  // ---- locus selection, reference code ---------------------------------------------------------------
  kga::LociiVectorArguments args;
  args.lowerOffset(g_opt.lower);
  args.upperOffset(g_opt.upper);
  args.lociiSpacing(g_opt.spacing);
  args.lociiCount(g_opt.count);
  args.minAlleleFrequency(g_opt.min_af);
  args.maxAlleleFrequency(g_opt.max_af);
  args.frequencySource(kgl::DataSourceEnum::Genome1000);
  auto t0 = Clock::now();
  kga::ContigLocusMap contig_locus_map = kga::InbreedSampling::getPopulationLocusMap(af_population, args);
  const double locus_seconds = std::chrono::duration<double>(Clock::now() - t0).count();
  const kga::LocusMap& locus_map = contig_locus_map.at(kContig);
  for (int k = 0; k < 6; ++k) {
    std::vector<uint32_t> selected;
    for (auto const& [offset, offset_ptr] : locus_map.at(kSuperPops[k])->getMap()) selected.push_back(uint32_t(offset));
    out.addU32(std::string("selected_offsets_") + kSuperPops[k], {selected.size()}, selected);
  }

  // ---- count-limited windows: RetrieveLociiVector::getLociiCount (kga_analysis_inbreed_locus.cpp:159-183), the walk the window
  // loop of InbreedingAnalysis::populationInbreeding makes (kga_analysis_inbreed_diploid.cpp:47-73) -- only with --count
  if (g_opt.count < 1000000000) {
    auto const& [af_genome_id, af_genome_ptr] = *af_population->getMap().begin();
    auto const& [af_contig_id, af_contig_ptr] = *af_genome_ptr->getMap().begin();
    for (int k = 0; k < 6; ++k) {
      std::vector<kgl::ContigOffset_t> counted = kga::RetrieveLociiVector::getLociiCount(af_contig_ptr, kSuperPops[k], args);
      std::vector<uint32_t> offsets32(counted.begin(), counted.end());
      out.addU32(std::string("counted_offsets_") + kSuperPops[k], {offsets32.size()}, offsets32);
    }
  }

  // ---- estimators: one task per genome on the reference thread pool -----------------------------------
  if (g_opt.grid > 1) {
    for (size_t i = 0; i < g_opt.grid; ++i) kglref::g_grid.push_back(-0.5 + double(i) / double(g_opt.grid - 1));
    out.addF64("ll_grid", {kglref::g_grid.size()}, kglref::g_grid);
  }
  if (g_opt.seed >= 0) {
    kglref::g_fixed_seed = unsigned(g_opt.seed);
    kglref::g_fixed_seed_enabled = true;
    // The start values every genome will draw: the same library calls the reference makes
    // (kga_analysis_inbreed_calc.cpp:165,181 and :237,251), on a generator seeded the same way.
    std::vector<double> hall, ll;
    { std::mt19937_64 gen(kglref::g_fixed_seed); std::uniform_real_distribution<> d(0.5, 0); for (int i = 0; i < 5; ++i) hall.push_back(d(gen)); }
    { std::mt19937_64 gen(kglref::g_fixed_seed); std::uniform_real_distribution<> d(0.5, -0.5); for (int i = 0; i < 5; ++i) ll.push_back(d(gen)); }
    out.addF64("hall_start_sequence", {5}, hall);
    out.addF64("ll_start_sequence", {5}, ll);
  }
  const size_t threads = g_opt.threads ? g_opt.threads : kel::WorkflowThreads::defaultThreads();
  std::vector<double> timing;
  for (auto const& algo_name : g_opt.algos) {
    auto algorithm_opt = kga::InbreedingCalculation::namedAlgorithm(algo_name);
    if (!algorithm_opt) { std::fprintf(stderr, "unknown algorithm %s\n", algo_name.c_str()); std::exit(2); }
    const kga::InbreedingAlgorithm algorithm = algorithm_opt.value();
    std::vector<GenomeOut> results(N);
    std::vector<double> repeat_seconds;
    for (size_t rep = 0; rep < g_opt.repeat; ++rep) {
    t0 = Clock::now();
    {
      kel::WorkflowThreads pool(threads);
      std::vector<std::pair<uint32_t, std::future<GenomeOut>>> futures;
      for (uint32_t g = 0; g < N; ++g) {
        auto genome_opt = diploid->getGenome(genome_ids[g]);
        if (!genome_opt) continue;
        auto contig_opt = genome_opt.value()->getContig(kContig);
        if (!contig_opt) continue;
        std::shared_ptr<const kgl::ContigDB> contig_ptr = contig_opt.value();
        const std::string super_pop = kSuperPops[flat.superpop[g]];
        std::shared_ptr<const kgl::ContigDB> locus_list = locus_map.at(super_pop);
        auto task = [algorithm](kgl::GenomeId_t genome_id, std::shared_ptr<const kgl::ContigDB> contig,
                                std::string pop, std::shared_ptr<const kgl::ContigDB> list) {
          kglref::tl_start_points.clear(); kglref::tl_end_points.clear();
          kglref::tl_grid_values.clear(); kglref::tl_evals.clear();
          GenomeOut go;
          go.results = algorithm(genome_id, contig, pop, list);
          go.starts = kglref::tl_start_points; go.ends = kglref::tl_end_points;
          go.grid = kglref::tl_grid_values; go.evals = kglref::tl_evals;
          return go;
        };
        futures.emplace_back(g, pool.enqueueFuture(task, genome_ids[g], contig_ptr, super_pop, locus_list));
      }
      for (auto& [g, future] : futures) results[g] = future.get();
    }
    repeat_seconds.push_back(std::chrono::duration<double>(Clock::now() - t0).count());
    }
    out.addF64(algo_name + "_repeat_seconds", {repeat_seconds.size()}, repeat_seconds);
    const double seconds = repeat_seconds.back();
    timing.push_back(seconds);
    std::fprintf(stderr, "[ref] %-14s %u genomes x %u loci  %.3f s  (%zu threads)  %.3e genotype-loci/s\n",
                 algo_name.c_str(), N, L, seconds, threads, double(N) * double(L) / seconds);

    std::vector<uint64_t> counts(size_t(N) * 5);
    std::vector<double> freqs(size_t(N) * 4), coeff(N);
    for (uint32_t g = 0; g < N; ++g) {
      const auto& r = results[g].results;
      counts[g * 5 + 0] = r.major_homo_count;  counts[g * 5 + 1] = r.major_hetero_count;
      counts[g * 5 + 2] = r.minor_homo_count;  counts[g * 5 + 3] = r.minor_hetero_count;
      counts[g * 5 + 4] = r.total_allele_count;
      freqs[g * 4 + 0] = r.major_homo_freq;  freqs[g * 4 + 1] = r.major_hetero_freq;
      freqs[g * 4 + 2] = r.minor_homo_freq;  freqs[g * 4 + 3] = r.minor_hetero_freq;
      coeff[g] = r.inbred_allele_sum;
    }
    out.addU64(algo_name + "_counts", {N, 5}, counts);   // majHom, majHet, minHom, minHet, total
    out.addF64(algo_name + "_freqs", {N, 4}, freqs);     // same order
    out.addF64(algo_name + "_coeff", {N}, coeff);
    if (algo_name == "Loglikelihood") {
      size_t runs = 0;
      for (auto const& r : results) runs = std::max(runs, r.starts.size());
      std::vector<double> starts(size_t(N) * runs, std::nan("")), ends(size_t(N) * runs, std::nan("")), evals(size_t(N) * runs, 0.0);
      for (uint32_t g = 0; g < N; ++g) for (size_t i = 0; i < results[g].starts.size(); ++i) {
        starts[g * runs + i] = results[g].starts[i];
        ends[g * runs + i] = results[g].ends[i];
        evals[g * runs + i] = double(results[g].evals[i]);
      }
      out.addF64("ll_starts", {N, runs}, starts);
      out.addF64("ll_ends", {N, runs}, ends);
      out.addF64("ll_evals", {N, runs}, evals);
      if (!kglref::g_grid.empty()) {
        const size_t G = kglref::g_grid.size();
        std::vector<double> grid(size_t(N) * G, std::nan(""));
        for (uint32_t g = 0; g < N; ++g) for (size_t i = 0; i < results[g].grid.size() && i < G; ++i) grid[g * G + i] = results[g].grid[i];
        out.addF64("ll_grid_values", {N, G}, grid);
      }
    }
  }
  out.addF64("seconds_per_algo", {timing.size()}, timing);
  out.addF64("meta", {4}, std::vector<double>{double(threads), double(std::thread::hardware_concurrency()), locus_seconds, 0.0});

  // ---- allele summaries (VariantDBVariant) ------------------------------------------------------------
  if (g_opt.variantdb) {
    t0 = Clock::now();
    kgl::VariantDBVariant variant_db(diploid);
    std::vector<uint64_t> by_variant(size_t(L) * 3, 0), by_genome(size_t(N) * 3, 0), pop(3, 0);
    std::vector<uint32_t> variant_present(L, 0);
    for (uint32_t l = 0; l < L; ++l) {
      // Only variants that exist in the population have a column (kgl_variant_db_variant.cpp:14-30).
      if (!variant_db.variantMap().contains(locus_variant[l]->HGVS())) continue;
      variant_present[l] = 1;
      auto s = variant_db.summaryByVariant(locus_variant[l]);
      by_variant[l * 3 + 0] = s.referenceHomozygous_; by_variant[l * 3 + 1] = s.minorHeterozygous_; by_variant[l * 3 + 2] = s.minorHomozygous_;
    }
    for (uint32_t g = 0; g < N; ++g) {
      if (!present[g]) continue;
      auto s = variant_db.summaryByGenome(genome_ids[g]);
      by_genome[g * 3 + 0] = s.referenceHomozygous_; by_genome[g * 3 + 1] = s.minorHeterozygous_; by_genome[g * 3 + 2] = s.minorHomozygous_;
    }
    auto s = variant_db.populationSummary();
    pop[0] = s.referenceHomozygous_; pop[1] = s.minorHeterozygous_; pop[2] = s.minorHomozygous_;
    const double seconds = std::chrono::duration<double>(Clock::now() - t0).count();
    std::fprintf(stderr, "[ref] VariantDBVariant   %zu variants x %zu genomes  %.3f s\n", variant_db.variantMap().size(), variant_db.genomeMap().size(), seconds);
    out.addU64("summary_by_variant", {L, 3}, by_variant);     // refHom, het, minorHom for the locus' "A>G" column
    out.addU32("variant_present", {L}, variant_present);
    out.addU64("summary_by_genome", {N, 3}, by_genome);
    out.addU64("summary_population", {3}, pop);
    out.addF64("variantdb_meta", {3}, std::vector<double>{double(variant_db.variantMap().size()), double(variant_db.genomeMap().size()), seconds});
  }
  // ---- CalcFWS (kga_PfEMP/kga_analysis_PfEMP_FWS.cpp:15-101) on a population whose variants carry the INFO AF ----------
  if (g_opt.fws) {
    t0 = Clock::now();
    kglref::BuiltPopulations pf = kglref::buildPopulations(flat, kgl::DataSourceEnum::Genome1000, kgl::DataSourceEnum::Genome1000, true);
    kga::CalcFWS calc_fws;
    calc_fws.calcFwsStatistics(pf.diploid);
    constexpr size_t B = kga::FWS_FREQUENCY_ARRAY_SIZE;
    std::vector<uint64_t> fws_genome(size_t(N) * B * 3, 0), fws_variant(size_t(L) * 3, 0);
    std::vector<uint32_t> fws_genome_present(N, 0), fws_variant_present(L, 0);
    for (uint32_t g = 0; g < N; ++g) {
      auto it = calc_fws.getGenomeMap().find(pf.genome_ids[g]);
      if (it == calc_fws.getGenomeMap().end()) continue;
      fws_genome_present[g] = 1;
      for (size_t b = 0; b < B; ++b) {
        const auto& s = it->second[b];
        fws_genome[(g * B + b) * 3 + 0] = s.referenceHomozygous_; fws_genome[(g * B + b) * 3 + 1] = s.minorHeterozygous_;
        fws_genome[(g * B + b) * 3 + 2] = s.minorHomozygous_;
      }
    }
    for (uint32_t l = 0; l < L; ++l) {
      auto it = calc_fws.getVariantMap().find(pf.locus_variant[l]->HGVS());
      if (it == calc_fws.getVariantMap().end()) continue;
      fws_variant_present[l] = 1;
      fws_variant[l * 3 + 0] = it->second.referenceHomozygous_; fws_variant[l * 3 + 1] = it->second.minorHeterozygous_;
      fws_variant[l * 3 + 2] = it->second.minorHomozygous_;
    }
    std::fprintf(stderr, "[ref] CalcFWS            %zu genomes x %zu bins  %.3f s\n", calc_fws.getGenomeMap().size(), B,
                 std::chrono::duration<double>(Clock::now() - t0).count());
    out.addU64("fws_genome", {N, B, 3}, fws_genome);            // refHom, het, minorHom per genome per AF bin
    out.addU32("fws_genome_present", {N}, fws_genome_present);
    out.addU64("fws_variant", {L, 3}, fws_variant);             // per "A>G" variant over all genomes
    out.addU32("fws_variant_present", {L}, fws_variant_present);
    if (flat.M() > 0) {                                         // one map entry per allele of a multi-allelic locus
      static const char* const kAlt[3] = {"G", "C", "T"};
      const size_t M = flat.M();
      std::vector<uint64_t> multi_variant(M * 3 * 3, 0);
      std::vector<uint32_t> multi_present(M * 3, 0);
      for (size_t m = 0; m < M; ++m)
        for (size_t a = 0; a < 3; ++a) {
          const kgl::Variant probe(kglref::kContig, flat.offsets[flat.multi_rows[m]], kgl::VariantPhase::UNPHASED, "",
                                   kgl::DNA5SequenceLinear(kgl::StringDNA5("A")), kgl::DNA5SequenceLinear(kgl::StringDNA5(kAlt[a])),
                                   pf.locus_variant[flat.multi_rows[m]]->evidence());
          auto it = calc_fws.getVariantMap().find(probe.HGVS());
          if (it == calc_fws.getVariantMap().end()) continue;
          multi_present[m * 3 + a] = 1;
          multi_variant[(m * 3 + a) * 3 + 0] = it->second.referenceHomozygous_; multi_variant[(m * 3 + a) * 3 + 1] = it->second.minorHeterozygous_;
          multi_variant[(m * 3 + a) * 3 + 2] = it->second.minorHomozygous_;
        }
      out.addU64("fws_multi_variant", {M, 3, 3}, multi_variant);      // refHom, het, minorHom per allele slot
      out.addU32("fws_multi_variant_present", {M, 3}, multi_present);
    }
    // HeteroHomoZygous::updateVariantAnalysisType (kga_PfEMP/kga_analysis_PfEMP_heterozygous.cpp:61-105), applied to every
    // offset of every genome exactly as analyzeVariantPopulation does (:14-58); the harness population has one contig.
    std::vector<uint64_t> hh(size_t(N) * 7, 0);
    for (uint32_t g = 0; g < N; ++g) {
      auto genome_opt = pf.diploid->getGenome(pf.genome_ids[g]);
      if (!genome_opt) continue;
      kga::VariantAnalysisType rec;
      for (auto const& [contig_id, contig_ptr] : genome_opt.value()->getMap())
        for (auto const& [offset, offset_ptr] : contig_ptr->getMap()) kga::HeteroHomoZygous::updateVariantAnalysisType(offset_ptr, rec);
      hh[g * 7 + 0] = rec.total_variants_; hh[g * 7 + 1] = rec.snp_count_; hh[g * 7 + 2] = rec.indel_count_;
      hh[g * 7 + 3] = rec.homozygous_minor_alleles_; hh[g * 7 + 4] = rec.heterozygous_minor_alleles_;
      hh[g * 7 + 5] = rec.heterozygous_reference_minor_alleles_; hh[g * 7 + 6] = rec.homozygous_reference_alleles_;
    }
    out.addU64("hetero_homo", {N, 7}, hh);   // total, snp, indel, homMinor, hetMinor, hetRefMinor, homRef
  }
  out.write(g_opt.out_path);
}

std::vector<std::string> split(const std::string& s, char d) {
  std::vector<std::string> r; std::stringstream ss(s); std::string t;
  while (std::getline(ss, t, d)) if (!t.empty()) r.push_back(t);
  return r;
}

}  // namespace

class HarnessEnv {
 public:
  HarnessEnv() = delete;
  inline static constexpr const char* VERSION = "1";
  inline static constexpr const char* MODULE_NAME = "kglRefHarness";
  inline static constexpr size_t MAX_ERROR_MESSAGES = 100000;
  inline static constexpr size_t MAX_WARNING_MESSAGES = 1000;

  static void executeApp() { run(); }

  [[nodiscard]] static bool parseCommandLine(int argc, char const** argv) {
    std::vector<std::string> pos;
    for (int i = 1; i < argc; ++i) {
      std::string a = argv[i];
      auto next = [&]() -> std::string { if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", a.c_str()); std::exit(2); } return argv[++i]; };
      if (a == "--algos") g_opt.algos = split(next(), ',');
      else if (a == "--min-af") g_opt.min_af = std::stod(next());
      else if (a == "--max-af") g_opt.max_af = std::stod(next());
      else if (a == "--spacing") g_opt.spacing = std::stoull(next());
      else if (a == "--count") g_opt.count = std::stoull(next());
      else if (a == "--lower") g_opt.lower = std::stoull(next());
      else if (a == "--upper") g_opt.upper = std::stoull(next());
      else if (a == "--threads") g_opt.threads = std::stoull(next());
      else if (a == "--grid") g_opt.grid = std::stoull(next());
      else if (a == "--seed") g_opt.seed = std::stol(next());
      else if (a == "--repeat") g_opt.repeat = std::max<size_t>(1, std::stoull(next()));
      else if (a == "--no-variantdb") g_opt.variantdb = false;
      else if (a == "--fws") g_opt.fws = true;
      else if (a == "--verbose") g_opt.quiet = false;
      else pos.push_back(a);
    }
    if (pos.size() != 2) { std::fprintf(stderr, "usage: kgl_ref_harness IN.flat OUT.tens [options]\n"); return false; }
    g_opt.in_path = pos[0]; g_opt.out_path = pos[1];
    // The reference logs one line per genome to stdout (kga_analysis_inbreed_calc.cpp:209,301,346,425).
    if (g_opt.quiet) { if (!std::freopen("/dev/null", "w", stdout)) return false; }
    return true;
  }

  [[nodiscard]] static std::unique_ptr<kel::ExecEnvLogger> createLogger() {
    const char* log_file = std::getenv("KGL_REF_LOG");
    return kel::ExecEnv::createLogger(MODULE_NAME, log_file ? log_file : "/tmp/kgl_ref_harness.log", MAX_ERROR_MESSAGES, MAX_WARNING_MESSAGES);
  }
};

int main(int argc, const char* argv[]) { return kel::ExecEnv::runApplication<HarnessEnv>(argc, argv); }
