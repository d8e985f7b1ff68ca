// plugin_harness.cpp -- TEST INFRASTRUCTURE (oracle/_ref build only). Not part of the product.
//
// Drives BOTH analyses through the reference's own plugin interface, on the same in-memory data, the way
// ExecutePackage::executeActive does (kgl_app/kgl_package.cpp:17-77): factory lookup by ident in
// VirtualAnalysis::analysis_factory_map_ (kgl_package_analysis.cpp:20-24), then initializeAnalysis once,
// fileReadAnalysis per data object, iterationAnalysis, finalizeAnalysis:
//   INBREED       the UNMODIFIED reference analysis, kga::InbreedAnalysis (kga_analytic/kga_inbreed/kga_analysis_inbreed.cpp)
//   INBREED_B200  the drop-in, kga::InbreedB200Analysis (kgl_gene_b200/host/kga_analysis_inbreed_b200.cpp) over libkgl_b200.so
// Each writes <work_dir>/<ident>/<OutputFile>.csv with the reference's own CSV writer; tests/test_plugin_dropin.py
// compares the two files. The populations, the PED resource and the parameter block are built in memory (no XML, VCF
// or PED files: those parsers need Boost, which this image lacks).
//
// It also dumps what the host flattener produced (KGLFLAT1) so that the test can check it cell by cell against the
// flat population the reference containers were built from.
#include "kel_exec_env_app.h"
#include "kgl_package_analysis_virtual.h"
#include "kgl_properties_resource.h"
#include "kgl_hsgenealogy_parser.h"
#include "kga_analysis_inbreed.h"
#include "kga_analysis_inbreed_b200.h"
#include "kga_analysis_pfemp_b200.h"
#include "kga_analysis_PfEMP_FWS.h"
#include "kga_analysis_PfEMP_heterozygous.h"
#include "kgl_variant_factory_1000_impl.h"
#include "kgl_variant_factory_vcf_evidence_analysis.h"

#include "flat_io.h"
#include "ref_population.h"

#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <set>

namespace kel = kellerberrin;
namespace kgl = kellerberrin::genome;
namespace kga = kellerberrin::genome::analysis;

// The registration point (kga_analytic/kga_analysis_factory.cpp:31-43), reduced to the two analyses under test.
kgl::VirtualAnalysis::AnalysisFactoryMap const kgl::VirtualAnalysis::analysis_factory_map_ = {
    {kga::InbreedAnalysis::IDENT, kga::InbreedAnalysis::factory},
    {kga::InbreedB200Analysis::IDENT, kga::InbreedB200Analysis::factory},
};

namespace {

struct Options {
  std::string in_path, work_dir;
  std::string algorithm{"Simple"};
  std::string min_af{"0.0"}, max_af{"1.0"}, spacing{"0"}, count{"1000"}, lower{"0"}, upper{"1000000000"};
  bool run_reference{true}, run_b200{true};
  bool pfemp{false};          // --pfemp: the kga_PfEMP consumer (HeteroHomoZygous) instead of the INBREED analysis
  bool vcf{false};            // --vcf: IN is a plain-text VCF, parsed by the reference's own 1000 Genomes parser
};
Options g_opt;

// kga::InbreedAnalysis looks the PED data up as resource type "genomeGenealogy" (kga_analysis_inbreed.cpp:29) while
// HsGenomeGenealogyData registers itself as "genomeAux" (kgl_hsgenome_aux.h:85). The tag base makes one object answer
// to the first type; dynamic_pointer_cast<HsGenomeGenealogyData> then cross-casts inside the complete object.
struct GenealogyTypeTag : public kgl::ResourceBase {
  GenealogyTypeTag() : kgl::ResourceBase(kgl::ResourceProperties::GENEALOGY_RESOURCE_ID_, "harnessPed") {}
};
struct HarnessGenealogy : public GenealogyTypeTag, public kgl::HsGenomeGenealogyData {
  HarnessGenealogy() : kgl::HsGenomeGenealogyData("harnessPed") {}
};

bool runAnalysis(const std::string& ident, const kgl::ActiveParameterList& parameters,
                 const std::shared_ptr<const kgl::AnalysisResources>& resources,
                 const std::vector<std::shared_ptr<const kgl::DataDB>>& data_files) {
  auto found = kgl::VirtualAnalysis::analysis_factory_map_.find(ident);
  if (found == kgl::VirtualAnalysis::analysis_factory_map_.end()) return false;
  std::unique_ptr<kgl::VirtualAnalysis> analysis = found->second();
  const std::string work_dir = g_opt.work_dir + "/" + ident;
  std::filesystem::create_directories(work_dir);
  if (!analysis->initializeAnalysis(work_dir, parameters, resources)) { std::fprintf(stderr, "[plugin] %s: initializeAnalysis failed\n", ident.c_str()); return false; }
  for (auto const& data : data_files)
    if (!analysis->fileReadAnalysis(data)) { std::fprintf(stderr, "[plugin] %s: fileReadAnalysis failed\n", ident.c_str()); return false; }
  if (!analysis->iterationAnalysis()) { std::fprintf(stderr, "[plugin] %s: iterationAnalysis failed\n", ident.c_str()); return false; }
  if (!analysis->finalizeAnalysis()) { std::fprintf(stderr, "[plugin] %s: finalizeAnalysis failed\n", ident.c_str()); return false; }
  std::fprintf(stderr, "[plugin] %s: wrote %s/harness_out.csv\n", ident.c_str(), work_dir.c_str());
  return true;
}

int g_exit_code = 0;

// ---- kga_PfEMP: HeteroHomoZygous (reference) and HeteroHomoB200 (product) on the same Pf7-style population -----------------
// The call order of kga::PfEMPAnalysis (kga_analysis_PfEMP.cpp:92-109 fileReadAnalysis, :129-150 finalizeAnalysis):
// analyzeVariantPopulation, location_summary, UpdateSampleLocation, write_variant_results. The Pf7 sample / FWS resources are
// built in memory (their file parsers and the physical-distance resource need Boost): genome g lives in city "City<g % 7>" of
// country "Country<(g % 7) / 3>"; every fifth genome fails QC, so some cities fall below MINIMUM_LOCATION_SAMPLES_ and take
// their country's aggregate. The reference's location_summary asks Pf7SampleLocation::sampleRadius (stubbed here, returns
// nothing), so its LocationSummaryMap is assembled from the reference's own aggregateResults with the same arithmetic
// (kga_analysis_PfEMP_heterozygous.cpp:301-352); UpdateSampleLocation and the CSV writer are the reference's.
void runPfEMP(const kglflat::Flat& flat) {
  const bool unphased = (flat.hdr.flags & kglflat::FLAG_UNPHASED) != 0;
  const auto source = unphased ? kgl::DataSourceEnum::Falciparum : kgl::DataSourceEnum::Genome1000;
  kglref::BuiltPopulations built = kglref::buildPopulations(flat, kgl::DataSourceEnum::Genome1000, source, true);
  kgl::Pf7SampleVector samples;
  kgl::Pf7FwsVector fws;
  kga::LocationSamplesMap locations;
  for (uint32_t g = 0; g < flat.N(); ++g) {
    kgl::Pf7SampleRecord rec;
    rec.Pf7Sample_id = built.genome_ids[g];
    rec.study_ = "Study" + std::to_string(g % 3);
    rec.location1_ = "City" + std::to_string(g % 7);
    rec.country_ = "Country" + std::to_string((g % 7) / 3);
    rec.year_ = std::to_string(2010 + g % 9);
    rec.qc_pass_ = (g % 5 == 4) ? "False" : "True";
    samples.push_back(rec);
    fws.push_back(kgl::Pf7FwsRecord{built.genome_ids[g], 0.5 + 0.001 * double(g % 400)});
    for (const std::string& loc : {rec.location1_, rec.country_}) {
      auto& l = locations[loc];
      l.location_type = loc[1] == 'i' ? kgl::LocationType::City : kgl::LocationType::Country;
      l.city = rec.location1_; l.country = rec.country_; l.region = "Region" + rec.country_.substr(7);
      l.samples.push_back(rec.Pf7Sample_id);
    }
  }
  auto sample_ptr = std::make_shared<const kgl::Pf7SampleResource>("harnessPf7Samples", samples);
  auto fws_ptr = std::make_shared<const kgl::Pf7FwsResource>("harnessPf7Fws", fws);
  std::set<kgl::GenomeId_t> pass;
  for (auto const& [id, rec] : sample_ptr->getMap()) if (rec.pass()) pass.insert(id);

  if (g_opt.run_reference) {
    kga::HeteroHomoZygous reference;
    reference.analyzeVariantPopulation(built.diploid, fws_ptr, sample_ptr);
    kga::LocationSummaryMap summary;
    for (auto const& [location, record] : locations) {
      auto aggregated = reference.aggregateResults(record.samples);
      kga::LocationSummary s;
      s.location_ = location; s.location_type_ = record.location_type; s.city_ = record.city; s.country_ = record.country; s.region_ = record.region;
      s.radii_samples_ = record.samples.size();
      for (auto const& id : record.samples) s.radii_samples_OK_ += pass.contains(id) ? 1 : 0;
      s.total_variants_ = aggregated.total_variants_;
      s.homozygous_reference_alleles_ = aggregated.homozygous_reference_alleles_;
      s.heterozygous_reference_minor_alleles_ = aggregated.heterozygous_reference_minor_alleles_;
      s.homozygous_minor_alleles_ = aggregated.homozygous_minor_alleles_;
      s.heterozygous_minor_alleles_ = aggregated.heterozygous_minor_alleles_;
      s.snp_count_ = aggregated.snp_count_; s.indel_count_ = aggregated.indel_count_;
      summary[location] = s;
    }
    reference.UpdateSampleLocation(summary);
    std::filesystem::create_directories(g_opt.work_dir + "/PFEMP");
    reference.write_variant_results(g_opt.work_dir + "/PFEMP/hetero_homo.csv", summary);
    std::fprintf(stderr, "[plugin] PFEMP reference: wrote %s/PFEMP/hetero_homo.csv\n", g_opt.work_dir.c_str());
    // CalcFWS as kga::PfEMPAnalysis drives it (kga_analysis_PfEMP.cpp:104 calcFwsStatistics; the two writers in finalizeAnalysis)
    kga::CalcFWS calc_fws;
    calc_fws.calcFwsStatistics(built.diploid);
    calc_fws.writeGenomeResults(fws_ptr, g_opt.work_dir + "/PFEMP/fws_genome.csv");
    calc_fws.writeVariantResults(g_opt.work_dir + "/PFEMP/fws_variant.csv");
  }
  if (g_opt.run_b200) {
    kga::HeteroHomoB200 product;
    if (!product.analyzeVariantPopulation(built.diploid, fws_ptr, sample_ptr)) { std::fprintf(stderr, "[plugin] PFEMP_B200: analyzeVariantPopulation failed\n"); g_exit_code = 6; return; }
    auto summary = product.location_summary(sample_ptr, locations, 0.0, fws_ptr);
    product.UpdateSampleLocation(summary, locations, sample_ptr);
    std::filesystem::create_directories(g_opt.work_dir + "/PFEMP_B200");
    product.write_variant_results(g_opt.work_dir + "/PFEMP_B200/hetero_homo.csv", summary);
    std::fprintf(stderr, "[plugin] PFEMP_B200: wrote %s/PFEMP_B200/hetero_homo.csv\n", g_opt.work_dir.c_str());
    kga::CalcFwsB200 calc_fws;
    if (calc_fws.calcFwsStatistics(built.diploid)) {
      calc_fws.writeGenomeResults(fws_ptr, g_opt.work_dir + "/PFEMP_B200/fws_genome.csv");
      calc_fws.writeVariantResults(g_opt.work_dir + "/PFEMP_B200/fws_variant.csv");
    } else {
      std::fprintf(stderr, "[plugin] PFEMP_B200: CalcFwsB200 refused the population\n");
    }
  }
}

// ---- N2 pin: the reference's own VCF parser (Genome1000VCFImpl over VCFReaderMT / ParseVCF, kgl_parser/kgl_variant_factory_1000_impl.cpp,
// kgl_variant_factory_readvcf_impl.cpp, kgl_variant_vcf_impl.cpp) reads a plain-text VCF into a PopulationDB; the product's
// flattener (flattenSelf: the population as its own locus list, frequency = the variant's INFO AF) turns it into the flat form,
// which tests/test_plugin_dropin.py compares cell by cell with what kgl_b200_vcf_ingest makes of the same file.
void runVcf() {
  auto population = std::make_shared<kgl::PopulationDB>("VCF_POPULATION", kgl::DataSourceEnum::Genome1000);
  kgl::ContigAliasMap alias_map;
  alias_map.setAlias("22", "22", "autosome");
  kgl::EvidenceInfoSet info_set;
  for (int k = 0; k < 6; ++k) info_set.insert(kglref::afFields(kgl::DataSourceEnum::Genome1000)[k]);
  kgl::Genome1000VCFImpl parser(population, nullptr, alias_map, info_set);
  parser.readParseVCFImpl(g_opt.in_path);
  auto allele_frequency = [](const kgl::Variant& variant) -> std::optional<double> {
    auto info_opt = kgl::InfoEvidenceAnalysis::getTypedInfoData<std::vector<double>>(variant, "AF");
    if (!info_opt) return std::nullopt;
    const size_t alt_index = variant.evidence().altVariantIndex();
    if (info_opt.value().size() <= alt_index) return std::nullopt;
    return info_opt.value()[alt_index];
  };
  auto flat_opt = kgl::b200::PopulationFlattener::flattenSelf(*population, "22", allele_frequency, false);
  if (!flat_opt) { std::fprintf(stderr, "[plugin] vcf: flatten failed\n"); g_exit_code = 7; return; }
  auto const& f = flat_opt.value();
  kglflat::Flat out;
  std::memset(&out.hdr, 0, sizeof out.hdr);
  out.hdr.n_genomes = uint32_t(f.nGenomes()); out.hdr.n_loci = uint32_t(f.nLoci()); out.hdr.n_superpop = 6;
  out.hdr.row_bytes = uint32_t(f.row_bytes); out.hdr.reserved[0] = uint32_t(f.nMulti());
  out.offsets = f.offsets; out.af = f.af; out.superpop = f.superpop; out.packed = f.packed;
  out.multi_rows = f.multi_rows; out.multi_af = f.multi_af; out.multi_cells = f.multi_cells;
  kglflat::writeFlat(g_opt.work_dir + "/flattened.flat", out);
  std::ofstream ids(g_opt.work_dir + "/flattened_genomes.txt");
  for (auto const& id : f.genome_ids) ids << id << '\n';
  std::fprintf(stderr, "[plugin] vcf: reference parser -> %zu genomes, %zu variant entries; flattened %zu loci (%zu multi-allelic)\n",
               population->getMap().size(), population->variantCount(), size_t(f.nLoci()), size_t(f.nMulti()));
}

void run() {
  if (g_opt.vcf) { runVcf(); return; }
  const kglflat::Flat flat = kglflat::readFlat(g_opt.in_path);
  if (g_opt.pfemp) { runPfEMP(flat); return; }
  const bool unphased = (flat.hdr.flags & kglflat::FLAG_UNPHASED) != 0;
  // AF data as a gnomAD 3.1 file (UnphasedMonoGenome), genotypes as 1000 Genomes (DiploidPhased) or Pf (DiploidUnphased).
  const auto diploid_source = unphased ? kgl::DataSourceEnum::Falciparum : kgl::DataSourceEnum::Genome1000;
  kglref::BuiltPopulations built = kglref::buildPopulations(flat, kgl::DataSourceEnum::Gnomad3_1, diploid_source);

  // ---- PED resource -------------------------------------------------------------------------------------
  auto genealogy = std::make_shared<HarnessGenealogy>();
  for (uint32_t g = 0; g < flat.N(); ++g) {
    const std::string super_pop = kglref::kSuperPops[flat.superpop[g]];
    kgl::HsGenealogyRecord record("FAM" + std::to_string(g), built.genome_ids[g], "0", "0", (g & 1) ? "1" : "2", "0",
                                  "POP_" + super_pop, "population of " + super_pop, super_pop, super_pop + " super population",
                                  "unrel", "0", "0", "0", "");
    genealogy->addGenealogyRecord(record);
  }
  genealogy->refreshPopulationLists();
  auto resources = std::make_shared<kgl::AnalysisResources>();
  resources->addResource(std::shared_ptr<const kgl::ResourceBase>(genealogy, static_cast<const GenealogyTypeTag*>(genealogy.get())));

  // ---- parameter blocks, the fields of kga_analysis_inbreed_args.h:164-172; --algo A,B,C makes one block per algorithm (output
  // harness_out.csv for a single algorithm, harness_out_<algorithm>.csv otherwise) -----------------------------------------
  std::vector<std::string> algorithms;
  { std::string t; for (char ch : g_opt.algorithm + ",") { if (ch == ',') { if (!t.empty()) algorithms.push_back(t); t.clear(); } else t += ch; } }
  std::vector<kgl::ParameterMap> blocks;
  for (auto const& algorithm : algorithms) {
    kgl::ParameterMap block;
    block.insert("AnalysisType", "FALSE");
    block.insert("OutputFile", algorithms.size() == 1 ? std::string("harness_out") : "harness_out_" + algorithm);
    block.insert("Algorithm", algorithm);
    block.insert("MinAlleleFreq", g_opt.min_af);
    block.insert("MaxAlleleFreq", g_opt.max_af);
    block.insert("LowerWindow", g_opt.lower);
    block.insert("UpperWindow", g_opt.upper);
    block.insert("LociiCount", g_opt.count);
    block.insert("SamplingDistance", g_opt.spacing);
    blocks.push_back(block);
  }
  kgl::ActiveParameterList parameters;
  parameters.addNamedParameterVector(kgl::NamedParameterVector{"HarnessBlock", kgl::ParameterVector{blocks}});

  const std::vector<std::shared_ptr<const kgl::DataDB>> data_files{built.diploid, built.af_population};

  // ---- the flattener on its own: dump what the product's host layer makes of the reference containers -----
  {
    auto af_filtered = built.af_population;   // the harness' AF variants are all SNP + PASS
    auto lookup = [&](const kgl::GenomeId_t& id) -> std::optional<std::string> {
      auto rec = genealogy->getGenomeGenealogyRecord(id);
      if (!rec) return std::nullopt;
      return rec.value().superPopulation();
    };
    auto flat_opt = kgl::b200::PopulationFlattener::flatten(*built.diploid, *af_filtered, lookup, unphased);
    if (!flat_opt) { std::fprintf(stderr, "[plugin] flatten failed\n"); g_exit_code = 3; return; }
    auto const& f = flat_opt.value();
    kglflat::Flat out;
    out.hdr = flat.hdr;
    out.hdr.n_genomes = uint32_t(f.nGenomes()); out.hdr.n_loci = uint32_t(f.nLoci()); out.hdr.n_superpop = 6;
    out.hdr.row_bytes = uint32_t(f.row_bytes); out.hdr.flags = f.unphased ? kglflat::FLAG_UNPHASED : 0;
    out.offsets = f.offsets; out.af = f.af; out.superpop = f.superpop; out.packed = f.packed;
    out.hdr.reserved[0] = uint32_t(f.nMulti());
    out.multi_rows = f.multi_rows; out.multi_af = f.multi_af; out.multi_cells = f.multi_cells;
    kglflat::writeFlat(g_opt.work_dir + "/flattened.flat", out);
    std::ofstream ids(g_opt.work_dir + "/flattened_genomes.txt");
    for (auto const& id : f.genome_ids) ids << id << '\n';
    std::fprintf(stderr, "[plugin] flattener: %zu genomes x %zu loci, %zu multi-allelic loci, %zu mixed-phase cells\n",
                 size_t(f.nGenomes()), size_t(f.nLoci()), size_t(f.nMulti()), f.mixed_phase_cells);
  }

  if (g_opt.run_reference && !runAnalysis(kga::InbreedAnalysis::IDENT, parameters, resources, data_files)) g_exit_code = 4;
  if (g_opt.run_b200 && !runAnalysis(kga::InbreedB200Analysis::IDENT, parameters, resources, data_files)) g_exit_code = 5;
}

}  // namespace

class PluginHarnessEnv {
 public:
  PluginHarnessEnv() = delete;
  inline static constexpr const char* VERSION = "1";
  inline static constexpr const char* MODULE_NAME = "kglPluginHarness";
  inline static constexpr size_t MAX_ERROR_MESSAGES = 100000;
  inline static constexpr size_t MAX_WARNING_MESSAGES = 1000;

  static void executeApp() { run(); }

  [[nodiscard]] static bool parseCommandLine(int argc, char const** argv) {
    std::vector<std::string> pos;
    for (int i = 1; i < argc; ++i) {
      std::string a = argv[i];
      auto next = [&]() -> std::string { if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", a.c_str()); std::exit(2); } return argv[++i]; };
      if (a == "--algo") g_opt.algorithm = next();
      else if (a == "--min-af") g_opt.min_af = next();
      else if (a == "--max-af") g_opt.max_af = next();
      else if (a == "--spacing") g_opt.spacing = next();
      else if (a == "--count") g_opt.count = next();
      else if (a == "--lower") g_opt.lower = next();
      else if (a == "--upper") g_opt.upper = next();
      else if (a == "--no-reference") g_opt.run_reference = false;
      else if (a == "--no-b200") g_opt.run_b200 = false;
      else if (a == "--pfemp") g_opt.pfemp = true;
      else if (a == "--vcf") g_opt.vcf = true;
      else pos.push_back(a);
    }
    if (pos.size() != 2) { std::fprintf(stderr, "usage: kgl_plugin_harness IN.flat WORK_DIR [options]\n"); return false; }
    g_opt.in_path = pos[0]; g_opt.work_dir = pos[1];
    std::filesystem::create_directories(g_opt.work_dir);
    // The reference logs one line per genome to stdout (kga_analysis_inbreed_calc.cpp:209,301,346,425).
    if (!std::getenv("KGL_HARNESS_VERBOSE")) { if (!std::freopen("/dev/null", "w", stdout)) return false; }
    return true;
  }

  [[nodiscard]] static std::unique_ptr<kel::ExecEnvLogger> createLogger() {
    const char* log_file = std::getenv("KGL_REF_LOG");
    return kel::ExecEnv::createLogger(MODULE_NAME, log_file ? log_file : "/tmp/kgl_plugin_harness.log", MAX_ERROR_MESSAGES, MAX_WARNING_MESSAGES);
  }
};

int main(int argc, const char* argv[]) {
  const int rc = kel::ExecEnv::runApplication<PluginHarnessEnv>(argc, argv);
  return rc != 0 ? rc : g_exit_code;
}
