// kga_analysis_inbreed_b200.cpp -- see kga_analysis_inbreed_b200.h.
#include "kga_analysis_inbreed_b200.h"

#include "kga_analysis_inbreed_execute.h"
#include "kga_analysis_inbreed_locus.h"
#include "kga_analysis_inbreed_calc.h"
#include "kgl_variant_filter_db_variant.h"
#include "kgl_variant_db_freq.h"
#include "kgl_properties_resource.h"
#include "kel_utility.h"

#include "kgl_b200.h"

#include <cstdlib>

namespace kga = kellerberrin::genome::analysis;
namespace kgl = kellerberrin::genome;
namespace b200 = kellerberrin::genome::b200;

namespace {

// InbreedingCalculation::inbreeding_algo_map_ (kga_analysis_inbreed_calc.h:103-118) -> C ABI algorithm code.
std::optional<int> algorithmCode(const std::string& name) {
  if (name == kga::InbreedingCalculation::SIMPLE_F) return KGL_B200_ALGO_SIMPLE;
  if (name == kga::InbreedingCalculation::RITLAND_LOCUS_F) return KGL_B200_ALGO_RITLAND;
  if (name == kga::InbreedingCalculation::HALL_ME_IBD) return KGL_B200_ALGO_HALLME;
  if (name == kga::InbreedingCalculation::LOGLIKELIHOOD_F) return KGL_B200_ALGO_LOGLIKELIHOOD;
  return std::nullopt;
}

}  // namespace

kga::InbreedB200Analysis::~InbreedB200Analysis() {
  if (context_ != nullptr) kgl_b200_destroy(context_);
}

bool kga::InbreedB200Analysis::ensureContext() {
  if (context_ != nullptr) return true;
  if (const char* env = std::getenv("KGL_B200_DEVICE")) device_ = std::atoi(env);
  const int rc = kgl_b200_create(device_, &context_);
  if (rc != KGL_B200_OK) {
    // No CPU fallback: the analysis is disabled by returning false (kgl_package_analysis.cpp:72-78).
    ExecEnv::log().error("InbreedB200Analysis; cannot create a device context on GPU {}: {}", device_, kgl_b200_last_error(nullptr));
    context_ = nullptr;
    return false;
  }
  return true;
}

// Setup the analytics to process VCF data (mirrors kga_analysis_inbreed.cpp:18-41).
bool kga::InbreedB200Analysis::initializeAnalysis(const std::string& work_directory,
                                                  const ActiveParameterList& named_parameters,
                                                  const std::shared_ptr<const AnalysisResources>& resource_ptr) {

  ExecEnv::log().info("Analysis Id: {} initialized with work directory: {} ({})", ident(), work_directory, kgl_b200_version());
  for (auto const& [parameter_ident, parameter_value] : named_parameters.getMap()) {

    ExecEnv::log().info("Initialize Analysis Id: {}, initialized with parameter: {}", ident(), parameter_ident);

  }

  // The reference asks for resource type GENEALOGY_RESOURCE_ID_ (kga_analysis_inbreed.cpp:29) while HsGenomeGenealogyData
  // registers itself as GENOMEAUX_RESOURCE_ID_ (kgl_hsgenome_aux.h:85); accept either, without the process-ending
  // critical() of getSingleResource.
  for (const char* resource_type : {ResourceProperties::GENEALOGY_RESOURCE_ID_, ResourceProperties::GENOMEAUX_RESOURCE_ID_}) {
    for (auto const& resource : resource_ptr->getResources(resource_type)) {
      if (auto ped = std::dynamic_pointer_cast<const HsGenomeGenealogyData>(resource)) { genealogy_data_ = ped; break; }
    }
    if (genealogy_data_) break;
  }
  if (not genealogy_data_) {
    ExecEnv::log().error("InbreedB200Analysis::initializeAnalysis; no genome genealogy (PED) resource supplied");
    return false;
  }
  work_directory_ = work_directory;

  for (auto const& parameter : InbreedArguments::extractParameters(named_parameters)) {

    parameter_output_vector_.emplace_back(InbreedParamOutput(parameter));

  }

  return ensureContext();

}

// This function superclasses the data objects and stores them for further use (mirrors kga_analysis_inbreed.cpp:43-86).
bool kga::InbreedB200Analysis::fileReadAnalysis(std::shared_ptr<const DataDB> data_object_ptr) {

  ExecEnv::log().info("Analysis: {}, begin processing data file", ident(), data_object_ptr->fileId());

  auto file_characteristic = data_object_ptr->dataCharacteristic();

  if (file_characteristic.data_structure == DataStructureEnum::DiploidPhased
      or file_characteristic.data_structure == DataStructureEnum::DiploidUnphased) {

    diploid_population_ = std::dynamic_pointer_cast<const PopulationDB>(data_object_ptr);
    diploid_is_unphased_ = file_characteristic.data_structure == DataStructureEnum::DiploidUnphased;

    if (not diploid_population_) {

      ExecEnv::log().error("InbreedB200Analysis::fileReadAnalysis, Analysis: {}, file: {} is not a Diploid Population", ident(), data_object_ptr->fileId());
      return false;

    }

  }

  if (file_characteristic.data_structure == DataStructureEnum::UnphasedMonoGenome) {

    unphased_population_ = std::dynamic_pointer_cast<const PopulationDB>(data_object_ptr);

    if (not unphased_population_) {

      ExecEnv::log().error("InbreedB200Analysis::fileReadAnalysis, Analysis: {}, file: {} is not an Unphased Population", ident(), data_object_ptr->fileId());
      return false;

    }

    // Only want SNP variants and variants that passed all VCF filters.
    unphased_population_ = unphased_population_->viewFilter(AndFilter(SNPFilter(), PassFilter()));

  }

  ExecEnv::log().info("Analysis: {}, completed data file: {}", ident(), data_object_ptr->fileId());

  return true;

}

// Perform the genetic analysis per iteration (mirrors kga_analysis_inbreed.cpp:88-116).
bool kga::InbreedB200Analysis::iterationAnalysis() {

  ExecEnv::log().info("Iteration Analysis called for Analysis Id: {}", ident());

  if (not diploid_population_ or not unphased_population_ or not genealogy_data_) {

    ExecEnv::log().error("InbreedB200Analysis::iterationAnalysis; necessary variant databases not supplied");
    return false;

  }

  // One walk over the variant database and ONE upload per iteration, shared by every parameter block (the reference walks
  // the database once per genome, locus and block): everything after it works on flat arrays that are resident on the GPU.
  bool ok = true;
  bool resident = false;
  for (auto& param_output : parameter_output_vector_) {

    // Set the allele frequency source for this population.
    param_output.getParameters().lociiArguments().frequencySource(unphased_population_->dataSource());
    if (param_output.getParameters().analyzeSynthetic()) {

      ok = ExecuteInbreedingAnalysis::executeAnalysis(diploid_population_, unphased_population_, genealogy_data_, param_output) and ok;

    } else {

      if (not resident) {

        if (not uploadPopulation(unphased_population_, *diploid_population_, *genealogy_data_, diploid_is_unphased_)) { ok = false; break; }
        resident = true;

      }
      ok = windowLoop(unphased_population_, param_output) and ok;

    }

  }

  // Clear the data structures.
  diploid_population_ = nullptr;
  unphased_population_ = nullptr;
  flat_.reset();

  return ok;

}

bool kga::InbreedB200Analysis::check(int rc, const char* what) {

  if (rc != KGL_B200_OK) ExecEnv::log().error("InbreedB200Analysis; {} failed [{}]: {}", what, rc, kgl_b200_last_error(context_));
  return rc == KGL_B200_OK;

}

// Flattens the populations (host/kgl_b200_flatten.cpp) and makes them resident on the device.
bool kga::InbreedB200Analysis::uploadPopulation(const std::shared_ptr<const PopulationDB>& unphased_ptr,
                                                const PopulationDB& diploid_population,
                                                const HsGenomeGenealogyData& ped_data,
                                                bool unphased_diploid) {

  if (not ensureContext()) return false;
  flat_.reset();
  auto super_population = [&ped_data](const GenomeId_t& genome_id) -> std::optional<std::string> {
    auto record_opt = ped_data.getGenomeGenealogyRecord(genome_id);
    if (not record_opt) return std::nullopt;
    return record_opt.value().superPopulation();
  };
  auto flat_opt = b200::PopulationFlattener::flatten(diploid_population, *unphased_ptr, super_population, unphased_diploid);
  if (not flat_opt) return false;
  flat_ = std::make_unique<b200::FlatContig>(std::move(flat_opt.value()));
  const b200::FlatContig& flat = *flat_;
  if (flat.nGenomes() == 0 or flat.nLoci() == 0) {

    ExecEnv::log().warn("InbreedB200Analysis::uploadPopulation; contig: {} has no genomes or no loci to analyse", flat.contig_id);
    return true;

  }

  if (not check(kgl_b200_upload_genotypes(context_, flat.nGenomes(), flat.nLoci(), flat.row_bytes, flat.packed.data()), "upload_genotypes")) return false;
  if (not check(kgl_b200_upload_loci(context_, flat.nLoci(), b200::kSuperPopCount, flat.af.data(), flat.offsets.data()), "upload_loci")) return false;
  if (not check(kgl_b200_set_genome_superpop(context_, flat.nGenomes(), flat.superpop.data()), "set_genome_superpop")) return false;
  if (not check(kgl_b200_set_unphased(context_, flat.unphased ? 1 : 0), "set_unphased")) return false;
  if (not check(kgl_b200_upload_multi_allelic(context_, flat.nMulti(), flat.multi_rows.data(), flat.multi_af.data(), flat.multi_cells.data()),
                "upload_multi_allelic")) return false;
  ExecEnv::log().info("InbreedB200Analysis; contig: {}, {} genomes x {} loci resident on the device ({} multi-allelic loci)", flat.contig_id,
                      flat.nGenomes(), flat.nLoci(), flat.nMulti());
  return true;

}

bool kga::InbreedB200Analysis::populationInbreeding(const std::shared_ptr<const PopulationDB>& unphased_ptr,
                                                    const PopulationDB& diploid_population,
                                                    const HsGenomeGenealogyData& ped_data,
                                                    bool unphased_diploid,
                                                    InbreedParamOutput& param_output) {

  if (not uploadPopulation(unphased_ptr, diploid_population, ped_data, unphased_diploid)) return false;
  return windowLoop(unphased_ptr, param_output);

}

// The window loop of InbreedingAnalysis::populationInbreeding (kga_analysis_inbreed_diploid.cpp:45-75) over the resident
// population: windows are defined on the "ALL" super-population by the reference's own RetrieveLociiVector (host, once per
// window); every super-population then selects its loci inside the window on the C ABI.
bool kga::InbreedB200Analysis::windowLoop(const std::shared_ptr<const PopulationDB>& unphased_ptr, InbreedParamOutput& param_output) {

  if (not flat_) return false;
  const b200::FlatContig& flat = *flat_;
  if (flat.nGenomes() == 0 or flat.nLoci() == 0) return true;

  auto algorithm_opt = algorithmCode(param_output.getParameters().inbreedingAlgorthim());
  if (not algorithm_opt) {

    ExecEnv::log().error("InbreedB200Analysis::windowLoop, Inbreeding algorithm not found: {}", param_output.getParameters().inbreedingAlgorthim());
    return false;

  }

  auto const& [af_genome_id, af_genome_ptr] = *unphased_ptr->getMap().begin();
  auto const& [contig_id, contig_ptr] = *af_genome_ptr->getMap().begin();

  // Windows are defined on the "ALL" super-population (kga_analysis_inbreed_diploid.cpp:48-51,69-73): the first LociiCount
  // accepted loci from the window's lower bound, RetrieveLociiVector::getLociiCount -- on the device, kgl_b200_count_loci.
  constexpr uint32_t kAllPop = b200::kSuperPopCount - 1;
  InbreedingParameters local_params = param_output.getParameters();
  uint64_t window_loci = 0, window_upper = 0;
  auto next_window = [&]() {
    auto const& a = local_params.lociiArguments();
    return check(kgl_b200_count_loci(context_, kAllPop, a.lowerOffset(), a.lociiSpacing(), a.lociiCount(), a.minAlleleFrequency(),
                                     a.maxAlleleFrequency(), &window_loci, &window_upper), "count_loci");
  };
  if (not next_window()) return false;
  if (window_loci == 0) return true;           // (the reference dereferences .back() of an empty vector here)
  local_params.lociiArguments().upperOffset(window_upper);

  std::vector<kgl_b200_locus_results> results(flat.nGenomes());
  while (local_params.lociiArguments().upperOffset() < param_output.getParameters().lociiArguments().upperOffset()
         and window_loci >= 100) {

    auto const& args = local_params.lociiArguments();
    if (not check(kgl_b200_select_loci(context_, args.lowerOffset(), args.upperOffset(), args.lociiSpacing(),
                                       args.minAlleleFrequency(), args.maxAlleleFrequency(), nullptr), "select_loci")) return false;
    if (not check(kgl_b200_run_inbreed(context_, algorithm_opt.value(), nullptr, results.data()), "run_inbreed")) return false;

    ResultsMap results_map;
    for (size_t g = 0; g < flat.nGenomes(); ++g) {

      LocusResults locus_results;
      locus_results.genome = flat.genome_ids[g];
      locus_results.major_hetero_count = results[g].major_hetero_count;  locus_results.major_hetero_freq = results[g].major_hetero_freq;
      locus_results.minor_hetero_count = results[g].minor_hetero_count;  locus_results.minor_hetero_freq = results[g].minor_hetero_freq;
      locus_results.minor_homo_count = results[g].minor_homo_count;      locus_results.minor_homo_freq = results[g].minor_homo_freq;
      locus_results.major_homo_count = results[g].major_homo_count;      locus_results.major_homo_freq = results[g].major_homo_freq;
      locus_results.total_allele_count = results[g].total_allele_count;  locus_results.inbred_allele_sum = results[g].inbred_allele_sum;
      results_map[locus_results.genome] = locus_results;

    }

    std::string result_ident = InbreedingResultColumn::generateIdent(contig_id, args.lowerOffset(), args.upperOffset());
    param_output.addColumn(InbreedingResultColumn(result_ident, results_map));

    local_params.lociiArguments().lowerOffset(local_params.lociiArguments().upperOffset());
    if (not next_window()) return false;
    if (window_loci == 0) break;
    local_params.lociiArguments().upperOffset(window_upper);

  }

  return true;

}

// All VCF data has been presented, finalize analysis and write results (mirrors kga_analysis_inbreed.cpp:121-160).
bool kga::InbreedB200Analysis::finalizeAnalysis() {

  ExecEnv::log().info("Finalize called for Analysis Id: {}", ident());

  return writeResults();

}

bool kga::InbreedB200Analysis::writeResults() {

  for (auto& param_output : parameter_output_vector_) {

    if (param_output.getParameters().analyzeSynthetic()) {

      InbreedingOutput::writeSynthetic(param_output, work_directory_);

    } else if (genealogy_data_) {

      InbreedingOutput::writePedResults(param_output, *genealogy_data_, work_directory_);

    } else {

      InbreedingOutput::writeNoPedResults(param_output, work_directory_);

    }

  }

  return true;

}
