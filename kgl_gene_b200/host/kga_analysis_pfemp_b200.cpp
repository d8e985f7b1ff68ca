// kga_analysis_pfemp_b200.cpp -- see kga_analysis_pfemp_b200.h.
#include "kga_analysis_pfemp_b200.h"

#include "kgl_b200_flatten.h"
#include "kgl_variant_factory_vcf_evidence_analysis.h"
#include "kel_exec_env.h"

#include "kgl_b200.h"

#include <cstdlib>
#include <fstream>
#include <set>

namespace kga = kellerberrin::genome::analysis;
namespace kgl = kellerberrin::genome;
namespace b200 = kellerberrin::genome::b200;
using kellerberrin::ExecEnv;

kga::HeteroHomoB200::~HeteroHomoB200() {
  if (context_ != nullptr) kgl_b200_destroy(context_);
}

bool kga::HeteroHomoB200::ensureContext() {
  if (context_ != nullptr) return true;
  int device = 0;
  if (const char* env = std::getenv("KGL_B200_DEVICE")) device = std::atoi(env);
  if (kgl_b200_create(device, &context_) != KGL_B200_OK) {
    ExecEnv::log().error("HeteroHomoB200; cannot create a device context on GPU {}: {}", device, kgl_b200_last_error(nullptr));
    context_ = nullptr;
    return false;
  }
  return true;
}

bool kga::HeteroHomoB200::analyzeVariantPopulation(const std::shared_ptr<const PopulationDB>& gene_population_ptr,
                                                   const std::shared_ptr<const Pf7FwsResource>& Pf7_fws_ptr,
                                                   const std::shared_ptr<const Pf7SampleResource>& Pf7_sample_ptr) {

  if (not ensureContext()) return false;

  // The records of every genome that has a sample record (:20-35); contigs are added below.
  std::set<ContigId_t> contigs;
  for (auto const& [genome_id, genome_ptr] : gene_population_ptr->getMap()) {

    auto record_iter = Pf7_sample_ptr->getMap().find(genome_id);
    if (record_iter == Pf7_sample_ptr->getMap().end()) {

      ExecEnv::log().error("HeteroHomoB200::analyzeVariantPopulation; Unexpected, could not find sample record for genome:{}", genome_id);
      continue;

    }
    auto const& [sample_id, sample_record] = *record_iter;
    double FWS_statistic = Pf7_fws_ptr->getFWS(genome_id);
    auto [genome_iter, result] = variant_analysis_map_.try_emplace(genome_id, genome_id, sample_record, FWS_statistic);
    for (auto const& [contig_id, contig_ptr] : genome_ptr->getMap()) {

      genome_iter->second.getMap().try_emplace(contig_id);
      contigs.insert(contig_id);

    }

  }

  // P7FrequencyFilter's field (kgl_variant_filter_Pf7.cpp:27-48): the allele's own INFO AF.
  auto allele_frequency = [](const Variant& variant) -> std::optional<double> {
    auto info_opt = InfoEvidenceAnalysis::getTypedInfoData<std::vector<double>>(variant, "AF");
    if (not info_opt) return std::nullopt;
    const size_t alt_index = variant.evidence().altVariantIndex();
    if (info_opt.value().size() <= alt_index) return std::nullopt;
    return info_opt.value()[alt_index];
  };

  const bool unphased = gene_population_ptr->dataSource() == DataSourceEnum::Falciparum;
  for (auto const& contig_id : contigs) {

    auto flat_opt = b200::PopulationFlattener::flattenSelf(*gene_population_ptr, contig_id, allele_frequency, unphased);
    if (not flat_opt) return false;
    const b200::FlatContig& flat = flat_opt.value();
    if (flat.nGenomes() == 0 or flat.nLoci() == 0) continue;
    auto check = [this](int rc, const char* what) {
      if (rc != KGL_B200_OK) ExecEnv::log().error("HeteroHomoB200; {} failed [{}]: {}", what, rc, kgl_b200_last_error(context_));
      return rc == KGL_B200_OK;
    };
    if (not check(kgl_b200_upload_genotypes(context_, flat.nGenomes(), flat.nLoci(), flat.row_bytes, flat.packed.data()), "upload_genotypes")) return false;
    if (not check(kgl_b200_upload_loci(context_, flat.nLoci(), b200::kSuperPopCount, flat.af.data(), flat.offsets.data()), "upload_loci")) return false;
    if (not check(kgl_b200_set_genome_superpop(context_, flat.nGenomes(), flat.superpop.data()), "set_genome_superpop")) return false;
    if (not check(kgl_b200_upload_multi_allelic(context_, flat.nMulti(), flat.multi_rows.data(), flat.multi_af.data(), flat.multi_cells.data()),
                  "upload_multi_allelic")) return false;
    std::vector<uint64_t> records(flat.nGenomes() * 7);
    if (not check(kgl_b200_run_hetero_homo(context_, 1, records.data()), "run_hetero_homo")) return false;
    for (size_t g = 0; g < flat.nGenomes(); ++g) {

      auto genome_iter = variant_analysis_map_.find(flat.genome_ids[g]);
      if (genome_iter == variant_analysis_map_.end()) continue;
      VariantAnalysisType& contig_count = genome_iter->second.getMap()[contig_id];
      const uint64_t* r = records.data() + g * 7;
      contig_count.total_variants_ += r[0];
      contig_count.snp_count_ += r[1];
      contig_count.indel_count_ += r[2];
      contig_count.homozygous_minor_alleles_ += r[3];
      contig_count.heterozygous_minor_alleles_ += r[4];
      contig_count.heterozygous_reference_minor_alleles_ += r[5];
      contig_count.homozygous_reference_alleles_ += r[6];

    }

  }

  return not variant_analysis_map_.empty();

}

kga::VariantAnalysisType kga::HeteroHomoB200::aggregateResults(const std::vector<GenomeId_t>& sample_vector) const {

  VariantAnalysisType analysis_summary;
  std::set<GenomeId_t> sample_set(sample_vector.begin(), sample_vector.end());
  for (auto const& genome_id : sample_set) {

    auto found = variant_analysis_map_.find(genome_id);
    if (found == variant_analysis_map_.end()) continue;
    for (auto const& [contig_id, het_hom_record] : found->second.getConstMap()) {

      analysis_summary.total_variants_ += het_hom_record.total_variants_;
      analysis_summary.heterozygous_reference_minor_alleles_ += het_hom_record.heterozygous_reference_minor_alleles_;
      analysis_summary.homozygous_minor_alleles_ += het_hom_record.homozygous_minor_alleles_;
      analysis_summary.heterozygous_minor_alleles_ += het_hom_record.heterozygous_minor_alleles_;
      analysis_summary.snp_count_ += het_hom_record.snp_count_;
      analysis_summary.indel_count_ += het_hom_record.indel_count_;
      analysis_summary.homozygous_reference_alleles_ += het_hom_record.homozygous_reference_alleles_;

    }

  }
  return analysis_summary;

}

kga::LocationSummaryMap kga::HeteroHomoB200::location_summary(const std::shared_ptr<const Pf7SampleResource>& Pf7_sample_ptr,
                                                              const LocationSamplesMap& location_samples, double radius_km,
                                                              const std::shared_ptr<const Pf7FwsResource>& Pf7_fws_ptr) const {

  std::set<GenomeId_t> pass_genomes;
  for (auto const& [genome_id, sample_record] : Pf7_sample_ptr->getMap()) if (sample_record.pass()) pass_genomes.insert(genome_id);

  LocationSummaryMap summary_map;
  for (auto const& [location, record] : location_samples) {

    std::vector<GenomeId_t> radii_passed;
    for (auto const& sample : record.samples) if (pass_genomes.contains(sample)) radii_passed.push_back(sample);
    auto aggregated = aggregateResults(record.samples);

    double hom_het_ratio{0.0};
    size_t total_heterozygous = aggregated.heterozygous_reference_minor_alleles_ + aggregated.heterozygous_minor_alleles_;
    if (total_heterozygous > 0) hom_het_ratio = static_cast<double>(aggregated.homozygous_minor_alleles_) / static_cast<double>(total_heterozygous);
    double variant_rate{0.0};
    if (not record.samples.empty()) variant_rate = static_cast<double>(aggregated.total_variants_) / static_cast<double>(record.samples.size());
    double monoclonal{0.0};
    if (not radii_passed.empty()) {

      auto mono_samples = Pf7_fws_ptr->filterFWS(FwsFilterType::GREATER_EQUAL, Pf7FwsResource::MONOCLONAL_FWS_THRESHOLD, radii_passed);
      monoclonal = static_cast<double>(mono_samples.size()) / static_cast<double>(radii_passed.size());

    }

    LocationSummary s;
    s.location_ = location; s.location_type_ = record.location_type; s.city_ = record.city; s.country_ = record.country;
    s.region_ = record.region; s.radius_km_ = radius_km; s.radii_samples_ = record.samples.size(); s.radii_samples_OK_ = radii_passed.size();
    s.monoclonal_Fst_ = monoclonal; s.hom_het_ratio_ = hom_het_ratio; s.total_variants_ = aggregated.total_variants_;
    s.variant_rate_ = variant_rate; s.homozygous_reference_alleles_ = aggregated.homozygous_reference_alleles_;
    s.heterozygous_reference_minor_alleles_ = aggregated.heterozygous_reference_minor_alleles_;
    s.homozygous_minor_alleles_ = aggregated.homozygous_minor_alleles_; s.heterozygous_minor_alleles_ = aggregated.heterozygous_minor_alleles_;
    s.snp_count_ = aggregated.snp_count_; s.indel_count_ = aggregated.indel_count_;
    summary_map[location] = s;

  }
  return summary_map;

}

void kga::HeteroHomoB200::UpdateSampleLocation(const LocationSummaryMap& location_summary, const LocationSamplesMap& location_samples,
                                               const std::shared_ptr<const Pf7SampleResource>& Pf7_sample_ptr) {

  // Index spaces of the C ABI: genomes in map order, locations in map order.
  std::map<GenomeId_t, uint32_t> genome_index;
  std::vector<uint64_t> records;
  std::vector<uint8_t> qc_pass;
  for (auto const& [genome_id, analysis_obj] : variant_analysis_map_) {

    genome_index[genome_id] = static_cast<uint32_t>(genome_index.size());
    const VariantAnalysisType a = aggregateResults({genome_id});
    records.insert(records.end(), {a.total_variants_, a.snp_count_, a.indel_count_, a.homozygous_minor_alleles_, a.heterozygous_minor_alleles_,
                                   a.heterozygous_reference_minor_alleles_, a.homozygous_reference_alleles_});

  }
  // radii_samples_OK_ counts QC-pass samples of the location whether or not they are in the analysed population (:281-299): the
  // member lists below hold the analysed genomes only, so samples outside it are appended as zero records.
  std::map<std::string, uint32_t> location_index;
  std::vector<uint64_t> location_begin{0};
  std::vector<uint32_t> location_members;
  auto index_of = [&](const GenomeId_t& sample) -> uint32_t {
    auto found = genome_index.find(sample);
    if (found != genome_index.end()) return found->second;
    const uint32_t idx = static_cast<uint32_t>(genome_index.size());
    genome_index[sample] = idx;
    records.insert(records.end(), 7, 0ull);
    return idx;
  };
  for (auto const& [location, record] : location_samples) {

    if (not location_summary.contains(location)) continue;
    location_index[location] = static_cast<uint32_t>(location_index.size());
    std::set<GenomeId_t> sample_set(record.samples.begin(), record.samples.end());
    for (auto const& sample : sample_set) location_members.push_back(index_of(sample));
    location_begin.push_back(location_members.size());

  }
  const size_t n = genome_index.size();
  qc_pass.assign(n, 0);
  for (auto const& [genome_id, idx] : genome_index) {
    auto rec = Pf7_sample_ptr->getMap().find(genome_id);
    qc_pass[idx] = (rec != Pf7_sample_ptr->getMap().end() and rec->second.pass()) ? 1 : 0;
  }
  const uint32_t none = static_cast<uint32_t>(location_index.size());
  std::vector<uint32_t> city(n, none), country(n, none);
  for (auto const& [genome_id, analysis_obj] : variant_analysis_map_) {

    const uint32_t g = genome_index[genome_id];
    if (auto f = location_index.find(analysis_obj.getCity()); f != location_index.end()) city[g] = f->second;
    else ExecEnv::log().error("HeteroHomoB200::UpdateSampleLocation; Unable to find the location record for sample/genome city: {}", analysis_obj.getCity());
    if (auto f = location_index.find(analysis_obj.getCountry()); f != location_index.end()) country[g] = f->second;

  }
  std::vector<double> fis(n, 0.0);
  const int rc = kgl_b200_location_fis(n, records.data(), none, location_begin.data(), location_members.data(), city.data(), country.data(),
                                       qc_pass.data(), MINIMUM_LOCATION_SAMPLES_, fis.data());
  if (rc != KGL_B200_OK) { ExecEnv::log().error("HeteroHomoB200::UpdateSampleLocation; kgl_b200_location_fis failed [{}]", rc); return; }
  for (auto& [genome_id, analysis_obj] : variant_analysis_map_) analysis_obj.setFIS(fis[genome_index[genome_id]]);

}

void kga::HeteroHomoB200::write_variant_results(const std::string& file_name, const LocationSummaryMap& location_summary) const {

  std::ofstream analysis_file(file_name);
  if (not analysis_file.good()) {

    ExecEnv::log().error("HeteroHomoB200::write_variant_results; Unable to open results file: {}", file_name);
    return;

  }
  if (variant_analysis_map_.empty()) return;

  const size_t contig_count = variant_analysis_map_.begin()->second.getConstMap().size();
  analysis_file << "Genome" << CSV_DELIMITER_ << "FWS" << CSV_DELIMITER_ << "FIS (inbreed)" << CSV_DELIMITER_ << "City" << CSV_DELIMITER_
                << "Country" << CSV_DELIMITER_ << "Region" << CSV_DELIMITER_ << "Study" << CSV_DELIMITER_ << "Year" << CSV_DELIMITER_ << "Hom/Het";
  for (size_t i = 0; i <= contig_count; ++i) {

    analysis_file << CSV_DELIMITER_ << "Contig" << CSV_DELIMITER_ << "Variant Count" << CSV_DELIMITER_ << "Hom Ref (A;A)" << CSV_DELIMITER_
                  << "Het Ref Minor (A;a)" << CSV_DELIMITER_ << "Hom Minor (a;a)" << CSV_DELIMITER_ << "Het Diff Minor (a;b)" << CSV_DELIMITER_
                  << "SNP" << CSV_DELIMITER_ << "Indel";

  }
  analysis_file << '\n';

  auto write_counts = [&analysis_file](const std::string& label, const VariantAnalysisType& c) {
    analysis_file << CSV_DELIMITER_ << label << CSV_DELIMITER_ << c.total_variants_ << CSV_DELIMITER_ << c.homozygous_reference_alleles_
                  << CSV_DELIMITER_ << c.heterozygous_reference_minor_alleles_ << CSV_DELIMITER_ << c.homozygous_minor_alleles_
                  << CSV_DELIMITER_ << c.heterozygous_minor_alleles_ << CSV_DELIMITER_ << c.snp_count_ << CSV_DELIMITER_ << c.indel_count_;
  };
  for (auto const& [genome_id, contig_map] : variant_analysis_map_) {

    auto aggregated = aggregateResults({genome_id});
    double hom_het_ratio{0.0};
    size_t total_heterozygous = aggregated.heterozygous_reference_minor_alleles_ + aggregated.heterozygous_minor_alleles_;
    if (total_heterozygous > 0) hom_het_ratio = static_cast<double>(aggregated.homozygous_minor_alleles_) / static_cast<double>(total_heterozygous);
    std::string region;
    if (auto iter = location_summary.find(contig_map.getCity()); iter != location_summary.end()) region = iter->second.region_;

    analysis_file << genome_id << CSV_DELIMITER_ << contig_map.getFWS() << CSV_DELIMITER_ << contig_map.getFIS() << CSV_DELIMITER_
                  << contig_map.getCity() << CSV_DELIMITER_ << contig_map.getCountry() << CSV_DELIMITER_ << region << CSV_DELIMITER_
                  << contig_map.getStudy() << CSV_DELIMITER_ << contig_map.getYear() << CSV_DELIMITER_ << hom_het_ratio;
    write_counts("Combined", aggregated);
    for (auto const& [contig_id, variant_counts] : contig_map.getConstMap()) write_counts(contig_id, variant_counts);
    analysis_file << '\n';

  }

}


// ---- CalcFwsB200 -----------------------------------------------------------------------------------------------------------

kga::CalcFwsB200::~CalcFwsB200() {
  if (context_ != nullptr) kgl_b200_destroy(context_);
}

std::pair<double, double> kga::CalcFwsB200::binRange(size_t bin) {
  // 5 % steps up to one half, then everything above (CalcFWS::getFrequency, kga_analysis_PfEMP_FWS.cpp:104-145)
  static constexpr double kEdges[FWS_FREQUENCY_ARRAY_SIZE + 1] = {0.0, 0.05, 0.10, 0.15, 0.20, 0.25, 0.30, 0.35, 0.40, 0.45, 0.5, 1.0};
  return bin < FWS_FREQUENCY_ARRAY_SIZE ? std::pair<double, double>{kEdges[bin], kEdges[bin + 1]} : std::pair<double, double>{0.0, 0.0};
}

bool kga::CalcFwsB200::calcFwsStatistics(const std::shared_ptr<const PopulationDB>& population) {

  if (context_ == nullptr) {
    int device = 0;
    if (const char* env = std::getenv("KGL_B200_DEVICE")) device = std::atoi(env);
    if (kgl_b200_create(device, &context_) != KGL_B200_OK) {
      ExecEnv::log().error("CalcFwsB200; cannot create a device context on GPU {}: {}", device, kgl_b200_last_error(nullptr));
      context_ = nullptr;
      return false;
    }
  }
  auto check = [this](int rc, const char* what) {
    if (rc != KGL_B200_OK) ExecEnv::log().error("CalcFwsB200; {} failed [{}]: {}", what, rc, kgl_b200_last_error(context_));
    return rc == KGL_B200_OK;
  };
  auto allele_frequency = [](const Variant& variant) -> std::optional<double> {        // P7FrequencyFilter's value (kgl_variant_filter_Pf7.cpp:22-48)
    auto info_opt = InfoEvidenceAnalysis::getTypedInfoData<std::vector<double>>(variant, "AF");
    if (not info_opt) return std::nullopt;
    const size_t alt_index = variant.evidence().altVariantIndex();
    if (info_opt.value().size() != variant.evidence().altVariantCount() or info_opt.value().size() <= alt_index) return std::nullopt;
    return info_opt.value()[alt_index];
  };

  std::set<ContigId_t> contigs;
  for (auto const& [genome_id, genome_ptr] : population->getMap())
    for (auto const& [contig_id, contig_ptr] : genome_ptr->getMap()) contigs.insert(contig_id);
  const size_t population_genomes = population->getMap().size();
  const bool unphased = population->dataSource() == DataSourceEnum::Falciparum;

  // Everything is collected first: a population the matrix cannot hold leaves the maps untouched.
  GenomeFWSMap genome_update;
  VariantFWSMap variant_update;
  for (auto const& [genome_id, genome_ptr] : population->getMap()) genome_update.try_emplace(genome_id, FwsFrequencyArray());
  double lower[FWS_FREQUENCY_ARRAY_SIZE], upper[FWS_FREQUENCY_ARRAY_SIZE];
  for (size_t b = 0; b < FWS_FREQUENCY_ARRAY_SIZE; ++b) std::tie(lower[b], upper[b]) = binRange(b);

  for (auto const& contig_id : contigs) {

    auto flat_opt = b200::PopulationFlattener::flattenSelf(*population, contig_id, allele_frequency, unphased);
    if (not flat_opt) return false;
    const b200::FlatContig& flat = flat_opt.value();
    if (flat.non_snp_entries > 0 or flat.too_many_alleles_skipped > 0) {
      ExecEnv::log().error("CalcFwsB200::calcFwsStatistics; contig: {} holds {} variant entries that are not SNPs and {} offsets with more than three alleles: filter the population first",
                           contig_id, flat.non_snp_entries, flat.too_many_alleles_skipped);
      return false;
    }
    const size_t N = flat.nGenomes(), L = flat.nLoci(), M = flat.nMulti();
    if (N == 0 or L == 0) continue;
    for (size_t i = 0; i < flat.multi_cells.size(); ++i)
      if (flat.multi_cells[i] == 0xFF) {
        ExecEnv::log().error("CalcFwsB200::calcFwsStatistics; contig: {}, a genome carries more than two variants at one offset", contig_id);
        return false;
      }
    if (not check(kgl_b200_upload_genotypes(context_, N, L, flat.row_bytes, flat.packed.data()), "upload_genotypes")) return false;
    if (not check(kgl_b200_upload_loci(context_, L, b200::kSuperPopCount, flat.af.data(), flat.offsets.data()), "upload_loci")) return false;
    if (not check(kgl_b200_set_genome_superpop(context_, N, flat.superpop.data()), "set_genome_superpop")) return false;
    if (not check(kgl_b200_upload_multi_allelic(context_, M, flat.multi_rows.data(), flat.multi_af.data(), flat.multi_cells.data()), "upload_multi_allelic")) return false;

    // updateGenomeFWSMap (:72-101) for the eleven bins at once
    std::vector<uint64_t> bins(FWS_FREQUENCY_ARRAY_SIZE * N * 4), bin_variants(FWS_FREQUENCY_ARRAY_SIZE);
    if (not check(kgl_b200_run_binned_genome_counts(context_, 0, FWS_FREQUENCY_ARRAY_SIZE, lower, upper, 1, bins.data(), bin_variants.data()), "run_binned_genome_counts")) return false;
    std::set<GenomeId_t> in_matrix(flat.genome_ids.begin(), flat.genome_ids.end());
    for (size_t b = 0; b < FWS_FREQUENCY_ARRAY_SIZE; ++b) {
      for (size_t g = 0; g < N; ++g) {
        const uint64_t* c = bins.data() + (b * N + g) * 4;
        AlleleSummmary& record = genome_update[flat.genome_ids[g]][b];
        record.referenceHomozygous_ += c[0] + c[3];          // a cell with some other allele has no copy of the column's variant
        record.minorHeterozygous_ += c[1];
        record.minorHomozygous_ += c[2];
      }
      for (auto& [genome_id, freq_array] : genome_update)   // a genome without this contig: every column of the bin is hom-ref
        if (not in_matrix.contains(genome_id)) freq_array[b].referenceHomozygous_ += bin_variants[b];
    }

    // updateVariantFWSMap (:41-70): one record per variant of the population
    std::vector<uint32_t> locus_counts(L * 4), multi_counts(M * 3 * 3);
    if (not check(kgl_b200_run_allele_count(context_, locus_counts.data(), nullptr), "run_allele_count")) return false;
    if (M > 0 and not check(kgl_b200_run_multi_allele_count(context_, multi_counts.data()), "run_multi_allele_count")) return false;
    std::vector<uint8_t> is_multi(L, 0);
    for (uint32_t row : flat.multi_rows) is_multi[row] = 1;
    const size_t absent = population_genomes - N;
    auto add_variant = [&variant_update, absent](const std::shared_ptr<const Variant>& variant, uint64_t none, uint64_t one, uint64_t two) {
      if (variant == nullptr or one + two == 0) return;      // only variants some genome carries have a column (kgl_variant_db_variant.cpp:14-30)
      AlleleSummmary& summary = variant_update[variant->HGVS()];
      summary.referenceHomozygous_ += none + absent;
      summary.minorHeterozygous_ += one;
      summary.minorHomozygous_ += two;
    };
    for (size_t l = 0; l < L; ++l)
      if (not is_multi[l]) add_variant(flat.locus_variant[l], uint64_t(locus_counts[l * 4]) + locus_counts[l * 4 + 3], locus_counts[l * 4 + 1], locus_counts[l * 4 + 2]);
    for (size_t m = 0; m < M; ++m)
      for (size_t a = 0; a < 3; ++a)
        add_variant(flat.multi_variant[m * 3 + a], multi_counts[(m * 3 + a) * 3], multi_counts[(m * 3 + a) * 3 + 1], multi_counts[(m * 3 + a) * 3 + 2]);

  }

  for (auto const& [genome_id, freq_array] : genome_update) {
    auto& target = genome_fws_map_[genome_id];
    for (size_t b = 0; b < FWS_FREQUENCY_ARRAY_SIZE; ++b) target[b] += freq_array[b];
  }
  for (auto const& [hgvs, summary] : variant_update) variant_fws_map_[hgvs] += summary;
  return true;

}

namespace {

// "hom / het" of the two CSV files: 0 when there is no heterozygous entry
double ratioOrZero(size_t numerator, size_t denominator) {
  return denominator > 0 ? static_cast<double>(numerator) / static_cast<double>(denominator) : 0.0;
}

}  // namespace

void kga::CalcFwsB200::writeGenomeResults(const std::shared_ptr<const Pf7FwsResource>& Pf7_fws_ptr, const std::string& file_name) const {

  std::ofstream out(file_name);
  if (not out.good()) {
    ExecEnv::log().error("CalcFwsB200::writeGenomeResults; Unable to open results file: {}", file_name);
    return;
  }
  static const char* const kBinColumns[] = {"LowerFreq", "UpperFreq", "Hom/Het", "Minor Hom/Het", "Variant Count", "Hom Ref (A;A)",
                                            "Het Ref Minor (A;a)", "Hom Minor (a;a)"};
  out << "Genome,FWS";
  for (size_t b = 0; b < FWS_FREQUENCY_ARRAY_SIZE; ++b) for (const char* column : kBinColumns) out << ',' << column;
  out << '\n';
  for (auto const& [genome_id, freq_array] : genome_fws_map_) {
    out << genome_id << ',' << Pf7_fws_ptr->getFWS(genome_id);
    for (size_t b = 0; b < FWS_FREQUENCY_ARRAY_SIZE; ++b) {
      const AlleleSummmary& s = freq_array[b];
      auto const [lower, upper] = binRange(b);
      out << ',' << lower << ',' << upper
          << ',' << ratioOrZero(s.minorHomozygous_ + s.referenceHomozygous_, s.minorHeterozygous_)
          << ',' << ratioOrZero(s.minorHomozygous_, s.minorHeterozygous_)
          << ',' << (s.minorHeterozygous_ + s.minorHomozygous_)
          << ',' << s.referenceHomozygous_ << ',' << s.minorHeterozygous_ << ',' << s.minorHomozygous_;
    }
    out << '\n';
  }

}

void kga::CalcFwsB200::writeVariantResults(const std::string& file_name) const {

  std::ofstream out(file_name);
  if (not out.good()) {
    ExecEnv::log().error("CalcFwsB200::writeVariantResults; Unable to open results file: {}", file_name);
    return;
  }
  out << "Variant,Hom/Het,Minor Hom/Het,Genome Count,Hom Ref (A;A),Het Ref Minor (A;a),Hom Minor (a;a)\n";
  for (auto const& [hgvs, s] : variant_fws_map_)
    out << hgvs << ',' << ratioOrZero(s.minorHomozygous_ + s.referenceHomozygous_, s.minorHeterozygous_)
        << ',' << ratioOrZero(s.minorHomozygous_, s.minorHeterozygous_)
        << ',' << (s.minorHeterozygous_ + s.minorHomozygous_)
        << ',' << s.referenceHomozygous_ << ',' << s.minorHeterozygous_ << ',' << s.minorHomozygous_ << '\n';

}
