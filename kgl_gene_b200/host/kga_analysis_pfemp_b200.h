// kga_analysis_pfemp_b200.h -- the Pf7 summary statistics of kga_PfEMP on the device: the second consumer of the allele-counting
// kernels behind KGL_Gene's analysis layer.
//
// Mirrors kga::HeteroHomoZygous (kga_analytic/kga_PfEMP/kga_analysis_PfEMP_heterozygous.h:97-131) -- same method names,
// argument meaning, record types (VariantAnalysisType, LocationSummary, LocationSummaryMap are the reference's own) and CSV
// layout -- the way kga::PfEMPAnalysis uses it (kga_analysis_PfEMP.cpp:92-109: analyzeVariantPopulation in fileReadAnalysis,
// location_summary + UpdateSampleLocation + write_variant_results in finalizeAnalysis). What changes underneath: the population
// is flattened once per contig (PopulationFlattener::flattenSelf) and the per-genome records come from one counting pass of
// the C ABI (kgl_b200_run_hetero_homo) instead of three filtered copies of every offset; Wright's F_IS is
// kgl_b200_location_fis. One difference in the interface: the samples of a location are passed in (the reference asks
// Pf7SampleLocation::sampleRadius, whose parser needs Boost and is not part of the hot path).
#ifndef KGA_ANALYSIS_PFEMP_B200_H
#define KGA_ANALYSIS_PFEMP_B200_H

#include "kga_analysis_PfEMP_heterozygous.h"
#include "kgl_pf7_fws_parser.h"
#include "kgl_pf7_sample_parser.h"

#include <map>
#include <memory>
#include <string>
#include <vector>

struct kgl_b200_ctx;

namespace kellerberrin::genome::analysis {

struct LocationSamples {
  LocationType location_type{LocationType::City};
  std::string city, country, region;
  std::vector<GenomeId_t> samples;           // what Pf7SampleLocation::sampleRadius(location, radius) returns
};
using LocationSamplesMap = std::map<std::string, LocationSamples>;

class HeteroHomoB200 {

public:

  HeteroHomoB200() = default;
  ~HeteroHomoB200();
  HeteroHomoB200(const HeteroHomoB200&) = delete;
  HeteroHomoB200& operator=(const HeteroHomoB200&) = delete;

  // HeteroHomoZygous::analyzeVariantPopulation (:15-58). False: no device, or a sample record is missing for every genome.
  [[nodiscard]] bool analyzeVariantPopulation(const std::shared_ptr<const PopulationDB>& gene_population_ptr,
                                              const std::shared_ptr<const Pf7FwsResource>& Pf7_fws_ptr,
                                              const std::shared_ptr<const Pf7SampleResource>& Pf7_sample_ptr);
  // If sample vector is empty the result is empty (as the reference's: it sums over the listed samples, :235-262).
  [[nodiscard]] VariantAnalysisType aggregateResults(const std::vector<GenomeId_t>& sample_vector) const;
  // HeteroHomoZygous::location_summary (:266-358) with the samples of every location supplied by the caller.
  [[nodiscard]] LocationSummaryMap location_summary(const std::shared_ptr<const Pf7SampleResource>& Pf7_sample_ptr,
                                                    const LocationSamplesMap& location_samples, double radius_km,
                                                    const std::shared_ptr<const Pf7FwsResource>& Pf7_fws_ptr) const;
  // HeteroHomoZygous::UpdateSampleLocation (:362-412) through kgl_b200_location_fis.
  void UpdateSampleLocation(const LocationSummaryMap& location_summary, const LocationSamplesMap& location_samples,
                            const std::shared_ptr<const Pf7SampleResource>& Pf7_sample_ptr);
  // HeteroHomoZygous::write_variant_results (:106-231): same columns, same formatting.
  void write_variant_results(const std::string& file_name, const LocationSummaryMap& location_summary) const;

  [[nodiscard]] const VariantAnalysisMap& getMap() const { return variant_analysis_map_; }

private:

  VariantAnalysisMap variant_analysis_map_;
  kgl_b200_ctx* context_{nullptr};
  constexpr static const char CSV_DELIMITER_ = ',';
  constexpr static const size_t MINIMUM_LOCATION_SAMPLES_ = 20;

  [[nodiscard]] bool ensureContext();

};

}  // namespace kellerberrin::genome::analysis

#endif  // KGA_ANALYSIS_PFEMP_B200_H
