// kga_analysis_inbreed_b200.h -- the B200 drop-in for KGL_Gene's INBREED analysis.
//
// A VirtualAnalysis (kgl_app/kgl_package_analysis_virtual.h:20-55) with the same four stages, the same parameter blocks
// (InbreedArguments::extractParameters, kga_analysis_inbreed_args.cpp:12), the same PED resource and the same CSV output
// (InbreedingOutput::writePedResults, kga_analysis_inbreed_output.cpp:183) as kga::InbreedAnalysis
// (kga_analytic/kga_inbreed/kga_analysis_inbreed.{h,cpp}). What changes is iterationAnalysis: instead of one CPU task per
// genome (InbreedingAnalysis::processResults, kga_analysis_inbreed_diploid.cpp:98-160) the population is flattened once
// (kgl_b200_flatten.h) and every window is one call sequence on the C ABI (include/kgl_b200.h). Register it next to the
// reference analysis in kga_analytic/kga_analysis_factory.cpp:31-43. Synthetic parameter blocks
// (AnalysisType = TRUE) are delegated to the reference's ExecuteInbreedingAnalysis unchanged.
#ifndef KGA_ANALYSIS_INBREED_B200_H
#define KGA_ANALYSIS_INBREED_B200_H

#include "kgl_package_analysis_virtual.h"
#include "kgl_hsgenealogy_parser.h"
#include "kgl_variant_db_population.h"
#include "kga_analysis_inbreed_args.h"
#include "kga_analysis_inbreed_output.h"

#include "kgl_b200_flatten.h"

struct kgl_b200_ctx;

namespace kellerberrin::genome::analysis {

class InbreedB200Analysis : public VirtualAnalysis {

public:

  InbreedB200Analysis() = default;
  ~InbreedB200Analysis() override;

  // The ident must match the ident used in the package XML.
  inline static const std::string IDENT{"INBREED_B200"};
  [[nodiscard]] std::string ident() const override { return IDENT; }
  [[nodiscard]] static std::unique_ptr<VirtualAnalysis> factory() { return std::make_unique<InbreedB200Analysis>(); }

  [[nodiscard]] bool initializeAnalysis(const std::string& work_directory,
                                        const ActiveParameterList& named_parameters,
                                        const std::shared_ptr<const AnalysisResources>& resource_ptr) override;
  [[nodiscard]] bool fileReadAnalysis(std::shared_ptr<const DataDB> data_object_ptr) override;
  [[nodiscard]] bool iterationAnalysis() override;
  [[nodiscard]] bool finalizeAnalysis() override;

  // InbreedingAnalysis::populationInbreeding (kga_analysis_inbreed_diploid.cpp:18-79) on the device: uploadPopulation (one
  // flatten + one upload) followed by windowLoop. iterationAnalysis calls the two halves itself, so that every parameter block
  // of an iteration works on the same resident population. Public so that a host program that already holds the populations
  // and the PED data can call them directly.
  [[nodiscard]] bool uploadPopulation(const std::shared_ptr<const PopulationDB>& unphased_ptr,
                                      const PopulationDB& diploid_population,
                                      const HsGenomeGenealogyData& ped_data,
                                      bool unphased_diploid);
  [[nodiscard]] bool windowLoop(const std::shared_ptr<const PopulationDB>& unphased_ptr, InbreedParamOutput& param_output);
  [[nodiscard]] bool populationInbreeding(const std::shared_ptr<const PopulationDB>& unphased_ptr,
                                          const PopulationDB& diploid_population,
                                          const HsGenomeGenealogyData& ped_data,
                                          bool unphased_diploid,
                                          InbreedParamOutput& param_output);

private:

  std::vector<InbreedParamOutput> parameter_output_vector_;
  std::string work_directory_;
  std::shared_ptr<const PopulationDB> diploid_population_;
  std::shared_ptr<const PopulationDB> unphased_population_;
  std::shared_ptr<const HsGenomeGenealogyData> genealogy_data_;
  bool diploid_is_unphased_{false};
  int device_{0};
  kgl_b200_ctx* context_{nullptr};
  std::unique_ptr<b200::FlatContig> flat_;               // the population that is resident on the device

  [[nodiscard]] bool ensureContext();
  [[nodiscard]] bool check(int rc, const char* what);
  [[nodiscard]] bool writeResults();

};

}  // namespace kellerberrin::genome::analysis

#endif  // KGA_ANALYSIS_INBREED_B200_H
