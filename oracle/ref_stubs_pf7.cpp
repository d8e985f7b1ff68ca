// ref_stubs_pf7.cpp -- TEST INFRASTRUCTURE (oracle/_ref/kgl_ref_harness only). The plugin harness links the reference's own
// kgl_pf7_fws_parser.cpp instead (it builds the Pf7 resources in memory to drive HeteroHomoZygous end to end).
#include "kgl_pf7_fws_parser.h"

namespace kgl = kellerberrin::genome;

// CalcFWS::writeGenomeResults (kga_PfEMP/kga_analysis_PfEMP_FWS.cpp:147) looks the published FWS of a sample up in the Pf7
// metadata resource, whose parser needs the Boost-based file IO. The harness only calls CalcFWS::calcFwsStatistics; the
// writer is never reached.
double kgl::Pf7FwsResource::getFWS(const GenomeId_t&) const { return 0.0; }

// HeteroHomoZygous::location_summary (kga_PfEMP/kga_analysis_PfEMP_heterozygous.cpp) walks the Pf7 sample metadata (city /
// country radii, FWS thresholds), whose parsers need the Boost-based file IO. The harness calls only the static per-offset
// rule HeteroHomoZygous::updateVariantAnalysisType (:61-105); never reached.
std::vector<kgl::GenomeId_t> kgl::Pf7FwsResource::filterFWS(FwsFilterType, double, const std::vector<GenomeId_t>& sample_vector) const { return sample_vector; }
