// kgl_b200_flatten.h -- host side of the drop-in: flattens KGL_Gene's variant database into the arrays the C ABI
// (include/kgl_b200.h) takes. Compiled against the KGL_Gene headers (C++23); see INTEGRATION.md.
//
// Input:  the diploid population (population -> genome -> contig -> offset -> variants; only NON-reference alleles are
//         stored, kgl_variant_db_population.h:33) and the allele-frequency "genome" (1 genome, 1 contig, SNP and PASS
//         filtered, kga_analysis_inbreed.cpp:67-81).
// Output: one FlatContig per analysed contig: locus table (offset + the INFO AF float per super-population), the 2-bit
//         loci-major genotype matrix and the super-population index of every genome column.
//
// The genotype code of (genome, locus) restates the classification in InbreedingCalculation::generateFrequencies
// (kga_analysis_inbreed_freq.cpp:452-543) for a locus with ONE alt allele in the AF list:
//   no SNP variant at the offset                                   -> 0   (MAJOR_HOMOZYGOUS if q > 0.01, decided on the device)
//   1 variant, analogous to the AF allele                          -> 1   (MAJOR_HETEROZYGOUS, :464-472)
//   2 variants, first analogous, front->homozygous(back)           -> 2   (MINOR_HOMOZYGOUS, :476-479)
//   2 variants, both analogous, same phase (unphased populations)  -> 2 with FlatContig::unphased set: the reference
//                                                                     classifies these MINOR_HETEROZYGOUS
//   anything else (first variant not in the AF list, second not found, more than 2 variants) -> 3 (dropped)
// Offsets whose AF entry lists several distinct alt alleles keep their row of the locus table (so that locus selection, the
// spacing chain and the window counts see them, kga_analysis_inbreed_locus.cpp:21-72) with "no value" in the frequency table,
// and are described by the side structures FlatContig::multi_* (include/kgl_b200.h, kgl_b200_upload_multi_allelic): per-allele
// frequencies and one byte per genome naming the one or two alleles it carries. The matrix holds 0 / 3 there.
#ifndef KGL_B200_FLATTEN_H
#define KGL_B200_FLATTEN_H

#include "kgl_variant_db_population.h"

#include <cstdint>
#include <functional>
#include <memory>
#include <optional>
#include <string>
#include <vector>

namespace kellerberrin::genome::b200 {

constexpr size_t kSuperPopCount = 6;                       // AFR, AMR, EAS, EUR, SAS, ALL (kgl_variant_db_freq.h:55-60)
extern const char* const kSuperPopCodes[kSuperPopCount];
// Index of a PED super-population code, or nullopt (the reference skips such genomes with an error, diploid.cpp:136).
std::optional<uint8_t> superPopIndex(const std::string& code);

struct FlatContig {
  ContigId_t contig_id;
  std::vector<GenomeId_t> genome_ids;                      // column order (std::map order of the population = reference row order)
  std::vector<uint8_t> superpop;                           // [n_genomes]
  std::vector<uint32_t> offsets;                           // [n_loci], strictly increasing
  std::vector<float> af;                                   // [kSuperPopCount][n_loci], NaN = no value for that super-population
  std::vector<uint8_t> packed;                             // [n_loci][row_bytes]
  uint64_t row_bytes{0};
  bool unphased{false};
  // multi-allelic loci: rows of the locus table, per allele slot frequencies [kSuperPopCount][n_multi][3] (NaN = none), side
  // cells [n_multi][n_genomes] (0 hom-ref; low nibble first variant's slot + 1, 4 = not in the list; high nibble the second
  // variant's, 0 = none; 0xFF = more than two variants)
  std::vector<uint32_t> multi_rows;
  std::vector<float> multi_af;
  std::vector<uint8_t> multi_cells;
  // the variants behind the rows (for consumers that report per variant, e.g. CalcFWS' HGVS-keyed map): the allele of an ordinary
  // row, [n_multi][3] alleles of a multi-allelic one (nullptr = slot not used)
  std::vector<std::shared_ptr<const Variant>> locus_variant;
  std::vector<std::shared_ptr<const Variant>> multi_variant;
  size_t too_many_alleles_skipped{0};                      // offsets with more than three distinct alt alleles (not a SNP locus)
  size_t non_snp_entries{0};                               // flattenSelf: variant entries that are not SNPs (not in the matrix)
  size_t mixed_phase_cells{0};                             // cells whose phase pattern contradicts `unphased` (coded 3)

  [[nodiscard]] uint64_t nGenomes() const { return genome_ids.size(); }
  [[nodiscard]] uint64_t nLoci() const { return offsets.size(); }
  [[nodiscard]] uint64_t nMulti() const { return multi_rows.size(); }
};

// genome id -> PED super-population code; nullopt = no PED record (genome is left out, as the reference does).
using SuperPopLookup = std::function<std::optional<std::string>(const GenomeId_t&)>;

class PopulationFlattener {

public:

  // af_population: 1 genome, 1 contig (checked; kga_analysis_inbreed_diploid.cpp:26,36). threads = 0: hardware_concurrency.
  // unphased_population: the diploid data is DataStructureEnum::DiploidUnphased (all variants share one phase, Pf7).
  static std::optional<FlatContig> flatten(const PopulationDB& diploid_population,
                                           const PopulationDB& af_population,
                                           const SuperPopLookup& super_population,
                                           bool unphased_population,
                                           size_t threads = 0);

  // A population that is its own locus list (kga_PfEMP: CalcFWS, HeteroHomoZygous): one row per offset at which a genome of
  // the contig carries a SNP, alleles = the distinct SNPs seen there, every frequency column = allele_frequency(allele)
  // (nullopt -> no value). Every genome that has the contig is a column; super-population 0 for all.
  static std::optional<FlatContig> flattenSelf(const PopulationDB& population, const ContigId_t& contig_id,
                                               const std::function<std::optional<double>(const Variant&)>& allele_frequency,
                                               bool unphased_population, size_t threads = 0);

};

}  // namespace kellerberrin::genome::b200

#endif  // KGL_B200_FLATTEN_H
