// kgl_b200_flatten.cpp -- see kgl_b200_flatten.h.
#include "kgl_b200_flatten.h"

#include "kel_exec_env.h"
#include "kgl_variant_db_freq.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <limits>
#include <map>
#include <thread>

namespace kgl = kellerberrin::genome;
namespace b200 = kellerberrin::genome::b200;
using kellerberrin::ExecEnv;

const char* const b200::kSuperPopCodes[b200::kSuperPopCount] = {
    kgl::FrequencyDatabaseRead::SUPER_POP_AFR_, kgl::FrequencyDatabaseRead::SUPER_POP_AMR_, kgl::FrequencyDatabaseRead::SUPER_POP_EAS_,
    kgl::FrequencyDatabaseRead::SUPER_POP_EUR_, kgl::FrequencyDatabaseRead::SUPER_POP_SAS_, kgl::FrequencyDatabaseRead::SUPER_POP_ALL_};

std::optional<uint8_t> b200::superPopIndex(const std::string& code) {
  for (size_t k = 0; k < kSuperPopCount; ++k)
    if (code == kSuperPopCodes[k]) return static_cast<uint8_t>(k);
  return std::nullopt;
}

namespace {

// Variant::analogous (kgl_variant_db.h:143) is equality of the HGVS string "{contig}:g.{offset}{ref}>{alt}"
// (kgl_variant_db.cpp:287-290); the same test on the fields, without building two strings per call.
bool analogous(const kgl::Variant& a, const kgl::Variant& b) {
  return a.offset() == b.offset() && a.contigId() == b.contigId() &&
         a.reference().getStringView() == b.reference().getStringView() &&
         a.alternate().getStringView() == b.alternate().getStringView();
}

}  // namespace

namespace {

constexpr size_t kSlots = 3;

// The locus table under construction: per row the first (or only) allele, and for rows with several alleles their list.
struct LocusTable {
  std::vector<std::shared_ptr<const kgl::Variant>> locus_allele;
  std::vector<std::array<float, b200::kSuperPopCount>> locus_af;
  std::vector<int64_t> multi_of;                                       // row -> index into the multi-allelic tables, -1
  std::vector<std::vector<std::shared_ptr<const kgl::Variant>>> multi_alleles;
  std::vector<std::array<float, b200::kSuperPopCount * kSlots>> multi_af_rows;
};

// One offset of the frequency source: its distinct alleles in the order of the variant array (AlleleFreqVector's duplicate test,
// freq.cpp:31-42) become one row of the locus table; allele_af(allele, k) = the frequency float of the allele for population k.
template <class AlleleAf>
bool addLocus(b200::FlatContig& flat, LocusTable& table, uint64_t offset, const kgl::OffsetDBArray& variants, const AlleleAf& allele_af) {
  if (variants.empty()) return true;
  std::vector<std::shared_ptr<const kgl::Variant>> alleles;
  for (auto const& v : variants) {
    bool seen = false;
    for (auto const& a : alleles) if (analogous(*v, *a)) { seen = true; break; }
    if (not seen) alleles.push_back(v);
  }
  if (alleles.size() > kSlots) { ++flat.too_many_alleles_skipped; return true; }
  if (offset > std::numeric_limits<uint32_t>::max()) {
    ExecEnv::log().error("PopulationFlattener; offset {} does not fit 32 bits", offset);
    return false;
  }
  const float kNone = std::numeric_limits<float>::quiet_NaN();
  std::array<float, b200::kSuperPopCount> row{};
  if (alleles.size() == 1) {
    for (size_t k = 0; k < b200::kSuperPopCount; ++k) row[k] = allele_af(*alleles.front(), k);
    table.multi_of.push_back(-1);
  } else {
    std::array<float, b200::kSuperPopCount * kSlots> mrow{};
    mrow.fill(kNone);
    for (size_t k = 0; k < b200::kSuperPopCount; ++k)
      for (size_t a = 0; a < alleles.size(); ++a) mrow[k * kSlots + a] = allele_af(*alleles[a], k);
    row.fill(kNone);                                                 // the frequency table is not used at such a row
    table.multi_of.push_back(static_cast<int64_t>(table.multi_alleles.size()));
    flat.multi_rows.push_back(static_cast<uint32_t>(flat.offsets.size()));
    table.multi_alleles.push_back(alleles);
    table.multi_af_rows.push_back(mrow);
  }
  flat.offsets.push_back(static_cast<uint32_t>(offset));
  table.locus_allele.push_back(alleles.front());
  table.locus_af.push_back(row);
  return true;
}

void finishLocusTable(b200::FlatContig& flat, const LocusTable& table) {
  const size_t L = flat.offsets.size(), M = flat.multi_rows.size();
  flat.af.resize(b200::kSuperPopCount * L);
  for (size_t l = 0; l < L; ++l)
    for (size_t k = 0; k < b200::kSuperPopCount; ++k) flat.af[k * L + l] = table.locus_af[l][k];
  flat.multi_af.resize(b200::kSuperPopCount * M * kSlots);
  for (size_t m = 0; m < M; ++m)
    for (size_t k = 0; k < b200::kSuperPopCount; ++k)
      for (size_t a = 0; a < kSlots; ++a) flat.multi_af[(k * M + m) * kSlots + a] = table.multi_af_rows[m][k * kSlots + a];
  flat.locus_variant = table.locus_allele;
  flat.multi_variant.assign(M * kSlots, nullptr);
  for (size_t m = 0; m < M; ++m)
    for (size_t a = 0; a < table.multi_alleles[m].size() && a < kSlots; ++a) flat.multi_variant[m * kSlots + a] = table.multi_alleles[m][a];
  if (flat.too_many_alleles_skipped > 0)
    ExecEnv::log().warn("PopulationFlattener; contig: {}, {} offsets with more than three alt alleles left out of the locus table",
                        flat.contig_id, flat.too_many_alleles_skipped);
}

// Genotype codes and side cells of every genome column. A thread owns whole 64-genome units, so no two threads touch the same byte.
void fillGenotypes(b200::FlatContig& flat, const LocusTable& table, const std::vector<std::shared_ptr<const kgl::ContigDB>>& genome_contig,
                   bool unphased_population, size_t threads) {
  const size_t N = flat.genome_ids.size(), L = flat.offsets.size(), M = flat.multi_rows.size();
  const size_t units = (N + 63) / 64;
  flat.row_bytes = 16 * units;
  flat.packed.assign(L * flat.row_bytes, 0);
  flat.multi_cells.assign(M * N, 0);
  if (N == 0 || L == 0) return;
  if (threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());
  threads = std::min(threads, units);
  std::atomic<size_t> mixed_phase{0};
  auto worker = [&](size_t t) {
    size_t mixed = 0;
    std::vector<const kgl::Variant*> snps;
    for (size_t u = t; u < units; u += threads) {
      for (size_t b = 0; b < 64 && u * 64 + b < N; ++b) {
        const size_t g = u * 64 + b;
        for (auto const& [offset, offset_ptr] : genome_contig[g]->getMap()) {
          if (offset > std::numeric_limits<uint32_t>::max()) break;
          auto it = std::lower_bound(flat.offsets.begin(), flat.offsets.end(), static_cast<uint32_t>(offset));
          if (it == flat.offsets.end() || *it != offset) continue;          // not a locus of the AF list
          const size_t l = static_cast<size_t>(it - flat.offsets.begin());
          snps.clear();
          for (auto const& v : offset_ptr->getVariantArray())
            if (v->isSNP()) snps.push_back(v.get());                        // the genome side is SNP filtered (freq.cpp:436)
          if (snps.empty()) continue;
          unsigned code = 3;
          uint8_t* unit = flat.packed.data() + l * flat.row_bytes + u * 16;
          if (table.multi_of[l] >= 0) {
            // several alt alleles: the side cell names the first and the second variant's allele (generateFrequencies looks the
            // FRONT of the offset array up first, freq.cpp:462); 4 = not in the locus' list; 0xFF = more than two variants, or a
            // same-allele pair whose phases contradict the population's phasing
            auto const& alleles = table.multi_alleles[table.multi_of[l]];
            auto slot_of = [&](const kgl::Variant& v) -> unsigned {
              for (size_t a = 0; a < alleles.size(); ++a) if (analogous(v, *alleles[a])) return static_cast<unsigned>(a) + 1;
              return 4;
            };
            unsigned cell = 0xFF;
            if (snps.size() == 1) cell = slot_of(*snps.front());
            else if (snps.size() == 2) {
              const unsigned s1 = slot_of(*snps.front()), s2 = slot_of(*snps.back());
              cell = s1 | (s2 << 4);
              if (s1 == s2 && s1 != 4) {
                const bool phased_pair = snps.front()->phaseId() != snps.back()->phaseId();
                if (phased_pair == unphased_population) { cell = 0xFF; ++mixed; }
              }
            }
            flat.multi_cells[static_cast<size_t>(table.multi_of[l]) * N + g] = static_cast<uint8_t>(cell);
          } else {
            const kgl::Variant& allele = *table.locus_allele[l];
            if (analogous(*snps.front(), allele)) {
              if (snps.size() == 1) code = 1;
              else if (snps.size() == 2 && analogous(*snps.back(), allele)) {
                const bool phased_pair = snps.front()->phaseId() != snps.back()->phaseId();   // Variant::homozygous
                if (phased_pair == !unphased_population) code = 2; else ++mixed;
              }
            }
          }
          if (code & 1u) unit[b >> 3] |= static_cast<uint8_t>(1u << (b & 7));
          if (code & 2u) unit[8 + (b >> 3)] |= static_cast<uint8_t>(1u << (b & 7));
        }
      }
    }
    mixed_phase += mixed;
  };
  std::vector<std::thread> pool;
  for (size_t t = 1; t < threads; ++t) pool.emplace_back(worker, t);
  worker(0);
  for (auto& th : pool) th.join();
  flat.mixed_phase_cells = mixed_phase;
  if (flat.mixed_phase_cells > 0)
    ExecEnv::log().warn("PopulationFlattener; contig: {}, {} allele pairs contradict the population's phasing and were dropped",
                        flat.contig_id, flat.mixed_phase_cells);
}

}  // namespace

std::optional<b200::FlatContig> b200::PopulationFlattener::flatten(const PopulationDB& diploid_population,
                                                                    const PopulationDB& af_population,
                                                                    const SuperPopLookup& super_population,
                                                                    bool unphased_population,
                                                                    size_t threads) {
  // Same preconditions as InbreedingAnalysis::populationInbreeding (kga_analysis_inbreed_diploid.cpp:26-41).
  if (af_population.getMap().size() != 1) {
    ExecEnv::log().error("PopulationFlattener::flatten; allele frequency population: {} has unexpected genome count: {}",
                         af_population.populationId(), af_population.getMap().size());
    return std::nullopt;
  }
  auto const& [af_genome_id, af_genome_ptr] = *af_population.getMap().begin();
  if (af_genome_ptr->getMap().size() != 1) {
    ExecEnv::log().error("PopulationFlattener::flatten; allele frequency genome: {} has more than 1 contig: {}", af_genome_id,
                         af_genome_ptr->getMap().size());
    return std::nullopt;
  }
  auto const& [contig_id, af_contig_ptr] = *af_genome_ptr->getMap().begin();

  FlatContig flat;
  flat.contig_id = contig_id;
  flat.unphased = unphased_population;

  // ---- locus table: one row per AF offset; offsets with several distinct alt alleles also get the side structures ---------
  LocusTable table;
  for (auto const& [offset, offset_ptr] : af_contig_ptr->getMap()) {
    const OffsetDBArray& variants = offset_ptr->getVariantArray();
    // AlleleFreqVector keeps, per allele, the first analogous variant that HAS a value for the super-population (freq.cpp:24-52).
    auto allele_af = [&variants](const Variant& allele, size_t k) -> float {
      for (auto const& v : variants) {
        if (not analogous(*v, allele)) continue;
        auto af_opt = FrequencyDatabaseRead::superPopFrequency(*v, kSuperPopCodes[k]);
        if (af_opt) return static_cast<float>(af_opt.value());            // stored as float by the parser: exact
      }
      return std::numeric_limits<float>::quiet_NaN();
    };
    if (not addLocus(flat, table, offset, variants, allele_af)) return std::nullopt;
  }
  finishLocusTable(flat, table);

  // ---- genome columns: genomes that have the contig, a PED record and a known super-population (diploid.cpp:121-140) -----
  std::vector<std::shared_ptr<const ContigDB>> genome_contig;
  for (auto const& [genome_id, genome_ptr] : diploid_population.getMap()) {
    auto contig_opt = std::const_pointer_cast<const GenomeDB>(genome_ptr)->getContig(contig_id);
    if (!contig_opt) continue;
    auto code_opt = super_population(genome_id);
    if (!code_opt) {
      ExecEnv::log().error("PopulationFlattener::flatten; genome sample: {} does not have a PED record", genome_id);
      continue;
    }
    auto index_opt = superPopIndex(code_opt.value());
    if (!index_opt) {
      ExecEnv::log().error("PopulationFlattener::flatten; locus set not found for super population: {}", code_opt.value());
      continue;
    }
    flat.genome_ids.push_back(genome_id);
    flat.superpop.push_back(index_opt.value());
    genome_contig.push_back(contig_opt.value());
  }
  fillGenotypes(flat, table, genome_contig, unphased_population, threads);
  return flat;
}

// A population that is its own locus list (the Pf7 analyses, kga_PfEMP: CalcFWS, HeteroHomoZygous): every offset at which some
// genome of the contig carries a SNP becomes a row, its alleles are the distinct SNPs seen there (first appearance in genome
// order), and the frequency columns all hold the allele's own INFO value `af_field` (what P7FrequencyFilter reads,
// kgl_variant_filter_Pf7.cpp:20-66; absent -> no value).
std::optional<b200::FlatContig> b200::PopulationFlattener::flattenSelf(const PopulationDB& population, const ContigId_t& contig_id,
                                                                        const std::function<std::optional<double>(const Variant&)>& allele_frequency,
                                                                        bool unphased_population, size_t threads) {
  FlatContig flat;
  flat.contig_id = contig_id;
  flat.unphased = unphased_population;
  std::vector<std::shared_ptr<const ContigDB>> genome_contig;
  std::map<ContigOffset_t, OffsetDBArray> union_map;
  for (auto const& [genome_id, genome_ptr] : population.getMap()) {
    auto contig_opt = std::const_pointer_cast<const GenomeDB>(genome_ptr)->getContig(contig_id);
    if (!contig_opt) continue;
    flat.genome_ids.push_back(genome_id);
    flat.superpop.push_back(0);
    genome_contig.push_back(contig_opt.value());
    for (auto const& [offset, offset_ptr] : contig_opt.value()->getMap()) {
      OffsetDBArray& alleles = union_map[offset];
      for (auto const& v : offset_ptr->getVariantArray()) {
        if (not v->isSNP()) { ++flat.non_snp_entries; continue; }
        bool seen = false;
        for (auto const& a : alleles) if (analogous(*v, *a)) { seen = true; break; }
        if (not seen) alleles.push_back(v);
      }
    }
  }
  LocusTable table;
  for (auto const& [offset, alleles] : union_map) {
    auto allele_af = [&allele_frequency](const Variant& allele, size_t) -> float {
      auto af_opt = allele_frequency(allele);
      return af_opt ? static_cast<float>(af_opt.value()) : std::numeric_limits<float>::quiet_NaN();
    };
    if (not addLocus(flat, table, offset, alleles, allele_af)) return std::nullopt;
  }
  finishLocusTable(flat, table);
  if (flat.non_snp_entries > 0)
    ExecEnv::log().warn("PopulationFlattener::flattenSelf; contig: {}, {} variant entries that are not SNPs are not in the matrix", contig_id,
                        flat.non_snp_entries);
  fillGenotypes(flat, table, genome_contig, unphased_population, threads);
  return flat;
}
