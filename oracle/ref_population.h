// ref_population.h -- TEST INFRASTRUCTURE (oracle/_ref builds only). Rebuilds a flattened population (flat_io.h, KGLFLAT1)
// through the REFERENCE's own containers: EvidenceFactory -> VariantEvidence -> Variant -> PopulationDB::addVariant
// (kgl_variant_factory_vcf_evidence.h:215-231, kgl_variant_db.h:50-63, kgl_variant_db_population.cpp:298), so that
// everything downstream of it is reference code. Only non-reference alleles are stored:
//   code 1 -> one "A>G" variant, code 2 -> two (phase A and B; both UNPHASED for an unphased population),
//   code 3 -> one "A>T" variant, an allele that is not in the AF list and therefore dropped (freq.cpp:462).
// Multi-allelic loci (flat_io.h, trailing section): allele slot a of the locus is the SNP "A>" + "GCT"[a]; the AF population
// holds one variant per slot at the offset (each with its own INFO block), a genome holds the one or two variants its cell
// names, FIRST variant first (generateFrequencies looks the front of the offset array up first, freq.cpp:462-511); slot 4 of
// a cell is the first base of "GCT" the locus does not list; 0xFF puts three variants at the offset (dropped: neither
// size() == 1 nor == 2).
#pragma once
#include "kgl_variant_db_population.h"
#include "kgl_variant_factory_vcf_evidence.h"
#include "kel_exec_env.h"

#include "flat_io.h"

#include <cmath>
#include <cstdio>
#include <memory>
#include <string>
#include <map>
#include <vector>

namespace kglref {

namespace kel = kellerberrin;
namespace kgl = kellerberrin::genome;

inline const char* const kSuperPops[6] = {"AFR", "AMR", "EAS", "EUR", "SAS", "ALL"};
inline const std::string kContig = "chrFlat";

// INFO field names the reference resolves per data source (kgl_variant_db_freq.h:84-96).
inline const char* const* afFields(kgl::DataSourceEnum source) {
  static const char* const genome1000[6] = {"AFR_AF", "AMR_AF", "EAS_AF", "EUR_AF", "SAS_AF", "AF"};
  static const char* const gnomad3_1[6] = {"AF_afr", "AF_amr", "AF_eas", "AF_nfe", "AF_sas", "AF"};
  return source == kgl::DataSourceEnum::Gnomad3_1 ? gnomad3_1 : genome1000;
}

inline std::string genomeName(uint32_t g) {
  char buf[32];
  std::snprintf(buf, sizeof buf, "G%07u", g);  // zero padded: lexicographic == numeric (std::map order)
  return buf;
}

struct BuiltPopulations {
  std::shared_ptr<kgl::PopulationDB> af_population, diploid;
  std::vector<kgl::GenomeId_t> genome_ids;
  std::vector<std::shared_ptr<const kgl::Variant>> locus_variant;   // phase-A copy of every locus' "A>G" variant
};

// diploid_evidence: the genotype population's own "A>G" variants carry the locus' INFO block (as the variants of a Pf7 VCF
// do) -- what P7FrequencyFilter (kgl_variant_filter_Pf7.cpp:20-66) reads in CalcFWS. The "A>T" stand-ins of code 3 stay
// without evidence: a variant without the AF field passes both frequency filters and therefore lands in no bin.
inline BuiltPopulations buildPopulations(const kglflat::Flat& flat, kgl::DataSourceEnum af_source, kgl::DataSourceEnum diploid_source,
                                         bool diploid_evidence = false) {
  const uint32_t N = flat.N(), L = flat.L();
  const bool unphased = (flat.hdr.flags & kglflat::FLAG_UNPHASED) != 0;
  const char* const* fields = afFields(af_source);
  BuiltPopulations out;

  // ---- INFO evidence: six Float AF fields, Number=A ----------------------------------------------
  kgl::EvidenceInfoSet info_set;
  kgl::VCFInfoRecordMap info_map;
  for (int k = 0; k < 6; ++k) {
    info_set.insert(fields[k]);
    info_map[fields[k]] = kgl::VCFInfoRecord{fields[k], "", "Float", "A", "", ""};
  }
  kgl::EvidenceFactory evidence_factory(info_set);
  evidence_factory.availableInfoFields(info_map);

  // ---- AF "genome": 1 genome, 1 contig (kga_analysis_inbreed_diploid.cpp:26,36) --------------------
  out.af_population = std::make_shared<kgl::PopulationDB>("AF_POPULATION", af_source);
  const std::vector<kgl::GenomeId_t> af_genome{"AF_GENOME"};
  std::vector<kgl::VariantEvidence> locus_evidence;
  if (diploid_evidence) locus_evidence.reserve(L);
  std::map<std::pair<uint32_t, uint32_t>, kgl::VariantEvidence> multi_evidence;   // (locus, allele slot) -> the allele's INFO, diploid side
  std::vector<int> multi_of(L, -1);
  for (uint32_t m = 0; m < flat.M(); ++m) multi_of[flat.multi_rows[m]] = int(m);
  static const char* const kAltBases[3] = {"G", "C", "T"};
  auto slots_of = [&](uint32_t m) {          // number of allele slots the locus lists (a slot is listed if any population has its AF)
    uint32_t n = 0;
    for (uint32_t a = 0; a < 3; ++a)
      for (int k = 0; k < 6; ++k) if (!std::isnan(flat.multiAf(k, m, a))) n = a + 1;
    return n;
  };
  for (uint32_t l = 0; l < L; ++l) {
    if (multi_of[l] >= 0) {
      const uint32_t m = uint32_t(multi_of[l]), n_slots = slots_of(m);
      for (uint32_t a = 0; a < n_slots; ++a) {
        std::string info;
        for (int k = 0; k < 6; ++k) {
          const float af = flat.multiAf(k, m, a);
          if (std::isnan(af)) continue;
          char buf[64];
          std::snprintf(buf, sizeof buf, "%s%s=%.9g", info.empty() ? "" : ";", fields[k], double(af));
          info += buf;
        }
        auto block = evidence_factory.createVariantEvidence(std::move(info));
        kgl::VariantEvidence evidence(l, af_source, true, block, nullptr, 0, 1);
        auto variant = std::make_shared<const kgl::Variant>(kContig, flat.offsets[l], kgl::VariantPhase::UNPHASED, "",
                                                            kgl::DNA5SequenceLinear(kgl::StringDNA5("A")),
                                                            kgl::DNA5SequenceLinear(kgl::StringDNA5(kAltBases[a])), evidence);
        if (!out.af_population->addVariant(variant, af_genome)) kel::ExecEnv::log().error("harness: AF addVariant failed at multi locus {}", l);
        // Pf7 style: every allele of the record is a variant of its own carrying its own frequency (its element of the Number=A list)
        if (diploid_evidence) multi_evidence.emplace(std::make_pair(l, a), kgl::VariantEvidence(l, diploid_source, true, block, nullptr, 0, 1));
      }
      if (diploid_evidence) locus_evidence.emplace_back(l, diploid_source, true, nullptr, nullptr, 0, 1);
      continue;
    }
    std::string info;
    for (int k = 0; k < 6; ++k) {
      const float af = flat.afAt(k, l);
      if (std::isnan(af)) continue;  // field absent for this variant -> superPopFrequency() == nullopt
      char buf[64];
      std::snprintf(buf, sizeof buf, "%s%s=%.9g", info.empty() ? "" : ";", fields[k], double(af));
      info += buf;
    }
    auto block = evidence_factory.createVariantEvidence(std::move(info));
    kgl::VariantEvidence evidence(l, af_source, true, block, nullptr, 0, 1);
    if (diploid_evidence) locus_evidence.emplace_back(l, diploid_source, true, block, nullptr, 0, 1);
    auto variant = std::make_shared<const kgl::Variant>(kContig, flat.offsets[l], kgl::VariantPhase::UNPHASED, "",
                                                        kgl::DNA5SequenceLinear(kgl::StringDNA5("A")),
                                                        kgl::DNA5SequenceLinear(kgl::StringDNA5("G")), evidence);
    if (!out.af_population->addVariant(variant, af_genome)) kel::ExecEnv::log().error("harness: AF addVariant failed at locus {}", l);
  }

  // ---- diploid population ----------------------------------------------------------------------------
  out.diploid = std::make_shared<kgl::PopulationDB>("DIPLOID", diploid_source);
  out.genome_ids.resize(N);
  for (uint32_t g = 0; g < N; ++g) out.genome_ids[g] = genomeName(g);
  out.locus_variant.resize(L);
  const kgl::VariantEvidence no_evidence(0, diploid_source, true, nullptr, nullptr, 0, 1);
  auto make = [&](uint32_t l, kgl::VariantPhase phase, const char* alt) {
    const bool with_info = diploid_evidence && alt[0] == 'G';
    const kgl::VariantEvidence* evidence = with_info ? &locus_evidence[l] : &no_evidence;
    if (diploid_evidence && multi_of[l] >= 0) {
      evidence = &no_evidence;
      for (uint32_t a = 0; a < 3; ++a)
        if (alt[0] == kAltBases[a][0]) { auto it = multi_evidence.find({l, a}); if (it != multi_evidence.end()) evidence = &it->second; }
    }
    return std::make_shared<const kgl::Variant>(kContig, flat.offsets[l], phase, "",
                                                kgl::DNA5SequenceLinear(kgl::StringDNA5("A")),
                                                kgl::DNA5SequenceLinear(kgl::StringDNA5(alt)), *evidence);
  };
  std::vector<kgl::GenomeId_t> first, second, other;
  for (uint32_t l = 0; l < L; ++l) {
    const auto phase_a0 = unphased ? kgl::VariantPhase::UNPHASED : kgl::VariantPhase::DIPLOID_PHASE_A;
    const auto phase_b0 = unphased ? kgl::VariantPhase::UNPHASED : kgl::VariantPhase::DIPLOID_PHASE_B;
    if (multi_of[l] >= 0) {
      const uint32_t m = uint32_t(multi_of[l]), n_slots = slots_of(m);
      out.locus_variant[l] = make(l, phase_a0, "G");
      auto alt_of = [&](unsigned slot1) -> const char* {     // slot1 = slot + 1; 4 = the first base the locus does not list
        if (slot1 <= 3) return kAltBases[slot1 - 1];
        return n_slots < 3 ? kAltBases[n_slots] : "N";
      };
      for (uint32_t g = 0; g < N; ++g) {
        const uint8_t cell = flat.multiCell(m, g);
        if (cell == 0) continue;
        const std::vector<kgl::GenomeId_t> one{out.genome_ids[g]};
        if (cell == 0xFF) {
          out.diploid->addVariant(make(l, phase_a0, "G"), one);
          out.diploid->addVariant(make(l, phase_b0, "G"), one);
          out.diploid->addVariant(make(l, phase_a0, "C"), one);
          continue;
        }
        out.diploid->addVariant(make(l, phase_a0, alt_of(cell & 15u)), one);
        if (cell >> 4) out.diploid->addVariant(make(l, phase_b0, alt_of(cell >> 4)), one);
      }
      continue;
    }
    first.clear(); second.clear(); other.clear();
    for (uint32_t g = 0; g < N; ++g) {
      switch (flat.code(l, g)) {
        case 1: first.push_back(out.genome_ids[g]); break;
        case 2: first.push_back(out.genome_ids[g]); second.push_back(out.genome_ids[g]); break;
        case 3: other.push_back(out.genome_ids[g]); break;
        default: break;
      }
    }
    const auto phase_a = unphased ? kgl::VariantPhase::UNPHASED : kgl::VariantPhase::DIPLOID_PHASE_A;
    const auto phase_b = unphased ? kgl::VariantPhase::UNPHASED : kgl::VariantPhase::DIPLOID_PHASE_B;
    auto va = make(l, phase_a, "G");
    out.locus_variant[l] = va;
    if (!first.empty() && !out.diploid->addVariant(va, first)) kel::ExecEnv::log().error("harness: addVariant A failed");
    if (!second.empty() && !out.diploid->addVariant(make(l, phase_b, "G"), second)) kel::ExecEnv::log().error("harness: addVariant B failed");
    if (!other.empty() && !out.diploid->addVariant(make(l, phase_a, "T"), other)) kel::ExecEnv::log().error("harness: addVariant X failed");
  }
  return out;
}

}  // namespace kglref
