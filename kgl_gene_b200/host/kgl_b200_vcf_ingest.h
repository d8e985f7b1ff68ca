/* kgl_b200_vcf_ingest.h -- VCF genotype columns straight into the packed 2-bit matrix.
 *
 * The reference parses a VCF into a PopulationDB (one shared_ptr<Variant> per non-reference allele per genome,
 * kgl_parser/kgl_variant_factory_1000_impl.cpp:63-145 for the phased 1000 Genomes files, kgl_variant_factory_pf_impl.cpp:73-330
 * for Pf7) and the analysis then walks that tree; the product's flattener turns the tree into the matrix of
 * include/kgl_b200.h. This ingest skips the tree: GT columns -> 2-bit codes, INFO -> float AF vectors, POS -> offsets, with
 * the reference parsers' rules:
 *   - GT = text up to the first ':' (1000_impl.cpp:166-178); alleles split at '|' (phased, :58) or '/' (Pf7);
 *     "." and "-" are the reference allele (:56-57,:228,:242); an index beyond the ALT list, a malformed number or a
 *     haploid GT on an autosome make the whole genotype reference (:193-216,:259-272); "<...>" abstract alts are reference.
 *   - 1000 Genomes: code = (A != 0) + (B != 0): one entry = het, two entries with phases A and B = hom-alt
 *     (Variant::homozygous, kgl_variant_db.cpp:287-290).
 *   - Pf7 (unphased): a genotype with any "." allele is skipped (pf_impl.cpp:139-152); every variant is UNPHASED, so the
 *     population carries the unphased flag.
 *   - FILTER: the AF side of the inbreeding path is SNP and PASS filtered (kga_analysis_inbreed.cpp:79): a record that is
 *     not PASS keeps its genotypes but gets no AF (NaN), so it can never be selected.
 *   - Records that the 2-bit matrix cannot represent are left out and counted: ALT lists with more than one allele and
 *     repeated POS (the flattener's multi_allelic_skipped), non-SNP REF/ALT (SNPFilter, kga_analysis_inbreed_freq.cpp:436).
 *   - offset = POS - 1 (kgl offsets are 0-based).
 *   - AF columns: INFO AFR_AF, AMR_AF, EAS_AF, EUR_AF, SAS_AF, AF (DataSourceEnum::Genome1000, kgl_variant_db_freq.h:87-92),
 *     parsed with strtof like the reference's std::stof (kgl_variant_factory_vcf_parse_info.cpp:232); absent -> NaN.
 * Plain text or gzip (zlib). Lines are parsed by a pool of threads, every thread packing whole rows.
 */
#ifndef KGL_B200_VCF_INGEST_H
#define KGL_B200_VCF_INGEST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kgl_b200_vcf kgl_b200_vcf;

typedef struct kgl_b200_vcf_stats {
  uint64_t records;                 /* data lines read */
  uint64_t kept;                    /* rows of the matrix */
  uint64_t skipped_multi_allelic;   /* more than one ALT allele, or a repeated POS */
  uint64_t skipped_non_snp;
  uint64_t not_pass;                /* kept, AF set to NaN */
  uint64_t malformed_genotypes;     /* treated as reference */
  uint64_t bytes;                   /* uncompressed bytes parsed */
  double seconds;
} kgl_b200_vcf_stats;

/* Returns 0 on success; on failure a message is copied to err. unphased != 0: Pf7 rules. n_threads 0 = all cores. */
int kgl_b200_vcf_ingest(const char* path, int unphased, int n_threads, kgl_b200_vcf** out, char* err, size_t err_len);
void kgl_b200_vcf_free(kgl_b200_vcf* v);

uint64_t kgl_b200_vcf_n_genomes(const kgl_b200_vcf* v);
uint64_t kgl_b200_vcf_n_loci(const kgl_b200_vcf* v);
uint64_t kgl_b200_vcf_row_bytes(const kgl_b200_vcf* v);
const uint8_t* kgl_b200_vcf_packed(const kgl_b200_vcf* v);      /* [n_loci][row_bytes], layout of include/kgl_b200.h */
const float* kgl_b200_vcf_af(const kgl_b200_vcf* v);            /* [6][n_loci] */
const uint32_t* kgl_b200_vcf_offsets(const kgl_b200_vcf* v);    /* [n_loci] */
const char* kgl_b200_vcf_genome_name(const kgl_b200_vcf* v, uint64_t i);
const char* kgl_b200_vcf_contig(const kgl_b200_vcf* v);
void kgl_b200_vcf_get_stats(const kgl_b200_vcf* v, kgl_b200_vcf_stats* stats);

#ifdef __cplusplus
}
#endif
#endif
