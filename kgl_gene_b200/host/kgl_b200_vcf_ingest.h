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
 *   - The records of one POS are one offset of the variant DB: its distinct SNP alleles (ALT lists with several alleles,
 *     repeated POS) are the locus' allele slots, in order of appearance. One SNP allele: an ordinary row (codes 0/1/2; 3 for
 *     more than two copies). Two or three: a multi-allelic row -- "no value" in the frequency table, codes 0 / 3 in the matrix,
 *     and the side structures of include/kgl_b200.h (kgl_b200_upload_multi_allelic): per-slot frequencies from the Number=A
 *     fields (indexed by the ALT's position, kgl_variant_db_freq.cpp:89-92) and one byte per genome naming the allele of
 *     phase A, then phase B (the order of addVariants, 1000_impl.cpp:118-139). Non-SNP alleles vanish from the genome's
 *     offset array (SNPFilter, kga_analysis_inbreed_freq.cpp:436); a POS without any SNP allele is left out.
 *   - A one-character ALT is a SNP allele whatever the character: the variant DB stores it as DNA5 (case folded, U = T, IUPAC
 *     codes and the spanning deletion "*" = N, kgl_alphabet_dna5.cpp:55-90), so "*" next to a base is a second allele slot.
 *   - offset = POS - 1 (kgl offsets are 0-based). The records must be sorted by POS (the variant DB sorts by itself; the locus
 *     table of a streamed ingest is searched by kgl_b200_select_loci): a POS below its predecessor's is an error.
 *   - AF columns: INFO AFR_AF, AMR_AF, EAS_AF, EUR_AF, SAS_AF, AF (DataSourceEnum::Genome1000, kgl_variant_db_freq.h:87-92),
 *     parsed with strtof like the reference's std::stof (kgl_variant_factory_vcf_parse_info.cpp:232); absent -> NaN.
 * Plain text or gzip (zlib), streamed in blocks: memory = the outputs + one block of parsed lines. Lines are parsed by a
 * pool of threads, every thread packing whole rows.
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
  uint64_t multi_allelic;           /* rows with two or three SNP alleles (kept, side structures) */
  uint64_t skipped_too_many_alleles;/* records of a POS with more than three SNP alleles */
  uint64_t skipped_non_snp;         /* records of a POS without any SNP allele */
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
uint64_t kgl_b200_vcf_n_multi(const kgl_b200_vcf* v);
const uint32_t* kgl_b200_vcf_multi_rows(const kgl_b200_vcf* v); /* [n_multi] */
const float* kgl_b200_vcf_multi_af(const kgl_b200_vcf* v);      /* [6][n_multi][3] */
const uint8_t* kgl_b200_vcf_multi_cells(const kgl_b200_vcf* v); /* [n_multi][n_genomes] */
const char* kgl_b200_vcf_genome_name(const kgl_b200_vcf* v, uint64_t i);
const char* kgl_b200_vcf_contig(const kgl_b200_vcf* v);
void kgl_b200_vcf_get_stats(const kgl_b200_vcf* v, kgl_b200_vcf_stats* stats);

#ifdef __cplusplus
}
#endif
#endif
