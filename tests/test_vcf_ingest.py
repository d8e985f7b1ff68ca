"""N2: VCF genotype columns -> packed 2-bit matrix without a PopulationDB (kgl_gene_b200/host/kgl_b200_vcf_ingest.cpp).
Round trips through VCF text, plus the reference parsers' edge rules (kgl_variant_factory_1000_impl.cpp:148-272,
kgl_variant_factory_pf_impl.cpp:139-152). The 1000 Genomes rules are pinned against the reference's own parser in
tests/test_plugin_dropin.py::test_vcf_ingest_equals_the_reference_vcf_parser (Genome1000VCFImpl / VCFReaderMT / ParseVCF compiled
into the harness); the Pf7 parser needs a GenomeReference (FASTA / GFF readers: Boost) and stays pinned by citation."""
import os

import numpy as np
import pytest

from kgl_gene_b200.synth import make_population
from kgl_gene_b200.vcf import ingest_vcf, write_vcf


def expected_codes(pop):
    c = pop.codes().copy()
    c[c == 3] = 0            # "." alleles are the reference allele for the 1000G parser
    return c


@pytest.mark.parametrize("suffix", [".vcf", ".vcf.gz"])
def test_round_trip_phased(tmp_path, suffix):
    pop, _ = make_population(131, 700, seed=4, missing_rate=0.01, missing_af_rate=0.05)
    path = str(tmp_path / ("p" + suffix))
    write_vcf(pop, path)
    got, names, contig, st = ingest_vcf(path, n_threads=3)
    assert contig == "22" and len(names) == 131 and names[5] == "G00005"
    assert st["records"] == 700 and st["kept"] == 700 and st["malformed_genotypes"] == 0
    assert np.array_equal(got.offsets, pop.offsets)
    assert np.array_equal(got.af.view(np.uint32), pop.af.view(np.uint32))          # float bits survive repr -> strtof
    assert np.array_equal(got.codes(), expected_codes(pop))
    assert got.packed.shape == pop.packed.shape and not got.unphased


def test_round_trip_unphased_pf7(tmp_path):
    pop, _ = make_population(64, 300, seed=6, missing_rate=0.05, unphased=True)
    path = str(tmp_path / "pf.vcf")
    write_vcf(pop, path)
    got, _, _, st = ingest_vcf(path, unphased=True)
    assert got.unphased and st["kept"] == 300
    assert np.array_equal(got.codes(), expected_codes(pop))                           # ./. genotypes are skipped: no variant


def test_edge_rules(tmp_path):
    pop, _ = make_population(4, 6, seed=1, missing_rate=0.0)
    path = str(tmp_path / "e.vcf")
    base = int(pop.offsets[-1]) + 100
    extra = [
        (5, f"22\t{base}\t.\tA\tG,T\t100\tPASS\tAF=0.1,0.2;EUR_AF=.,0.4\tGT\t0|1\t1|2\t0|0\t2|2"),   # two SNP alleles: a multi-allelic row
        (5, f"22\t{base + 10}\t.\tAT\tA\t100\tPASS\tAF=0.1\tGT\t0|1\t0|0\t0|0\t1|1"),              # indel: left out
        (5, f"22\t{base + 20}\t.\tA\tG\t100\tq10\tAF=0.25\tGT\t0|1\t1|1\t0|0\t.|1"),                # not PASS: kept, AF -> NaN
        (5, f"22\t{base + 30}\t.\tA\tG\t100\tPASS\tEUR_AF=0.5;AF=0.125;DP=7\tGT:DP\t1|0:3\t0|2:1\t1:9\t-|1:2"),
        (5, f"22\t{base + 40}\t.\tC\tT\t100\tPASS\tAF=0.3\tGT\t0|1\t0|0\t0|0\t1|1"),                # repeated POS: one offset, two alleles
        (5, f"22\t{base + 40}\t.\tC\tA\t100\tPASS\tAF=0.1\tGT\t0|0\t0|1\t0|0\t1|0"),
        (5, f"22\t{base + 50}\t.\tA\tG,AT\t100\tPASS\tAF=0.2,0.01\tGT\t0|1\t1|2\t2|2\t2|0"),          # SNP + indel: the indel vanishes
    ]
    write_vcf(pop, path, extra_lines=extra)
    got, _, _, st = ingest_vcf(path, n_threads=1)
    assert st["records"] == 13 and st["kept"] == 11 and st["multi_allelic"] == 2 and st["skipped_non_snp"] == 1
    assert st["not_pass"] == 1
    assert got.offsets.tolist()[6:] == [base - 1, base + 20 - 1, base + 30 - 1, base + 40 - 1, base + 50 - 1]
    # the multi-allelic rows: no value in the frequency table, 0 / 3 in the matrix, the alleles in the side structures
    assert got.multi_rows.tolist() == [6, 9]
    assert np.all(np.isnan(got.af[:, 6])) and got.codes()[6].tolist() == [3, 3, 0, 3]
    assert got.multi_cells[0].tolist() == [1, 1 | (2 << 4), 0, 2 | (2 << 4)]            # 0|1, 1|2, 0|0, 2|2
    assert got.multi_af[5, 0].tolist()[:2] == [np.float32(0.1), np.float32(0.2)] and np.isnan(got.multi_af[5, 0, 2])
    # EUR_AF=.,0.4: "." inside a list is the parser's MISSING_VALUE_FLOAT_ (the lowest float), a value -- not an absent field
    assert got.multi_af[3, 0, 0] == np.finfo(np.float32).min and got.multi_af[3, 0, 1] == np.float32(0.4)
    assert got.multi_cells[1].tolist() == [1, 2, 0, 0xFF]                               # T, T and A: three variants at the offset
    assert np.all(np.isnan(got.af[:, 7]))                                               # not PASS
    assert got.codes()[7].tolist() == [1, 2, 0, 1]                                      # ".|1": "." is the reference allele
    # 1|0 -> het; 0|2 -> index beyond the ALT list: whole genotype reference; "1" haploid on an autosome: reference; -|1 het
    assert got.codes()[8].tolist() == [1, 0, 0, 1] and st["malformed_genotypes"] == 2
    assert got.af[5, 8] == np.float32(0.125) and got.af[3, 8] == np.float32(0.5) and np.isnan(got.af[0, 8])
    # G + an insertion: one SNP allele -> an ordinary row; the insertion is not a SNP and vanishes from the genome's offset array
    assert got.codes()[10].tolist() == [1, 1, 0, 0] and got.af[5, 10] == np.float32(0.2)


def test_round_trip_multi_allelic(tmp_path):
    """A population with multi-allelic loci through VCF text and back: allele slots = ALT order, frequencies per slot from the
    Number=A fields, side cells = (phase A allele, phase B allele)."""
    from kgl_gene_b200.synth import add_multi_allelic
    for unphased in (False, True):
        pop, _ = make_population(70, 500, seed=8, missing_rate=0.0, unphased=unphased)
        add_multi_allelic(pop, 60, seed=9, unknown_rate=0.0, three_rate=0.0)
        path = str(tmp_path / f"m{int(unphased)}.vcf.gz")
        write_vcf(pop, path)
        got, _, _, st = ingest_vcf(path, unphased=unphased, n_threads=2)
        assert st["kept"] == 500 and st["multi_allelic"] == 60
        assert np.array_equal(got.offsets, pop.offsets) and np.array_equal(got.codes(), pop.codes())
        assert np.array_equal(got.af.view(np.uint32), pop.af.view(np.uint32))
        assert np.array_equal(got.multi_rows, pop.multi_rows) and np.array_equal(got.multi_cells, pop.multi_cells)
        # a "." inside a Number=A list comes back as the reference parser's missing-value float (the lowest float), an absent
        # field as NaN: a slot that is NaN for a population that has values for other slots is written as "."
        lowest = np.finfo(np.float32).min
        want = pop.multi_af.copy()
        n_slots = (~np.isnan(pop.multi_af)).any(axis=0).cumsum(axis=1).argmax(axis=1) + 1          # slots the locus lists
        for k in range(6):
            for m in range(pop.n_multi):
                if not np.all(np.isnan(want[k, m, :n_slots[m]])):
                    want[k, m, :n_slots[m]] = np.where(np.isnan(want[k, m, :n_slots[m]]), lowest, want[k, m, :n_slots[m]])
        both = ~(np.isnan(got.multi_af) | np.isnan(want))
        assert np.array_equal(np.isnan(got.multi_af), np.isnan(want)) and np.array_equal(got.multi_af[both], want[both])


def test_ingested_population_feeds_the_oracle(tmp_path):
    """The ingested matrix is a valid input of the hot path: same allele counts as the population it was written from."""
    import oracle_py as O
    pop, _ = make_population(70, 400, seed=8, missing_rate=0.0)
    path = str(tmp_path / "o.vcf.gz")
    write_vcf(pop, path)
    got, _, _, _ = ingest_vcf(path)
    a, b = O.allele_count(got), O.allele_count(pop)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
