"""Drop-in tests through the reference's own plugin interface (oracle/_ref/kgl_plugin_harness).

The harness links the UNMODIFIED reference INBREED analysis (kga::InbreedAnalysis and everything under it) and the
product's host layer (kgl_gene_b200/host: PopulationFlattener + kga::InbreedB200Analysis over libkgl_b200.so) into one
program, builds the populations through the reference's containers, and drives both analyses through
VirtualAnalysis::{initialize,fileRead,iteration,finalize}Analysis. Each writes its CSV with the reference's CSV writer.

CPU part: the flattener must reproduce, bit for bit, the flat population the reference containers were built from.
GPU part: the two CSV files must agree (same header, same window columns, same genomes; coefficients to the 6
significant digits the reference prints).
"""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

HARNESS = os.path.join(ROOT, "oracle", "_ref", "kgl_plugin_harness")
needs_harness = pytest.mark.skipif(not os.path.exists(HARNESS), reason="oracle/_ref/kgl_plugin_harness not built (make -C oracle plugin)")


def run_harness(tmp_path, pop, *extra):
    src = os.path.join(tmp_path, "in.flat")
    work = os.path.join(tmp_path, "work")
    pop.write(src)
    env = dict(os.environ, KGL_REF_LOG=os.path.join(tmp_path, "harness.log"))
    r = subprocess.run([HARNESS, src, work, *extra], capture_output=True, text=True, timeout=1200, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    return work


def read_csv(path):
    with open(path) as f:
        lines = [ln.rstrip("\n") for ln in f]
    header, columns = lines[0], lines[1].split(",")
    rows = {}
    for ln in lines[2:]:
        p = ln.split(",")
        rows[p[0]] = (p[1:9], [float(x) for x in p[9:] if x != ""])
    return header, columns, rows


@needs_harness
@pytest.mark.parametrize("unphased", [False, True])
def test_flattener_reproduces_the_flat_population(tmp_path, unphased):
    from kgl_gene_b200.flatfile import FlatPopulation
    from kgl_gene_b200.synth import make_population
    pop, _ = make_population(131, 1500, seed=17, spectrum="sfs", missing_rate=0.02, unphased=unphased, missing_af_rate=0.05)
    work = run_harness(str(tmp_path), pop, "--no-reference", "--no-b200")
    got = FlatPopulation.read(os.path.join(work, "flattened.flat"))
    # genomes without any non-reference allele never enter a PopulationDB: compare the columns that exist
    ids = [ln.strip() for ln in open(os.path.join(work, "flattened_genomes.txt"))]
    cols = np.array([int(i[1:]) for i in ids])
    # loci whose AF genome has no variant... every locus has one here; a locus with NaN AF for all pops still has a variant
    assert np.array_equal(got.offsets, pop.offsets)
    assert np.array_equal(got.af.view(np.uint32), pop.af.view(np.uint32))
    assert np.array_equal(got.superpop, pop.superpop[cols])
    assert got.unphased == unphased
    assert np.array_equal(got.codes(), pop.codes()[:, cols])


@needs_harness
@pytest.mark.parametrize("seed", [int(x) for x in os.environ.get("KGL_FLATTEN_FUZZ_SEEDS", "1,2,3,4,5,6,7,8").split(",")])
def test_flattener_on_random_populations(tmp_path, seed):
    """The flattener against the reference containers on populations drawn at random: shape, spectrum, phase, missing cells and
    frequencies, super-population order, multi-allelic sites (with alleles outside the list and cells of more than two variants).
    Everything it emits -- offsets, frequency bits, codes, side structures -- bit for bit what the containers were built from."""
    from kgl_gene_b200.flatfile import FlatPopulation
    from kgl_gene_b200.synth import add_multi_allelic, make_population
    rng = np.random.default_rng(seed)
    n, l = int(rng.choice([1, 7, 64, 65, 150])), int(rng.choice([1, 40, 700, 2000]))
    pop, _ = make_population(n, l, seed=int(rng.integers(1, 10**6)), spectrum=str(rng.choice(["sfs", "dense"])),
                             missing_rate=float(rng.choice([0.0, 0.02])), missing_af_rate=float(rng.choice([0.0, 0.05])),
                             grouped=bool(rng.integers(0, 2)), unphased=bool(rng.integers(0, 2)))
    if l >= 40 and rng.integers(0, 2):
        add_multi_allelic(pop, int(rng.choice([1, l // 10])), seed=int(rng.integers(1, 10**6)))
    work = run_harness(str(tmp_path), pop, "--no-reference", "--no-b200")
    got = FlatPopulation.read(os.path.join(work, "flattened.flat"))
    ids = [ln.strip() for ln in open(os.path.join(work, "flattened_genomes.txt"))]
    cols = np.array([int(i[1:]) for i in ids], dtype=np.int64)
    carries = (pop.codes() != 0).any(axis=0)
    assert np.array_equal(cols, np.flatnonzero(carries))          # genomes without any non-reference allele never enter a PopulationDB
    assert np.array_equal(got.offsets, pop.offsets)
    assert np.array_equal(got.af.view(np.uint32), pop.af.view(np.uint32))
    assert np.array_equal(got.superpop, pop.superpop[cols])
    assert got.unphased == pop.unphased or not carries.any()
    assert np.array_equal(got.codes(), pop.codes()[:, cols])
    assert got.n_multi == pop.n_multi
    if pop.n_multi:
        assert np.array_equal(got.multi_rows, pop.multi_rows)
        assert np.array_equal(got.multi_af.view(np.uint32), pop.multi_af.view(np.uint32))
        assert np.array_equal(got.multi_cells, pop.multi_cells[:, cols])


@needs_harness
@pytest.mark.gpu
@pytest.mark.parametrize("algo,rtol,atol", [("Simple", 2e-6, 1e-9), ("RitlandLocus", 2e-6, 1e-9),
                                            # the reference's optimiser stops at xtol_abs = 1e-6 (kga_analysis_inbreed_calc.cpp:131-143)
                                            ("Loglikelihood", 2e-6, 5e-6)])
def test_plugin_csv_matches_reference_plugin(tmp_path, algo, rtol, atol):
    from kgl_gene_b200.synth import make_population
    pop, _ = make_population(96, 4000, seed=23, spectrum="dense", missing_rate=0.01, grouped=False)
    work = run_harness(str(tmp_path), pop, "--algo", algo, "--spacing", "20", "--count", "600")
    h_ref, c_ref, r_ref = read_csv(os.path.join(work, "INBREED", "harness_out.csv"))
    h_new, c_new, r_new = read_csv(os.path.join(work, "INBREED_B200", "harness_out.csv"))
    assert h_ref == h_new                      # parameter header line
    assert c_ref == c_new and len(c_ref) > 9 + 2   # same window columns, several windows
    assert sorted(r_ref) == sorted(r_new)      # same genomes
    for g, (meta, vals) in r_ref.items():
        meta2, vals2 = r_new[g]
        assert meta == meta2 and len(vals) == len(vals2)
        assert np.allclose(vals2, vals, rtol=rtol, atol=atol), (g, vals, vals2)


@needs_harness
def test_flattener_emits_the_multi_allelic_side_structures(tmp_path):
    """Offsets with several alt alleles in the AF population: the row stays in the locus table (frequency "no value", matrix
    codes 0 / 3) and the side structures name every allele's frequencies and every genome's one or two alleles, bit for bit
    what the reference containers were built from."""
    from kgl_gene_b200.flatfile import FlatPopulation
    from kgl_gene_b200.synth import add_multi_allelic, make_population
    pop, _ = make_population(77, 1200, seed=19, spectrum="sfs", missing_rate=0.01, missing_af_rate=0.03)
    add_multi_allelic(pop, 140, seed=20)
    work = run_harness(str(tmp_path), pop, "--no-reference", "--no-b200")
    got = FlatPopulation.read(os.path.join(work, "flattened.flat"))
    ids = [ln.strip() for ln in open(os.path.join(work, "flattened_genomes.txt"))]
    cols = np.array([int(i[1:]) for i in ids])
    assert np.array_equal(got.offsets, pop.offsets)
    assert np.array_equal(got.af.view(np.uint32), pop.af.view(np.uint32))
    assert np.array_equal(got.codes(), pop.codes()[:, cols])
    assert np.array_equal(got.multi_rows, pop.multi_rows)
    assert np.array_equal(got.multi_af.view(np.uint32), pop.multi_af.view(np.uint32))
    assert np.array_equal(got.multi_cells, pop.multi_cells[:, cols])


@needs_harness
@pytest.mark.gpu
def test_plugin_multi_allelic_contig_and_shared_upload(tmp_path):
    """A contig with multi-allelic sites (normal for 1000 Genomes), three parameter blocks in one iteration: the drop-in
    flattens and uploads ONCE (kga_analysis_inbreed_b200.cpp, iterationAnalysis) and every block's CSV equals the unmodified
    reference analysis' -- spacing > 0, so a dropped multi-allelic locus would shift the whole accept chain."""
    from kgl_gene_b200.synth import add_multi_allelic, make_population
    pop, _ = make_population(90, 4000, seed=31, spectrum="dense", missing_rate=0.01, grouped=False)
    add_multi_allelic(pop, 400, seed=32)
    work = run_harness(str(tmp_path), pop, "--algo", "Simple,RitlandLocus,Loglikelihood", "--spacing", "20", "--count", "600")
    log = open(os.path.join(str(tmp_path), "harness.log")).read()
    assert log.count("resident on the device") == 1 and "400 multi-allelic loci" in log
    for algo, atol in (("Simple", 1e-9), ("RitlandLocus", 1e-9), ("Loglikelihood", 5e-6)):
        h_ref, c_ref, r_ref = read_csv(os.path.join(work, "INBREED", f"harness_out_{algo}.csv"))
        h_new, c_new, r_new = read_csv(os.path.join(work, "INBREED_B200", f"harness_out_{algo}.csv"))
        assert h_ref == h_new and c_ref == c_new and len(c_ref) > 9 + 2 and sorted(r_ref) == sorted(r_new)
        for g, (meta, vals) in r_ref.items():
            assert np.allclose(r_new[g][1], vals, rtol=2e-6, atol=atol), (algo, g)


@needs_harness
@pytest.mark.gpu
def test_plugin_unphased_population(tmp_path):
    from kgl_gene_b200.synth import make_population
    pop, _ = make_population(64, 3000, seed=29, spectrum="dense", missing_rate=0.0, unphased=True)
    work = run_harness(str(tmp_path), pop, "--algo", "Simple", "--spacing", "0", "--count", "1000")
    _, c_ref, r_ref = read_csv(os.path.join(work, "INBREED", "harness_out.csv"))
    _, c_new, r_new = read_csv(os.path.join(work, "INBREED_B200", "harness_out.csv"))
    assert c_ref == c_new and sorted(r_ref) == sorted(r_new)
    for g, (_, vals) in r_ref.items():
        assert np.allclose(vals, r_new[g][1], rtol=2e-6, atol=2e-8)


def _compare_fws_csv(work, tag):
    """CalcFwsB200 (host/kga_analysis_pfemp_b200.cpp over kgl_b200_run_binned_genome_counts, ..._allele_count, ..._multi_allele_count)
    against the reference's CalcFWS (kga_analysis_PfEMP_FWS.cpp, calcFwsStatistics + the two writers): per-genome records in the
    eleven allele-frequency bins and the HGVS-keyed per-variant records, both CSV files byte for byte."""
    for name in ("fws_genome.csv", "fws_variant.csv"):
        ref = open(os.path.join(work, "PFEMP", name)).read().splitlines()
        new = open(os.path.join(work, "PFEMP_B200", name)).read().splitlines()
        assert len(ref) > 1 and ref[0] == new[0], (tag, name)
        assert ref == new, (tag, name, len(ref), len(new), [(a, b) for a, b in zip(ref, new) if a != b][:2])


@needs_harness
@pytest.mark.gpu
@pytest.mark.parametrize("unphased", [True, False], ids=["pf7-unphased", "phased"])
def test_pfemp_hetero_homo_csv_equals_reference_writer(tmp_path, unphased):
    """kga_PfEMP's second consumer of the counting kernels (rows a17/a18): HeteroHomoB200 (host/kga_analysis_pfemp_b200.cpp) over
    kgl_b200_run_hetero_homo + kgl_b200_location_fis against the reference's own HeteroHomoZygous -- analyzeVariantPopulation,
    location aggregates, UpdateSampleLocation (Wright's F_IS incl. the city -> country fallback) and write_variant_results -- on a
    Pf7-style population with multi-allelic sites: the two CSV files must be identical, F_IS column included."""
    from kgl_gene_b200.synth import add_multi_allelic, make_population
    pop, _ = make_population(170, 2500, seed=41, spectrum="sfs", missing_rate=0.01, unphased=unphased, grouped=False)
    add_multi_allelic(pop, 250, seed=42, three_rate=0.0)
    work = run_harness(str(tmp_path), pop, "--pfemp")
    ref = open(os.path.join(work, "PFEMP", "hetero_homo.csv")).read().splitlines()
    new = open(os.path.join(work, "PFEMP_B200", "hetero_homo.csv")).read().splitlines()
    assert len(ref) > 150 and ref[0] == new[0]
    assert ref == new
    fis = np.array([float(ln.split(",")[2]) for ln in ref[1:]])
    het_diff = np.array([int(ln.split(",")[14]) for ln in ref[1:]])
    assert np.count_nonzero(fis) > 100 and het_diff.sum() > 0          # F_IS is exercised; so is "Het Diff Minor (a;b)"
    _compare_fws_csv(work, unphased)
    assert len(open(os.path.join(work, "PFEMP", "fws_variant.csv")).read().splitlines()) > 2500      # one line per variant: more than the offsets


@needs_harness
@pytest.mark.gpu
@pytest.mark.parametrize("seed", [int(x) for x in os.environ.get("KGL_FUZZ_SEEDS", "5,6,7").split(",")])
def test_pfemp_hetero_homo_csv_on_random_populations(tmp_path, seed):
    """HeteroHomoB200 against the reference's HeteroHomoZygous writer on populations drawn at random: the two CSV files are identical."""
    from kgl_gene_b200.synth import add_multi_allelic, make_population
    rng = np.random.default_rng(seed)
    n, l = int(rng.choice([9, 64, 170, 300])), int(rng.choice([50, 800, 2500]))
    pop, _ = make_population(n, l, seed=int(rng.integers(1, 10**6)), spectrum=str(rng.choice(["sfs", "dense"])),
                             missing_rate=float(rng.choice([0.0, 0.01, 0.05])), unphased=bool(rng.integers(0, 2)), grouped=bool(rng.integers(0, 2)))
    if l >= 800 and rng.integers(0, 3):
        add_multi_allelic(pop, int(rng.choice([3, l // 10])), seed=int(rng.integers(1, 10**6)), three_rate=0.0)
    work = run_harness(str(tmp_path), pop, "--pfemp")
    ref = open(os.path.join(work, "PFEMP", "hetero_homo.csv")).read().splitlines()
    new = open(os.path.join(work, "PFEMP_B200", "hetero_homo.csv")).read().splitlines()
    assert len(ref) > 1 and ref == new, (seed, n, l, pop.n_multi, [i for i, (a, b) in enumerate(zip(ref, new)) if a != b][:3])
    _compare_fws_csv(work, (seed, n, l, pop.n_multi))


@needs_harness
@pytest.mark.gpu
def test_calc_fws_b200_refuses_what_the_matrix_cannot_hold(tmp_path):
    """A genome with three variants at one offset (side cell 0xFF: the flat form does not say which): CalcFwsB200 refuses the
    population instead of writing records that differ from the reference's; HeteroHomoB200, which only needs the number of entries
    there, still equals the reference file."""
    from kgl_gene_b200.synth import add_multi_allelic, make_population
    pop, _ = make_population(60, 800, seed=51, spectrum="sfs", unphased=True)
    add_multi_allelic(pop, 120, seed=52, three_rate=0.02)
    assert (pop.multi_cells == 0xFF).any()
    work = run_harness(str(tmp_path), pop, "--pfemp")
    assert os.path.exists(os.path.join(work, "PFEMP", "fws_genome.csv"))
    assert not os.path.exists(os.path.join(work, "PFEMP_B200", "fws_genome.csv")) and not os.path.exists(os.path.join(work, "PFEMP_B200", "fws_variant.csv"))
    assert "more than two variants at one offset" in open(os.path.join(str(tmp_path), "harness.log")).read()
    ref = open(os.path.join(work, "PFEMP", "hetero_homo.csv")).read().splitlines()
    new = open(os.path.join(work, "PFEMP_B200", "hetero_homo.csv")).read().splitlines()
    assert ref == new


def _carried_alleles(pop, col):
    """{offset: list per genome of the sorted frequency values (column `col`) of the alleles the genome carries there}."""
    out = {}
    codes = pop.codes()
    multi_of = {int(r): m for m, r in enumerate(pop.multi_rows)} if pop.n_multi else {}
    for l in range(pop.n_loci):
        if l in multi_of:
            m = multi_of[l]
            rows = []
            for c in pop.multi_cells[m]:
                c = int(c)
                assert c != 0xFF and (c & 15) != 4 and (c >> 4) != 4
                slots = [s - 1 for s in (c & 15, c >> 4) if s]
                rows.append(sorted(float(pop.multi_af[col, m, s]) for s in slots))
        else:
            a = float(pop.af[col, l])
            rows = [[a] * int(c) if c < 3 else None for c in codes[l]]
        out[int(pop.offsets[l])] = rows
    return out


@needs_harness
def test_vcf_ingest_equals_the_reference_vcf_parser(tmp_path):
    """N2 pinned against the reference's own parser: the same plain-text VCF (bi- and multi-allelic sites, a repeated POS, "."
    alleles, a GT with extra FORMAT fields) goes (a) through Genome1000VCFImpl / VCFReaderMT / ParseVCF
    (kgl_variant_factory_1000_impl.cpp:63-272, compiled into the harness) into a PopulationDB, flattened by the product's
    flattener, and (b) through kgl_b200_vcf_ingest. Every genome must carry the same alleles at every offset."""
    from kgl_gene_b200.flatfile import FlatPopulation
    from kgl_gene_b200.synth import add_multi_allelic, make_population
    from kgl_gene_b200.vcf import ingest_vcf, write_vcf
    pop, _ = make_population(60, 600, seed=12, missing_rate=0.01)
    add_multi_allelic(pop, 70, seed=13, unknown_rate=0.0, three_rate=0.0)
    rng = np.random.default_rng(3)
    nan = np.isnan(pop.multi_af[5]) & (np.arange(3)[None, :] < 2)          # every listed allele gets an "AF" value (the comparison key)
    pop.multi_af[5][nan] = rng.uniform(0.01, 0.3, size=int(nan.sum())).astype(np.float32)
    base = int(pop.offsets[-1]) + 100
    extra = [
        (599, f"22\t{base}\t.\tC\tT\t100\tPASS\tAF=0.31\tGT\t" + "\t".join(["0|1", "1|1", "0|0"] * 20)),          # repeated POS:
        (599, f"22\t{base}\t.\tC\tA\t100\tPASS\tAF=0.11\tGT\t" + "\t".join(["0|0", "0|0", "1|0"] * 20)),          #   one offset, two alleles
        (599, f"22\t{base + 10}\t.\tA\tG\t100\tPASS\tAF=0.125;DP=7\tGT:DP\t" + "\t".join(["1|0:3", ".|1:1", "0|0:9"] * 20)),
    ]
    path = str(tmp_path / "pin.vcf")
    write_vcf(pop, path, extra_lines=extra)
    ours, names, _, st = ingest_vcf(path, n_threads=2)
    assert st["multi_allelic"] == 71
    work = os.path.join(str(tmp_path), "work")
    env = dict(os.environ, KGL_REF_LOG=os.path.join(str(tmp_path), "harness.log"))
    r = subprocess.run([HARNESS, path, work, "--vcf"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    ref = FlatPopulation.read(os.path.join(work, "flattened.flat"))
    ids = [ln.strip() for ln in open(os.path.join(work, "flattened_genomes.txt"))]
    cols = [names.index(i) for i in ids]
    want = _carried_alleles(ref, 5)
    got = _carried_alleles(ours, 5)
    assert set(want) <= set(got) and len(want) > 500
    for offset, rows in got.items():
        if offset in want:
            assert [rows[c] for c in cols] == want[offset], offset
        else:                                  # an offset no genome carries a SNP at: absent from the variant DB
            assert all(r == [] for r in rows), offset
    # genomes that are absent from the variant DB carry nothing anywhere
    absent = [g for g in range(len(names)) if g not in cols]
    for rows in got.values():
        assert all(rows[g] == [] for g in absent)


def test_vcf_ingest_refuses_an_unsorted_file(tmp_path):
    """The locus table must be sorted and unique (select_loci searches it): a record whose POS lies before the previous one is an
    error, not a silently unsorted table (the reference's variant DB sorts by itself, a streamed ingest cannot)."""
    from kgl_gene_b200.synth import make_population
    from kgl_gene_b200.vcf import ingest_vcf, write_vcf
    pop, _ = make_population(5, 40, seed=3)
    late = f"22\t{int(pop.offsets[3]) + 2}\t.\tA\tG\t100\tPASS\tAF=0.2\tGT\t" + "\t".join(["0|1"] * 5)
    path = str(tmp_path / "unsorted.vcf")
    write_vcf(pop, path, extra_lines=[(20, late)])
    with pytest.raises(RuntimeError, match="not sorted by POS"):
        ingest_vcf(path, n_threads=1)


KINDS = os.environ.get("KGL_VCF_FUZZ_KINDS", "indel,symbolic,format,noaf,mnp,mixed_alt,star,lower,three,unphased_gt,haploid").split(",")


@needs_harness
@pytest.mark.parametrize("seed", [int(x) for x in os.environ.get("KGL_VCF_FUZZ_SEEDS", "1,2,3,4,5,6").split(",")])
def test_vcf_ingest_on_random_files(tmp_path, seed):
    """N2 on VCF files drawn at random: population shape, missing calls, multi-allelic sites, and spliced-in records of the kinds
    a 1000 Genomes file holds -- two SNP records at one POS (one genome may carry both alleles), an indel record between them,
    indels, MNPs, symbolic and spanning-deletion ("*") alleles, lower-case bases, three alternate alleles, an ALT list that mixes a SNP
    with an indel, '/'-separated and haploid calls, extra FORMAT fields with half-missing calls, a record without AF.
    kgl_b200_vcf_ingest against the reference's own parser (compiled into the harness), allele by allele for every genome."""
    from kgl_gene_b200.flatfile import FlatPopulation
    from kgl_gene_b200.synth import add_multi_allelic, make_population
    from kgl_gene_b200.vcf import ingest_vcf, write_vcf
    rng = np.random.default_rng(seed)
    n, l = int(rng.choice([3, 20, 60])), int(rng.choice([50, 300, 600]))
    pop, _ = make_population(n, l, seed=int(rng.integers(1, 10**6)), missing_rate=float(rng.choice([0.0, 0.01])))
    n_multi = int(rng.choice([0, 5, 40]))
    if n_multi:
        add_multi_allelic(pop, min(n_multi, l // 4), seed=int(rng.integers(1, 10**6)), unknown_rate=0.0, three_rate=0.0)
        nan = np.isnan(pop.multi_af[5]) & (np.arange(3)[None, :] < 2)
        pop.multi_af[5][nan] = rng.uniform(0.01, 0.3, size=int(nan.sum())).astype(np.float32)

    def gts(choices):
        return "\t".join(rng.choice(choices, size=n))

    def pair():        # two records at one POS: per genome a combination with at most two alternate alleles in total
        combos = rng.choice(5, size=n)
        first = ["0|1", "1|1", "0|0", "0|0", "0|1"]
        second = ["0|0", "0|0", "1|0", "0|0", "1|0"]
        return "\t".join(first[c] for c in combos), "\t".join(second[c] for c in combos)

    extra = []
    base = int(pop.offsets[-1]) + 100
    for i in range(int(rng.integers(1, 4))):
        a, b = pair()
        pos = base + 40 * i
        extra.append((l - 1, f"22\t{pos}\t.\tC\tT\t100\tPASS\tAF=0.31\tGT\t{a}"))
        if rng.integers(0, 2):
            extra.append((l - 1, f"22\t{pos}\t.\tCA\tC\t100\tPASS\tAF=0.21\tGT\t" + gts(["0|1", "0|0"])))
        extra.append((l - 1, f"22\t{pos}\t.\tC\tA\t100\tPASS\tAF=0.11\tGT\t{b}"))
    for kind in rng.choice(KINDS, size=int(rng.integers(2, 6))):
        row = int(rng.integers(0, l))
        pos = int(pop.offsets[row]) + 1 + int(rng.integers(1, 9))         # between two loci of the generated population
        line = {"indel": f"22\t{pos}\t.\tAT\tA\t100\tPASS\tAF=0.2\tGT\t" + gts(["0|1", "1|1", "0|0"]),
                "symbolic": f"22\t{pos}\t.\tA\t<CN0>\t100\tPASS\tAF=0.2\tGT\t" + gts(["0|1", "1|1", "0|0"]),
                "format": f"22\t{pos}\t.\tA\tG\t100\tPASS\tAF=0.125;DP=7\tGT:DP\t" + gts(["1|0:3", ".|1:1", "0|0:9"]),
                "noaf": f"22\t{pos}\t.\tA\tG\t100\tPASS\tDP=7\tGT\t" + gts(["1|0", "0|1", "0|0"]),
                "mnp": f"22\t{pos}\t.\tAC\tGT\t100\tPASS\tAF=0.3\tGT\t" + gts(["1|0", "0|1", "0|0"]),
                "mixed_alt": f"22\t{pos}\t.\tA\tG,AT\t100\tPASS\tAF=0.3,0.1\tGT\t" + gts(["1|0", "0|2", "0|0", "1|1", "2|0"]),
                "star": f"22\t{pos}\t.\tA\tG,*\t100\tPASS\tAF=0.3,0.1\tGT\t" + gts(["1|0", "0|2", "0|0", "1|1"]),
                "lower": f"22\t{pos}\t.\ta\tg\t100\tPASS\tAF=0.3\tGT\t" + gts(["1|0", "0|1", "0|0"]),
                "three": f"22\t{pos}\t.\tA\tG,C,T\t100\tPASS\tAF=0.3,0.1,0.05\tGT\t" + gts(["1|0", "0|2", "0|0", "3|0", "0|3"]),
                "unphased_gt": f"22\t{pos}\t.\tA\tG\t100\tPASS\tAF=0.3\tGT\t" + gts(["1/0", "0/1", "0/0", "1/1"]),
                "haploid": f"22\t{pos}\t.\tA\tG\t100\tPASS\tAF=0.3\tGT\t" + gts(["1", "0", "0|1"])}[str(kind)]
        if row < l - 1 and not any(e[1].split("\t")[1] == str(pos) for e in extra):      # after the last row come the pairs above
            extra.append((row, line))
    extra.sort(key=lambda e: (e[0], int(e[1].split("\t")[1])))                              # a VCF is sorted by POS
    path = str(tmp_path / "random.vcf")
    write_vcf(pop, path, extra_lines=extra)
    block = int(rng.choice([0, 0, 64, 700, 5000]))          # the ingest streams the file in blocks: small ones cut lines and POS groups
    if block:
        os.environ["KGL_B200_VCF_BLOCK_BYTES"] = str(block)
    try:
        ours, names, _, st = ingest_vcf(path, n_threads=int(rng.choice([1, 3])))
    finally:
        os.environ.pop("KGL_B200_VCF_BLOCK_BYTES", None)
    work = os.path.join(str(tmp_path), "work")
    env = dict(os.environ, KGL_REF_LOG=os.path.join(str(tmp_path), "harness.log"))
    r = subprocess.run([HARNESS, path, work, "--vcf"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    ref = FlatPopulation.read(os.path.join(work, "flattened.flat"))
    ids = [ln.strip() for ln in open(os.path.join(work, "flattened_genomes.txt"))]
    cols = [names.index(i) for i in ids]

    def key(rows):      # NaN ("no AF") compares equal to itself
        return [None if r is None else [-1.0 if x != x else x for x in r] for r in rows]

    want, got = _carried_alleles(ref, 5), _carried_alleles(ours, 5)
    assert set(want) <= set(got), (seed, sorted(set(want) - set(got))[:5])
    assert np.all(np.diff(ours.offsets.astype(np.int64)) > 0)                    # the locus table stays sorted and unique
    for offset, rows in got.items():
        if offset in want:
            assert key([rows[c] for c in cols]) == key(want[offset]), (seed, offset)
        else:
            assert all(r == [] for r in rows), (seed, offset)


@needs_harness
@pytest.mark.gpu
@pytest.mark.parametrize("seed", [int(x) for x in os.environ.get("KGL_FUZZ_SEEDS", "5,6,7").split(",")])
def test_plugin_csv_on_random_populations_and_parameters(tmp_path, seed):
    """The INBREED_B200 analysis against the unmodified INBREED analysis on populations and parameter blocks drawn at random
    (phase, missing cells and frequencies, multi-allelic sites, sampling distance, window size, frequency range, offsets range):
    same header, same window columns, same genomes, same class counts, coefficients to the digits the reference prints."""
    from kgl_gene_b200.synth import add_multi_allelic, make_population
    rng = np.random.default_rng(seed)
    n, l = int(rng.choice([40, 96, 130])), int(rng.choice([1500, 4000, 9000]))
    pop, _ = make_population(n, l, seed=int(rng.integers(1, 10**6)), spectrum=str(rng.choice(["sfs", "dense"])),
                             missing_rate=float(rng.choice([0.0, 0.01])), missing_af_rate=float(rng.choice([0.0, 0.03])),
                             grouped=bool(rng.integers(0, 2)), unphased=bool(rng.integers(0, 3) == 0))
    if rng.integers(0, 2):
        add_multi_allelic(pop, int(rng.choice([20, 200])), seed=int(rng.integers(1, 10**6)))
    args = ["--algo", "Simple,RitlandLocus,Loglikelihood", "--spacing", str(int(rng.choice([0, 20, 300]))),
            "--count", str(int(rng.choice([150, 600, 5000]))), "--min-af", str(float(rng.choice([0.0, 0.01, 0.1]))),
            "--max-af", str(float(rng.choice([1.0, 0.5])))]
    if rng.integers(0, 2):
        lo, hi = sorted(rng.integers(0, l, size=2).tolist())
        args += ["--lower", str(int(pop.offsets[lo])), "--upper", str(int(pop.offsets[hi]))]
    work = run_harness(str(tmp_path), pop, *args)
    n_values = n_far = 0
    for algo, atol in (("Simple", 1e-9), ("RitlandLocus", 1e-9), ("Loglikelihood", 5e-6)):
        ref_csv = os.path.join(work, "INBREED", f"harness_out_{algo}.csv")
        new_csv = os.path.join(work, "INBREED_B200", f"harness_out_{algo}.csv")
        if not os.path.exists(ref_csv):
            # no window: the first count-limited window already ends at or beyond UpperOffset (kga_analysis_inbreed_diploid.cpp:53),
            # writePedResults has nothing to write (kga_analysis_inbreed_output.cpp:189) -- the drop-in must not write either
            assert not os.path.exists(new_csv), (seed, args, algo)
            continue
        h_ref, c_ref, r_ref = read_csv(ref_csv)
        h_new, c_new, r_new = read_csv(new_csv)
        assert h_ref == h_new and c_ref == c_new, (seed, args, algo)
        assert sorted(r_ref) == sorted(r_new), (seed, args, algo)
        for g, (meta, vals) in r_ref.items():
            meta2, vals2 = r_new[g]
            assert meta == meta2 and len(vals) == len(vals2), (seed, args, algo, g)
            close = np.isclose(vals2, vals, rtol=2e-6, atol=atol, equal_nan=True)
            if algo == "Loglikelihood" and pop.unphased:
                # parity unpinned: on unphased populations the reference's Nelder-Mead run regularly stops at the bound
                # or at a lower likelihood than the maximiser (checked against the oracle's objective); counted, not compared
                n_values += close.size; n_far += int((~close).sum())
            else:
                assert close.all(), (seed, args, algo, g, vals, vals2)
    assert n_far <= 0.1 * n_values, (seed, args, n_far, n_values)
