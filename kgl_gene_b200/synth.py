"""Seeded synthetic populations.

Loci are biallelic SNPs at a fixed spacing; allele frequencies per super-population follow either a site-frequency
spectrum (Beta(0.2, 2) clipped to [1e-4, 0.9999]) or the dense stress law U(0.05, 0.5), rounded through float32 as the
reference stores INFO floats (kgl_variant_factory_vcf_parse_info.cpp:232). Genotypes follow the reference's own
class law {q^2+Fpq, 2pq(1-F), p^2+Fpq} (AlleleFreqVector::unadjustedAlleleClassFrequencies,
kga_analysis_inbreed_freq.cpp:127-205) with a per-genome F, drawn with a counter-based splitmix64 stream per cell so
that the numpy generator here, the C oracle and the device generator (kgl_b200_synth_genotypes) emit the same bits.
"""
from __future__ import annotations

import numpy as np

from .flatfile import FlatPopulation, pack_codes, row_bytes_for

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def mix64(z: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def make_loci(n_loci: int, seed: int, spectrum: str = "sfs", spacing: int = 10, first_offset: int = 1000,
              n_pop: int = 6, missing_af_rate: float = 0.0):
    """Returns (offsets uint32 [L], af float32 [n_pop, L])."""
    rng = np.random.default_rng(seed)
    offsets = (first_offset + spacing * np.arange(n_loci, dtype=np.uint64)).astype(np.uint32)
    if spectrum == "sfs":
        af = np.clip(rng.beta(0.2, 2.0, size=(n_pop, n_loci)), 1e-4, 0.9999)
    elif spectrum == "dense":
        af = rng.uniform(0.05, 0.5, size=(n_pop, n_loci))
    else:
        raise ValueError(spectrum)
    af = af.astype(np.float32)
    if missing_af_rate > 0:
        af[rng.random(af.shape) < missing_af_rate] = np.nan
    return offsets, af


def make_genomes(n_genomes: int, seed: int, grouped: bool = True):
    """Returns (superpop uint8 [N] over AFR..SAS, inbreeding float64 [N] ~ U(-0.05, 0.25))."""
    rng = np.random.default_rng(seed + 7919)
    if grouped:   # genomes of one super-population are contiguous, as the host flattener lays columns out
        superpop = (np.arange(n_genomes) * 5 // max(n_genomes, 1)).astype(np.uint8)
    else:
        superpop = (np.arange(n_genomes) % 5).astype(np.uint8)
    return superpop, rng.uniform(-0.05, 0.25, size=n_genomes)


def synth_codes(seed: int, af: np.ndarray, superpop: np.ndarray, inbreeding: np.ndarray,
                missing_rate: float = 0.001, locus_base: int = 0) -> np.ndarray:
    """uint8 [L, N] genotype codes; bit-identical to the device generator kgl_b200_synth_genotypes (and to the CPU checker used by the tests)."""
    n_loci = af.shape[1]
    n = superpop.shape[0]
    a = af[superpop.astype(np.int64), :].T.astype(np.float64)           # [L, N]
    p = np.where(np.isnan(a), 0.0, np.clip(a, 0.0, 1.0))
    q = 1.0 - p
    F = inbreeding[None, :]
    pq = p * q
    t0 = q * q + F * pq
    t1 = t0 + (2.0 * pq) * (1.0 - F)
    loc = (np.arange(n_loci, dtype=np.uint64) + np.uint64(locus_base))[:, None] << np.uint64(32)
    key = np.uint64(seed) ^ loc ^ np.arange(n, dtype=np.uint64)[None, :]
    h1 = mix64(key)
    h2 = mix64(h1 ^ np.uint64(0xD6E8FEB86659FD93))
    u = (h1 >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    codes = (u >= t0).astype(np.uint8) + (u >= t1).astype(np.uint8)
    codes[(h2 >> np.uint64(40)) < np.uint64(int(missing_rate * 16777216.0))] = 3
    return codes


def make_population(n_genomes: int, n_loci: int, seed: int = 20261018, spectrum: str = "sfs",
                    missing_rate: float = 0.001, unphased: bool = False, grouped: bool = True,
                    missing_af_rate: float = 0.0) -> tuple[FlatPopulation, np.ndarray]:
    """Small/medium populations on the host (numpy). Returns (population, true inbreeding per genome)."""
    offsets, af = make_loci(n_loci, seed, spectrum, missing_af_rate=missing_af_rate)
    superpop, inbreeding = make_genomes(n_genomes, seed, grouped)
    codes = synth_codes(seed, af, superpop, inbreeding, missing_rate)
    pop = FlatPopulation(offsets, af, superpop, pack_codes(codes), n_genomes, unphased)
    assert pop.row_bytes == row_bytes_for(n_genomes)
    return pop, inbreeding


def add_multi_allelic(pop: FlatPopulation, n_multi: int, seed: int, unknown_rate: float = 0.01, three_rate: float = 0.003) -> FlatPopulation:
    """Turns `n_multi` loci of `pop` into multi-allelic loci (two or three alternate alleles, FlatPopulation.multi_*): per
    super-population allele frequencies, some alleles without a frequency for a population, a few loci whose frequencies
    sum above 1 (invalid for that population), just above 1 (valid, normalised) or above 0.99 (rare major allele); every
    genome's two haplotypes are drawn from the locus' allele frequencies of its population. A few cells carry an allele that
    is not in the list (slot 4) or more than two variants (0xFF). In place; returns pop."""
    rng = np.random.default_rng(seed)
    n_pop, n_loci, n = pop.af.shape[0], pop.n_loci, pop.n_genomes
    rows = np.sort(rng.choice(n_loci, size=min(n_multi, n_loci), replace=False)).astype(np.uint32)
    m_count = rows.shape[0]
    n_alleles = rng.integers(2, 4, size=m_count)
    af = np.full((n_pop, m_count, 3), np.nan, dtype=np.float32)
    for m in range(m_count):
        a = int(n_alleles[m])
        for k in range(n_pop):
            total = rng.beta(0.6, 1.6) * 0.9 + 1e-3
            kind = rng.integers(0, 40)
            if kind == 0:
                total = 1.2                      # sum > 1 + 1e-5: the vector is invalid for this population
            elif kind == 1:
                total = 1.000004                 # sum in (1, 1 + 1e-5]: valid, the class frequencies are normalised by the sum
            elif kind == 2:
                total = 0.995                    # q <= 0.01: a hom-ref genome is dropped
            parts = rng.dirichlet(np.ones(a)) * total
            af[k, m, :a] = parts.astype(np.float32)
            if rng.random() < 0.08:
                af[k, m, rng.integers(0, a)] = np.nan        # this allele has no frequency for the population
            if rng.random() < 0.02:
                af[k, m, :] = np.nan                         # nor has any: empty vector
    cells = np.zeros((m_count, n), dtype=np.uint8)
    sp = pop.superpop.astype(np.int64)
    for m in range(m_count):
        a = int(n_alleles[m])
        p = np.nan_to_num(af[:, m, :a].astype(np.float64), nan=0.0)           # [n_pop, a]
        tot = p.sum(axis=1, keepdims=True)
        p = np.where(tot > 1.0, p / np.maximum(tot, 1e-300), p)
        cum = np.cumsum(np.concatenate([np.maximum(0.0, 1.0 - p.sum(axis=1, keepdims=True)), p], axis=1), axis=1)   # ref first
        u = rng.random((2, n))
        hap = (u[:, :, None] >= cum[sp][None, :, :]).sum(axis=2)              # 0 = ref, 1..a = allele slot + 1
        hap = np.minimum(hap, a)
        h1, h2 = hap[0], hap[1]
        first = np.where(h1 > 0, h1, h2)
        second = np.where((h1 > 0) & (h2 > 0), h2, 0)
        cell = (first | (second << 4)).astype(np.uint8)
        if a < 3:
            unk = (rng.random(n) < unknown_rate) & (cell != 0)
            swap_first = rng.random(n) < 0.5
            cell = np.where(unk & swap_first, (cell & 0xF0) | 4, cell)
            cell = np.where(unk & ~swap_first & (cell >> 4 != 0), (cell & 0x0F) | (4 << 4), cell).astype(np.uint8)
        cell = np.where(rng.random(n) < three_rate, 0xFF, cell).astype(np.uint8)
        cells[m] = cell
    pop.af[:, rows] = np.nan
    pop.packed[rows] = pack_codes(np.where(cells != 0, 3, 0).astype(np.uint8))
    pop.multi_rows, pop.multi_af, pop.multi_cells = rows, af, cells
    return pop
