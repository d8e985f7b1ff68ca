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
