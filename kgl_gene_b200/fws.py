"""Host-side mirror of the reference's Pf7 summary statistics that sit on top of the allele-counting kernels.

  CalcFWS              kga_analytic/kga_PfEMP/kga_analysis_PfEMP_FWS.cpp:15-145
      updateVariantFWSMap  -> per-variant AlleleSummmary over all genomes        = kgl_b200_run_allele_count (locus_counts)
      updateGenomeFWSMap   -> per-genome AlleleSummmary in 11 allele-frequency bins = kgl_b200_run_binned_genome_counts
  HeteroHomoZygous     kga_analytic/kga_PfEMP/kga_analysis_PfEMP_heterozygous.cpp:61-105 (updateVariantAnalysisType),
                       :362-412 (UpdateSampleLocation: Wright's FIS against the location aggregate)

The product form of the HeteroHomoZygous half is the C++ class `HeteroHomoB200` (kgl_gene_b200/host/kga_analysis_pfemp_b200.{h,cpp})
over `kgl_b200_run_hetero_homo` and `kgl_b200_location_fis` (include/kgl_b200.h); this module is the ctypes-side mirror the
tests use to check the same bookkeeping from Python.

The reference rebuilds a VariantDBVariant twelve times and copies the population eleven times for this; here it is one raw
pass plus one masked pass of the streaming kernel per bin. Everything in this module is integer bookkeeping on the kernels'
outputs; nothing here touches the CPU oracle.
"""
from __future__ import annotations

import numpy as np

# CalcFWS::getFrequency (kga_analysis_PfEMP_FWS.cpp:104-145): eleven bins [lower, upper)
FWS_BINS = [(0.0, 0.05), (0.05, 0.10), (0.10, 0.15), (0.15, 0.20), (0.20, 0.25), (0.25, 0.30), (0.30, 0.35), (0.35, 0.40),
            (0.40, 0.45), (0.45, 0.5), (0.5, 1.0)]


def calc_fws(ctx, pop: int = 0, bins=FWS_BINS, n_multi: int = 0):
    """Returns dict(variant_summary uint32[L][3], present bool[L], genome_bins uint64[n_bins][N][3], bin_variants uint64[n_bins]) and,
    with n_multi multi-allelic loci uploaded, multi_variant_summary uint32[M][3][3] / multi_present bool[M][3] -- one variant per
    listed allele there (the rows of those loci in variant_summary are not variants).

    variant_summary[l] = {referenceHomozygous_, minorHeterozygous_, minorHomozygous_} of locus l over all genomes; `present`
    marks the loci that are variants of the population (carried by some genome) -- only those have an entry in the
    reference's variant_fws_map_. genome_bins[b][g] is genome g's AlleleSummmary over the variants whose AF lies in bin b.
    """
    lc, _ = ctx.allele_count(want_loci=True, want_genomes=False)
    lower = [b[0] for b in bins]
    upper = [b[1] for b in bins]
    counts, rows = ctx.binned_genome_counts(lower, upper, pop=pop, present_only=True)
    # AlleleSummmary counts copies of THIS variant (kgl_variant_db_variant.cpp:73-103): a genome that carries some other
    # allele at the offset (code 3) has none and is referenceHomozygous_ for the column -- pinned against the reference's
    # own CalcFWS in tests/golden (ref_fws_genome, ref_fws_variant).
    variant_summary = lc[:, :3].copy()
    variant_summary[:, 0] += lc[:, 3]
    genome_bins = counts[:, :, :3].copy()
    genome_bins[:, :, 0] += counts[:, :, 3]
    out = {"variant_summary": variant_summary, "present": (lc[:, 1] + lc[:, 2]) > 0,
           "genome_bins": genome_bins, "bin_variants": rows}
    if n_multi:
        mc = ctx.multi_allele_count(n_multi)               # genomes with 0 / 1 / 2 copies per allele slot
        out["multi_variant_summary"] = mc
        out["multi_present"] = (mc[:, :, 1] + mc[:, :, 2]) > 0
    return out


def hetero_homo_summary(genome_counts: np.ndarray, other_allele_entries: int = 1) -> dict:
    """HeteroHomoZygous::updateVariantAnalysisType (kga_analysis_PfEMP_heterozygous.cpp:61-105) for a biallelic SNP matrix, from
    the per-genome code counts uint64[N][4] of kgl_b200_run_allele_count: a het cell is one variant entry at its offset, a
    hom-alt cell two, a code-3 cell `other_allele_entries` entries of some other allele (1 = the flattener's usual case and
    what the reference harness builds; 0 = the cell carries nothing, e.g. a missing call).

      total_variants_ = snp_count_                       = n1 + 2 n2 + n3  (every entry is a SNP here; indel_count_ = 0)
      heterozygous_reference_minor_alleles_              = n1 + n3         (offsets with exactly one entry, :85-87)
      homozygous_minor_alleles_                          = n2              (UniqueUnphasedFilter over two identical entries, :93-96)
      heterozygous_minor_alleles_                        = 0               (needs two different alts at one offset)

    Pinned against the reference's own translation unit (tests/golden, ref_hetero_homo).
    """
    n1 = genome_counts[:, 1].astype(np.uint64)
    n2 = genome_counts[:, 2].astype(np.uint64)
    n3 = genome_counts[:, 3].astype(np.uint64) * np.uint64(1 if other_allele_entries else 0)
    total = n1 + 2 * n2 + n3
    return {"total_variants": total, "snp_count": total.copy(), "indel_count": np.zeros_like(total),
            "heterozygous_reference_minor_alleles": n1 + n3, "homozygous_minor_alleles": n2,
            "heterozygous_minor_alleles": np.zeros_like(total), "homozygous_reference_alleles": np.zeros_like(total)}


def wrights_fis(summary: dict, location_of_genome: np.ndarray) -> np.ndarray:
    """HeteroHomoZygous::UpdateSampleLocation (:362-412): F_IS = (H_exp - H_obs) / H_exp, H = heterozygous entries / total
    entries, H_exp from the aggregate of the genome's location group, H_obs from the genome itself; 0 where either total is 0."""
    het = (summary["heterozygous_minor_alleles"] + summary["heterozygous_reference_minor_alleles"]).astype(np.float64)
    tot = summary["total_variants"].astype(np.float64)
    loc = np.asarray(location_of_genome)
    groups, inv = np.unique(loc, return_inverse=True)
    g_het = np.bincount(inv, weights=het, minlength=len(groups))
    g_tot = np.bincount(inv, weights=tot, minlength=len(groups))
    out = np.zeros(tot.shape[0], dtype=np.float64)
    ok = (g_tot[inv] > 0) & (tot > 0)
    h_exp = np.where(ok, g_het[inv] / np.where(g_tot[inv] > 0, g_tot[inv], 1.0), 0.0)
    h_obs = np.where(ok, het / np.where(tot > 0, tot, 1.0), 0.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        out = np.where(ok, (h_exp - h_obs) / h_exp, 0.0)
    return out
