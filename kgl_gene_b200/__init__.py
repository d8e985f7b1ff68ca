"""kgl_gene_b200 -- B200-native population-genotype hot path for KGL_Gene."""
__all__ = ["flatfile", "synth"]
