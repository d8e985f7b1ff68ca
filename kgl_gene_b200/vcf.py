"""VCF -> flat population without a PopulationDB in between: ctypes wrapper over
kgl_gene_b200/host/kgl_b200_vcf_ingest.cpp (libkgl_b200_host.so, plain C++17 + zlib, no GPU involved), and a writer that
renders a FlatPopulation as a 1000 Genomes style VCF (used by the tests and to hand synthetic populations to the
reference's own parser)."""
from __future__ import annotations

import ctypes as C
import gzip
import os

import numpy as np

from .flatfile import SUPER_POPULATIONS, FlatPopulation

HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(HERE, "libkgl_b200_host.so")
_AF_KEYS = ("AFR_AF", "AMR_AF", "EAS_AF", "EUR_AF", "SAS_AF", "AF")      # kgl_variant_db_freq.h:87-92 (Genome1000)


class VcfStats(C.Structure):
    _fields_ = [("records", C.c_uint64), ("kept", C.c_uint64), ("multi_allelic", C.c_uint64), ("skipped_too_many_alleles", C.c_uint64),
                ("skipped_non_snp", C.c_uint64),
                ("not_pass", C.c_uint64), ("malformed_genotypes", C.c_uint64), ("bytes", C.c_uint64), ("seconds", C.c_double)]


_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise RuntimeError(f"{HOST_LIB_PATH} is missing: build it with `python kgl_gene_b200/build.py`")
        lib = C.CDLL(HOST_LIB_PATH)
        for name, res in (("n_genomes", C.c_uint64), ("n_loci", C.c_uint64), ("row_bytes", C.c_uint64), ("packed", C.c_void_p),
                          ("af", C.c_void_p), ("offsets", C.c_void_p), ("contig", C.c_char_p), ("n_multi", C.c_uint64),
                          ("multi_rows", C.c_void_p), ("multi_af", C.c_void_p), ("multi_cells", C.c_void_p)):
            fn = getattr(lib, "kgl_b200_vcf_" + name)
            fn.restype, fn.argtypes = res, [C.c_void_p]
        lib.kgl_b200_vcf_genome_name.restype, lib.kgl_b200_vcf_genome_name.argtypes = C.c_char_p, [C.c_void_p, C.c_uint64]
        lib.kgl_b200_vcf_free.restype, lib.kgl_b200_vcf_free.argtypes = None, [C.c_void_p]
        lib.kgl_b200_vcf_get_stats.restype, lib.kgl_b200_vcf_get_stats.argtypes = None, [C.c_void_p, C.POINTER(VcfStats)]
        _lib = lib
    return _lib


def ingest_vcf(path: str, unphased: bool = False, n_threads: int = 0, superpop=None):
    """Returns (FlatPopulation, genome names, contig, stats dict). superpop: uint8[N] PED super-population indices
    (default: everybody in "ALL")."""
    lib = _load()
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    rc = lib.kgl_b200_vcf_ingest(path.encode(), C.c_int(int(unphased)), C.c_int(int(n_threads)), C.byref(h), err, C.c_size_t(512))
    if rc != 0:
        raise RuntimeError("VCF ingest failed: " + err.value.decode(errors="replace"))
    try:
        n, l, rb = lib.kgl_b200_vcf_n_genomes(h), lib.kgl_b200_vcf_n_loci(h), lib.kgl_b200_vcf_row_bytes(h)
        packed = np.ctypeslib.as_array(C.cast(lib.kgl_b200_vcf_packed(h), C.POINTER(C.c_uint8)), shape=(l, rb)).copy() if l else np.zeros((0, rb), np.uint8)
        af = np.ctypeslib.as_array(C.cast(lib.kgl_b200_vcf_af(h), C.POINTER(C.c_float)), shape=(6, l)).copy() if l else np.zeros((6, 0), np.float32)
        offsets = np.ctypeslib.as_array(C.cast(lib.kgl_b200_vcf_offsets(h), C.POINTER(C.c_uint32)), shape=(l,)).copy() if l else np.zeros(0, np.uint32)
        m = lib.kgl_b200_vcf_n_multi(h)
        multi = None
        if m:
            multi = (np.ctypeslib.as_array(C.cast(lib.kgl_b200_vcf_multi_rows(h), C.POINTER(C.c_uint32)), shape=(m,)).copy(),
                     np.ctypeslib.as_array(C.cast(lib.kgl_b200_vcf_multi_af(h), C.POINTER(C.c_float)), shape=(6, m, 3)).copy(),
                     np.ctypeslib.as_array(C.cast(lib.kgl_b200_vcf_multi_cells(h), C.POINTER(C.c_uint8)), shape=(m, n)).copy())
        names = [lib.kgl_b200_vcf_genome_name(h, i).decode() for i in range(n)]
        contig = lib.kgl_b200_vcf_contig(h).decode()
        st = VcfStats()
        lib.kgl_b200_vcf_get_stats(h, C.byref(st))
        stats = {k: getattr(st, k) for k, _ in VcfStats._fields_}
    finally:
        lib.kgl_b200_vcf_free(h)
    sp = np.full(n, SUPER_POPULATIONS.index("ALL"), dtype=np.uint8) if superpop is None else np.ascontiguousarray(superpop, dtype=np.uint8)
    pop = FlatPopulation(offsets, af, sp, packed, int(n), bool(unphased))
    if multi is not None:
        pop.multi_rows, pop.multi_af, pop.multi_cells = multi
    return pop, names, contig, stats


def write_vcf(pop: FlatPopulation, path: str, contig: str = "22", names=None, missing_as: str = ".", extra_lines=()):
    """Renders `pop` as VCF text (gzip when path ends in .gz). Codes 0/1/2 -> 0|0, 0|1 (alternating with 1|0), 1|1;
    code 3 -> `missing_as` on both alleles (the 1000G parser maps "." to the reference allele). With pop.unphased the
    separator is '/' (Pf7). extra_lines: [(after_row_index, text_line)] raw records to splice in (tests)."""
    codes = pop.codes()
    n = pop.n_genomes
    names = names or [f"G{i:05d}" for i in range(n)]
    sep = "/" if pop.unphased else "|"
    gts = {0: f"0{sep}0", 1: f"0{sep}1", 2: f"1{sep}1", 3: f"{missing_as}{sep}{missing_as}"}
    alt_het = f"1{sep}0"
    extra = {}
    for after, text in extra_lines:
        extra.setdefault(after, []).append(text)
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "wt") as f:
        f.write("##fileformat=VCFv4.1\n")
        for k in _AF_KEYS:
            f.write(f'##INFO=<ID={k},Number=A,Type=Float,Description="allele frequency">\n')
        f.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(names) + "\n")
        for text in extra.get(-1, []):
            f.write(text + "\n")
        multi_of = {int(r): m for m, r in enumerate(pop.multi_rows)} if pop.n_multi else {}
        alt_bases = "GCT"
        for l in range(pop.n_loci):
            if l in multi_of:
                # a multi-allelic locus: ALT = its allele slots, Number=A frequencies per slot ("." = none), genotypes from the
                # side cells (phase A = first variant, phase B = second; slot 4 and 0xFF cannot be written: not generated here)
                m = multi_of[l]
                n_slots = max(a + 1 for a in range(3) if not np.all(np.isnan(pop.multi_af[:, m, a])))
                info = ";".join(f"{k}=" + ",".join("." if np.isnan(pop.multi_af[i, m, a]) else repr(float(pop.multi_af[i, m, a])) for a in range(n_slots))
                                for i, k in enumerate(_AF_KEYS) if not np.all(np.isnan(pop.multi_af[i, m, :n_slots]))) or "."
                cells = []
                for g, c in enumerate(pop.multi_cells[m]):
                    a, b = int(c) & 15, int(c) >> 4
                    if pop.unphased:
                        cells.append(f"{a}/{b}" if b else f"0/{a}" if a else "0/0")
                    else:
                        cells.append(f"{a}|{b}" if b else ((f"{a}|0" if (g + l) % 2 else f"0|{a}") if a else "0|0"))
                f.write(f"{contig}\t{int(pop.offsets[l]) + 1}\t.\tA\t{','.join(alt_bases[:n_slots])}\t100\tPASS\t{info}\tGT\t" + "\t".join(cells) + "\n")
                for text in extra.get(l, []):
                    f.write(text + "\n")
                continue
            info = ";".join(f"{k}={float(pop.af[i, l])!r}" for i, k in enumerate(_AF_KEYS) if not np.isnan(pop.af[i, l])) or "."
            row = codes[l]
            cells = [(alt_het if (c == 1 and (g + l) % 2) else gts[int(c)]) for g, c in enumerate(row)]
            f.write(f"{contig}\t{int(pop.offsets[l]) + 1}\t.\tA\tG\t100\tPASS\t{info}\tGT\t" + "\t".join(cells) + "\n")
            for text in extra.get(l, []):
                f.write(text + "\n")
