"""Times the phases of the end-to-end step of bench.py separately (development tool)."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kgl_gene_b200.capi import KglB200, RESULT_DTYPE
from kgl_gene_b200.flatfile import row_bytes_for
from kgl_gene_b200.synth import make_genomes, make_loci
import ctypes as C

n, l = 2504, 1_100_000
rb = row_bytes_for(n)
ctx = KglB200(0)
offsets, af = make_loci(l, 1)
superpop, f = make_genomes(n, 1)
ctx.upload_loci(af, offsets); ctx.set_genome_superpop(superpop); ctx.synth_genotypes(1, n, l, f, missing_rate=0.001)
h_packed = torch.empty((l, rb), dtype=torch.uint8, pin_memory=True)
ctx._check(ctx.lib.kgl_b200_download_genotypes(ctx.h, C.c_uint64(l * rb), C.c_void_p(h_packed.data_ptr())), "dl")
h_af = torch.from_numpy(af).pin_memory(); af_np = h_af.numpy()
h_lc = torch.empty((l, 4), dtype=torch.int32, pin_memory=True)
h_res = torch.empty((n * RESULT_DTYPE.itemsize,), dtype=torch.uint8, pin_memory=True)

def t(name, fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    print(f"{name:28s} {(time.perf_counter() - t0) / reps * 1e3:8.3f} ms")

t("upload_genotypes (704 MB)", lambda: ctx.upload_genotypes_ptr(h_packed.data_ptr(), n, l, rb))
t("upload_loci (26 MB + offsets)", lambda: ctx.upload_loci(af_np, offsets))
t("set_genome_superpop", lambda: ctx.set_genome_superpop(superpop))
t("select_loci", lambda: ctx.select_loci())
t("count_and_inbreed_into", lambda: ctx.count_and_inbreed_into(h_lc.data_ptr(), h_res.data_ptr()))
x = torch.empty(704_000_000, dtype=torch.uint8, device="cuda")
t("torch H2D copy 704 MB pinned", lambda: x.copy_(h_packed.view(-1)[:704_000_000], non_blocking=True))
