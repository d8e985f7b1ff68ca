"""BASELINE config 5 on N GPUs (development tool; launch with torchrun, one rank per GPU): 100k genomes x 1M SNPs pairwise
kinship, the int8 Gram matrix on tcgen05 and the popcount IBS tiles, tile-sharded (the packed matrix is replicated,
every rank keeps the tiles it computed, no collective in the data path). Checked through size-independent invariants:
the Gram diagonal (n1 + 4 n2 per genome, summed over the ranks that own the diagonal tiles) and IBS0 + IBS1 + IBS2 = valid
on the last slab of every rank. Prints one JSON line on rank 0."""
import json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kgl_gene_b200.capi import KglB200
from kgl_gene_b200.shards import _RawCudaArray, block_tile_coords, tiles_of_rank
from kgl_gene_b200.synth import make_genomes, make_loci

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
l = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
mode = sys.argv[3] if len(sys.argv) > 3 else "both"
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
SEED = 5
offsets, af = make_loci(l, SEED)
superpop, f = make_genomes(n, SEED)
pair_loci = n * (n + 1) / 2 * l
out = {"workload": f"{n} genomes x {l} SNPs pairwise kinship, tiles dealt to {world} GPU(s), matrix replicated", "n_gpus": world}


def fresh():
    ctx = KglB200(local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.upload_loci(af, offsets)
    ctx.set_genome_superpop(superpop)
    ctx.synth_genotypes(SEED, n, l, f, missing_rate=0.001)
    torch.cuda.synchronize()
    return ctx, stream


def timed(stream, fn):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    fn()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


if mode in ("both", "gram"):
    ctx, stream = fresh()
    ctx.enqueue_gram_tiles(rank, world)                      # warm-up: sample-major copy, code matrix
    ms = timed(stream, lambda: ctx.enqueue_gram_tiles(rank, world))
    k_ms = ctx.last_gram_kernel_ms()
    ptr, count, ld = ctx.gram_buffer()
    g = torch.as_tensor(_RawCudaArray(ptr, count, "<i4"), device=dev).view(ld, ld)
    diag = torch.diagonal(g)[:n].to(torch.int64).clone()
    if world > 1:
        dist.all_reduce(diag, op=dist.ReduceOp.SUM)          # every diagonal cell is owned by exactly one rank
    _, gc = ctx.allele_count(want_loci=False, want_genomes=True)
    ok = bool(np.array_equal(diag.cpu().numpy(), (gc[:, 1] + 4 * gc[:, 2]).astype(np.int64)))
    out["gram_i8"] = {"ms": ms, "kernel_ms_rank0": k_ms, "pair_loci_per_s": pair_loci / (ms * 1e-3), "diag_matches_allele_counts": ok,
                      "free_hbm_gb": torch.cuda.mem_get_info()[0] / 1e9}
    assert ok
    del g
    ctx.close()
    torch.cuda.empty_cache()

if mode in ("both", "ibs", "ibs_popcount"):
    ctx, stream = fresh()
    side, n_up = ctx.ibs_tile_grid()
    SLAB = 8192
    tensor = mode != "ibs_popcount"
    ctx.set_ibs_tensor_cores(tensor)
    # tensor-core form: whole 256 x 256 blocks of tiles per rank (explicit tile lists); popcount form: single tiles, strided
    coords = block_tile_coords(n, rank, world) if tensor else None
    mine = int(coords.shape[0]) if tensor else tiles_of_rank(n_up, rank, world)

    def sweep():
        done = 0
        while done < mine:
            k = min(SLAB, mine - done)
            if tensor:
                ctx.enqueue_ibs_tile_list(coords[done:done + k])
            else:
                ctx.enqueue_ibs_tiles(rank + done * world, world, k)
            done += k
        return k

    if tensor:
        ctx.enqueue_ibs_tile_list(coords[: min(SLAB, mine)])   # warm-up: sample-major planes, masked planes, code matrix, class counts
    else:
        ctx.enqueue_ibs_tiles(rank, world, min(SLAB, mine))
    torch.cuda.synchronize()
    last = [0]
    ms = timed(stream, lambda: last.__setitem__(0, sweep()))
    ptr, cnt = ctx.ibs_tiles_buffer()
    t = torch.as_tensor(_RawCudaArray(ptr, cnt, "<u4"), device=dev).view(-1, 64, 64, 4)[: last[0]].to(torch.int64)
    ok = bool(torch.equal(t[..., :3].sum(-1), t[..., 3])) and bool((t >= 0).all()) and bool((t[..., 3] <= l).all())
    okt = torch.tensor([int(ok)], device=dev)
    if world > 1:
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    out["ibs_tensor" if ctx.ibs_used_tensor_cores() else "ibs_popcount"] = {
        "ms": ms, "pair_loci_per_s": pair_loci / (ms * 1e-3), "tiles": int(n_up), "tiles_this_rank": int(mine),
        "ibs_classes_sum_to_valid": bool(okt.item()), "free_hbm_gb": torch.cuda.mem_get_info()[0] / 1e9}
    assert ok
    ctx.close()

if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
