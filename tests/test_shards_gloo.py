"""Host-side partitioning logic (kgl_gene_b200/shards.py) on two gloo ranks, CPU only.

The kernels cannot run here, so each rank's device result is stood in for by the CPU oracle on that rank's shard; what is
tested is the product's host logic around it: the locus-shard boundaries, the additivity contract of the all-reduced
per-genome partial sums, the dealing of upper-triangle tiles to ranks and the gather / assembly of the IBS matrix.
"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_py as O
from conftest import results_matrix
from kgl_gene_b200 import shards
from kgl_gene_b200.flatfile import FlatPopulation
from kgl_gene_b200.synth import make_population

WORLD = 2


def _cut_tiles(full, n_genomes, rank, world):
    """What kgl_b200_run_ibs_tiles(first=rank, stride=world) returns, cut out of a full matrix."""
    side = shards.tile_side(n_genomes)
    padded = np.zeros((side * 64, side * 64, 4), dtype=np.uint32)
    padded[:n_genomes, :n_genomes] = full
    coords = shards.upper_tile_coords(n_genomes)[rank::world]
    return np.stack([padded[ti * 64:(ti + 1) * 64, tj * 64:(tj + 1) * 64] for ti, tj in coords]) if len(coords) else np.zeros((0, 64, 64, 4), np.uint32)


def _worker(rank, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        pop, _ = make_population(150, 5000, seed=5, missing_rate=0.01)
        sel = O.select_all_pops(pop, spacing=0)

        # ---- locus-sharded Simple estimator: partial sums are additive over shards ----
        l0, l1 = shards.locus_shard(pop.n_loci, rank, WORLD)
        shard = FlatPopulation(pop.offsets[l0:l1], pop.af[:, l0:l1], pop.superpop, pop.packed[l0:l1], pop.n_genomes, pop.unphased)
        counts, freqs = results_matrix(O.inbreed(shard, sel[:, l0:l1], "Simple"))
        part = torch.from_numpy(np.concatenate([counts[:, :4].astype(np.float64), freqs], axis=1))
        dist.all_reduce(part, op=dist.ReduceOp.SUM)
        part = part.numpy()
        want = O.inbreed(pop, sel, "Simple")
        wc, wf = results_matrix(want)
        assert np.array_equal(part[:, :4].astype(np.uint64), wc[:, :4])
        assert np.allclose(part[:, 4:], wf, rtol=1e-12, atol=0)
        n = part[:, :4].sum(axis=1)
        o_hom, e_hom = part[:, 0] + part[:, 2], part[:, 4] + part[:, 6]
        coeff = (o_hom - e_hom) / (n - e_hom)                       # processSimple, kga_analysis_inbreed_calc.cpp:335-344
        assert np.allclose(coeff, want["inbred_allele_sum"], rtol=1e-8, atol=1e-12)   # cancellation in O_hom - E_hom

        # ---- IBS tiles dealt round-robin, gathered on rank 0 ----
        full = O.ibs(pop)
        mine = _cut_tiles(full, pop.n_genomes, rank, WORLD)
        assert mine.shape[0] == shards.tiles_of_rank(shards.n_upper_tiles(pop.n_genomes), rank, WORLD)
        got = shards.gather_ibs(pop.n_genomes, mine)
        if rank == 0:
            assert np.array_equal(got, full)
        else:
            assert got is None
        open(os.path.join(tmp, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


def test_two_rank_partitioning_gloo(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(WORLD))


@pytest.mark.parametrize("n_loci,world", [(1, 1), (255, 2), (256, 2), (1_100_000, 8), (1000, 3), (80_000_000, 8)])
def test_locus_shards_cover_exactly(n_loci, world):
    edges = [shards.locus_shard(n_loci, r, world) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == n_loci
    for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
        assert a1 == b0 and a0 <= a1
    assert all(b % 256 == 0 for b, _ in edges)


@pytest.mark.parametrize("n_genomes,world", [(1, 1), (64, 2), (65, 2), (150, 3), (2504, 8), (500, 8)])
def test_tile_dealing_is_a_partition(n_genomes, world):
    n_up = shards.n_upper_tiles(n_genomes)
    counts = [shards.tiles_of_rank(n_up, r, world) for r in range(world)]
    assert sum(counts) == n_up and max(counts) - min(counts) <= 1
    coords = shards.upper_tile_coords(n_genomes)
    assert coords.shape == (n_up, 2) and np.all(coords[:, 0] <= coords[:, 1])
    # assemble() inverts the dealing
    rng = np.random.default_rng(1)
    sym = rng.integers(0, 1000, size=(n_genomes, n_genomes, 4), dtype=np.uint32)
    sym = np.triu(sym.transpose(2, 0, 1)).transpose(1, 2, 0)
    sym = sym + np.triu(sym.transpose(2, 0, 1), 1).transpose(2, 1, 0)
    blocks = [_cut_tiles(sym, n_genomes, r, world) for r in range(world)]
    assert np.array_equal(shards.assemble_ibs(n_genomes, blocks), sym)


@pytest.mark.parametrize("n_genomes,world", [(1, 1), (64, 2), (257, 2), (700, 3), (2504, 8), (500, 8)])
def test_block_dealing_is_a_partition(n_genomes, world):
    """Whole 256 x 256 blocks of 64 x 64 tiles per rank (the unit of the tensor-core IBS form): every upper-triangle tile exactly
    once, and assemble_ibs_coords inverts the dealing."""
    n_up = shards.n_upper_tiles(n_genomes)
    coords = [shards.block_tile_coords(n_genomes, r, world) for r in range(world)]
    allc = np.concatenate([c for c in coords if c.size]) if n_up else np.zeros((0, 2), np.uint32)
    assert allc.shape[0] == n_up and np.all(allc[:, 0] <= allc[:, 1])
    assert len({(int(a), int(b)) for a, b in allc}) == n_up
    rng = np.random.default_rng(2)
    sym = rng.integers(0, 1000, size=(n_genomes, n_genomes, 4), dtype=np.uint32)
    sym = np.triu(sym.transpose(2, 0, 1)).transpose(1, 2, 0)
    sym = sym + np.triu(sym.transpose(2, 0, 1), 1).transpose(2, 1, 0)
    tiles = []
    for c in coords:
        t = np.zeros((c.shape[0], 64, 64, 4), dtype=np.uint32)
        for i, (ti, tj) in enumerate(c.tolist()):
            a0, b0 = ti * 64, tj * 64
            a1, b1 = min(n_genomes, a0 + 64), min(n_genomes, b0 + 64)
            t[i, : a1 - a0, : b1 - b0] = sym[a0:a1, b0:b1]
        tiles.append(t)
    assert np.array_equal(shards.assemble_ibs_coords(n_genomes, coords, tiles), sym)
