"""KGLFLAT1 / KGLTENS1 containers (see oracle/flat_io.h for the byte layout).

KGLFLAT1 is the flattened population the host flattener hands to the C-ABI: locus offsets, the float
allele-frequency vectors per super-population, the super-population of every genome and the 2-bit packed
loci-major genotype matrix (include/kgl_b200.h). KGLTENS1 is a bag of named little-endian arrays.
"""
from __future__ import annotations

import json
import struct
from dataclasses import dataclass

import numpy as np

SUPER_POPULATIONS = ("AFR", "AMR", "EAS", "EUR", "SAS", "ALL")  # kgl_variant_db_freq.h:55-60,66-71
FLAG_UNPHASED = 1

_DTYPES = {"f64": np.float64, "u64": np.uint64, "u32": np.uint32, "f32": np.float32, "u8": np.uint8}


def row_bytes_for(n_genomes: int) -> int:
    """Bytes per locus row: one 128-bit unit ({u64 lo, u64 hi}) per 64 genomes."""
    return 16 * ((int(n_genomes) + 63) // 64)


@dataclass
class FlatPopulation:
    offsets: np.ndarray      # uint32 [L], strictly increasing contig offsets
    af: np.ndarray           # float32 [6, L], NaN = no frequency for that super-population
    superpop: np.ndarray     # uint8 [N], index into SUPER_POPULATIONS
    packed: np.ndarray       # uint8 [L, row_bytes]
    n_genomes: int
    unphased: bool = False
    # Multi-allelic loci (include/kgl_b200.h, kgl_b200_upload_multi_allelic). At the listed rows `af` is NaN for every
    # super-population and `packed` holds only 0 (hom-ref) and 3 (see multi_cells).
    multi_rows: np.ndarray | None = None     # uint32 [M] rows of the locus table, ascending
    multi_af: np.ndarray | None = None       # float32 [6, M, 3] per allele slot, NaN = none
    multi_cells: np.ndarray | None = None    # uint8 [M, N]: 0 hom-ref; low nibble first variant's slot + 1 (4: not in the list), high
                                             # nibble the second's (0: none); 0xFF: more than two variants

    @property
    def n_multi(self) -> int:
        return 0 if self.multi_rows is None else int(self.multi_rows.shape[0])

    @property
    def n_loci(self) -> int:
        return int(self.offsets.shape[0])

    @property
    def row_bytes(self) -> int:
        return int(self.packed.shape[1])

    def codes(self) -> np.ndarray:
        """Unpack to uint8 [L, N] genotype codes (0 hom-ref, 1 het, 2 hom-alt, 3 dropped). Small inputs only."""
        return unpack_codes(self.packed, self.n_genomes)

    def write(self, path: str) -> None:
        hdr = struct.pack("<8s5I9I", b"KGLFLAT1", self.n_genomes, self.n_loci, self.af.shape[0], self.row_bytes,
                          FLAG_UNPHASED if self.unphased else 0, self.n_multi, *([0] * 8))
        assert len(hdr) == 64
        with open(path, "wb") as f:
            f.write(hdr)
            f.write(np.ascontiguousarray(self.offsets, dtype="<u4").tobytes())
            f.write(np.ascontiguousarray(self.af, dtype="<f4").tobytes())
            f.write(np.ascontiguousarray(self.superpop, dtype=np.uint8).tobytes())
            f.write(np.ascontiguousarray(self.packed, dtype=np.uint8).tobytes())
            if self.n_multi:
                f.write(np.ascontiguousarray(self.multi_rows, dtype="<u4").tobytes())
                f.write(np.ascontiguousarray(self.multi_af, dtype="<f4").tobytes())
                f.write(np.ascontiguousarray(self.multi_cells, dtype=np.uint8).tobytes())

    @staticmethod
    def read(path: str) -> "FlatPopulation":
        with open(path, "rb") as f:
            hdr = f.read(64)
            magic, n, l, npop, rb, flags, n_multi = struct.unpack("<8s6I", hdr[:32])
            if magic != b"KGLFLAT1":
                raise ValueError(f"{path}: bad magic")
            offsets = np.frombuffer(f.read(4 * l), dtype="<u4").copy()
            af = np.frombuffer(f.read(4 * npop * l), dtype="<f4").reshape(npop, l).copy()
            superpop = np.frombuffer(f.read(n), dtype=np.uint8).copy()
            packed = np.frombuffer(f.read(l * rb), dtype=np.uint8).reshape(l, rb).copy()
            pop = FlatPopulation(offsets, af, superpop, packed, n, bool(flags & FLAG_UNPHASED))
            if n_multi:
                pop.multi_rows = np.frombuffer(f.read(4 * n_multi), dtype="<u4").copy()
                pop.multi_af = np.frombuffer(f.read(4 * npop * n_multi * 3), dtype="<f4").reshape(npop, n_multi, 3).copy()
                pop.multi_cells = np.frombuffer(f.read(n_multi * n), dtype=np.uint8).reshape(n_multi, n).copy()
        return pop


def pack_codes(codes: np.ndarray) -> np.ndarray:
    """uint8 [L, N] codes -> packed uint8 [L, row_bytes] (unit = u64 lo-plane then u64 hi-plane, little endian)."""
    codes = np.asarray(codes, dtype=np.uint8)
    n_loci, n = codes.shape
    units = (n + 63) // 64
    padded = np.zeros((n_loci, units * 64), dtype=np.uint8)
    padded[:, :n] = codes
    lo = np.packbits((padded & 1).reshape(n_loci, units, 64), axis=2, bitorder="little")        # [L, units, 8]
    hi = np.packbits(((padded >> 1) & 1).reshape(n_loci, units, 64), axis=2, bitorder="little")
    return np.ascontiguousarray(np.concatenate([lo, hi], axis=2).reshape(n_loci, units * 16))


def unpack_codes(packed: np.ndarray, n_genomes: int) -> np.ndarray:
    packed = np.asarray(packed, dtype=np.uint8)
    n_loci, rb = packed.shape
    units = rb // 16
    u = packed.reshape(n_loci, units, 16)
    lo = np.unpackbits(u[:, :, :8], axis=2, bitorder="little").reshape(n_loci, units * 64)
    hi = np.unpackbits(u[:, :, 8:], axis=2, bitorder="little").reshape(n_loci, units * 64)
    return (lo | (hi << 1))[:, :n_genomes].astype(np.uint8)


def read_tensors(path: str) -> dict[str, np.ndarray]:
    with open(path, "rb") as f:
        blob = f.read()
    if blob[:8] != b"KGLTENS1":
        raise ValueError(f"{path}: bad magic")
    (jl,) = struct.unpack("<Q", blob[8:16])
    entries = json.loads(blob[16:16 + jl].decode())
    base = 16 + jl
    out = {}
    for e in entries:
        dt = np.dtype(_DTYPES[e["dtype"]]).newbyteorder("<")
        a = np.frombuffer(blob, dtype=dt, count=int(np.prod(e["shape"], dtype=np.int64)), offset=base + e["offset"])
        out[e["name"]] = a.reshape(e["shape"]).copy()
    return out


def write_tensors(path: str, arrays: dict[str, np.ndarray]) -> None:
    names = {np.dtype(v): k for k, v in _DTYPES.items()}
    entries, chunks, off = [], [], 0
    for name, a in arrays.items():
        a = np.ascontiguousarray(a)
        raw = a.astype(a.dtype.newbyteorder("<")).tobytes()
        pad = (-len(raw)) % 8
        entries.append({"name": name, "dtype": names[np.dtype(a.dtype.type)], "shape": list(a.shape), "offset": off, "nbytes": len(raw)})
        chunks.append(raw + b"\0" * pad)
        off += len(raw) + pad
    js = json.dumps(entries).encode()
    js += b" " * ((-len(js)) % 8)
    with open(path, "wb") as f:
        f.write(b"KGLTENS1" + struct.pack("<Q", len(js)) + js + b"".join(chunks))
