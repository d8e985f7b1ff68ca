// flat_io.h -- TEST INFRASTRUCTURE (oracle side). Not part of the product.
//
// Two tiny binary containers shared by the reference harness (ref_harness.cpp), the Python tests
// (tests/flatfile.py mirrors them with numpy) and bench.py's cpu_baseline leg.
//
//  * KGLFLAT1: a flattened population = what the product's flattener hands to the C-ABI
//      header (64 B) | u32 offsets[L] | f32 af[6][L] (NaN = no AF for that super-population)
//      | u8 superpop[N] (index into AFR,AMR,EAS,EUR,SAS,ALL) | u8 packed[L][row_bytes]
//    `packed` is the product's canonical loci-major layout (include/kgl_b200.h): per locus a row of
//    128-bit units, unit u = {u64 lo-plane, u64 hi-plane} of genomes 64u..64u+63, code = lo + 2*hi
//    (0 hom-ref, 1 het, 2 hom-alt, 3 dropped/missing).
//    Multi-allelic loci (optional trailing section; header.reserved[0] = M > 0): u32 multi_rows[M] | f32 multi_af[6][M][3] |
//    u8 multi_cells[M][N]. At the listed rows of the locus table the main `af` is NaN and `packed` holds only 0 (hom-ref) and 3
//    (see multi_cells). multi_af[k][m][a] = frequency of allele slot a (0..2) for super-population k, NaN = none. multi_cells:
//    0 = hom-ref; low nibble = first variant's allele slot + 1 (4 = an allele that is not in the locus' list), high nibble =
//    the second variant's (0 = there is none); 0xFF = more than two variants at the offset.
//  * KGLTENS1: named little-endian arrays: "KGLTENS1" | u64 json_len | json | raw data (8-byte aligned)
//      json = [{"name":..,"dtype":"f64|u64|u32|f32|u8","shape":[..],"offset":..}, ...]
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

namespace kglflat {

constexpr uint32_t FLAG_UNPHASED = 1u;  // Pf7-style population: every variant carries VariantPhase::UNPHASED

struct Header {
  char magic[8];
  uint32_t n_genomes;
  uint32_t n_loci;
  uint32_t n_superpop;
  uint32_t row_bytes;
  uint32_t flags;
  uint32_t reserved[9];
};
static_assert(sizeof(Header) == 64, "KGLFLAT1 header is 64 bytes");

struct Flat {
  Header hdr{};
  std::vector<uint32_t> offsets;
  std::vector<float> af;          // [6][L]
  std::vector<uint8_t> superpop;  // [N]
  std::vector<uint8_t> packed;    // [L][row_bytes]
  std::vector<uint32_t> multi_rows;   // [M]
  std::vector<float> multi_af;        // [6][M][3]
  std::vector<uint8_t> multi_cells;   // [M][N]

  uint32_t M() const { return hdr.reserved[0]; }
  float multiAf(uint32_t pop, uint32_t m, uint32_t a) const { return multi_af[(size_t(pop) * M() + m) * 3 + a]; }
  uint8_t multiCell(uint32_t m, uint32_t g) const { return multi_cells[size_t(m) * hdr.n_genomes + g]; }
  uint32_t N() const { return hdr.n_genomes; }
  uint32_t L() const { return hdr.n_loci; }
  float afAt(uint32_t pop, uint32_t locus) const { return af[size_t(pop) * hdr.n_loci + locus]; }
  // 2-bit code of genome g at locus l.
  unsigned code(uint32_t l, uint32_t g) const {
    const uint8_t* row = packed.data() + size_t(l) * hdr.row_bytes;
    const uint8_t* unit = row + size_t(g / 64) * 16;
    uint64_t lo, hi;
    std::memcpy(&lo, unit, 8);
    std::memcpy(&hi, unit + 8, 8);
    const unsigned bit = g % 64;
    return unsigned((lo >> bit) & 1u) | (unsigned((hi >> bit) & 1u) << 1);
  }
};

inline Flat readFlat(const std::string& path) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("cannot open " + path);
  Flat fl;
  auto rd = [&](void* p, size_t n) {
    if (n && std::fread(p, 1, n, f) != n) { std::fclose(f); throw std::runtime_error("short read " + path); }
  };
  rd(&fl.hdr, sizeof(Header));
  if (std::memcmp(fl.hdr.magic, "KGLFLAT1", 8) != 0) { std::fclose(f); throw std::runtime_error("bad magic " + path); }
  const size_t L = fl.hdr.n_loci, N = fl.hdr.n_genomes;
  fl.offsets.resize(L);
  fl.af.resize(size_t(fl.hdr.n_superpop) * L);
  fl.superpop.resize(N);
  fl.packed.resize(L * size_t(fl.hdr.row_bytes));
  rd(fl.offsets.data(), L * 4);
  rd(fl.af.data(), fl.af.size() * 4);
  rd(fl.superpop.data(), N);
  rd(fl.packed.data(), fl.packed.size());
  if (const size_t M = fl.M()) {
    fl.multi_rows.resize(M);
    fl.multi_af.resize(size_t(fl.hdr.n_superpop) * M * 3);
    fl.multi_cells.resize(M * N);
    rd(fl.multi_rows.data(), M * 4);
    rd(fl.multi_af.data(), fl.multi_af.size() * 4);
    rd(fl.multi_cells.data(), fl.multi_cells.size());
  }
  std::fclose(f);
  return fl;
}

inline void writeFlat(const std::string& path, const Flat& fl) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) throw std::runtime_error("cannot write " + path);
  Header h = fl.hdr;
  std::memcpy(h.magic, "KGLFLAT1", 8);
  std::fwrite(&h, sizeof(Header), 1, f);
  std::fwrite(fl.offsets.data(), 4, fl.offsets.size(), f);
  std::fwrite(fl.af.data(), 4, fl.af.size(), f);
  std::fwrite(fl.superpop.data(), 1, fl.superpop.size(), f);
  std::fwrite(fl.packed.data(), 1, fl.packed.size(), f);
  if (fl.M()) {
    std::fwrite(fl.multi_rows.data(), 4, fl.multi_rows.size(), f);
    std::fwrite(fl.multi_af.data(), 4, fl.multi_af.size(), f);
    std::fwrite(fl.multi_cells.data(), 1, fl.multi_cells.size(), f);
  }
  std::fclose(f);
}

class TensorWriter {
 public:
  void add(const std::string& name, const char* dtype, std::vector<size_t> shape, const void* data, size_t bytes) {
    Entry e{name, dtype, std::move(shape), blob_.size(), bytes};
    const auto* p = static_cast<const uint8_t*>(data);
    blob_.insert(blob_.end(), p, p + bytes);
    while (blob_.size() % 8) blob_.push_back(0);
    entries_.push_back(std::move(e));
  }
  void addF64(const std::string& n, std::vector<size_t> s, const std::vector<double>& v) { add(n, "f64", std::move(s), v.data(), v.size() * 8); }
  void addU64(const std::string& n, std::vector<size_t> s, const std::vector<uint64_t>& v) { add(n, "u64", std::move(s), v.data(), v.size() * 8); }
  void addU32(const std::string& n, std::vector<size_t> s, const std::vector<uint32_t>& v) { add(n, "u32", std::move(s), v.data(), v.size() * 4); }
  void write(const std::string& path) const {
    std::string json = "[";
    for (size_t i = 0; i < entries_.size(); ++i) {
      const auto& e = entries_[i];
      if (i) json += ",";
      json += "{\"name\":\"" + e.name + "\",\"dtype\":\"" + e.dtype + "\",\"shape\":[";
      for (size_t d = 0; d < e.shape.size(); ++d) { if (d) json += ","; json += std::to_string(e.shape[d]); }
      json += "],\"offset\":" + std::to_string(e.offset) + ",\"nbytes\":" + std::to_string(e.bytes) + "}";
    }
    json += "]";
    while (json.size() % 8) json += " ";
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot write " + path);
    const uint64_t jl = json.size();
    std::fwrite("KGLTENS1", 1, 8, f);
    std::fwrite(&jl, 8, 1, f);
    std::fwrite(json.data(), 1, json.size(), f);
    std::fwrite(blob_.data(), 1, blob_.size(), f);
    std::fclose(f);
  }
 private:
  struct Entry { std::string name, dtype; std::vector<size_t> shape; size_t offset, bytes; };
  std::vector<Entry> entries_;
  std::vector<uint8_t> blob_;
};

}  // namespace kglflat
