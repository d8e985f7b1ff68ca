// kgl_b200_vcf_ingest.cpp -- see kgl_b200_vcf_ingest.h. Standalone C++17 (zlib, threads); no reference headers needed.
//
// The file is streamed: blocks of text are inflated, cut at line ends and parsed by a pool of threads (one row of bits per
// line, written by the thread that parsed it); the main thread then walks the parsed lines in file order, groups the records of
// one POS and appends the group's row to the growing outputs. Memory = the outputs + one block of parsed lines.
#include "kgl_b200_vcf_ingest.h"

#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <thread>
#include <vector>

namespace {

constexpr int kPops = 6;
constexpr int kSlots = 3;                                        // a SNP has at most three alternate alleles
const char* const kAfKeys[kPops] = {"AFR_AF", "AMR_AF", "EAS_AF", "EUR_AF", "SAS_AF", "AF"};
const float kNone = std::numeric_limits<float>::quiet_NaN();

}  // namespace

struct kgl_b200_vcf {
  std::vector<std::string> genomes;
  std::string contig;
  uint64_t n_genomes = 0, row_bytes = 0;
  std::vector<uint8_t> packed;          // [n_loci][row_bytes]
  std::vector<float> af_rows;           // [n_loci][6] while reading; transposed into af at the end
  std::vector<float> af;                // [6][n_loci]
  std::vector<uint32_t> offsets;
  std::vector<uint32_t> multi_rows;
  std::vector<float> multi_af_rows;     // [n_multi][6][3] while reading
  std::vector<float> multi_af;          // [6][n_multi][3]
  std::vector<uint8_t> multi_cells;     // [n_multi][n_genomes]
  kgl_b200_vcf_stats stats{};
};

namespace {

// One ALT allele of a record.
struct Alt {
  char base = 0;                        // the alternate base of a SNP as DNA5 reads it (A C G T N), 0 otherwise (indel, abstract "<...>", ".")
  float af[kPops];
};

// One parsed data line.
struct Line {
  uint32_t offset = 0;
  char ref = 0;                         // REF base when it is a single base, else 0
  bool pass = true, usable = false;     // usable: at least nine columns
  uint32_t malformed = 0;
  std::vector<Alt> alts;
  bool simple = false;                  // one ALT: `bits` holds the copies of that allele as 2-bit codes (0, 1, 2)
  std::vector<uint8_t> bits;            // row_bytes (simple lines)
  std::vector<uint8_t> gt;              // [n_genomes][2] allele index of phase A / B, 0 = reference (lines with several ALTs)
};

// One allele token -> alt index (0 = reference); returns false when the token is not a number.
// DNA5::convertChar (kgl_alphabet_dna5.cpp:55-90): case folded, U read as T, anything that is not a nucleotide becomes N.
inline char dna5_base(char ch) {
  switch (ch & ~0x20) {
    case 'A': return 'A';
    case 'C': return 'C';
    case 'G': return 'G';
    case 'T': case 'U': return 'T';
    default: return ch == 0 ? 0 : 'N';
  }
}

inline bool allele_index(const char* b, const char* e, uint32_t& out) {
  if (b == e) return false;
  if (e - b == 1 && (*b == '.' || *b == '-')) { out = 0; return true; }
  if (*b == '<') { out = 0; return true; }                       // abstract alt: counted, treated as reference (:222-232)
  uint32_t v = 0;
  for (const char* p = b; p < e; ++p) {
    if (*p < '0' || *p > '9') return false;
    v = v * 10 + (uint32_t)(*p - '0');
    if (v > 1000000) return false;
  }
  out = v;
  return true;
}

// INFO: the Number=A frequency fields, one value per ALT (kgl_variant_db_freq.cpp:89-92: indexed by altVariantIndex).
void parse_info(const char* b, const char* e, std::vector<Alt>& alts) {
  for (Alt& a : alts) for (int k = 0; k < kPops; ++k) a.af[k] = kNone;
  const char* p = b;
  while (p < e) {
    const char* semi = static_cast<const char*>(std::memchr(p, ';', (size_t)(e - p)));
    const char* fe = semi ? semi : e;
    const char* eq = static_cast<const char*>(std::memchr(p, '=', (size_t)(fe - p)));
    if (eq) {
      const size_t klen = (size_t)(eq - p);
      for (int k = 0; k < kPops; ++k) {
        if (std::strlen(kAfKeys[k]) == klen && std::memcmp(p, kAfKeys[k], klen) == 0) {
          const char* v = eq + 1;
          for (size_t a = 0; a < alts.size() && v <= fe; ++a) {
            const char* comma = static_cast<const char*>(std::memchr(v, ',', (size_t)(fe - v)));
            const char* ve = comma ? comma : fe;
            // VCFInfoParser::convertToFloat (vcf_parse_info.cpp:208-270): std::stof; "." / "NaN" / anything that is not a
            // number -> MISSING_VALUE_FLOAT_ = the lowest float, which IS a value (the allele is in the frequency list, its
            // frequency clamps to 0, freq.cpp:47); out of range -> the smallest / largest float
            const std::string val(v, (size_t)(ve - v));
            char* endp = nullptr;
            errno = 0;
            float f = std::strtof(val.c_str(), &endp);
            const bool is_nan_text = val.size() == 3 && (val[0] | 0x20) == 'n' && (val[1] | 0x20) == 'a' && (val[2] | 0x20) == 'n';
            if (endp == val.c_str() || is_nan_text) f = std::numeric_limits<float>::lowest();
            else if (errno == ERANGE) f = (std::fabs(f) < 1.0f) ? std::numeric_limits<float>::min() : std::numeric_limits<float>::max();
            alts[a].af[k] = f;
            if (!comma) break;
            v = comma + 1;
          }
        }
      }
    }
    p = fe + 1;
  }
}

inline void set_code(std::vector<uint8_t>& bits, uint64_t g, unsigned code) {
  if (code & 1u) bits[(g >> 6) * 16 + ((g & 63) >> 3)] |= (uint8_t)(1u << (g & 7));
  if (code & 2u) bits[(g >> 6) * 16 + 8 + ((g & 63) >> 3)] |= (uint8_t)(1u << (g & 7));
}
inline unsigned get_code(const std::vector<uint8_t>& bits, uint64_t g) {
  const unsigned lo = (bits[(g >> 6) * 16 + ((g & 63) >> 3)] >> (g & 7)) & 1u;
  const unsigned hi = (bits[(g >> 6) * 16 + 8 + ((g & 63) >> 3)] >> (g & 7)) & 1u;
  return lo | (hi << 1);
}

// Parses one data line. Fields: CHROM POS ID REF ALT QUAL FILTER INFO FORMAT samples...
void parse_line(const char* b, const char* e, uint64_t n_genomes, uint64_t row_bytes, bool unphased, Line& ln, std::string* contig) {
  const char* f[10];
  int nf = 0;
  const char* p = b;
  f[nf++] = p;
  while (nf < 10 && p < e) {
    const char* t = static_cast<const char*>(std::memchr(p, '\t', (size_t)(e - p)));
    if (!t) break;
    p = t + 1;
    f[nf++] = p;
  }
  ln = Line();
  if (nf < 9) return;
  ln.usable = true;
  if (contig && contig->empty()) contig->assign(f[0], (size_t)(f[1] - 1 - f[0]));
  ln.offset = (uint32_t)std::strtoul(std::string(f[1], (size_t)(f[2] - 1 - f[1])).c_str(), nullptr, 10);
  ln.offset = ln.offset > 0 ? ln.offset - 1 : 0;
  const size_t ref_len = (size_t)(f[4] - 1 - f[3]);
  ln.ref = ref_len == 1 ? f[3][0] : 0;
  {
    const char* a = f[4];
    const char* alt_end = f[5] - 1;
    while (a <= alt_end) {                                       // Utility::charTokenizer(alt, ',') (1000_impl.cpp:87)
      const char* comma = static_cast<const char*>(std::memchr(a, ',', (size_t)(alt_end - a)));
      const char* ae = comma ? comma : alt_end;
      Alt alt;
      // a one-character ALT is a SNP of the variant DB whatever the character: StringDNA5 folds case, reads U as T and turns
      // everything else -- IUPAC codes, the spanning-deletion "*" -- into N (DNA5::convertChar, kgl_alphabet_dna5.cpp:55-90)
      const char alt_base = ae - a == 1 ? dna5_base(*a) : 0;
      const bool snp = ln.ref != 0 && alt_base != 0 && *a != '.' && alt_base != dna5_base(ln.ref);
      alt.base = snp ? alt_base : 0;
      ln.alts.push_back(alt);
      if (!comma) break;
      a = comma + 1;
    }
  }
  {
    // Utility::toupper(filter) == "PASS" (1000_impl.cpp:73)
    const size_t fl = (size_t)(f[7] - 1 - f[6]);
    ln.pass = fl == 4 && (f[6][0] | 0x20) == 'p' && (f[6][1] | 0x20) == 'a' && (f[6][2] | 0x20) == 's' && (f[6][3] | 0x20) == 's';
  }
  parse_info(f[7], f[8] - 1, ln.alts);
  if (!ln.pass) for (Alt& a : ln.alts) for (int k = 0; k < kPops; ++k) a.af[k] = kNone;
  const uint32_t n_alts = (uint32_t)ln.alts.size();
  ln.simple = n_alts == 1;
  if (ln.simple) ln.bits.assign(row_bytes, 0); else ln.gt.assign(n_genomes * 2, 0);
  if (nf < 10) return;                                         // no sample columns
  const char sep = unphased ? '/' : '|';
  p = f[9];
  for (uint64_t g = 0; g < n_genomes && p <= e; ++g) {
    // fast path: the three-character genotype "a|b" with single-digit / '.' alleles, ended by a tab or the line end
    if (ln.simple && e - p >= 3 && p[1] == sep && (p + 3 == e || p[3] == '\t')) {
      const char ca = p[0], cb = p[2];
      const bool da = ca == '0' || ca == '1', db = cb == '0' || cb == '1';
      const bool ma = ca == '.' || ca == '-', mb = cb == '.' || cb == '-';
      if ((da || ma) && (db || mb)) {
        unsigned code = (unsigned)(ca == '1') + (unsigned)(cb == '1');
        if (unphased && (ca == '.' || cb == '.')) code = 0;
        set_code(ln.bits, g, code);
        if (p + 3 == e) break;
        p += 4;
        continue;
      }
    }
    const char* t = static_cast<const char*>(std::memchr(p, '\t', (size_t)(e - p)));
    const char* ge = t ? t : e;
    const char* colon = static_cast<const char*>(std::memchr(p, ':', (size_t)(ge - p)));
    const char* gt_e = colon ? colon : ge;
    while (gt_e > p && (gt_e[-1] == ' ' || gt_e[-1] == '\r')) --gt_e;
    uint32_t a = 0, bb = 0;
    const char* s = static_cast<const char*>(std::memchr(p, sep, (size_t)(gt_e - p)));
    if (!s && unphased) s = static_cast<const char*>(std::memchr(p, '|', (size_t)(gt_e - p)));
    if (s) {
      const bool ok = allele_index(p, s, a) && allele_index(s + 1, gt_e, bb);
      const bool missing = (s - p == 1 && *p == '.') || (gt_e - s - 1 == 1 && s[1] == '.');
      if (!ok || a > n_alts || bb > n_alts) { ++ln.malformed; a = bb = 0; }     // beyond the ALT list / not a number: reference (:259-272)
      else if (unphased && missing) a = bb = 0;                                 // Pf7: genotype skipped (pf_impl.cpp:139-152)
    } else if (gt_e > p) {
      if (!(gt_e - p == 1 && (*p == '.' || *p == '-'))) ++ln.malformed;         // haploid GT on an autosome: reference (:193-203)
    }
    if (ln.simple) set_code(ln.bits, g, (unsigned)(a != 0) + (unsigned)(bb != 0));
    else { ln.gt[g * 2] = (uint8_t)a; ln.gt[g * 2 + 1] = (uint8_t)bb; }
    if (!t) break;
    p = t + 1;
  }
}

// ---- the records of one POS -> one row of the locus table ---------------------------------------------------------------------
// The variant DB would hold, at this offset, one variant per (non-reference allele, phase) of every genome, in the order
// (line, phase A, phase B) (1000_impl.cpp:118-139); the genome side of the analysis keeps the SNPs (freq.cpp:436). The distinct
// SNP alleles of the group, in order of appearance, are the locus' allele slots.
struct Group {
  std::vector<Line*> lines;
};

void emit_group(kgl_b200_vcf& v, const Group& grp, bool unphased) {
  kgl_b200_vcf_stats& st = v.stats;
  const uint64_t N = v.n_genomes;
  struct Slot { char base; float af[kPops]; };
  std::vector<Slot> slots;
  // (line, alt) -> slot + 1, 0 = not a SNP (the allele vanishes from the genome's SNP-filtered offset array)
  std::vector<std::vector<uint8_t>> slot_of(grp.lines.size());
  bool any_not_pass = false;
  for (size_t i = 0; i < grp.lines.size(); ++i) {
    const Line& ln = *grp.lines[i];
    slot_of[i].assign(ln.alts.size(), 0);
    any_not_pass = any_not_pass || !ln.pass;
    for (size_t a = 0; a < ln.alts.size(); ++a) {
      const Alt& alt = ln.alts[a];
      if (!alt.base) continue;
      size_t s = 0;
      for (; s < slots.size(); ++s) if (slots[s].base == alt.base) break;
      if (s == slots.size()) {
        Slot n; n.base = alt.base;
        for (int k = 0; k < kPops; ++k) n.af[k] = alt.af[k];
        slots.push_back(n);
      } else {
        // AlleleFreqVector keeps, per allele, the first variant that HAS a value for the population (freq.cpp:24-52)
        for (int k = 0; k < kPops; ++k) if (std::isnan(slots[s].af[k])) slots[s].af[k] = alt.af[k];
      }
      slot_of[i][a] = (uint8_t)(s + 1);
    }
    st.malformed_genotypes += ln.malformed;
  }
  if (slots.empty()) { st.skipped_non_snp += grp.lines.size(); return; }
  if (slots.size() > (size_t)kSlots) { st.skipped_too_many_alleles += grp.lines.size(); return; }
  if (any_not_pass) ++st.not_pass;
  const size_t row = v.offsets.size();
  v.offsets.push_back(grp.lines.front()->offset);
  v.packed.resize((row + 1) * v.row_bytes, 0);
  uint8_t* out_bits = v.packed.data() + row * v.row_bytes;
  v.af_rows.resize((row + 1) * kPops, kNone);
  ++st.kept;

  const bool one_simple_line = grp.lines.size() == 1 && grp.lines[0]->simple;
  if (slots.size() == 1 && one_simple_line) {                  // the common case: the parse thread's row as it is
    std::memcpy(out_bits, grp.lines[0]->bits.data(), v.row_bytes);
    for (int k = 0; k < kPops; ++k) v.af_rows[row * kPops + k] = slots[0].af[k];
    return;
  }
  // general case: every genome's SNP alleles at this offset, in the order the variant DB would hold them
  const bool multi = slots.size() > 1;
  std::vector<uint8_t> cells(multi ? N : 0, 0);
  std::vector<uint8_t> bits(v.row_bytes, 0);
  for (uint64_t g = 0; g < N; ++g) {
    uint8_t carried[4];
    int n = 0;
    for (size_t i = 0; i < grp.lines.size(); ++i) {
      const Line& ln = *grp.lines[i];
      uint32_t a = 0, b = 0;
      if (ln.simple) { const unsigned c = get_code(ln.bits, g); a = c >= 1; b = c == 2; }   // copies of ALT 1 (order is irrelevant)
      else { a = ln.gt[g * 2]; b = ln.gt[g * 2 + 1]; }
      for (uint32_t idx : {a, b}) {
        if (idx == 0) continue;
        const uint8_t s = slot_of[i][idx - 1];
        if (s == 0) continue;                                   // not a SNP: filtered from the genome's offset array
        if (n < 4) carried[n] = s;
        ++n;
      }
    }
    if (n == 0) continue;
    if (!multi) {
      set_code(bits, g, n == 1 ? 1u : (n == 2 ? 2u : 3u));     // more than two copies: dropped
    } else {
      uint8_t cell = 0xFF;
      if (n == 1) cell = carried[0];
      else if (n == 2) cell = (uint8_t)(carried[0] | (carried[1] << 4));
      cells[g] = cell;
      set_code(bits, g, 3u);
    }
  }
  std::memcpy(out_bits, bits.data(), v.row_bytes);
  if (!multi) {
    for (int k = 0; k < kPops; ++k) v.af_rows[row * kPops + k] = slots[0].af[k];
    return;
  }
  (void)unphased;
  ++st.multi_allelic;
  const size_t m = v.multi_rows.size();
  v.multi_rows.push_back((uint32_t)row);
  v.multi_af_rows.resize((m + 1) * kPops * kSlots, kNone);
  for (size_t s = 0; s < slots.size(); ++s)
    for (int k = 0; k < kPops; ++k) v.multi_af_rows[(m * kPops + k) * kSlots + s] = slots[s].af[k];
  v.multi_cells.insert(v.multi_cells.end(), cells.begin(), cells.end());
}

}  // namespace

extern "C" {

int kgl_b200_vcf_ingest(const char* path, int unphased, int n_threads, kgl_b200_vcf** out, char* err, size_t err_len) {
  auto fail = [&](const std::string& m) { if (err && err_len) { std::snprintf(err, err_len, "%s", m.c_str()); } return 1; };
  if (!path || !out) return fail("null argument");
  *out = nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  gzFile gz = gzopen(path, "rb");                              // transparently reads plain text too
  if (!gz) return fail(std::string("cannot open ") + path);
  gzbuffer(gz, 1 << 20);
  auto v = new kgl_b200_vcf();
  int nt = n_threads > 0 ? n_threads : (int)std::max(1u, std::thread::hardware_concurrency());

  size_t kBlock = 32u << 20;                                   // text per block
  if (const char* e = std::getenv("KGL_B200_VCF_BLOCK_BYTES")) {   // test hook: small blocks put the block boundaries everywhere
    const unsigned long long b = std::strtoull(e, nullptr, 10);
    if (b >= 64) kBlock = (size_t)b;
  }
  std::string block, carry;
  std::vector<char> buf(std::min<size_t>(8u << 20, kBlock));
  bool have_header = false, eof = false;
  std::vector<Line> parsed;
  std::vector<std::unique_ptr<Line>> held;                     // lines of a POS group that a block boundary cuts
  Group open;                                                  // the group being collected (lines of `held` / `parsed`)
  uint64_t total_bytes = 0;
  while (!eof) {
    // ---- next block: the carried-over partial line + fresh text, cut at the last line end ----
    block.swap(carry);
    carry.clear();
    do {                                                       // at least one read: a line longer than a block keeps growing
      const int n = gzread(gz, buf.data(), (unsigned)buf.size());
      if (n < 0) { int e = 0; const std::string m = std::string("read error: ") + gzerror(gz, &e); gzclose(gz); delete v; return fail(m); }
      if (n == 0) { eof = true; break; }
      block.append(buf.data(), (size_t)n);
      total_bytes += (size_t)n;
    } while (block.size() < kBlock);
    if (!eof) {
      const size_t last_nl = block.rfind('\n');
      if (last_nl == std::string::npos) { carry.swap(block); continue; }
      carry.assign(block, last_nl + 1, std::string::npos);
      block.resize(last_nl + 1);
    }
    // ---- header lines, then the line index of the block ----
    size_t pos = 0;
    while (!have_header && pos < block.size()) {
      size_t nl = block.find('\n', pos);
      if (nl == std::string::npos) nl = block.size();
      if (block[pos] != '#') break;
      if (block.compare(pos, 6, "#CHROM") == 0) {
        std::string line = block.substr(pos, nl - pos);
        if (!line.empty() && line.back() == '\r') line.pop_back();
        size_t p = 0; int col = 0;
        while (p <= line.size()) {
          size_t t = line.find('\t', p);
          if (t == std::string::npos) t = line.size();
          if (col >= 9) v->genomes.emplace_back(line.substr(p, t - p));
          ++col; p = t + 1;
        }
        have_header = true;
        v->n_genomes = v->genomes.size();
        v->row_bytes = 16 * ((v->n_genomes + 63) / 64);
      }
      pos = nl + 1;
    }
    if (!have_header) {
      if (eof || pos < block.size()) { gzclose(gz); delete v; return fail("no #CHROM header line"); }
      continue;
    }
    std::vector<std::pair<size_t, size_t>> lines;
    while (pos < block.size()) {
      size_t nl = block.find('\n', pos);
      if (nl == std::string::npos) nl = block.size();
      size_t end = nl;
      if (end > pos && block[end - 1] == '\r') --end;
      if (end > pos && block[pos] != '#') lines.emplace_back(pos, end);
      pos = nl + 1;
    }
    // ---- parse the block's lines on the pool ----
    const size_t n_lines = lines.size();
    parsed.assign(n_lines, Line());
    if (n_lines && v->contig.empty())
      parse_line(block.data() + lines[0].first, block.data() + lines[0].second, v->n_genomes, v->row_bytes, unphased != 0, parsed[0], &v->contig);
    std::atomic<size_t> next{0};
    auto work = [&]() {
      for (;;) {
        const size_t i0 = next.fetch_add(128);
        if (i0 >= n_lines) break;
        const size_t i1 = std::min(n_lines, i0 + 128);
        for (size_t i = i0; i < i1; ++i)
          parse_line(block.data() + lines[i].first, block.data() + lines[i].second, v->n_genomes, v->row_bytes, unphased != 0, parsed[i], nullptr);
      }
    };
    const int use = (int)std::min<size_t>((size_t)nt, std::max<size_t>(1, n_lines / 64));
    std::vector<std::thread> pool;
    for (int t = 1; t < use; ++t) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    // ---- group by POS in file order (the records of a POS need not be adjacent to be one offset of the variant DB, but a
    // sorted VCF keeps them together) and emit every completed group ----
    v->stats.records += n_lines;
    for (size_t i = 0; i < n_lines; ++i) {
      Line& ln = parsed[i];
      if (!ln.usable) { ++v->stats.skipped_non_snp; continue; }
      if (!open.lines.empty() && open.lines.front()->offset != ln.offset) {
        if (ln.offset < open.lines.front()->offset) {
          // the locus table must be sorted and unique (kgl_b200_select_loci searches it); the variant DB sorts by itself, a
          // streamed ingest cannot -- refuse instead of handing over a table that selects the wrong loci
          const std::string msg = "VCF is not sorted by POS: record at POS " + std::to_string((unsigned long long)ln.offset + 1) +
                                  " follows POS " + std::to_string((unsigned long long)open.lines.front()->offset + 1);
          gzclose(gz); delete v;
          return fail(msg);
        }
        emit_group(*v, open, unphased != 0); open.lines.clear(); held.clear();
      }
      open.lines.push_back(&ln);
    }
    // the open group may continue in the next block: its lines move out of `parsed`, which the next block overwrites
    for (Line*& l : open.lines)
      if (l >= parsed.data() && l < parsed.data() + parsed.size()) { held.push_back(std::make_unique<Line>(std::move(*l))); l = held.back().get(); }
  }
  if (!open.lines.empty()) emit_group(*v, open, unphased != 0);
  gzclose(gz);
  if (!have_header) { delete v; return fail("no #CHROM header line"); }
  // ---- final layouts ----
  const size_t L = v->offsets.size(), M = v->multi_rows.size();
  v->af.assign((size_t)kPops * L, kNone);
  for (size_t l = 0; l < L; ++l) for (int k = 0; k < kPops; ++k) v->af[(size_t)k * L + l] = v->af_rows[l * kPops + k];
  v->af_rows.clear(); v->af_rows.shrink_to_fit();
  v->multi_af.assign((size_t)kPops * M * kSlots, kNone);
  for (size_t m = 0; m < M; ++m) for (int k = 0; k < kPops; ++k) for (int s = 0; s < kSlots; ++s)
    v->multi_af[((size_t)k * M + m) * kSlots + s] = v->multi_af_rows[(m * kPops + k) * kSlots + s];
  v->multi_af_rows.clear(); v->multi_af_rows.shrink_to_fit();
  for (uint32_t r : v->multi_rows) for (int k = 0; k < kPops; ++k) v->af[(size_t)k * L + r] = kNone;   // the table is not used at such a row
  v->stats.bytes = total_bytes;
  v->stats.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (std::getenv("KGL_B200_VCF_TIMING")) std::fprintf(stderr, "vcf ingest: %.3f s, %zu rows (%zu multi-allelic), %d threads\n", v->stats.seconds, L, M, nt);
  *out = v;
  return 0;
}

void kgl_b200_vcf_free(kgl_b200_vcf* v) { delete v; }
uint64_t kgl_b200_vcf_n_genomes(const kgl_b200_vcf* v) { return v ? v->n_genomes : 0; }
uint64_t kgl_b200_vcf_n_loci(const kgl_b200_vcf* v) { return v ? v->offsets.size() : 0; }
uint64_t kgl_b200_vcf_row_bytes(const kgl_b200_vcf* v) { return v ? v->row_bytes : 0; }
const uint8_t* kgl_b200_vcf_packed(const kgl_b200_vcf* v) { return v ? v->packed.data() : nullptr; }
const float* kgl_b200_vcf_af(const kgl_b200_vcf* v) { return v ? v->af.data() : nullptr; }
const uint32_t* kgl_b200_vcf_offsets(const kgl_b200_vcf* v) { return v ? v->offsets.data() : nullptr; }
uint64_t kgl_b200_vcf_n_multi(const kgl_b200_vcf* v) { return v ? v->multi_rows.size() : 0; }
const uint32_t* kgl_b200_vcf_multi_rows(const kgl_b200_vcf* v) { return v ? v->multi_rows.data() : nullptr; }
const float* kgl_b200_vcf_multi_af(const kgl_b200_vcf* v) { return v ? v->multi_af.data() : nullptr; }
const uint8_t* kgl_b200_vcf_multi_cells(const kgl_b200_vcf* v) { return v ? v->multi_cells.data() : nullptr; }
const char* kgl_b200_vcf_genome_name(const kgl_b200_vcf* v, uint64_t i) { return (v && i < v->genomes.size()) ? v->genomes[i].c_str() : ""; }
const char* kgl_b200_vcf_contig(const kgl_b200_vcf* v) { return v ? v->contig.c_str() : ""; }
void kgl_b200_vcf_get_stats(const kgl_b200_vcf* v, kgl_b200_vcf_stats* stats) { if (v && stats) *stats = v->stats; }

}  // extern "C"
