// kgl_b200_vcf_ingest.cpp -- see kgl_b200_vcf_ingest.h. Standalone C++17 (zlib, threads); no reference headers needed.
#include "kgl_b200_vcf_ingest.h"

#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

struct kgl_b200_vcf {
  std::vector<std::string> genomes;
  std::string contig;
  uint64_t n_genomes = 0, row_bytes = 0;
  std::vector<uint8_t> packed;
  std::vector<float> af;            // [6][n_loci]
  std::vector<uint32_t> offsets;
  kgl_b200_vcf_stats stats{};
};

namespace {

constexpr int kPops = 6;
const char* const kAfKeys[kPops] = {"AFR_AF", "AMR_AF", "EAS_AF", "EUR_AF", "SAS_AF", "AF"};

struct Row {
  uint32_t offset = 0;
  float af[kPops];
  int status = 0;                   // 0 kept, 1 multi-allelic, 2 non-SNP
  bool pass = true;
  uint32_t malformed = 0;
  std::vector<uint8_t> bits;        // row_bytes
};

// One allele token -> alt index (0 = reference); returns false when the token is not a number.
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

void parse_info(const char* b, const char* e, float (&af)[kPops]) {
  for (int k = 0; k < kPops; ++k) af[k] = std::numeric_limits<float>::quiet_NaN();
  const char* p = b;
  while (p < e) {
    const char* semi = static_cast<const char*>(std::memchr(p, ';', (size_t)(e - p)));
    const char* fe = semi ? semi : e;
    const char* eq = static_cast<const char*>(std::memchr(p, '=', (size_t)(fe - p)));
    if (eq) {
      const size_t klen = (size_t)(eq - p);
      for (int k = 0; k < kPops; ++k) {
        if (std::strlen(kAfKeys[k]) == klen && std::memcmp(p, kAfKeys[k], klen) == 0) {
          // first value of a Number=A field (biallelic rows only reach the matrix)
          std::string val(eq + 1, (size_t)(fe - eq - 1));
          const size_t comma = val.find(',');
          if (comma != std::string::npos) val.resize(comma);
          char* endp = nullptr;
          const float f = std::strtof(val.c_str(), &endp);
          if (endp != val.c_str()) af[k] = f;
        }
      }
    }
    p = fe + 1;
  }
}

// Parses one data line into `row`. Fields: CHROM POS ID REF ALT QUAL FILTER INFO FORMAT samples...
void parse_line(const char* b, const char* e, uint64_t n_genomes, uint64_t row_bytes, bool unphased, Row& row, std::string* contig) {
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
  row.bits.assign(row_bytes, 0);
  row.status = 0; row.malformed = 0;
  if (nf < 9) { row.status = 2; return; }
  if (contig && contig->empty()) contig->assign(f[0], (size_t)(f[1] - 1 - f[0]));
  row.offset = (uint32_t)std::strtoul(std::string(f[1], (size_t)(f[2] - 1 - f[1])).c_str(), nullptr, 10);
  row.offset = row.offset > 0 ? row.offset - 1 : 0;
  const size_t ref_len = (size_t)(f[4] - 1 - f[3]), alt_len = (size_t)(f[5] - 1 - f[4]);
  const char* alt = f[4];
  if (std::memchr(alt, ',', alt_len)) { row.status = 1; return; }
  if (ref_len != 1 || alt_len != 1 || *alt == '.' || *alt == '<' || *alt == '*') { row.status = 2; return; }
  {
    // Utility::toupper(filter) == "PASS" (1000_impl.cpp:73)
    const size_t fl = (size_t)(f[7] - 1 - f[6]);
    row.pass = fl == 4 && (f[6][0] | 0x20) == 'p' && (f[6][1] | 0x20) == 'a' && (f[6][2] | 0x20) == 's' && (f[6][3] | 0x20) == 's';
  }
  parse_info(f[7], f[8] - 1, row.af);
  if (!row.pass) for (int k = 0; k < kPops; ++k) row.af[k] = std::numeric_limits<float>::quiet_NaN();
  if (nf < 10) return;                                         // no sample columns
  const char sep = unphased ? '/' : '|';
  p = f[9];
  for (uint64_t g = 0; g < n_genomes && p <= e; ++g) {
    // fast path: the three-character genotype "a|b" with single-digit / '.' alleles, ended by a tab or the line end
    if (e - p >= 3 && p[1] == sep && (p + 3 == e || p[3] == '\t')) {
      const char ca = p[0], cb = p[2];
      const bool da = ca == '0' || ca == '1', db = cb == '0' || cb == '1';
      const bool ma = ca == '.' || ca == '-', mb = cb == '.' || cb == '-';
      if ((da || ma) && (db || mb)) {
        unsigned code = (unsigned)(ca == '1') + (unsigned)(cb == '1');
        if (unphased && (ca == '.' || cb == '.')) code = 0;
        if (code & 1u) row.bits[(g >> 6) * 16 + ((g & 63) >> 3)] |= (uint8_t)(1u << (g & 7));
        if (code & 2u) row.bits[(g >> 6) * 16 + 8 + ((g & 63) >> 3)] |= (uint8_t)(1u << (g & 7));
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
    unsigned code = 0;
    const char* s = static_cast<const char*>(std::memchr(p, sep, (size_t)(gt_e - p)));
    if (!s && unphased) s = static_cast<const char*>(std::memchr(p, '|', (size_t)(gt_e - p)));
    if (s) {
      uint32_t a = 0, bb = 0;
      const bool ok = allele_index(p, s, a) && allele_index(s + 1, gt_e, bb);
      const bool missing = (s - p == 1 && *p == '.') || (gt_e - s - 1 == 1 && s[1] == '.');
      if (!ok || a > 1 || bb > 1) { ++row.malformed; code = 0; }                 // beyond the ALT list / not a number: reference
      else if (unphased && missing) code = 0;                                   // Pf7: genotype skipped
      else code = (a != 0) + (bb != 0);
    } else if (gt_e > p) {
      if (!(gt_e - p == 1 && (*p == '.' || *p == '-'))) ++row.malformed;        // haploid GT on an autosome: reference (:193-203)
    }
    if (code & 1u) row.bits[(g >> 6) * 16 + ((g & 63) >> 3)] |= (uint8_t)(1u << (g & 7));
    if (code & 2u) row.bits[(g >> 6) * 16 + 8 + ((g & 63) >> 3)] |= (uint8_t)(1u << (g & 7));
    if (!t) break;
    p = t + 1;
  }
}

bool read_all(const char* path, std::string& data, std::string& err) {
  gzFile f = gzopen(path, "rb");                               // transparently reads plain text too
  if (!f) { err = std::string("cannot open ") + path; return false; }
  gzbuffer(f, 1 << 20);
  std::vector<char> buf(8u << 20);
  for (;;) {
    const int n = gzread(f, buf.data(), (unsigned)buf.size());
    if (n < 0) { int e = 0; err = std::string("read error: ") + gzerror(f, &e); gzclose(f); return false; }
    if (n == 0) break;
    data.append(buf.data(), (size_t)n);
  }
  gzclose(f);
  return true;
}

}  // namespace

extern "C" {

int kgl_b200_vcf_ingest(const char* path, int unphased, int n_threads, kgl_b200_vcf** out, char* err, size_t err_len) {
  auto fail = [&](const std::string& m) { if (err && err_len) { std::snprintf(err, err_len, "%s", m.c_str()); } return 1; };
  if (!path || !out) return fail("null argument");
  *out = nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  std::string data, e;
  if (!read_all(path, data, e)) return fail(e);
  const double t_read = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  auto v = new kgl_b200_vcf();
  // header
  size_t pos = 0;
  bool have_header = false;
  while (pos < data.size()) {
    size_t nl = data.find('\n', pos);
    if (nl == std::string::npos) nl = data.size();
    if (data[pos] != '#') break;
    if (data.compare(pos, 6, "#CHROM") == 0) {
      std::string line = data.substr(pos, nl - pos);
      if (!line.empty() && line.back() == '\r') line.pop_back();
      size_t p = 0; int col = 0;
      while (p <= line.size()) {
        size_t t = line.find('\t', p);
        if (t == std::string::npos) t = line.size();
        if (col >= 9) v->genomes.emplace_back(line.substr(p, t - p));
        ++col; p = t + 1;
      }
      have_header = true;
    }
    pos = nl + 1;
  }
  if (!have_header) { delete v; return fail("no #CHROM header line"); }
  v->n_genomes = v->genomes.size();
  v->row_bytes = 16 * ((v->n_genomes + 63) / 64);
  // line index of the data section
  std::vector<std::pair<size_t, size_t>> lines;
  while (pos < data.size()) {
    size_t nl = data.find('\n', pos);
    if (nl == std::string::npos) nl = data.size();
    size_t end = nl;
    if (end > pos && data[end - 1] == '\r') --end;
    if (end > pos) lines.emplace_back(pos, end);
    pos = nl + 1;
  }
  const size_t n_lines = lines.size();
  std::vector<Row> rows(n_lines);
  int nt = n_threads > 0 ? n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
  nt = (int)std::min<size_t>((size_t)nt, std::max<size_t>(1, n_lines / 64));
  if (n_lines) { parse_line(data.data() + lines[0].first, data.data() + lines[0].second, v->n_genomes, v->row_bytes, unphased != 0, rows[0], &v->contig); }
  std::atomic<size_t> next{1};
  auto work = [&]() {
    for (;;) {
      const size_t i0 = next.fetch_add(256);
      if (i0 >= n_lines) break;
      const size_t i1 = std::min(n_lines, i0 + 256);
      for (size_t i = i0; i < i1; ++i)
        parse_line(data.data() + lines[i].first, data.data() + lines[i].second, v->n_genomes, v->row_bytes, unphased != 0, rows[i], nullptr);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; ++t) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
  const double t_parse = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  // repeated POS: the variant DB would hold several variants at the offset -> not representable, all of them are dropped
  std::vector<uint8_t> drop(n_lines, 0);
  for (size_t i = 0; i < n_lines; ++i) {
    if (rows[i].status != 0) continue;
    const bool dup_prev = i > 0 && rows[i - 1].status != 2 && rows[i - 1].offset == rows[i].offset;
    const bool dup_next = i + 1 < n_lines && rows[i + 1].status != 2 && rows[i + 1].offset == rows[i].offset;
    if (dup_prev || dup_next) drop[i] = 1;
  }
  kgl_b200_vcf_stats& st = v->stats;
  st.records = n_lines; st.bytes = data.size();
  size_t kept = 0;
  for (size_t i = 0; i < n_lines; ++i) {
    if (rows[i].status == 1 || drop[i]) ++st.skipped_multi_allelic;
    else if (rows[i].status == 2) ++st.skipped_non_snp;
    else ++kept;
  }
  v->offsets.reserve(kept);
  v->packed.resize(kept * v->row_bytes);
  v->af.assign((size_t)kPops * kept, std::numeric_limits<float>::quiet_NaN());
  size_t r = 0;
  for (size_t i = 0; i < n_lines; ++i) {
    if (rows[i].status != 0 || drop[i]) continue;
    v->offsets.push_back(rows[i].offset);
    std::memcpy(v->packed.data() + r * v->row_bytes, rows[i].bits.data(), v->row_bytes);
    for (int k = 0; k < kPops; ++k) v->af[(size_t)k * kept + r] = rows[i].af[k];
    if (!rows[i].pass) ++st.not_pass;
    st.malformed_genotypes += rows[i].malformed;
    ++r;
  }
  st.kept = kept;
  st.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (std::getenv("KGL_B200_VCF_TIMING")) std::fprintf(stderr, "vcf ingest: read %.3f s, index+parse %.3f s, pack %.3f s (%d threads)\n", t_read, t_parse - t_read, st.seconds - t_parse, nt);
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
const char* kgl_b200_vcf_genome_name(const kgl_b200_vcf* v, uint64_t i) { return (v && i < v->genomes.size()) ? v->genomes[i].c_str() : ""; }
const char* kgl_b200_vcf_contig(const kgl_b200_vcf* v) { return v ? v->contig.c_str() : ""; }
void kgl_b200_vcf_get_stats(const kgl_b200_vcf* v, kgl_b200_vcf_stats* stats) { if (v && stats) *stats = v->stats; }

}  // extern "C"
