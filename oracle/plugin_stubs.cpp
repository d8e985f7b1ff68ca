// plugin_stubs.cpp -- TEST INFRASTRUCTURE (oracle/_ref build only).
// The file-reader factories behind kel_io/kel_basic_io.cpp (kel_file_io.cpp and kel_bzip_workflow.cpp need
// Boost.Iostreams / zlib plumbing that this image lacks). The plugin harness builds its PED data in memory and never
// opens a file through them; a call returns "could not open".
//
// For the VCF parser pin (plugin_harness.cpp --vcf: the reference's own Genome1000VCFImpl / PfVCFImpl, VCFReaderMT and ParseVCF
// read a PLAIN TEXT VCF) two pieces are supplied here: a line reader for uncompressed text behind TextStreamIO::getStreamIO, and
// the VCF header reader (kgl_variant_factory_vcf_parse_header.cpp needs boost/tokenizer): genome names from the #CHROM line and
// the ##INFO declarations (ID, Number, Type, Description) -- what EvidenceFactory::availableInfoFields needs to know the fields.
// Everything that turns a data line into variants is the reference's own code.
#include "kel_file_io.h"
#include "kel_bzip_workflow.h"
#include "kgl_variant_factory_vcf_parse_header.h"

#include <fstream>

namespace kel = kellerberrin;
namespace kgl = kellerberrin::genome;

namespace {

class HarnessTextStream : public kel::BaseStreamIO {
 public:
  bool open(const std::string& file_name) override { in_.open(file_name); line_ = 0; return in_.good(); }
  void close() override { in_.close(); }
  kel::IOLineRecord readLine() override {
    std::string text;
    if (!std::getline(in_, text)) return kel::IOLineRecord::createEOFMarker();
    if (!text.empty() && text.back() == '\r') text.pop_back();
    return kel::IOLineRecord(++line_, std::move(text));
  }
 private:
  std::ifstream in_;
  size_t line_{0};
};

}  // namespace

std::optional<std::unique_ptr<kel::BaseStreamIO>> kel::TextStreamIO::getStreamIO(const std::string& file_name) {
  auto stream = std::make_unique<HarnessTextStream>();
  if (!stream->open(file_name)) return std::nullopt;
  return std::optional<std::unique_ptr<kel::BaseStreamIO>>(std::move(stream));
}

bool kgl::VCFParseHeader::parseHeader(const std::string& file_name) {
  std::ifstream in(file_name);
  if (!in.good()) return false;
  std::string line;
  while (std::getline(in, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.rfind("##", 0) == 0) {
      const size_t eq = line.find('=');
      if (eq != std::string::npos) vcf_header_info_.emplace_back(line.substr(2, eq - 2), line.substr(eq + 1));
      continue;
    }
    if (line.rfind(FIELD_NAME_FRAGMENT_, 0) == 0) {
      size_t p = 0, col = 0;
      while (p <= line.size()) {
        size_t t = line.find(RECORD_FIELD_LIST_SEPARATOR_, p);
        if (t == std::string::npos) t = line.size();
        if (col >= SKIP_FIELD_NAMES_) vcf_genomes_.push_back(line.substr(p, t - p));
        ++col; p = t + 1;
      }
    }
    break;
  }
  return true;
}

bool kgl::VCFParseHeader::parseVcfHeader(const VCFHeaderInfo& header, VCFContigMap&, VCFInfoRecordMap& vcf_info_map) {
  for (auto const& [key, value] : header) {
    if (key != "INFO") continue;                       // ##INFO=<ID=AF,Number=A,Type=Float,Description="...">
    auto field = [&value](const std::string& name) -> std::string {
      const size_t at = value.find(name + "=");
      if (at == std::string::npos) return "";
      size_t b = at + name.size() + 1, e;
      if (b < value.size() && value[b] == '"') { ++b; e = value.find('"', b); }
      else e = value.find_first_of(",>", b);
      return value.substr(b, e == std::string::npos ? std::string::npos : e - b);
    };
    VCFInfoRecord record{field("ID"), field("Description"), field("Type"), field("Number"), field("Source"), field("Version")};
    if (!record.ID.empty()) vcf_info_map[record.ID] = record;
  }
  return true;
}

bool kgl::VCFParseHeader::checkVCFReferenceContigs(const VCFContigMap&, std::shared_ptr<const GenomeReference>) { return true; }

std::optional<std::unique_ptr<kel::BaseStreamIO>> kel::GZStreamIO::getStreamIO(const std::string&) { return std::nullopt; }
std::optional<std::unique_ptr<kel::BaseStreamIO>> kel::BZ2StreamIO::getStreamIO(const std::string&) { return std::nullopt; }
std::optional<std::unique_ptr<kel::BaseStreamIO>> kel::BGZStreamIO::getStreamIO(const std::string&, size_t) { return std::nullopt; }
bool kel::BGZStreamIO::verify(const std::string&, bool) { return false; }
