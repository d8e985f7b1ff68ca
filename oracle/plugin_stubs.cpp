// plugin_stubs.cpp -- TEST INFRASTRUCTURE (oracle/_ref build only).
// The file-reader factories behind kel_io/kel_basic_io.cpp (kel_file_io.cpp and kel_bzip_workflow.cpp need
// Boost.Iostreams / zlib plumbing that this image lacks). The plugin harness builds its PED data in memory and never
// opens a file through them; a call returns "could not open".
#include "kel_file_io.h"
#include "kel_bzip_workflow.h"

namespace kel = kellerberrin;

std::optional<std::unique_ptr<kel::BaseStreamIO>> kel::TextStreamIO::getStreamIO(const std::string&) { return std::nullopt; }
std::optional<std::unique_ptr<kel::BaseStreamIO>> kel::GZStreamIO::getStreamIO(const std::string&) { return std::nullopt; }
std::optional<std::unique_ptr<kel::BaseStreamIO>> kel::BZ2StreamIO::getStreamIO(const std::string&) { return std::nullopt; }
std::optional<std::unique_ptr<kel::BaseStreamIO>> kel::BGZStreamIO::getStreamIO(const std::string&, size_t) { return std::nullopt; }
bool kel::BGZStreamIO::verify(const std::string&, bool) { return false; }
