// cobs/util/fs.hpp -- filesystem alias (reference: cobs/util/fs.hpp:14-22, experimental::filesystem)
#pragma once
#include <filesystem>
#include <system_error>

namespace cobs {
namespace fs = std::filesystem;
using std::error_code;
} // namespace cobs
