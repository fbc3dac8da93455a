// cobs/file/file_io_exception.hpp -- thrown for bad magic words / versions
// (reference: cobs/file/file_io_exception.hpp:16-33, cobs/file/header.hpp:23-53)
#pragma once
#include <stdexcept>
#include <string>

namespace cobs {

class FileIOException : public std::runtime_error
{
public:
    explicit FileIOException(const std::string& msg) : std::runtime_error(msg), msg_(msg) { }
    std::string& message() { return msg_; }

private:
    std::string msg_;
};

} // namespace cobs
