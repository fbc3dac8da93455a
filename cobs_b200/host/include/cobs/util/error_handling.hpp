// cobs/util/error_handling.hpp -- the reference's error conventions on the query path:
// assert_exit() prints and exit(EXIT_FAILURE)s (cobs/util/error_handling.cpp:19-28), die()
// terminates unless set_die_with_exception(true) (tlx/die/core.hpp:58).
#pragma once
#include <stdexcept>
#include <string>

namespace cobs {

void print_errno(const std::string& msg);
void exit_error(const std::string& msg);
void assert_exit(bool cond, const std::string& msg);
void exit_error_errno(const std::string& msg);

template <class E>
void assert_throw(bool cond, const std::string& msg) {
    if (!cond) throw E(msg);
}

//! exception thrown by die() when enabled (mirrors tlx::DieException)
class DieException : public std::runtime_error
{
public:
    explicit DieException(const std::string& msg) : std::runtime_error(msg) { }
};

//! switch die() from std::terminate to throwing DieException; returns the old value
bool set_die_with_exception(bool b);
//! print "DIE: msg" to stderr and terminate (or throw DieException)
[[noreturn]] void die_with_message(const std::string& msg);

} // namespace cobs
