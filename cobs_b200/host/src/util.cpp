// settings, error handling, timer and header sniffing of the host-side COBS mirror
#include <cobs/file/file_io_exception.hpp>
#include <cobs/settings.hpp>
#include <cobs/util/error_handling.hpp>
#include <cobs/util/file.hpp>
#include <cobs/util/timer.hpp>

#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <mutex>
#include <thread>

namespace cobs {

// ---- settings (reference: cobs/settings.cpp:15-19) -------------------------------------

static int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return v && *v ? std::atoi(v) : dflt;
}

size_t gopt_threads = std::thread::hardware_concurrency();
bool gopt_load_complete_index = false;
bool gopt_disable_cache = false;
int gopt_gpu_device = env_int("COBS_GPU_DEVICE", 0);
unsigned gopt_gpus = static_cast<unsigned>(std::max(1, env_int("COBS_GPUS", 1)));

// ---- error handling (reference: cobs/util/error_handling.cpp:15-32) --------------------

void print_errno(const std::string& msg) {
    std::cerr << msg + ": " << std::strerror(errno) << std::endl;
}

void exit_error(const std::string& msg) {
    std::cerr << msg << std::endl;
    std::exit(EXIT_FAILURE);
}

void assert_exit(bool cond, const std::string& msg) {
    if (!cond) exit_error(msg);
}

void exit_error_errno(const std::string& msg) {
    exit_error(msg + ": " + std::strerror(errno));
}

static bool s_die_with_exception = false;

bool set_die_with_exception(bool b) {
    bool old = s_die_with_exception;
    s_die_with_exception = b;
    return old;
}

void die_with_message(const std::string& msg) {
    if (s_die_with_exception) throw DieException(msg);
    std::cerr << "DIE: " << msg << std::endl;
    std::terminate();
}

// ---- timer (reference: cobs/util/timer.cpp:23-89) ---------------------------------------

Timer::Entry& Timer::find_or_create(const char* name) {
    for (Entry& e : timers_)
        if (e.name == name) return e;
    timers_.push_back(Entry { name, 0.0 });
    return timers_.back();
}

void Timer::active(const char* timer) {
    stop();
    running_ = timer;
}

void Timer::stop() {
    auto now = std::chrono::steady_clock::now();
    if (running_) {
        double s = std::chrono::duration<double>(now - start_).count();
        find_or_create(running_).seconds += s;
        total_ += s;
    }
    start_ = now;
    running_ = nullptr;
}

void Timer::reset() {
    timers_.clear();
    total_ = 0;
}

double Timer::get(const char* timer) {
    return find_or_create(timer).seconds;
}

void Timer::add(const char* timer, double seconds) {
    find_or_create(timer).seconds += seconds;
    total_ += seconds;
}

//! several threads may fold their timers into a shared one (reference: timer.cpp:67-75)
static std::mutex s_timer_add_mutex;

Timer& Timer::operator += (const Timer& b) {
    std::unique_lock<std::mutex> lock(s_timer_add_mutex);
    for (const Entry& t : b.timers_) find_or_create(t.name.c_str()).seconds += t.seconds;
    total_ += b.total_;
    return *this;
}

void Timer::print(const char* info, std::ostream& os) const {
    os << "TIMER info=" << info;
    for (const Entry& t : timers_) os << ' ' << t.name << '=' << t.seconds;
    os << " total=" << total_ << std::endl;
}

void Timer::print(const char* info) const {
    print(info, std::cerr);
}

// ---- header sniffing ---------------------------------------------------------------------

const std::string ClassicIndexHeader::magic_word = "CLASSIC_INDEX";
const uint32_t ClassicIndexHeader::version = 1;
const std::string ClassicIndexHeader::file_extension = ".cobs_classic";
const std::string CompactIndexHeader::magic_word = "COMPACT_INDEX";
const uint32_t CompactIndexHeader::version = 1;
const std::string CompactIndexHeader::file_extension = ".cobs_compact";

bool file_has_magic(const fs::path& p, const std::string& magic_word, uint32_t version) {
    std::error_code ec;
    if (!fs::is_regular_file(p, ec)) return false;
    std::ifstream is(p.string(), std::ios::in | std::ios::binary);
    std::string want = "COBS:" + magic_word;
    std::string got(want.size(), '\0');
    uint32_t v = 0;
    is.read(&got[0], static_cast<std::streamsize>(got.size()));
    is.read(reinterpret_cast<char*>(&v), sizeof(v));
    return is.good() && got == want && v == version;
}

} // namespace cobs
