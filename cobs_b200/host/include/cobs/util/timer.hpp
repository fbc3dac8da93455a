// cobs/util/timer.hpp -- named phase timers printed as "TIMER info=<tag> <phase>=<sec> ... total="
// Same public interface as the reference (cobs/util/timer.hpp:19-63); add() is an extension so
// that phases measured with CUDA events on the device can be accounted.
#pragma once
#include <chrono>
#include <ostream>
#include <string>
#include <vector>

namespace cobs {

class Timer
{
public:
    Timer() = default;

    //! start the named phase (stops the running one)
    void active(const char* timer);
    void stop();
    void reset();
    //! seconds accumulated under `timer`
    double get(const char* timer);
    void print(const char* info, std::ostream& os) const;
    //! prints to stderr
    void print(const char* info) const;
    //! merge another timer's phases
    Timer& operator += (const Timer& b);

    //! extension: account `seconds` measured elsewhere (CUDA events) to `timer`
    void add(const char* timer, double seconds);

private:
    struct Entry {
        std::string name;
        double seconds;
    };
    std::vector<Entry> timers_;
    double total_ = 0;
    const char* running_ = nullptr;
    std::chrono::steady_clock::time_point start_;

    Entry& find_or_create(const char* name);
};

} // namespace cobs
