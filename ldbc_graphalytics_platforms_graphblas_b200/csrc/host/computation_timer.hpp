// computation_timer.hpp -- scope stopwatch of the wrappers: prints "<name> starts" when a phase begins and
// "<name> duration: <seconds>s" (rounded to milliseconds) when it ends, one tab of indentation per nesting
// level -- the log lines of the reference's ComputationTimer (include/computation_timer.hpp:17-52), which
// the harness tees into runner.logs.
#pragma once

#include <sys/time.h>

#include <cmath>
#include <cstdio>
#include <string>

class ComputationTimer {
    std::string label_;
    int depth_;
    double began_; // seconds since the epoch

    static double wall()
    {
        timeval tv;
        gettimeofday(&tv, nullptr);
        return (double)tv.tv_sec + 1e-6 * (double)tv.tv_usec;
    }
    void say(const char *what, double seconds) const
    {
        std::string tabs((size_t)depth_, '\t');
        if (seconds < 0.0) std::printf("%s%s %s\n", tabs.c_str(), label_.c_str(), what);
        else std::printf("%s%s %s: %gs\n", tabs.c_str(), label_.c_str(), what, seconds);
        std::fflush(stdout);
    }

  public:
    explicit ComputationTimer(std::string label, int depth = 0) : label_(std::move(label)), depth_(depth), began_(wall())
    {
        say("starts", -1.0);
    }
    ComputationTimer(std::string label, const ComputationTimer &outer) : ComputationTimer(std::move(label), outer.depth_ + 1) {}
    ComputationTimer(const ComputationTimer &) = delete;
    ComputationTimer &operator=(const ComputationTimer &) = delete;
    ~ComputationTimer()
    {
        const double elapsed = wall() - began_;
        say("duration", elapsed > 0.0 ? std::round(elapsed * 1000.0) / 1000.0 : 0.0);
    }
};
