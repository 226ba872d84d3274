// computation_timer.hpp -- RAII wall-clock scope timer printing
// "<name> starts" / "<name> duration: <s>s" to stdout (tab-indented per nesting
// level, rounded to milliseconds), the log lines the reference's
// ComputationTimer emits (include/computation_timer.hpp:17-52).
#pragma once

#include <chrono>
#include <cmath>
#include <iostream>
#include <string>

class ComputationTimer {
    using clock = std::chrono::steady_clock;
    int level_;
    std::string name_;
    clock::time_point t0_;

    void indent() const
    {
        for (int i = 0; i < level_; i++) std::cout << '\t';
    }

  public:
    explicit ComputationTimer(std::string name, int level = 0) : level_(level), name_(std::move(name)), t0_(clock::now())
    {
        indent();
        std::cout << name_ << " starts" << std::endl;
    }
    ComputationTimer(std::string name, const ComputationTimer &parent) : ComputationTimer(std::move(name), parent.level_ + 1) {}
    ~ComputationTimer()
    {
        const double s = std::chrono::duration<double>(clock::now() - t0_).count();
        indent();
        std::cout << name_ << " duration: " << std::round(s * 1000.0) / 1000.0 << "s" << std::endl;
    }
};
