// utils.h -- flag parsing, clock and error helper of the six per-algorithm
// binaries.  Mirrors the reference's wrapper interface (include/utils.h:12-55,
// src/utils.cpp:8-53): same struct, same function names, same flag semantics
// (`--key value` pairs in any order, unknown keys ignored, defaults below).
#pragma once

#include <ctime>
#include <stdexcept>
#include <string>

#include "gxb200.h"

struct BenchmarkParameters {
    bool binary = false;
    std::string input_dir;
    std::string output_file;
    bool directed = false;
    unsigned long source_vertex = 0;
    double damping_factor = 0.0;
    int max_iteration = 0;
    unsigned long thread_num = 1; // accepted for CLI compatibility; the GPU path has no use for it
};

BenchmarkParameters ParseBenchmarkParameters(int argc, char **argv);

struct ConverterParameters {
    std::string data_dir;
    bool weighted = false;
    bool directed = false;
};

ConverterParameters ParseConverterParameters(int argc, char **argv);

time_t GetCurrentMilliseconds();

// OK(call): the reference's macro throws std::runtime_error on a GraphBLAS error
// (utils.h:45-55), which aborts the job with a non-zero exit status.  Same here
// for the gx_* status codes, with the library's message attached.
#define OK(method)                                                                             \
    do {                                                                                       \
        int info__ = (method);                                                                 \
        if (info__ != GX_OK)                                                                   \
            throw std::runtime_error(std::string("gxb200 error [") + std::to_string(info__) +  \
                                     "]  " + gx_last_error());                                 \
    } while (0)
