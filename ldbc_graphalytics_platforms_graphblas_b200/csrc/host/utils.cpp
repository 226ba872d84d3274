// utils.cpp -- command-line flags of the per-algorithm binaries and the converter, and the epoch clock behind
// the "Processing starts/ends at" lines.  Flag semantics are the reference's (src/utils.cpp:19-53): `--key value`
// pairs in any order, unknown keys skipped, booleans spelled "true".  Table-driven: one row per flag.
#include "utils.h"

#include <sys/time.h>

#include <cstring>

namespace {

bool spelled_true(const char *text) { return text != nullptr && std::strcmp(text, "true") == 0; }

template <class Params>
struct FlagRow {
    const char *name;
    void (*store)(Params &, const char *);
};

// Walks argv once; a flag in the very last slot has no value and is skipped rather than read past the array.
template <class Params, size_t N>
Params collect(int argc, char **argv, const FlagRow<Params> (&rows)[N])
{
    Params out;
    for (int slot = 1; slot + 1 < argc; ++slot)
        for (const FlagRow<Params> &row : rows)
            if (std::strcmp(argv[slot], row.name) == 0) {
                row.store(out, argv[slot + 1]);
                break;
            }
    return out;
}

const FlagRow<BenchmarkParameters> kJobFlags[] = {
    {"--input-dir", [](BenchmarkParameters &b, const char *v) { b.input_dir = v; }},
    {"--output-file", [](BenchmarkParameters &b, const char *v) { b.output_file = v; }},
    {"--binary", [](BenchmarkParameters &b, const char *v) { b.binary = spelled_true(v); }},
    {"--directed", [](BenchmarkParameters &b, const char *v) { b.directed = spelled_true(v); }},
    {"--source-vertex", [](BenchmarkParameters &b, const char *v) { b.source_vertex = std::stoul(v); }},
    {"--max-iteration", [](BenchmarkParameters &b, const char *v) { b.max_iteration = std::stoi(v); }},
    {"--damping-factor", [](BenchmarkParameters &b, const char *v) { b.damping_factor = std::stod(v); }},
    {"--threadnum", [](BenchmarkParameters &b, const char *v) { b.thread_num = std::stoul(v); }},
};

const FlagRow<ConverterParameters> kConverterFlags[] = {
    {"--data-dir", [](ConverterParameters &c, const char *v) { c.data_dir = v; }},
    {"--weighted", [](ConverterParameters &c, const char *v) { c.weighted = spelled_true(v); }},
    {"--directed", [](ConverterParameters &c, const char *v) { c.directed = spelled_true(v); }},
};

} // namespace

BenchmarkParameters ParseBenchmarkParameters(int argc, char **argv) { return collect(argc, argv, kJobFlags); }

ConverterParameters ParseConverterParameters(int argc, char **argv) { return collect(argc, argv, kConverterFlags); }

time_t GetCurrentMilliseconds()
{
    timeval now;
    gettimeofday(&now, nullptr);
    return (time_t)now.tv_sec * 1000 + (time_t)(now.tv_usec / 1000);
}
