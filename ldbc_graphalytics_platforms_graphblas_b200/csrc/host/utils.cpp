#include "utils.h"

#include <chrono>
#include <cstring>

time_t GetCurrentMilliseconds()
{
    using namespace std::chrono;
    return (time_t)duration_cast<milliseconds>(system_clock::now().time_since_epoch()).count();
}

static bool is_true(const char *v) { return v && std::strcmp(v, "true") == 0; }

BenchmarkParameters ParseBenchmarkParameters(int argc, char **argv)
{
    BenchmarkParameters p;
    // every argv slot is tried as a key (the reference scans all i, src/utils.cpp:22-50); a key
    // in the last slot has no value and is ignored instead of reading past argv
    for (int i = 1; i + 1 < argc; i++) {
        const std::string key = argv[i];
        const char *value = argv[i + 1];
        if (key == "--binary") p.binary = is_true(value);
        else if (key == "--input-dir") p.input_dir = value;
        else if (key == "--output-file") p.output_file = value;
        else if (key == "--directed") p.directed = is_true(value);
        else if (key == "--source-vertex") p.source_vertex = std::stoul(value);
        else if (key == "--damping-factor") p.damping_factor = std::stod(value);
        else if (key == "--max-iteration") p.max_iteration = std::stoi(value);
        else if (key == "--threadnum") p.thread_num = std::stoul(value);
    }
    return p;
}

ConverterParameters ParseConverterParameters(int argc, char **argv)
{
    ConverterParameters p;
    for (int i = 1; i + 1 < argc; i++) {
        const std::string key = argv[i];
        if (key == "--data-dir") p.data_dir = argv[i + 1];
        else if (key == "--weighted") p.weighted = is_true(argv[i + 1]);
        else if (key == "--directed") p.directed = is_true(argv[i + 1]);
    }
    return p;
}
