// Host build of csrc/decimal_to_double.cuh (the device tokenizer's FP64 conversion) checked against strtod on random
// decimal strings of the shapes edge-weight files carry; run by tests/test_host_side.py.  Prints the totals and, per
// input shape, how many values the conversion left to the host fallback; exit status 1 on any wrong double.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <random>
#include <string>
#include "decimal_to_double.cuh"
static const uint64_t POW5[] = {
#include "pow5_table.inc"
};
int main(int argc, char **argv) {
    double p10[23]; p10[0] = 1.0; for (int i = 1; i < 23; i++) p10[i] = p10[i-1] * 10.0;
    std::mt19937_64 rng(12345);
    long N = argc > 1 ? atol(argv[1]) : 2000000;
    long undecided = 0, bad = 0, total = 0; long und_kind[8] = {0};
    char buf[128];
    for (long it = 0; it < N; it++) {
        int kind = it % 8;
        if (kind == 0) { // random double printed with 17 digits
            uint64_t b = rng(); b &= 0x7FFFFFFFFFFFFFFFull; b = (b % 0x7FE0000000000000ull) + 0x0010000000000000ull; double d; memcpy(&d,&b,8);
            snprintf(buf, sizeof buf, "%.17g", d);
        } else if (kind == 1) { double d = (double)(rng() >> 11) * (1.0 / 9007199254740992.0); snprintf(buf, sizeof buf, "%.17g", d); }
        else if (kind == 2) { double d = (double)(rng() >> 11) * (1.0 / 9007199254740992.0); snprintf(buf, sizeof buf, "%.16e", d); }
        else if (kind == 3) { snprintf(buf, sizeof buf, "%llu.%llu", (unsigned long long)(rng()%1000000), (unsigned long long)(rng()%100000000000000ull)); }
        else if (kind == 4) { snprintf(buf, sizeof buf, "%llue%d", (unsigned long long)rng(), (int)(rng()%600) - 300); }
        else if (kind == 5) { double d = (double)(rng() % 100000) / 1000.0; snprintf(buf, sizeof buf, "%g", d); }
        else if (kind == 6) { snprintf(buf, sizeof buf, "%llu%llu.%llue-%d", (unsigned long long)(rng()%1000000000ull), (unsigned long long)(rng()%1000000000ull), (unsigned long long)(rng()%10000), (int)(rng()%40)); }
        else { // halfway cases: 2^53 + odd, large integers
            uint64_t m = (1ull<<53) + (rng() % 100000) ; snprintf(buf, sizeof buf, "%llu%s", (unsigned long long)m, (rng()&1) ? "5" : "0"); }
        std::string s = buf; s.push_back('\n');
        uint64_t p = 0; double got = 0;
        int st = gx::parse_double_token(s, p, p10, POW5, &got);
        char *e; double ref = strtod(buf, &e);
        total++;
        if (st != 0) { undecided++; und_kind[kind]++; continue; }
        if ((size_t)(e - buf) != p) { bad++; if (bad < 10) printf("LEN %s: %zu vs %llu\n", buf, (size_t)(e-buf), (unsigned long long)p); continue; }
        if (memcmp(&got, &ref, 8) != 0) { bad++; if (bad < 10) printf("MISMATCH %s: got %.17g ref %.17g\n", buf, got, ref); }
    }
    printf("total %ld undecided %ld bad %ld\n", total, undecided, bad); for (int k = 0; k < 8; k++) printf("kind %d undecided %ld\n", k, und_kind[k]);
    return bad ? 1 : 0;
}
