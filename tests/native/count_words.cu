// Host-side check of the count-carrying 64-bit words of the cross-GPU exchange (bayesrrcmd_b200/csrc/sweep_dense.cuh: kCnt1,
// line_count, line_value): the sum of `n` signed values, each added together with kCnt1, must decode to (n, sum) whenever
// |sum| < 2^55 and n <= 255 -- negative sums borrow from the count field and the decoding has to undo that.
#include <cstdio>
#include <cstdint>
#include <random>
#include "sweep_dense.cuh"

int main()
{
    using namespace brr;
    std::mt19937_64 rng(12345);
    long long bad = 0, cases = 0;
    const long long lim = (1LL << 52);                     // per-GPU reduced sums stay below 2^52 (publish_dots)
    for (int trial = 0; trial < 200000; ++trial) {
        const unsigned n = 1u + unsigned(rng() % 7u);      // peers of a round (world <= 8)
        unsigned long long word = 0;
        long long sum = 0;
        for (unsigned k = 0; k < n; ++k) {
            long long v;
            switch (trial % 4) {
            case 0: v = (long long)(rng() % (2 * lim)) - lim; break;            // full range
            case 1: v = -(long long)(rng() % lim); break;                       // all negative
            case 2: v = (long long)(rng() % 3) - 1; break;                      // around zero
            default: v = (k & 1) ? lim - 1 : -(lim - 1); break;                 // extremes
            }
            sum += v;
            word += (unsigned long long)v + kCnt1;
            // partial words decode to the partial count as well (a reader sees them while it polls)
            if (line_count(word) != k + 1) ++bad;
        }
        ++cases;
        if (line_count(word) != n || line_value(word) != sum) ++bad;
    }
    // a zero word is "nobody yet"
    if (line_count(0ULL) != 0u || line_value(0ULL) != 0) ++bad;
    std::printf("cases %lld bad %lld\n", cases, bad);
    return bad ? 1 : 0;
}
