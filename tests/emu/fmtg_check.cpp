/* TEST SCAFFOLDING: csrc/xm_fmtg.h against the C library's "%g" on 32-bit floats.
 *   fmtg_check STRIDE [THREADS]     every STRIDE-th bit pattern (1: all 2^32), exit status 0 when none differs */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#include "../../xenomapper_b200/csrc/xm_fmtg.h"

int main(int argc, char **argv)
{
    const uint64_t stride = argc > 1 ? strtoull(argv[1], nullptr, 10) : 65537;
    const int nt = argc > 2 ? atoi(argv[2]) : 1;
    std::atomic<uint64_t> bad(0), done(0);
    auto work = [&](int t) {
        char a[32], b[64];
        for (uint64_t u = (uint64_t)t * stride; u < (1ull << 32); u += stride * (uint64_t)nt) {
            const uint32_t bits = (uint32_t)u;
            float f;
            memcpy(&f, &bits, 4);
            const int n = xm::fmt_g(f, a);
            a[n] = 0;
            snprintf(b, sizeof b, "%g", (double)f);
            if (strcmp(a, b) != 0 && bad.fetch_add(1) < 20) fprintf(stderr, "0x%08x: fmt_g \"%s\"  printf \"%s\"\n", bits, a, b);
            done.fetch_add(1);
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back(work, t);
    for (auto &x : th) x.join();
    printf("%llu floats checked, %llu differ\n", (unsigned long long)done.load(), (unsigned long long)bad.load());
    return bad.load() ? 1 : 0;
}
