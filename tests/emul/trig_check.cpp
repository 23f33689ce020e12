// Exhaustive host check: olf::trig::sinf_exact / cosf_exact (the source the device compiles) == this machine's libm sinf / cosf
// on EVERY float in [0, 6.5] (the rBRIEF angles are in [0, 2 pi]).  argv[1] = stride (1 = exhaustive).  Prints mismatch counts.
#include "../../orb_line_slam_b200/csrc/sincosf_exact.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
int main(int argc, char** argv) {
    const uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 1;
    const float hi = 6.5f; uint32_t uh; memcpy(&uh, &hi, 4);
    const int nt = (int)std::max(1u, std::thread::hardware_concurrency());
    std::vector<long> bad(nt, 0), tot(nt, 0);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back([&, t] {
        for (uint64_t u = (uint64_t)t * stride; u <= uh; u += (uint64_t)nt * stride) {
            float f; const uint32_t v = (uint32_t)u; memcpy(&f, &v, 4);
            ++tot[t];
            if (olf::trig::sinf_exact(f) != sinf(f) || olf::trig::cosf_exact(f) != cosf(f)) ++bad[t];
        }
    });
    for (auto& x : th) x.join();
    long b = 0, n = 0; for (int t = 0; t < nt; ++t) { b += bad[t]; n += tot[t]; }
    printf("%ld %ld\n", n, b);
    return b ? 1 : 0;
}
