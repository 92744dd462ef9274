// Stress of cc_worker_pool (collectivecrossing_b200/csrc/cc_workers.h): many rounds of varying size, every index must run exactly once
// per round, run() must not return before all of them have.  Built with -fsanitize=thread where the toolchain has it.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../collectivecrossing_b200/csrc/cc_workers.h"

int main(int argc, char **argv) {
    const int rounds = argc > 1 ? atoi(argv[1]) : 3000;
    unsigned rng = 12345u;
    auto next = [&] { rng = rng * 1664525u + 1013904223u; return rng >> 8; };
    for (int pools = 0; pools < 3; ++pools) {
        cc_worker_pool pool;
        std::vector<std::atomic<int>> hits(16);
        long long plain[16] = {0};   // written without synchronisation of its own: the pool's hand-over must order it
        for (int r = 0; r < rounds; ++r) {
            const int n = 1 + (int)(next() % 12);
            for (auto &h : hits) h.store(0);
            pool.run(n, [&](int w) {
                hits[w].fetch_add(1);
                plain[w] += w + 1;
                if ((w + r) % 7 == 0) for (volatile int spin = 0; spin < 2000; ++spin) {}
            });
            for (int w = 0; w < 16; ++w)
                if (hits[w].load() != (w < n ? 1 : 0)) { printf("round %d n %d: index %d ran %d times\n", r, n, w, hits[w].load()); return 1; }
        }
        long long total = 0;
        for (long long v : plain) total += v;
        if (total <= 0) return 1;
    }
    printf("ok %d rounds x 3 pools\n", rounds);
    return 0;
}
