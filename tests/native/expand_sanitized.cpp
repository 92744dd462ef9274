// cc_expand.cpp under AddressSanitizer / UBSan: every crew size class, both row types, ragged env counts, destinations at every
// alignment, several thread counts — against a naive restatement of observations.py:62-94, with guard bytes around the destination.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../collectivecrossing_b200/csrc/cc_expand.cpp"

int cc_fail(int code, const char *, ...) { return code; }

template <typename T>
static int check(int B, int E, int64_t n, int misalign, int threads, unsigned seed) {
    cc_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.num_boarding = B; cfg.num_exiting = E; cfg.door_left = 7; cfg.door_right = 10; cfg.division_y = 4;
    const int A = B + E, L = 6 + 4 * A;
    std::vector<int8_t> table((size_t)n * A * 4);
    for (auto &v : table) { seed = seed * 1664525u + 1013904223u; v = (int8_t)((seed >> 16) % 100) - 1; }
    const size_t bytes = (size_t)n * A * L * sizeof(T), guard = 192;
    std::vector<unsigned char> raw(bytes + 2 * guard + 64 + sizeof(T), 0xA5);
    unsigned char *base = raw.data() + guard;
    base += (64 - (reinterpret_cast<uintptr_t>(base) & 63)) & 63;
    base += misalign * (int)sizeof(T);   // rows of T stay T-aligned, as in any array of T
    if (cc_expand_obs_host(&cfg, n, table.data(), base, sizeof(T) == 4 ? CC_OBS_FP32 : CC_OBS_INT8, threads) != CC_OK) return 1;
    const T *obs = reinterpret_cast<const T *>(base);
    for (int64_t e = 0; e < n; ++e)
        for (int i = 0; i < A; ++i) {
            const T *row = obs + ((size_t)e * A + i) * L;
            const int8_t *t = table.data() + (size_t)e * A * 4;
            bool ok = row[0] == (T)t[4 * i] && row[1] == (T)t[4 * i + 1] && row[2] == (T)8 && row[3] == (T)4 && row[4] == (T)7 && row[5] == (T)10;
            for (int j = 0; j < A && ok; ++j)
                for (int k = 0; k < 4; ++k) ok = ok && row[6 + 4 * j + k] == (j == i ? (T)-1 : (T)t[4 * j + k]);
            if (!ok) { printf("A=%d n=%lld misalign=%d threads=%d: env %lld row %d differs\n", A, (long long)n, misalign, threads, (long long)e, i); return 1; }
        }
    for (size_t k = 0; k < raw.size(); ++k) {
        const unsigned char *p = raw.data() + k;
        if ((p < base || p >= base + bytes) && *p != 0xA5) { printf("A=%d n=%lld misalign=%d: byte outside the destination written\n", A, (long long)n, misalign); return 1; }
    }
    return 0;
}

int main() {
    const int crews[][2] = {{1, 0}, {3, 2}, {5, 3}, {7, 5}, {48, 16}, {100, 28}};
    int runs = 0;
    for (const auto &c : crews)
        for (int64_t n : {1, 33, 2049, 5000})
            for (int mis : {0, 1, 3, 7})
                for (int threads : {1, 3}) {
                    if ((c[0] + c[1]) > 16 && n > 100) continue;   // big crews: small batches are enough
                    if (check<float>(c[0], c[1], n, mis, threads, 7u * (unsigned)n + (unsigned)mis)) return 1;
                    if (check<int8_t>(c[0], c[1], n, mis, threads, 11u * (unsigned)n + (unsigned)mis)) return 1;
                    runs += 2;
                }
    printf("ok %d runs\n", runs);
    return 0;
}
