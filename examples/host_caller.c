/* host_caller.c — the C ABI (include/ccb200.h) driven by a plain C program: no CUDA headers, no Python, no torch.
 *
 * What the reference's rollout scripts do with one Python env (scripts/run_greedy_policy_demo.py:60-95:
 * policy -> env.step -> reset on done) for N envs on a B200, with every buffer in ordinary host memory:
 *     cc_create -> cc_reset -> K x cc_step_host(policy in the kernel, auto-reset) -> cc_stats_read
 * and, with `--dump FILE`, the final state, the last step's outputs and the statistics written to FILE so that
 * tests/test_gpu_c_caller.py can compare them with the oracle bit for bit.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/host_caller.c -Lcollectivecrossing_b200/csrc -lccb200 \
 *       -Wl,-rpath,$PWD/collectivecrossing_b200/csrc -o host_caller && ./host_caller 65536 200
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "ccb200.h"

#define CHECK(call)                                                                     \
    do {                                                                                \
        int rc_ = (call);                                                               \
        if (rc_ != CC_OK) {                                                             \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, cc_last_error());             \
            return 1;                                                                   \
        }                                                                               \
    } while (0)

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int main(int argc, char **argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 4096;
    const int steps = argc > 2 ? atoi(argv[2]) : 100;
    const uint64_t seed = argc > 3 ? strtoull(argv[3], NULL, 10) : 7;
    const char *dump = (argc > 5 && strcmp(argv[4], "--dump") == 0) ? argv[5] : NULL;
    if (n < 1 || steps < 1) { fprintf(stderr, "usage: %s [envs] [steps] [seed] [--dump file]\n", argv[0]); return 2; }

    /* the README environment (reference README.md:48-64), lowered to absolute geometry as utils/geometry.py:20-47 does:
     * centre 12/2 = 6, tram_length 9 -> tram x in [2, 10]; door 5..7 relative -> 7..9 */
    cc_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.width = 12; cfg.height = 8; cfg.division_y = 4;
    cfg.tram_left = 2; cfg.tram_right = 10; cfg.door_left = 7; cfg.door_right = 9;
    cfg.boarding_dest_y = 8; cfg.exiting_dest_y = 0;
    cfg.num_boarding = 5; cfg.num_exiting = 3;
    cfg.max_steps = 100;
    cfg.reward_kind = CC_REWARD_DEFAULT;
    cfg.terminated_kind = CC_TERM_INDIVIDUAL_AT_DESTINATION;
    cfg.reward_params[0] = 15.0; cfg.reward_params[1] = 10.0; cfg.reward_params[2] = 5.0; cfg.reward_params[3] = 0.1;
    const int A = cfg.num_boarding + cfg.num_exiting, L = 6 + 4 * A;

    if (cc_abi_version() != CCB200_ABI_VERSION) { fprintf(stderr, "header / library ABI mismatch\n"); return 1; }
    cc_handle *h = NULL;
    CHECK(cc_create(&cfg, n, 0, 0, seed, &h));
    CHECK(cc_reset(h, NULL, NULL, CC_OBS_NONE, NULL));

    /* ordinary (pageable) host memory: what a numpy caller hands over */
    float *obs = malloc((size_t)n * A * L * sizeof(float)), *reward = malloc((size_t)n * A * sizeof(float));
    uint8_t *agent_flags = malloc((size_t)n * A), *env_flags = malloc((size_t)n);
    int8_t *applied = malloc((size_t)n * A);
    if (!obs || !reward || !agent_flags || !env_flags || !applied) { fprintf(stderr, "out of memory\n"); return 1; }
    cc_step_io io;
    memset(&io, 0, sizeof io);
    io.obs = obs; io.reward = reward; io.agent_flags = agent_flags; io.env_flags = env_flags; io.actions_out = applied;
    io.obs_dtype = CC_OBS_FP32; io.reward_dtype = CC_REWARD_F32;
    io.policy = CC_POLICY_GREEDY;   /* baseline_policies/greedy_policy.py at epsilon 0, evaluated inside the step kernel */
    io.auto_reset = 1;

    const double t0 = now_s();
    for (int t = 0; t < steps; ++t) CHECK(cc_step_host(h, &io));
    const double dt = now_s() - t0;
    CHECK(cc_check_error(h, NULL));

    /* every row is [x_i, y_i, door centre, division_y, door left, door right, blocks...] with the own block masked (observations.py:62-94) */
    for (int64_t e = 0; e < n; ++e)
        for (int i = 0; i < A; ++i) {
            const float *row = obs + ((size_t)e * A + i) * L;
            if (row[2] != 8.f || row[3] != 4.f || row[4] != 7.f || row[5] != 9.f || row[6 + 4 * i] != -1.f || row[6 + 4 * i + 3] != -1.f) {
                fprintf(stderr, "env %lld agent %d: malformed observation row\n", (long long)e, i);
                return 1;
            }
        }
    cc_stats st;
    CHECK(cc_stats_read(h, &st, NULL));
    int64_t call[5];
    CHECK(cc_last_host_call(h, call));
    printf("%lld envs x %d steps: %.3f ms per step, %.1f M agent-steps/s through host buffers (%s; %lld chunks, %lld host threads)\n",
           (long long)n, steps, 1e3 * dt / steps, 1e-6 * (double)n * A * steps / dt, cc_last_kernel_name(h), (long long)call[0], (long long)call[2]);
    printf("episodes %lld (terminated %lld, truncated %lld), arrivals %lld, mean length %.2f, mean return %.3f\n", (long long)st.episodes,
           (long long)st.terminated_all, (long long)st.truncated_all, (long long)st.arrivals,
           st.episodes ? (double)st.episode_length_sum / (double)st.episodes : 0.0, st.episodes ? st.episode_return_sum / (double)st.episodes : 0.0);

    if (dump) {
        int8_t *x = malloc((size_t)n * A), *y = malloc((size_t)n * A);
        uint8_t *flags = malloc((size_t)n * A);
        int32_t *step = malloc((size_t)n * sizeof(int32_t));
        if (!x || !y || !flags || !step) { fprintf(stderr, "out of memory\n"); return 1; }
        CHECK(cc_get_state_host(h, x, y, flags, step));
        FILE *f = fopen(dump, "wb");
        if (!f) { perror(dump); return 1; }
        /* x, y, flags, step | obs, reward, agent_flags, env_flags, applied actions of the last step | stats */
        fwrite(x, 1, (size_t)n * A, f); fwrite(y, 1, (size_t)n * A, f); fwrite(flags, 1, (size_t)n * A, f); fwrite(step, sizeof(int32_t), (size_t)n, f);
        fwrite(obs, sizeof(float), (size_t)n * A * L, f); fwrite(reward, sizeof(float), (size_t)n * A, f);
        fwrite(agent_flags, 1, (size_t)n * A, f); fwrite(env_flags, 1, (size_t)n, f); fwrite(applied, 1, (size_t)n * A, f);
        fwrite(&st, sizeof st, 1, f);
        fclose(f);
        free(x); free(y); free(flags); free(step);
    }
    cc_destroy(h);
    free(obs); free(reward); free(agent_flags); free(env_flags); free(applied);
    return 0;
}
