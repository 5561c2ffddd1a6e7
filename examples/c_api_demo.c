/* c_api_demo.c -- the C ABI of libg2048.so from plain C (no Python, no torch): host-buffer context API.
 *   gcc -O2 -Iinclude examples/c_api_demo.c -L2048_q-learning_b200 -lg2048 -Wl,-rpath,'$ORIGIN/../2048_q-learning_b200' -o examples/c_api_demo
 * Runs 4,096 penalty-flavour envs: reset, a few explicit choose_action / step / update_q_value rounds (the loop of
 * QLearningBase/Agent/main.py:91-101, batched), then a fused rollout, and prints counters.  Exit code 0 = ok. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "g2048.h"

#define N 4096
#define CHECK(call)                                                          \
    do {                                                                     \
        int rc_ = (call);                                                    \
        if (rc_ != 0) {                                                      \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, g2048_last_error()); \
            return 1;                                                        \
        }                                                                    \
    } while (0)

int main(void) {
    if (g2048_device_count() < 1) {
        fprintf(stderr, "no CUDA device: libg2048 has no CPU fallback\n");
        return 2;
    }
    g2048_ctx* ctx = g2048_ctx_create(0, N, 1u << 23);
    if (!ctx) {
        fprintf(stderr, "g2048_ctx_create: %s\n", g2048_last_error());
        return 1;
    }
    static uint64_t boards[N], aux[N], prev[N];
    static int32_t score[N], move_score[N];
    static uint8_t actions[N], flags[N], maxlvl[N], done[N];
    static double reward[N];
    static float reward32[N];
    const uint64_t seed = 0x2048;
    for (int i = 0; i < N; ++i) aux[i] = G2048_AUX_INIT;
    CHECK(g2048_ctx_env_reset(ctx, boards, score, NULL, NULL, N, seed, 0, 0));
    for (uint64_t t = 0; t < 32; ++t) {
        memcpy(prev, boards, sizeof boards);
        CHECK(g2048_ctx_choose_action(ctx, boards, actions, N, 0.5, seed, t, 0));
        CHECK(g2048_ctx_env_step(ctx, boards, aux, score, actions, NULL, reward, flags, maxlvl, move_score, N,
                                 G2048_FLAVOUR_PENALTY, seed, t, 0));
        for (int i = 0; i < N; ++i) {
            reward32[i] = (float)reward[i];
            done[i] = (flags[i] & G2048_FLAG_DONE) != 0;
        }
        CHECK(g2048_ctx_qtable_update(ctx, prev, actions, reward32, boards, done, N, 0.1f, 0.99f, G2048_MODE_DETERMINISTIC));
        CHECK(g2048_ctx_env_reset(ctx, boards, score, done, NULL, N, seed, t + 1, 0));
    }
    int64_t states = g2048_ctx_qtable_size(ctx);
    int64_t counters[G2048_N_COUNTERS];
    CHECK(g2048_ctx_rollout_qlearn(ctx, boards, aux, score, N, 256, G2048_FLAVOUR_PENALTY, 0.1f, 0.99f, 0.1, seed, 32, 0,
                                   counters));
    printf("explicit loop: %lld states; fused rollout: %lld steps, %lld episodes, max tile %lld, %lld states now\n",
           (long long)states, (long long)counters[G2048_C_STEPS], (long long)counters[G2048_C_EPISODES],
           1ll << counters[G2048_C_MAXLVL], (long long)g2048_ctx_qtable_size(ctx));
    int ok = counters[G2048_C_STEPS] == (int64_t)N * 256 && states > 0 && counters[G2048_C_DROPPED] == 0;
    g2048_ctx_destroy(ctx);
    return ok ? 0 : 3;
}
