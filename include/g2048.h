/*
 * g2048.h -- C ABI of libg2048.so: the B200 (sm_100a) batched 2048 environment,
 * epsilon-greedy action selection and tabular Q-learning update.
 *
 * This is the drop-in boundary for the hot path of Rocco9999/2048_Q-Learning.
 * The reference is pure Python and has no FFI of its own; each entry point cites
 * the reference interface it replaces (paths relative to the reference root):
 *   ENV-P  QLearningBase/environment/Game2048_env.py            (penalty flavour)
 *   ENV-N  Deep_QLearning/environment/Game2048_nopenalty_env.py (nopenalty flavour)
 *   AGENT  QLearningBase/Agent/main.py
 *   DQN    Deep_QLearning/main_dir/Dqn8TestNOPERCNN.py, mainDQL_CNN_step2.py
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference adds.
 *
 * Conventions
 *  - Plain pointers and sizes only.  g2048_* functions taking `stream` work on
 *    DEVICE pointers owned by the caller (e.g. torch tensors' data_ptr()) and are
 *    asynchronous on that CUDA stream (a cudaStream_t passed as void*, NULL = the
 *    default stream).  g2048_ctx_* functions take HOST pointers, stage through
 *    device buffers owned by the context and return when the results are in host
 *    memory.
 *  - Return value: 0 = ok, > 0 = a cudaError_t, < 0 = a G2048_ERR_* code;
 *    g2048_last_error() describes the last failure of the calling thread.
 *  - No CPU fallback exists: every compute entry point needs a CUDA device.
 *  - A board is a uint64: cell (r,c) = nibble 4r+c = log2(tile), 0 = empty.
 *  - One host thread (or process) per GPU; handles are not thread-safe.
 */
#ifndef G2048_H
#define G2048_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
#define G2048_API extern "C" __attribute__((visibility("default")))
#else
#define G2048_API __attribute__((visibility("default")))
#endif

#define G2048_VERSION 200

/* error codes (< 0) */
#define G2048_ERR_ARG (-1)      /* bad argument (null pointer, bad flavour/mode, capacity not 2^k ...) */
#define G2048_ERR_NOINIT (-2)   /* g2048_init(device) was not called for the current device */
#define G2048_ERR_NOMEM (-3)    /* scratch buffer too small / allocation failed */
#define G2048_ERR_PEER (-4)     /* a peer GPU did not reach a barrier of the exchange in time: the ranks are out of step */

/* env flavours */
#define G2048_FLAVOUR_PENALTY 0    /* ENV-P: shaped float64 reward, stall penalty, lagged done */
#define G2048_FLAVOUR_NOPENALTY 1  /* ENV-N + caller-commit protocol, integer reward, full-board quirk */

/* flags byte written by env_step: bit0 valid, bit1 game_over, bit2 done, bits 4-7 legal-move mask of the
 * resulting board (bit 4+a set iff action a would move it; mainDQL_CNN_step2.py:169-174) */
#define G2048_FLAG_VALID 1
#define G2048_FLAG_GAME_OVER 2
#define G2048_FLAG_DONE 4

/* per-env persistent state of ENV-P (Game2048_env.py:84-95; survives reset(), :187-191), one uint64:
 * bits 0-7 log2(previous_max) | 8-15 consecutive_action (0xFF = None) | 16-23 index into the stall
 * penalty sequence (last_consecutive_penalty) | 32-63 consecutive_count */
#define G2048_AUX_INIT 0x000000000000FF01ull

/* Philox streams: counter = (env id, step, stream), key = seed */
#define G2048_STREAM_STEP 0       /* x0,x1 spawn of the move; x2 epsilon test; x3 random action */
#define G2048_STREAM_RESET 1      /* explicit env_reset: x0..x3 = the two spawns */
#define G2048_STREAM_QUIRK 2      /* reserved (the ENV-N full-board spawn reuses the step's x0, x1) */
#define G2048_STREAM_AUTORESET 3  /* in-rollout reset after done */

/* counters (int64[G2048_N_COUNTERS], accumulated with atomics; the caller zeroes them) */
#define G2048_C_STEPS 0
#define G2048_C_VALID 1
#define G2048_C_EPISODES 2
#define G2048_C_SCORE 3      /* sum of merge scores */
#define G2048_C_MAXLVL 4     /* max log2(tile) seen */
#define G2048_C_REWARD_FX 5  /* sum of trunc(reward * 2^20): order-independent checksum of the rewards */
#define G2048_C_INSERTS 6    /* new Q-table states */
#define G2048_C_DROPPED 7    /* lookups that hit the probe limit (table too full) */
#define G2048_C_LOST 8       /* Q-updates that were NOT applied: the state had no slot (table full).  The fused rollout
                                applies every other update -- a compare-and-swap that loses a race is re-issued on the
                                winner's value -- so this stays 0 unless DROPPED is non-zero */
#define G2048_C_RETRIED 9    /* fused rollout: updates whose first compare-and-swap lost a race against another env and
                                were re-applied (a contention measure; always 0 with one env) */
#define G2048_N_COUNTERS 16  /* 10-15 reserved (zero) */

/* Q-learning modes */
#define G2048_MODE_ATOMIC 0         /* q <- q + lr (target - q) as an atomic CAS loop (order among duplicates unspecified);
                                       no sort: the fast choice for batches with few collisions.  A value hit by thousands
                                       of records at once costs one L2 round trip per record: use DETERMINISTIC there */
#define G2048_MODE_DETERMINISTIC 1  /* sort by (state, action), duplicates applied one after another in ascending env order */

#define G2048_QTABLE_SLOT_BYTES 32  /* key u64 | float q[4] | meta u64 (unused) */

/* one-hot dtypes */
#define G2048_DTYPE_F32 0
#define G2048_DTYPE_BF16 1

/* ------------------------------------------------------------------ library */
G2048_API int g2048_version(void);
G2048_API const char* g2048_last_error(void);
/* Builds the 64K-entry row LUT and the reward tables and uploads them to `device`; idempotent. */
G2048_API int g2048_init(int device);
G2048_API int g2048_device_count(void);
/* Host-only: the tables g2048_init uploads, for inspection and CPU-side tests (no GPU needed).  row[65536] and
 * merged[65536] are returned in PLAIN row order (the device copy is bank-swizzled), mscore[256],
 * reward_valid[16*16*256] indexed (level*16 + d)*256 + score/4, reward_invalid[2*16*16] indexed
 * game_over*256 + level*16 + d, pen[32].  Any pointer may be NULL. */
G2048_API void g2048_host_tables(uint16_t* row, uint8_t* merged, uint32_t* mscore, double* reward_valid,
                                 double* reward_invalid, double* pen);
/* pinned host memory for the ctx API (cudaHostAlloc / cudaFreeHost) */
G2048_API void* g2048_host_alloc(size_t bytes);
G2048_API void g2048_host_free(void* p);

/* ------------------------------------------------------------------ env (device pointers) */
/* ENV-P/ENV-N reset() -> Game2048.__init__ (Game2048_env.py:11-14, :187-191): empty board + two spawns,
 * score = 0, aux untouched.  mask: NULL = all, else only envs with mask[i] != 0.
 * replay_draws: NULL = Philox(seed, env id, episode_idx, STREAM_RESET), else uint8[n][4] =
 * {k_a, is4_a, k_b, is4_b} as drawn by the reference (np.random.randint / random() >= 0.9). */
G2048_API int g2048_env_reset(uint64_t* boards, int32_t* score, const uint8_t* mask, const uint8_t* replay_draws,
                              int64_t n, uint64_t seed, uint64_t episode_idx, uint64_t env_id_base, void* stream);

/* Game2048_env.step(action) for n envs (ENV-P Game2048_env.py:97-129; ENV-N
 * Game2048_nopenalty_env.py:106-120 with the caller's commit of mainDQL_CNN_step2.py:237 folded in).
 * In/out: boards, aux (ENV-P, may be NULL for ENV-N), score (may be NULL).
 * replay_draws: NULL = Philox(seed, env id, step_idx), else uint8[n][4] = {k, is4, k_quirk, is4_quirk}
 * recorded from the reference (bit-exact replay).  Outputs (each may be NULL): reward_f64 / reward_f32,
 * flags, maxlvl (log2 of the returned max_number), move_score. */
G2048_API int g2048_env_step(uint64_t* boards, uint64_t* aux, int32_t* score, const uint8_t* actions,
                             const uint8_t* replay_draws, double* reward_f64, float* reward_f32, uint8_t* flags,
                             uint8_t* maxlvl, int32_t* move_score, int64_t n, int flavour, uint64_t seed,
                             uint64_t step_idx, uint64_t env_id_base, void* stream);

/* Game2048.move(action, trial=True) (Game2048_nopenalty_env.py:53-66): the move without the spawn.
 * out_boards (may be NULL) receives the moved boards, moved[n] / move_score[n] (each may be NULL) the
 * reference's (moved, score) return pair.  boards_in is not modified. */
G2048_API int g2048_move_trial(const uint64_t* boards_in, const uint8_t* actions, uint64_t* out_boards, uint8_t* moved,
                               int32_t* move_score, int64_t n, void* stream);

/* game.move(a, trial=True) for a in 0..3 -> 4-bit mask (mainDQL_CNN_step2.py:169-174); mask 0 == is_game_over */
G2048_API int g2048_legal_mask(const uint64_t* boards, uint8_t* out, int64_t n, void* stream);

/* np.int64[n][16] raw tile values (the reference's board arrays) <-> packed boards.  bad_count (device
 * int64, may be NULL) counts cells that are not 0 or a power of two in 2..32768. */
G2048_API int g2048_pack_i64(const int64_t* tiles, uint64_t* boards, int64_t n, int64_t* bad_count, void* stream);
G2048_API int g2048_unpack_i64(const uint64_t* boards, int64_t* tiles, int64_t n, void* stream);

/* DQNAgent.encode_state (Dqn8TestNOPERCNN.py:271-277): out[n][16 level][4][4], float32 or bfloat16 */
G2048_API int g2048_encode_onehot(const uint64_t* boards, void* out, int64_t n, int dtype, void* stream);

/* DQNAgent.act / act_ripetitive (Dqn8TestNOPERCNN.py:312-336) on network outputs qvalues[n][4]:
 * explore iff x2 < floor(eps * 2^32) (Philox STREAM_STEP); legal_mask NULL = act (uniform over 4 / plain argmax), else
 * act_ripetitive (uniform over legal moves / argmax over legal moves; no legal move = act). */
G2048_API int g2048_select_action(const float* qvalues, const uint8_t* legal_mask, uint8_t* actions, int64_t n,
                                  double eps, uint64_t seed, uint64_t step_idx, uint64_t env_id_base, void* stream);

/* One env-side step of the DQN driver loop (mainDQL_CNN_step2.py:163-237) for n nopenalty envs in ONE launch:
 * act_ripetitive / act on the network outputs qvalues[n][4] (legal_in NULL = act), env.step with the caller's commit,
 * the driver's terminal bonus (:202-213: +100 for a tile >= 2048, +50 for two tiles >= 1024; opts bit 0), reset of
 * the finished games (opts bit 1; Philox(seed, env id, reset_idx, STREAM_RESET) like g2048_env_reset), legal-move
 * mask and one-hot encoding (DQNAgent.encode_state) of the boards the envs continue from.
 * Outputs (each may be NULL): actions, state_out (boards before the step), next_state_out (boards after the step,
 * before any reset -- what `remember` stores), reward, done, legal_out, onehot_out[n][16][4][4] of `dtype`. */
#define G2048_DQN_TERMINAL_BONUS 1u
#define G2048_DQN_AUTO_RESET 2u
G2048_API int g2048_dqn_env_step(uint64_t* boards, int32_t* score, const float* qvalues, const uint8_t* legal_in,
                                 uint8_t* actions, uint64_t* state_out, uint64_t* next_state_out, float* reward,
                                 uint8_t* done, uint8_t* legal_out, void* onehot_out, int dtype, int64_t n, double eps,
                                 uint32_t opts, uint64_t seed, uint64_t step_idx, uint64_t reset_idx,
                                 uint64_t env_id_base, void* stream);

/* ------------------------------------------------------------------ fused rollouts (boards in registers) */
/* k_steps env steps per env under the uniform-random policy (action = x3 >> 30), in-kernel reset on done. */
G2048_API int g2048_rollout_random(uint64_t* boards, uint64_t* aux, int32_t* score, int64_t n, int64_t k_steps,
                                   int flavour, uint64_t seed, uint64_t step_base, uint64_t env_id_base,
                                   int64_t* counters, void* stream);

/* The loop of main.py:91-101 for n envs and k_steps steps each, fused: epsilon-greedy choose_action
 * (main.py:34-38) -> env step -> update_q_value (main.py:40-43) on the HBM hash table, asynchronously: each env
 * applies q <- q + lr (target - q) with ONE atomic from the value it read (a state that is not in the table yet is
 * inserted together with its first update by a 128-bit compare-and-swap).  EVERY update is applied exactly once: one
 * that loses its race against another env is applied to the winner's value -- in place for launches of fewer than
 * 16,384 envs, else through a list that the same call groups by address and applies after the rollout, every group as
 * one sequential chain (still asynchronous on `stream`: when the stream has passed the call, the table is complete).
 * counters[G2048_C_RETRIED] counts those second applications, counters[G2048_C_LOST] stays 0 unless the table is full.
 * Nothing stale is ever summed, no thread spins on a contended value.  With one env nothing races: N = 1 is exactly
 * the reference's sequential order.  Envs that finish (done) are reset in place (STREAM_AUTORESET); every state an
 * env reads is in the table when the call is done (defaultdict semantics, main.py:16).
 * The list and its work buffers belong to the device (about 36 bytes per 4 env steps of the launch, allocated on
 * first use and grown on demand -- the only calls that may synchronise the device); up to four such launches may be in
 * flight per device at a time. */
G2048_API int g2048_rollout_qlearn(uint64_t* boards, uint64_t* aux, int32_t* score, void* table, uint64_t capacity,
                                   int64_t n, int64_t k_steps, int flavour, float lr, float gamma, double eps,
                                   uint64_t seed, uint64_t step_base, uint64_t env_id_base, int64_t* counters,
                                   void* stream);

/* g2048_rollout_qlearn on ONE Q-table spread over the GPUs of the box: the global slot range is cut into n_shards
 * (a power of two) shards of slots_per_shard (a power of two) slots, shards[j] (HOST array of device pointers) being
 * shard j -- this GPU's own memory or a peer's, mapped with g2048_peer_open.  Every GPU runs the call on its own env
 * shard at the same time; lookups are plain loads and updates single atomic compare-and-swaps that travel over NVLink 5 /
 * NVSwitch to the owner's L2 (system scope), so all envs of the box learn the same table (q_table, main.py:16) with
 * no exchange step at all; lost races are deferred and applied like in g2048_rollout_qlearn (every update is applied).  n_shards * slots_per_shard <= 2^31.  Each shard is an ordinary slot array: clear, size and
 * export it with the g2048_qtable_* calls on its owner. */
G2048_API int g2048_rollout_qlearn_sharded(uint64_t* boards, uint64_t* aux, int32_t* score, const void* const* shards,
                                           int n_shards, uint64_t slots_per_shard, int64_t n, int64_t k_steps,
                                           int flavour, float lr, float gamma, double eps, uint64_t seed,
                                           uint64_t step_base, uint64_t env_id_base, int64_t* counters, void* stream);
/* q_table[state] (main.py:16) on a sharded table; see g2048_qtable_lookup. */
G2048_API int g2048_qtable_lookup_sharded(const void* const* shards, int n_shards, uint64_t slots_per_shard,
                                          const uint64_t* keys, int64_t n, float* rows, uint8_t* found, int insert,
                                          void* stream);

/* One synchronous batched Q-learning step: every env chooses from and bootstraps
 * on the table as it is at step start (target_i = r_i + gamma max Q[s'_i] (1 - done_i)); afterwards every
 * (state, action) receives its targets one after another, q <- q + lr (target_i - q) -- the reference's own
 * update applied in sequence -- in ascending env order (DETERMINISTIC) or in unspecified order (ATOMIC).
 * rec_key/rec_action/rec_target (each may be NULL) export the (state, action, target) records, e.g. for the
 * cross-GPU exchange; apply = 0 only emits them and leaves the Q values untouched. */
G2048_API size_t g2048_qlearn_scratch_bytes(int64_t n);
G2048_API int g2048_qlearn_step(uint64_t* boards, uint64_t* aux, int32_t* score, void* table, uint64_t capacity,
                                int64_t n, int flavour, float lr, float gamma, double eps, int mode, int apply,
                                uint64_t seed, uint64_t step_idx, uint64_t env_id_base, int64_t* counters,
                                uint64_t* rec_key, uint8_t* rec_action, float* rec_target, void* scratch,
                                size_t scratch_bytes, void* stream);

/* ---------------------------------------------- synchronous step across GPUs: records over NVLink peer memory */
/* One transition of a synchronous step as it travels between GPUs (16 bytes): the state key, and the action
 * (bits 0-1) with the float32 bits of the TD target in bits 32-63. */
typedef struct g2048_record {
    uint64_t key;
    uint64_t action_target;
} g2048_record;
#define G2048_MAX_PEERS 16
#define G2048_IPC_HANDLE_BYTES 64

/* g2048_qlearn_step with apply = 0, writing the records packed (16-byte aligned `records[n]`): choose_action
 * (main.py:34-38) + env step + the target of update_q_value (main.py:41-42) for every env, table values untouched. */
G2048_API int g2048_qlearn_emit(uint64_t* boards, uint64_t* aux, int32_t* score, void* table, uint64_t capacity,
                                int64_t n, int flavour, float gamma, double eps, uint64_t seed, uint64_t step_idx,
                                uint64_t env_id_base, int64_t* counters, g2048_record* records, void* stream);
/* Apply n_lists record lists (HOST arrays `lists[j]` = device pointer, `counts[j]` = records in it), list after
 * list and record after record: q <- q + lr (target - q) (main.py:43).  The pointers may be peer memory of other
 * GPUs (g2048_peer_open): the kernel reads each record directly from its owner over NVLink, so the "all-gather"
 * of the exchange step and the table lookups of the apply are one kernel and no gathered copy exists.  With the
 * lists in rank order the DETERMINISTIC result equals the single-GPU step over all envs. scratch: sized by
 * g2048_qlearn_scratch_bytes(sum of counts). */
G2048_API int g2048_qtable_apply_records(void* table, uint64_t capacity, const g2048_record* const* lists,
                                         const int64_t* counts, int n_lists, float lr, int mode, void* scratch,
                                         size_t scratch_bytes, void* stream);
/* The exact synchronous step on ONE table sharded over the GPUs (see g2048_rollout_qlearn_sharded for the shard
 * list), owner computes: g2048_qlearn_emit_owned advances the envs like g2048_qlearn_emit, looking s and s' up
 * wherever their slots live, and appends the record of each transition to owner_lists[j] (HOST array of n_shards
 * device pointers, room for n records each -- K * n when K steps are exchanged at once) of the GPU j that owns the slot of s; owner_counts[j] (device, zeroed by
 * the caller) counts them.  A record holds ((owner-local slot * 4 + action) << idx_bits) | (record_index_base + i)
 * and the float32 target; idx_bits >= log2(total envs of the job), record_index_base = first global env index of
 * this rank.  After a barrier, g2048_qtable_apply_owned on GPU j sorts the lists every rank wrote for j (read in place,
 * local or peer memory; counts[] on the HOST, e.g. from g2048_peer_read_u64) and applies them to ITS shard, record
 * after record in global env order: q <- q + lr (target - q) (main.py:43).  The result equals the single-GPU
 * deterministic step bit for bit, and every GPU sorts only its share of the records.  A second barrier must separate
 * the apply from the next emit (which reads remote shards).
 * Exchange every K steps: call emit K times (record_index_base = k * total envs + first env of the rank, idx_bits for
 * K * total envs, counts zeroed once) before the barrier and the apply; the table's values are then frozen for the K
 * steps.  carry_slot[n] / carry_row[n][4] (device, both or neither) receive the slot and row of the state every env
 * continues from; with use_carry = 1 (steps 2..K of a window) the kernel takes them from there instead of looking the
 * state up again -- one remote request less per env step. */
G2048_API int g2048_qlearn_emit_owned(uint64_t* boards, uint64_t* aux, int32_t* score, const void* const* shards,
                                      int n_shards, uint64_t slots_per_shard, int64_t n, int flavour, float gamma,
                                      double eps, uint64_t seed, uint64_t step_idx, uint64_t env_id_base,
                                      uint64_t record_index_base, int idx_bits, int64_t* counters,
                                      g2048_record* const* owner_lists, uint64_t* owner_counts, uint32_t* carry_slot,
                                      float* carry_row, int use_carry, void* stream);
G2048_API int g2048_qtable_apply_owned(void* shard, uint64_t slots_per_shard, const g2048_record* const* lists,
                                       const int64_t* counts, int n_lists, int idx_bits, float lr, void* scratch,
                                       size_t scratch_bytes, void* stream);
/* n 8-byte values at device (local or peer) addresses src[j] -> host_out[j]; waits for the stream. */
G2048_API int g2048_peer_read_u64(const uint64_t* const* src, int n, uint64_t* host_out, void* stream);
G2048_API int g2048_peer_memset(void* dev_ptr, int value, size_t bytes, void* stream);

/* Device memory other processes of this box can map (CUDA IPC): alloc (zero-filled) + 64-byte handle to send to
 * the peers; open/close on the peers' side; free by the owner. */
G2048_API int g2048_peer_alloc(size_t bytes, void** dev_ptr, void* ipc_handle_out);
G2048_API int g2048_peer_open(const void* ipc_handle, void** dev_ptr);
G2048_API int g2048_peer_close(void* dev_ptr);
G2048_API int g2048_peer_free(void* dev_ptr);
/* Stream-ordered barrier between the `world` GPUs: flags[j] (HOST array of device pointers) = rank j's flag
 * block (uint64[world], zero at start, in peer memory); epoch must grow by one per barrier.  Work enqueued
 * before it on every rank is visible to work enqueued after it on every rank.  If a peer does not arrive within
 * timeout_ns (0 = 5 s), *timed_out (an int in device or pinned host memory, may be NULL) is set to 1 + the missing
 * rank and the kernel returns instead of hanging.  A time-out is fatal for the exchange: while the flag is set,
 * g2048_qtable_apply_records and g2048_qtable_apply_owned on this device apply NOTHING (the peers' lists are not
 * valid), and the caller must stop (dist.py raises at the step). */
G2048_API int g2048_peer_barrier(uint64_t* const* flags, int rank, int world, uint64_t epoch, uint64_t timeout_ns,
                                 int* timed_out, void* stream);

/* The same exact step, ROUTED (round 2): no GPU ever touches another GPU's shard.  The owner of a state is the GPU that
 * holds its home slot, owner(key) = top bits of the global home slot, exactly as in the sharded table above; a probe
 * sequence stays inside the shard.  Per step (update_q_value of every env on ONE table, main.py:40-43; choose_action on
 * the table as it stands at step start, main.py:34-38):
 *   1. every env takes (slot, row) of its state from the owner's last answer, chooses, steps, and appends the key of s'
 *      (after a game over also the fresh board) to its list for the owner of that key          -- barrier --
 *   2. every owner pulls the lists written for it (coalesced NVLink reads), finds-or-inserts the keys in its own shard
 *      and pushes {slot, max Q} into the requester's answer buffer (coalesced NVLink writes)   -- barrier --
 *   3. every env forms r + gamma max Q(s') and pushes its record (one 8-byte word) into the sort input of the owner
 *      of s, at a place that makes the input ascending in the global env index                  -- barrier --
 *   4. every owner sorts (stable, on the slot and action bits) and applies the records for its shard, each (s, a) in
 *      ascending global env index, and pushes the rows as they are now for every request of 2.   -- barrier --
 * Same table as the single-GPU deterministic g2048_qlearn_step, bit for bit (states with a non-zero value; the fresh
 * board after a game over is inserted one step earlier here).  Only bulk lists cross NVLink: 40 bytes per env step (key, {slot, max Q},
 * 8-byte record, row).
 *
 * g2048_routed_buffer_bytes(world, cap): size of the zero-filled buffer every rank must allocate with
 * g2048_peer_alloc and share with all peers (cap = the largest env count of any rank).  g2048_routed_create: rank's
 * view; peer_buffers[j] (HOST array) = rank j's buffer as mapped in this process, `shard` = this rank's
 * slots_per_shard * 32 bytes of table (slots_per_shard <= 2^30), n_total = envs of all ranks (global env ids must stay below it; rank j's env
 * ids must all be smaller than rank j + 1's -- the owners apply the lists in rank order, which is then env order).
 * g2048_routed_prime: looks the envs' current boards up (call once after reset, on every rank, before the first step;
 * again whenever the boards were changed from outside).  g2048_routed_step: one env step of this rank's n envs; every
 * rank must call it the same number of times; *applied (host, may be NULL) = records applied to this rank's shard.
 * The step synchronises the stream once (the sort needs the record count).  A peer that does not reach a barrier
 * within 5 s makes the call (or the next one) return G2048_ERR_PEER; nothing of that step is applied on this GPU. */
typedef struct g2048_routed g2048_routed;
G2048_API size_t g2048_routed_buffer_bytes(int world, int64_t cap);
G2048_API g2048_routed* g2048_routed_create(int rank, int world, int64_t cap, int64_t n_total, void* const* peer_buffers,
                                            void* shard, uint64_t slots_per_shard);
G2048_API void g2048_routed_destroy(g2048_routed* r);
G2048_API int g2048_routed_prime(g2048_routed* r, const uint64_t* boards, int64_t n, void* stream);
G2048_API int g2048_routed_step(g2048_routed* r, uint64_t* boards, uint64_t* aux, int32_t* score, int64_t n, int flavour,
                                float lr, float gamma, double eps, uint64_t seed, uint64_t step_idx, uint64_t env_id_base,
                                int64_t* counters, int64_t* applied, void* stream);

/* ------------------------------------------------------------------ Q-table (device pointers) */
/* QLearningAgent.q_table (main.py:16): `table` is capacity * 32 bytes of device memory, 32-byte aligned,
 * capacity = 2^k <= 2^31 slots (64 GiB). */
G2048_API size_t g2048_qtable_bytes(uint64_t capacity);
G2048_API int g2048_qtable_clear(void* table, uint64_t capacity, void* stream);
/* q_table[state] for n states -> rows[n][4]; found[n] (may be NULL); insert != 0 creates missing zero rows
 * like the reference's defaultdict. */
G2048_API int g2048_qtable_lookup(void* table, uint64_t capacity, const uint64_t* keys, int64_t n, float* rows,
                                  uint8_t* found, int insert, void* stream);
/* choose_action (main.py:34-38) for n states with Philox(seed, env id, step_idx) draws */
G2048_API int g2048_choose_action(void* table, uint64_t capacity, const uint64_t* boards, uint8_t* actions, int64_t n,
                                  double eps, uint64_t seed, uint64_t step_idx, uint64_t env_id_base, void* stream);
/* update_q_value (main.py:40-43) for n transitions as ONE synchronous batch (N = 1: the reference). */
G2048_API int g2048_qtable_update(void* table, uint64_t capacity, const uint64_t* s, const uint8_t* a, const float* r,
                                  const uint64_t* s2, const uint8_t* done, int64_t n, float lr, float gamma, int mode,
                                  void* scratch, size_t scratch_bytes, void* stream);
/* Q[key][a] <- Q + lr (target - Q) for n records (e.g. all-gathered from the other ranks), same modes. */
G2048_API int g2048_qtable_apply_targets(void* table, uint64_t capacity, const uint64_t* keys, const uint8_t* a,
                                         const float* target, int64_t n, float lr, int mode, void* scratch,
                                         size_t scratch_bytes, void* stream);
/* number of states -> *count (device int64) */
G2048_API int g2048_qtable_size(const void* table, uint64_t capacity, int64_t* count, void* stream);
/* stats[3] (device int64): number of states, sum over the states of the probes that come BEFORE their slot in their
 * key's probe sequence (home slot, its partner in the same 64 bytes, then the next pair ...), and the largest such
 * number; a lookup of a stored state costs 1 + that many probes, so mean probe length = 1 + sum / n. */
G2048_API int g2048_qtable_probe_stats(const void* table, uint64_t capacity, int64_t* stats, void* stream);
/* compact (key, row) pairs into keys[max_out], rows[max_out][4]; *count (device int64, zeroed by the caller)
 * receives the number of states (may exceed max_out: the excess is not written) */
G2048_API int g2048_qtable_export(const void* table, uint64_t capacity, uint64_t* keys, float* rows, int64_t max_out,
                                  int64_t* count, void* stream);

/* ------------------------------------------------------------------ host-buffer API */
typedef struct g2048_ctx g2048_ctx;
/* device buffers for up to max_envs envs, one stream, and (table_capacity > 0) a Q-table in HBM.  Contexts for up to
 * 64 envs (the N = 1 drop-in adapters) stage through one pinned, mapped host block that the kernels read and write
 * directly, so a call costs its kernel and one stream synchronisation instead of a dozen small copies. */
G2048_API g2048_ctx* g2048_ctx_create(int device, int64_t max_envs, uint64_t table_capacity);
G2048_API void g2048_ctx_destroy(g2048_ctx* ctx);
G2048_API int g2048_ctx_env_reset(g2048_ctx* ctx, uint64_t* boards, int32_t* score, const uint8_t* mask,
                                  const uint8_t* replay_draws, int64_t n, uint64_t seed, uint64_t episode_idx,
                                  uint64_t env_id_base);
G2048_API int g2048_ctx_env_step(g2048_ctx* ctx, uint64_t* boards, uint64_t* aux, int32_t* score,
                                 const uint8_t* actions, const uint8_t* replay_draws, double* reward_f64,
                                 uint8_t* flags, uint8_t* maxlvl, int32_t* move_score, int64_t n, int flavour,
                                 uint64_t seed, uint64_t step_idx, uint64_t env_id_base);
G2048_API int g2048_ctx_legal_mask(g2048_ctx* ctx, const uint64_t* boards, uint8_t* out, int64_t n);
G2048_API int g2048_ctx_move_trial(g2048_ctx* ctx, const uint64_t* boards_in, const uint8_t* actions,
                                   uint64_t* out_boards, uint8_t* moved, int32_t* move_score, int64_t n);
G2048_API int g2048_ctx_rollout_random(g2048_ctx* ctx, uint64_t* boards, uint64_t* aux, int32_t* score, int64_t n,
                                       int64_t k_steps, int flavour, uint64_t seed, uint64_t step_base,
                                       uint64_t env_id_base, int64_t* counters);
G2048_API int g2048_ctx_rollout_qlearn(g2048_ctx* ctx, uint64_t* boards, uint64_t* aux, int32_t* score, int64_t n,
                                       int64_t k_steps, int flavour, float lr, float gamma, double eps, uint64_t seed,
                                       uint64_t step_base, uint64_t env_id_base, int64_t* counters);
G2048_API int g2048_ctx_qtable_lookup(g2048_ctx* ctx, const uint64_t* keys, int64_t n, float* rows, uint8_t* found,
                                      int insert);
G2048_API int g2048_ctx_choose_action(g2048_ctx* ctx, const uint64_t* boards, uint8_t* actions, int64_t n, double eps,
                                      uint64_t seed, uint64_t step_idx, uint64_t env_id_base);
G2048_API int g2048_ctx_qtable_update(g2048_ctx* ctx, const uint64_t* s, const uint8_t* a, const float* r,
                                      const uint64_t* s2, const uint8_t* done, int64_t n, float lr, float gamma,
                                      int mode);
G2048_API int64_t g2048_ctx_qtable_size(g2048_ctx* ctx);
G2048_API int64_t g2048_ctx_qtable_export(g2048_ctx* ctx, uint64_t* keys, float* rows, int64_t max_out);
G2048_API int g2048_ctx_qtable_clear(g2048_ctx* ctx);
/* raw handles for callers that mix both APIs */
G2048_API void* g2048_ctx_table(g2048_ctx* ctx);
G2048_API uint64_t g2048_ctx_table_capacity(g2048_ctx* ctx);
G2048_API void* g2048_ctx_stream(g2048_ctx* ctx);

#endif /* G2048_H */
