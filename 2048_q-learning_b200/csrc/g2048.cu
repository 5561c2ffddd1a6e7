// g2048.cu -- kernels and C ABI of libg2048.so (see include/g2048.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC
//
// Kernels
//   k_env_step        one step per env, state in HBM (HBM-bound: 22 / 38 B per step)
//   k_rollout_random  K steps per env with the board in registers, row LUT staged in shared memory by
//                     one bulk-async (TMA) copy, Philox spawn, in-kernel reset
//   k_rollout_qlearn  the same loop with epsilon-greedy choose_action and the TD update on the HBM hash
//                     table fused in: a table visit is one load + ONE atomic (a new state is inserted together with
//                     its first update by a 128-bit CAS), settled one step later; every update is applied
//   k_qlearn_phase_a / k_q_update_phase_a / k_keys_to_records + k_apply_atomic / CUB radix sort +
//   k_segment_apply   synchronous batched update, atomic and deterministic modes
//   k_q_lookup, k_choose_action, k_q_size, k_q_export, k_legal_mask, k_pack, k_unpack, k_onehot,
//   k_select_action
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cuda_bf16.h>

#include "g2048_device.cuh"

using namespace g2048;

// ============================================================================================ host state
namespace {

thread_local char g_err[512] = "";
int fail(int code, const char* what) {
    if (code > 0)
        snprintf(g_err, sizeof g_err, "%s: %s (%s)", what, cudaGetErrorString((cudaError_t)code),
                 cudaGetErrorName((cudaError_t)code));
    else
        snprintf(g_err, sizeof g_err, "%s (code %d)", what, code);
    return code;
}
#define CK(expr)                                       \
    do {                                               \
        cudaError_t e_ = (expr);                       \
        if (e_ != cudaSuccess) return fail((int)e_, #expr); \
    } while (0)

constexpr int kMaxDevices = 64;
constexpr size_t kLutRowBytes = 65536 * sizeof(uint16_t);
constexpr size_t kLutMergedBytes = 65536;
constexpr size_t kLutCoreBytes = kLutRowBytes + kLutMergedBytes + 256 * sizeof(uint32_t);  // row | merged | mscore
constexpr size_t kLutBytes = kLutCoreBytes + kHotDoubles * sizeof(double);               // + hot reward tables
#ifndef G2048_EXP_THREADS
#define G2048_EXP_THREADS 768
#endif
constexpr int kRolloutThreads = 1024;
constexpr int kQlearnThreads = G2048_EXP_THREADS;   // fused Q-learning rollout
constexpr int kSmallRolloutThreads = 128;   // small batches: LUT through L1, more registers per thread
constexpr long long kEnvStepSmemLutMinEnvs = 1 << 19;   // g2048_env_step: stage the LUT in shared memory from this batch size on

struct DeviceState {
    bool ready = false;
    Tables tables{};
    void* lut = nullptr;      // kLutBytes, 128-byte aligned
    int sm_count = 0;
    unsigned long long* queue = nullptr;   // kQueueSlots env-queue heads of the fused Q-learning rollout, one per launch in flight
    unsigned queue_next = 0;
    // deferred-update lists of the fused Q-learning rollout (+ the sort buffers that apply them): a small ring, so that
    // launches on different streams do not share one; grown on demand
    struct DeferSet { void* buf = nullptr; size_t bytes = 0; } defer[4];
    unsigned defer_next = 0;
    int* abort_flag = nullptr;         // device int, set by a g2048_peer_barrier that timed out: peer record lists are not consumed any more
};
constexpr unsigned kQueueSlots = 256;
constexpr int64_t kDeferMinEnvs = 16384;   // fused rollouts of fewer envs settle lost races in place (no list)
DeviceState g_dev[kMaxDevices];
std::mutex g_mu;

// ---- host-side table construction -------------------------------------------------------------------
// Row table: move_left on one row (Game2048_env.py:25-41) for all 65,536 rows.
void build_row_tables(std::vector<uint16_t>& row, std::vector<uint8_t>& merged, std::vector<uint32_t>& mscore) {
    row.resize(65536);
    merged.resize(65536);
    mscore.resize(256);
    for (unsigned m = 0; m < 256; ++m) {  // merged byte -> score (sum of 2^level over its two nibbles) | hi level << 24
        unsigned hi = m >> 4, lo = m & 15;
        mscore[m] = ((hi ? 1u << hi : 0u) + (lo ? 1u << lo : 0u)) | (hi << 24);
    }
    for (unsigned r = 0; r < 65536; ++r) {
        int t[4] = {(int)(r & 15), (int)((r >> 4) & 15), (int)((r >> 8) & 15), (int)((r >> 12) & 15)};
        int packed[4], n = 0;
        for (int c = 0; c < 4; ++c)
            if (t[c]) packed[n++] = t[c];
        int out[4] = {0, 0, 0, 0}, m = 0, mg[2] = {0, 0}, k = 0;
        for (int i = 0; i < n; ++i) {
            bool pair = (i + 1 < n) && packed[i] == packed[i + 1] && packed[i] < 15;  // 2^16 does not fit a nibble
            if (pair) { out[m++] = packed[i] + 1; mg[k++] = packed[i] + 1; ++i; }
            else out[m++] = packed[i];
        }
        unsigned p = lut_index2(r) & 0xFFFFu;  // bank-swizzled position of row r
        row[p] = (uint16_t)(out[0] | (out[1] << 4) | (out[2] << 8) | (out[3] << 12));
        int hi = mg[0] > mg[1] ? mg[0] : mg[1], lo = mg[0] > mg[1] ? mg[1] : mg[0];
        merged[p] = (uint8_t)((hi << 4) | lo);
    }
}
// update_and_normalize (Game2048_env.py:197-205), same libm calls as CPython's math.log2
double normalize_reward(double reward) {
    if (reward >= 0) return std::fmin(std::log2(reward + 1), 10);
    return -std::fmin(std::log2(std::fabs(reward - 1)), 10);
}
// calculate_reward (Game2048_env.py:136-184) tabulated over (level, progress step d, score/4).
// The float64 expression order follows the reference statement by statement so that every entry has the
// bits the Python code produces on the same host.
void build_reward_tables(std::vector<double>& valid, std::vector<double>& invalid, std::vector<double>& pen) {
    valid.assign(16 * 16 * 256, 0.0);
    invalid.assign(2 * 16 * 16, 0.0);
    for (int lvl = 1; lvl < 16; ++lvl)
        for (int d = 0; d < 16; ++d) {
            if (d > lvl - 1) continue;  // previous level = lvl - d >= 1
            double current_level = (double)lvl;
            double bonus = 0;
            if (d > 0) bonus = (current_level - (double)(lvl - d)) * std::pow(current_level, 1.2);  // :148-149
            for (int over = 0; over < 2; ++over) {
                double reward = 0;
                if (over) {
                    if (lvl == 9 || lvl == 10 || lvl == 11) reward = bonus + std::pow(current_level, 1.2);  // :156-158
                    else reward -= std::log2((double)((1ll << lvl) + 1));                                      // :160
                } else {
                    reward -= 0.1 * current_level;                                                             // :164
                }
                invalid[over * 256 + lvl * 16 + d] = normalize_reward(reward);
            }
            for (int s4 = 0; s4 < 256; ++s4) {
                double reward = (double)(4 * s4);                                   // :168
                if (bonus > 0) reward += bonus;                                     // :171-172
                else if (bonus == 0) reward += current_level * 0.05;                // :173-174
                if (lvl >= 9) reward += std::pow(current_level, 1.2) * 2;           // :176-177
                valid[(lvl * 16 + d) * 256 + s4] = normalize_reward(reward);
            }
        }
    pen.assign(32, -10.0);
    double p = -1;                                                                  // :95
    pen[0] = p;
    for (int i = 1; i < 32; ++i) {                                                  // :124
        double q = p * 1.1;
        p = q > -10 ? q : -10;
        pen[i] = p;
    }
}

int current_device_state(DeviceState** out) {
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices || !g_dev[dev].ready) return fail(G2048_ERR_NOINIT, "g2048_init(device) not called");
    *out = &g_dev[dev];
    return 0;
}
inline cudaStream_t S(void* stream) { return (cudaStream_t)stream; }
inline int grid_for(int64_t n, int block, int sm_count, int per_sm = 8) {
    int64_t g = (n + block - 1) / block;
    int64_t cap = (int64_t)sm_count * per_sm;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}
inline u64 eps_threshold(double eps) {
    if (!(eps > 0)) return 0;
    if (eps >= 1) return 1ull << 32;
    return (u64)(eps * 4294967296.0);
}
inline bool pow2(uint64_t c) { return c && !(c & (c - 1)) && c <= (1ull << 31); }   // table capacity: 2^k <= 2^31 slots

}  // namespace

// ============================================================================================ kernels
namespace {

__device__ __forceinline__ Lut global_lut(const Tables& T) { return Lut{T.lut_row, T.lut_merged, T.lut_mscore, nullptr}; }

// Stage the 213 KB LUT blob (row table, merged levels, score table, hot reward tables) into dynamic shared
// memory with one bulk-async copy (TMA, UBLKCP) completing
// on an mbarrier; every thread then waits on phase 0.
__device__ __forceinline__ Lut stage_lut(const Tables& T, unsigned char* smem) {
    __shared__ __align__(8) unsigned long long mbar;
    u32 bar = (u32)__cvta_generic_to_shared(&mbar);
    u32 dst = (u32)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((u32)kLutBytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                     "l"(T.lut_row), "r"((u32)kLutBytes), "r"(bar)
                     : "memory");
    }
    u32 ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar)
            : "memory");
    }
    return Lut{reinterpret_cast<const uint16_t*>(smem), smem + kLutRowBytes,
               reinterpret_cast<const uint32_t*>(smem + kLutRowBytes + kLutMergedBytes),
               reinterpret_cast<const double*>(smem + kLutCoreBytes)};
}

template <bool REPLAY>
__global__ void k_env_reset(u64* boards, int* score, const uint8_t* mask, const uint8_t* draws, long long n, u64 seed,
                            u64 episode, u64 id_base) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (mask && !mask[i]) continue;
        u64 b;
        if (REPLAY) {
            uchar4 d = reinterpret_cast<const uchar4*>(draws)[i];
            b = fresh_board<true>(d.x, d.y, d.z, d.w);
        } else {
            Draw4 x = philox(seed, id_base + (u64)i, episode, G2048_STREAM_RESET);
            b = fresh_board<false>(x.x0, x.x1, x.x2, x.x3);
        }
        boards[i] = b;
        if (score) score[i] = 0;
    }
}

// SMEM_LUT: one persistent 1024-thread CTA per SM with the row tables staged in shared memory (large batches).  Through
// L1 every lane of a LUT gather touches its own line (32 wavefronts per lookup; measured 42-46 G steps/s, L1-bound);
// the swizzled shared-memory copy needs ~2.7 wavefronts per lookup.
template <int FLAVOUR, bool REPLAY, bool SMEM_LUT>
__global__ void __launch_bounds__(SMEM_LUT ? kRolloutThreads : 256, 1)
k_env_step(Tables T, u64* boards, u64* aux, int* score, const uint8_t* actions, const uint8_t* draws, double* rew64,
           float* rew32, uint8_t* flags, uint8_t* maxlvl, int* move_score, long long n, u64 seed, u64 t, u64 id_base) {
    extern __shared__ __align__(128) unsigned char smem[];
    Lut L = SMEM_LUT ? stage_lut(T, smem) : global_lut(T);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        Env e;
        env_load(e, boards[i], (FLAVOUR == G2048_FLAVOUR_PENALTY && aux) ? aux[i] : G2048_AUX_INIT, score ? score[i] : 0);
        int a = actions[i] & 3;
        StepOut o;
        if (REPLAY) {
            uchar4 d = reinterpret_cast<const uchar4*>(draws)[i];
            if (FLAVOUR == G2048_FLAVOUR_PENALTY) penalty_step<true>(e, a, d.x, d.y, L, T, o);
            else nopenalty_step<true>(e, a, d.x, d.y, d.z, d.w, L, o);
        } else {
            Draw4 x = philox(seed, id_base + (u64)i, t, G2048_STREAM_STEP);
            philox_step<FLAVOUR>(e, a, x, seed, id_base + (u64)i, t, L, T, o);
        }
        boards[i] = e.board;
        if (FLAVOUR == G2048_FLAVOUR_PENALTY && aux) aux[i] = env_to_aux(e);
        if (score) score[i] = e.score;
        if (rew64) rew64[i] = o.reward;
        if (rew32) rew32[i] = (float)o.reward;
        if (flags)
            flags[i] = (uint8_t)((o.valid ? G2048_FLAG_VALID : 0) | (o.game_over ? G2048_FLAG_GAME_OVER : 0) |
                                 (o.done ? G2048_FLAG_DONE : 0) | (legal_mask(e.board) << 4));
        if (maxlvl) maxlvl[i] = (uint8_t)o.maxlvl;
        if (move_score) move_score[i] = o.move_score;
    }
}

// SMEM_LUT is a template parameter, not a run-time flag: with the address space of the LUT known the lookups are
// LDS with 32-bit addresses (a run-time choice made them generic LD.E with 64-bit address arithmetic).
template <int FLAVOUR, bool SMEM_LUT>
__global__ void __launch_bounds__(SMEM_LUT ? kRolloutThreads : kSmallRolloutThreads, 1)
k_rollout_random(Tables T, u64* boards, u64* aux, int* score, long long n, long long k_steps, u64 seed, u64 step_base,
                 u64 id_base, long long* counters) {
    extern __shared__ __align__(128) unsigned char smem[];
    Lut L = SMEM_LUT ? stage_lut(T, smem) : global_lut(T);
    Counters c;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        Env e;
        env_load(e, boards[i], (FLAVOUR == G2048_FLAVOUR_PENALTY && aux) ? aux[i] : G2048_AUX_INIT, score ? score[i] : 0);
        u64 id = id_base + (u64)i;
        for (long long k = 0; k < k_steps; ++k) {
            u64 t = step_base + (u64)k;
            Draw4 x = philox(seed, id, t, G2048_STREAM_STEP);
            StepOut o;
            philox_step<FLAVOUR>(e, (int)(x.x3 >> 30), x, seed, id, t, L, T, o);
            c.add(o);
            if (o.done) philox_autoreset(e, seed, id, t);
        }
        boards[i] = e.board;
        if (FLAVOUR == G2048_FLAVOUR_PENALTY && aux) aux[i] = env_to_aux(e);
        if (score) score[i] = e.score;
    }
    flush_counters(c, counters);
}

// main.py:91-101 fused.  Per step: Philox -> epsilon-greedy from the carried row of s -> env step ->
// look s' up -> q <- q + lr (target - q) on Q[s][a] -> carry (slot, row) of s' as the next s.
//
// What bounds it (tools/membench5.cu, profiles/r02_membench.txt): the table lives in 32-64 GiB of HBM, every lookup is
// an L2 miss, and the memory system charges PER REQUEST, not per byte: ~29 ps for a load that misses (35 G/s) and
// about as much again for EVERY write-type request (store, RED or CAS alike, even to a sector that is already dirty):
// load + insert CAS + update CAS = 12.2 G table visits/s, load + ONE write = 17.5 G/s.  With two requests per visit
// the kernel is bound by round trips instead (1024 envs in flight per SM), so nothing is waited for where it is issued:
//   * the lookup is plain 256-bit load(s) and writes nothing: a state that is not in the table yet is carried as
//     `fresh` (slot = the empty slot that ended its probe sequence, zero row);
//   * one step later the update of Q[s][a] goes out as ONE atomic: a 32-bit CAS on the value if s is in the table, or
//     -- s fresh -- a 128-bit CAS {0, 0, 0} -> {key, q0, q1} on the first half of the slot that inserts the state and
//     applies its first update together (a fresh row is zero, so a greedy policy picks action 0 and the value shares
//     the half with the key; for actions 2, 3 the key is inserted by a 64-bit CAS and the value follows);
//   * the result of that atomic is looked at one step later, after the next lookup's wait, when it has long arrived.
//     A lost race (another env changed the value or inserted the same state first) is NOT retried in place: thousands
//     of envs leave the same few start states every step, and a compare-and-swap loop serialises at one success per
//     round trip per address, slower than those updates arrive (measured: 2.6 ms per launch instead of 1.0).  The
//     update is appended to a list as (slot * 4 + action, target) and g2048_rollout_qlearn applies the list after the
//     rollout: sorted by address, every run of records one after another (k_segment_apply).  EVERY update is applied
//     exactly once (counters[RETRIED] counts the deferred ones, LOST stays 0), updates of one Q value compose like the
//     reference's sequential loop (a contraction towards the targets);
//   * small launches (no list: D.count == NULL) settle everything in place with compare-and-swap loops; with one env
//     nothing ever races, and N = 1 is the reference's sequential order exactly;
//   * defaultdict semantics (main.py:16, :41): states that are read but never updated -- the s' of a terminal
//     transition, the state an env sits in when the launch ends -- are inserted on their own;
//   * warps take their 32 envs at a time from a queue (one atomicAdd per warp and rollout), so a warp that met long
//     probe sequences plays fewer envs instead of holding the whole grid up.
struct Deferred {
    ulonglong2* rec;               // [cap] {slot * 4 + action (all ones = unused), float bits of the target}
    unsigned long long* count;     // appended so far (may exceed cap: the excess was applied in place); NULL = no list
    unsigned long long cap;
};
enum { kPendNone = 0, kPendUpd32 = 1, kPendMerged = 2, kPendInsert = 3 };
struct PendingUpdate {
    u64 key;        // merged / insert: the state being inserted
    u64 ret_key;    // merged / insert: key found in the slot (0 = the insert went through)
    u64 ret_q01;    // merged: {q0, q1} found; 32-bit CAS: the value found (low word)
    u32 slot, assumed;
    float target;
    int kind, a;
};
// An update that could not go out as its step's single atomic: q <- q + lr (target - q) on Q[slot][a], `seen` = the
// last value seen there.
struct LateUpdate {
    bool pending = false;
    u32 slot = 0;
    int a = 0;
    float target = 0.f, seen = 0.f;
};
// Look at the result of the update atomic issued one step ago; what is left to do comes back in `late`.
// `cur_slot` follows if the state the env still sits in had to move to another slot.
template <class TAB>
__device__ __forceinline__ void settle_update(const TAB& tab, PendingUpdate& P, u64 cur_key, u32& cur_slot, Counters& c,
                                              LateUpdate& late) {
    if (P.kind == kPendUpd32) {
        if ((u32)P.ret_q01 != P.assumed) {       // another env changed the value first
            c.retried += 1;
            late.pending = true; late.slot = P.slot; late.a = P.a; late.target = P.target;
            late.seen = __uint_as_float((u32)P.ret_q01);
        }
    } else if (P.kind != kPendNone) {            // merged insert + update, or insert alone (the update still to come)
        u32 s = P.slot;
        float seen = 0.f;
        bool todo = (P.kind == kPendInsert);
        if (P.ret_key == 0) {
            c.inserts += 1;
        } else if (P.ret_key == P.key) {         // another env inserted the same state first
            todo = true;
            if (P.kind == kPendMerged) seen = __uint_as_float(P.a == 0 ? (u32)P.ret_q01 : (u32)(P.ret_q01 >> 32));
        } else {                                 // another state took the slot (rare): find a place for ours now
            float4 r2;
            s = table_find<true>(tab, P.key, r2, c.inserts);
            seen = q_at(r2, P.a);
            todo = true;
            if (cur_key == P.key) cur_slot = s;
        }
        if (todo) {
            c.retried += (P.kind == kPendMerged);
            if (s != kNoSlot) { late.pending = true; late.slot = s; late.a = P.a; late.target = P.target; late.seen = seen; }
            else { c.dropped += 1; c.lost += 1; }
        }
    }
    P.kind = kPendNone;
}
// places a warp reserved in the list and will not use: marked "no record" (all ones), so the list needs no clearing
__device__ __forceinline__ void release_places(const Deferred& D, unsigned long long cur, unsigned long long end, unsigned live) {
    const unsigned rank = __popc(live & ((1u << (threadIdx.x & 31)) - 1u)), step = __popc(live);   // the lanes that are here
    for (unsigned long long j = cur + rank; j < end && j < D.cap; j += step)
        D.rec[j] = make_ulonglong2(~0ull, 0ull);
}
// The late updates of a warp (all its `live` lanes, lane 0 among them, call this together): appended to the list of deferred updates -- the warp
// reserves 64 places at a time with one atomicAdd and fills them from a cursor it carries in registers; places it does
// not use stay all ones -- or, without a list or beyond its end, applied in place by a compare-and-swap loop.
template <class TAB>
__device__ __forceinline__ void flush_late(const TAB& tab, const Deferred& D, const LateUpdate& late, float lr,
                                           unsigned long long& cur, unsigned long long& end, unsigned live) {
    const unsigned m = __ballot_sync(live, late.pending);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    unsigned long long j = ~0ull;
    if (D.count) {
        const int cnt = __popc(m);
        if (cur + (unsigned long long)cnt > end) {
            release_places(D, cur, end, live);             // what is left of the old reservation stays unused
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(D.count, 64ull);
            cur = __shfl_sync(live, base, 0);
            end = cur + 64;
        }
        j = cur + (unsigned long long)__popc(m & ((1u << lane) - 1u));
        cur += (unsigned long long)cnt;
    }
    if (late.pending) {
        if (j < D.cap) {
            D.rec[j] = make_ulonglong2((u64)late.slot * 4 + (u64)late.a, (u64)__float_as_uint(late.target));
        } else {
            q_update_atomic<TAB::kSys>(&tab.at(late.slot)->q[late.a], late.seen, lr, late.target);
        }
    }
}
// threads per CTA: a local table is bound by requests, not by round trips, and runs best with 768 threads of 80 registers;
// a table sharded over the GPUs waits for NVLink round trips and wants every thread it can get (1024 x 64)
template <class TAB> struct QlearnThreads { static constexpr int value = kQlearnThreads; };
template <> struct QlearnThreads<ShardedTable> { static constexpr int value = kRolloutThreads; };
template <int FLAVOUR, bool SMEM_LUT, class TAB>
__global__ void __launch_bounds__(SMEM_LUT ? QlearnThreads<TAB>::value : kSmallRolloutThreads, 1)
k_rollout_qlearn(Tables T, u64* boards, u64* aux, int* score, const __grid_constant__ TAB table, long long n,
                 long long k_steps, float lr, float gamma, u64 eps_thresh, u64 seed, u64 step_base, u64 id_base,
                 long long* counters, unsigned long long* queue, const __grid_constant__ Deferred D) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ Slot* shard_base[G2048_MAX_PEERS];
    const auto tab = table.view(shard_base);
    constexpr bool SYS = decltype(tab)::kSys;
    Lut L = SMEM_LUT ? stage_lut(T, smem) : global_lut(T);
    Counters c;
    const int lane = threadIdx.x & 31;
    unsigned long long dcur = 0, dend = 0;         // the warp's reserved places in the list of deferred updates
    for (;;) {
        long long i = 0;
        if (lane == 0) i = (long long)atomicAdd(queue, 32ull);
        i = __shfl_sync(0xFFFFFFFFu, i, 0);
        if (i >= n) break;
        i += lane;
        const unsigned live = __ballot_sync(0xFFFFFFFFu, i < n);   // the last batch may be ragged: the lanes without an env
        if (i >= n) continue;                                      // wait at the top, the others play on under this mask
        Env e;
        env_load(e, boards[i], (FLAVOUR == G2048_FLAVOUR_PENALTY && aux) ? aux[i] : G2048_AUX_INIT, score ? score[i] : 0);
        const u64 id = id_base + (u64)i;
        float4 row;
        bool fresh;
        u32 slot = table_probe(tab, e.board, row, fresh, c.dropped);
        PendingUpdate P;
        P.kind = kPendNone;
#ifdef G2048_EXP_TIMING
        u64 tm0, tm1, tm2, tm3, tm_probe = 0, tm_settle = 0, tm_upd = 0, tm_done = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm0));
        const u64 tm_start = tm0;
#define TM(acc) do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm1)); acc += tm1 - tm0; tm0 = tm1; } while (0)
#else
#define TM(acc) do { } while (0)
#endif
        for (long long k = 0; k < k_steps; ++k) {
            const u64 t = step_base + (u64)k;
            Draw4 x = philox(seed, id, t, G2048_STREAM_STEP);
            const int a = choose_action(row, x, eps_thresh);
            const u64 s_board = e.board;
            StepOut o;
            philox_step<FLAVOUR>(e, a, x, seed, id, t, L, T, o);
            c.add(o);
            float4 row2 = row;
            u32 slot2 = slot;
            bool fresh2 = fresh;
            const bool same = (e.board == s_board);   // an invalid move leaves s' == s: nothing to look up
            TM(tm_done);
            if (!same) slot2 = table_probe(tab, e.board, row2, fresh2, c.dropped);
            __syncwarp(live);
            TM(tm_probe);
            // the atomic issued one step ago has arrived by now (a whole lookup went by)
            LateUpdate late;
            if (P.kind != kPendNone) {
                settle_update(tab, P, s_board, slot, c, late);
                if (same) slot2 = slot;
            }
            __syncwarp(live);
            flush_late(tab, D, late, lr, dcur, dend, live);
            TM(tm_settle);
            if (slot == kNoSlot) {
                c.lost += 1;                       // the state has no slot (table full): the update cannot be stored
            } else {
                const float q = q_at(row, a);
                const float target = td_target(gamma, (float)o.reward, max4(row2), o.done);
                const float nq = td_apply(q, lr, target);
                Slot* sp = tab.at(slot);
                P.slot = slot; P.a = a; P.target = target;
                if (!fresh) {
                    P.assumed = __float_as_uint(q);
                    P.ret_q01 = cas32<SYS>(reinterpret_cast<u32*>(&sp->q[a]), P.assumed, __float_as_uint(nq));
                    P.kind = kPendUpd32;
                } else if (a < 2 && decltype(tab)::kMerge128) {
                    P.key = s_board;
                    cas_w0<SYS>(sp, 0ull, 0ull, s_board, (u64)__float_as_uint(nq) << (32 * a), P.ret_key, P.ret_q01);
                    P.kind = kPendMerged;
                } else if (D.count) {              // key and value in different halves: the value follows the insert
                    P.key = s_board;
                    P.ret_key = cas64<SYS>(&sp->key, 0ull, s_board);
                    P.kind = kPendInsert;
                } else {                           // no list: insert now, then the usual 32-bit CAS
                    slot = insert_at(tab, slot, s_board, c.inserts, c.dropped);
                    if (same) slot2 = slot;
                    if (slot != kNoSlot) {
                        P.slot = slot;
                        P.assumed = 0u;
                        P.ret_q01 = cas32<SYS>(reinterpret_cast<u32*>(&tab.at(slot)->q[a]), 0u, __float_as_uint(nq));
                        P.kind = kPendUpd32;
                    } else {
                        c.lost += 1;
                    }
                }
                if (same) { q_set(row2, a, nq); fresh2 = false; }
            }
            row = row2;
            slot = slot2;
            fresh = fresh2;
            __syncwarp(live);
            TM(tm_upd);
            if (o.done) {
                // update_q_value read Q[next_state] (main.py:41): the terminal state is in the table from now on;
                // state = env.reset() is looked up at once (main.py:81-82, :92)
                if (fresh && slot != kNoSlot) insert_at(tab, slot, e.board, c.inserts, c.dropped);
                philox_autoreset(e, seed, id, t);
                slot = table_probe(tab, e.board, row, fresh, c.dropped);
            }
        }
        {
            LateUpdate late;
            if (P.kind != kPendNone) settle_update(tab, P, e.board, slot, c, late);
            __syncwarp(live);
            flush_late(tab, D, late, lr, dcur, dend, live);
        }
        if (fresh && slot != kNoSlot) insert_at(tab, slot, e.board, c.inserts, c.dropped);   // the launch ends: reading a state creates it
        boards[i] = e.board;
        if (FLAVOUR == G2048_FLAVOUR_PENALTY && aux) aux[i] = env_to_aux(e);
        if (score) score[i] = e.score;
#ifdef G2048_EXP_TIMING
        __syncwarp(live);
        TM(tm_done);
        if (lane == 0) {
            long long dt = (long long)(tm1 - tm_start);
            atomicMax(counters + 10, dt);
            atomicAdd((unsigned long long*)(counters + 11), (unsigned long long)dt);
            atomicAdd((unsigned long long*)(counters + 13), 1ull);
            unsigned sm;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
            if (dt > 400000) printf("slow rollout: sm %u warp %d env %lld total %lld us: probe %llu settle %llu update %llu step+done %llu\n", sm,
                                    (int)(threadIdx.x >> 5), i, dt / 1000, tm_probe / 1000, tm_settle / 1000, tm_upd / 1000, tm_done / 1000);
        }
#endif
    }
#undef TM
    if (D.count) release_places(D, dcur, dend, 0xFFFFFFFFu);
    flush_counters(c, counters);
}

// Synchronous step, phase A: choose + env step + bootstrap on the snapshot; emits one record per env:
// sort key = slot * 4 + action (all ones = no slot), TD target, and optionally (state key, action).
template <int FLAVOUR>
__global__ void __launch_bounds__(256)
k_qlearn_phase_a(Tables T, u64* boards, u64* aux, int* score, Slot* tab, u64 mask, long long n, float lr, float gamma,
                 u64 eps_thresh, u64 seed, u64 t, u64 id_base, long long* counters, u64* sortkey, float* target_out,
                 u64* rec_key, uint8_t* rec_action, float* rec_target, ulonglong2* rec_packed) {
    Lut L = global_lut(T);
    Counters c;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        Env e;
        env_load(e, boards[i], (FLAVOUR == G2048_FLAVOUR_PENALTY && aux) ? aux[i] : G2048_AUX_INIT, score ? score[i] : 0);
        u64 id = id_base + (u64)i;
        float4 row, row2;
        u64 s_key = e.board;
        u32 slot = table_find<true>(tab, mask, s_key, row, c.inserts);
        c.dropped += (slot == kNoSlot);
        Draw4 x = philox(seed, id, t, G2048_STREAM_STEP);
        int a = choose_action(row, x, eps_thresh);
        StepOut o;
        philox_step<FLAVOUR>(e, a, x, seed, id, t, L, T, o);
        c.add(o);
        u32 slot2 = table_find<true>(tab, mask, e.board, row2, c.inserts);
        c.dropped += (slot2 == kNoSlot);
        float target = td_target(gamma, (float)o.reward, max4(row2), o.done);
        if (sortkey) sortkey[i] = slot == kNoSlot ? ~0ull : ((u64)slot * 4 + (u64)a);
        if (target_out) target_out[i] = target;
        if (rec_key) rec_key[i] = s_key;
        if (rec_action) rec_action[i] = (uint8_t)a;
        if (rec_target) rec_target[i] = target;
        if (rec_packed) rec_packed[i] = pack_record(s_key, a, target);
        if (o.done) philox_autoreset(e, seed, id, t);
        boards[i] = e.board;
        if (FLAVOUR == G2048_FLAVOUR_PENALTY && aux) aux[i] = env_to_aux(e);
        if (score) score[i] = e.score;
    }
    flush_counters(c, counters);
}

// update_q_value phase A on given transitions (teacher-forced): reads only, emits sort key + TD target
__global__ void __launch_bounds__(256)
k_q_update_phase_a(Slot* tab, u64 mask, const u64* s, const uint8_t* a, const float* r, const u64* s2,
                   const uint8_t* done, long long n, float gamma, u64* sortkey, float* target_out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float4 row, row2;
        u32 ins = 0;
        table_find<true>(tab, mask, s2[i], row2, ins);
        u32 slot = table_find<true>(tab, mask, s[i], row, ins);
        int act = a[i] & 3;
        sortkey[i] = slot == kNoSlot ? ~0ull : ((u64)slot * 4 + (u64)act);
        target_out[i] = td_target(gamma, r[i], max4(row2), done[i] != 0);
    }
}
// (key, action, target) records -> sort key (find-or-insert the key)
__global__ void __launch_bounds__(256)
k_keys_to_records(Slot* tab, u64 mask, const u64* keys, const uint8_t* a, long long n, u64* sortkey) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float4 row;
        u32 ins = 0;
        u32 slot = table_find<true>(tab, mask, keys[i], row, ins);
        sortkey[i] = slot == kNoSlot ? ~0ull : ((u64)slot * 4 + (u64)(a[i] & 3));
    }
}
// ---- exact synchronous step on a SHARDED table: owner computes -------------------------------------------------
// Phase A of the synchronous step with the table spread over the GPUs: s and s' are looked up (found-or-inserted)
// wherever their slots live, and the record of the transition is appended to the list of the GPU that OWNS the slot
// of s.  A record carries the owner-local sort key directly -- ((local slot * 4 + action) << idx_bits) | global
// env index -- so the owner neither looks anything up nor depends on the order of the appends: sorting the composite
// key groups each (state, action) and orders its records by global env id, exactly the single-GPU order.
struct OwnedLists {
    ulonglong2* list[G2048_MAX_PEERS];   // list[j]: this rank's records for owner j (room for n each)
    unsigned long long* count;           // count[j], zeroed by the caller
    int idx_bits;
};
template <int FLAVOUR>
__global__ void __launch_bounds__(256)
k_qlearn_emit_owned(Tables T, u64* boards, u64* aux, int* score, const __grid_constant__ ShardedTable table, long long n,
                    float gamma, u64 eps_thresh, u64 seed, u64 t, u64 id_base, u64 rec_base, long long* counters,
                    const __grid_constant__ OwnedLists out, u32* carry_slot, float4* carry_row, int use_carry) {
    __shared__ Slot* shard_base[G2048_MAX_PEERS];
    __shared__ ulonglong2* list_base[G2048_MAX_PEERS];
    if (threadIdx.x < G2048_MAX_PEERS) list_base[threadIdx.x] = out.list[threadIdx.x];
    const auto tab = table.view(shard_base);   // (syncs the block)
    Lut L = global_lut(T);
    Counters c;
    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = blockIdx.x * (long long)blockDim.x + (threadIdx.x & ~31); i0 < n; i0 += stride) {   // warp-uniform
        const long long i = i0 + lane;
        int owner = -1;
        u64 key = 0;
        float target = 0.f;
        if (i < n) {
            Env e;
            env_load(e, boards[i], (FLAVOUR == G2048_FLAVOUR_PENALTY && aux) ? aux[i] : G2048_AUX_INIT, score ? score[i] : 0);
            u64 id = id_base + (u64)i;
            float4 row, row2;
            u32 slot;
            if (use_carry) {   // inside a window the values are frozen: (slot, row) of s are those found for s' one step ago
                slot = carry_slot[i];
                row = carry_row[i];
            } else {
                slot = table_find<true>(tab, e.board, row, c.inserts);
                c.dropped += (slot == kNoSlot);
            }
            Draw4 x = philox(seed, id, t, G2048_STREAM_STEP);
            int a = choose_action(row, x, eps_thresh);
            const u64 s_board = e.board;
            StepOut o;
            philox_step<FLAVOUR>(e, a, x, seed, id, t, L, T, o);
            c.add(o);
            u32 slot2 = slot;
            row2 = row;
            if (e.board != s_board) {   // an invalid move leaves s' == s: nothing to look up
                slot2 = table_find<true>(tab, e.board, row2, c.inserts);
                c.dropped += (slot2 == kNoSlot);
            }
            target = td_target(gamma, (float)o.reward, max4(row2), o.done);
            if (slot != kNoSlot) {
                owner = (int)((u64)slot >> tab.shift);
                key = (((((u64)slot & tab.low) << 2) | (u64)a) << out.idx_bits) | (rec_base + (u64)i);
            }
            if (o.done) {
                philox_autoreset(e, seed, id, t);
                if (carry_slot) {   // the next step continues from the fresh board: find it now
                    slot2 = table_find<true>(tab, e.board, row2, c.inserts);
                    c.dropped += (slot2 == kNoSlot);
                }
            }
            if (carry_slot) { carry_slot[i] = slot2; carry_row[i] = row2; }
            boards[i] = e.board;
            if (FLAVOUR == G2048_FLAVOUR_PENALTY && aux) aux[i] = env_to_aux(e);
            if (score) score[i] = e.score;
        }
        // one atomicAdd per (warp, owner): the lanes that share an owner take consecutive places
        unsigned peers = __match_any_sync(0xFFFFFFFFu, owner);
        int leader = __ffs(peers) - 1;
        unsigned long long base = 0;
        if (lane == leader && owner >= 0) base = atomicAdd(&out.count[owner], (unsigned long long)__popc(peers));
        base = __shfl_sync(0xFFFFFFFFu, base, leader);
        if (owner >= 0)
            list_base[owner][base + __popc(peers & ((1u << lane) - 1))] = make_ulonglong2(key, (u64)__float_as_uint(target));
    }
    flush_counters(c, counters);
}
// ---- record lists that may live in the HBM of OTHER GPUs (NVLink peer memory) ---------------------------------
// Up to kMaxLists lists of 16-byte records {key, action | target bits << 32}; list j holds the records of rank j in
// ascending env order, so walking the lists in order visits ascending GLOBAL env ids.
struct RecordLists {
    const ulonglong2* ptr[G2048_MAX_PEERS];
    long long end[G2048_MAX_PEERS];   // exclusive prefix ends
    int n_lists;
};
// The all-gather fused into its consumer: every rank pulls each record straight out of its owner's memory
// (system-scope loads, never cached in L1), finds-or-inserts the state in ITS replica and writes the sort key +
// target for the deterministic apply.  Consecutive threads read consecutive records: 512 B per warp over NVLink.
__global__ void __launch_bounds__(256)
k_peer_records_to_sortkeys(Slot* tab, u64 mask, RecordLists R, long long n, u64* sortkey, float* target, const int* abort_flag) {
    const bool aborted = abort_flag && *(const volatile int*)abort_flag != 0;   // a peer never reached the barrier: its list is not valid
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (aborted) { sortkey[i] = ~0ull; target[i] = 0.f; continue; }
        int j = 0;
        while (j + 1 < R.n_lists && i >= R.end[j]) ++j;
        long long local = i - (j ? R.end[j - 1] : 0);
        u64 key, at;
        asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(key), "=l"(at) : "l"(R.ptr[j] + local) : "memory");
        float4 row;
        u32 ins = 0;
        u32 slot = table_find<true>(tab, mask, key, row, ins);
        sortkey[i] = slot == kNoSlot ? ~0ull : ((u64)slot * 4 + (at & 3));
        target[i] = __uint_as_float((u32)(at >> 32));
    }
}
// owner-computes form: the owner pulls its lists out of every rank's memory (NVLink peer reads, coalesced) into the
// sort buffers; the records already carry the sort key
__global__ void __launch_bounds__(256)
k_gather_owned(RecordLists R, long long n, u64* sortkey, float* target, const int* abort_flag) {
    const bool aborted = abort_flag && *(const volatile int*)abort_flag != 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (aborted) { sortkey[i] = ~0ull; target[i] = 0.f; continue; }
        int j = 0;
        while (j + 1 < R.n_lists && i >= R.end[j]) ++j;
        long long local = i - (j ? R.end[j - 1] : 0);
        u64 key, tb;
        asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(key), "=l"(tb) : "l"(R.ptr[j] + local) : "memory");
        sortkey[i] = key;
        target[i] = __uint_as_float((u32)tb);
    }
}
// Barrier between the GPUs of one box through flags in peer memory: rank r stores `epoch` into flags[r] of every
// peer (release, system scope) and waits until all of its own flags reached `epoch` (acquire).  Epochs only grow,
// so the flags never need a reset.  A peer that never arrives is reported after `timeout_ns` instead of hanging.
struct PeerFlags { u64* ptr[G2048_MAX_PEERS]; };
__global__ void k_peer_barrier(PeerFlags F, int rank, int world, u64 epoch, u64 timeout_ns, int* timed_out, int* abort_flag) {
    int t = threadIdx.x;
    if (t >= world) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(F.ptr[t] + rank), "l"(epoch) : "memory");
    u64 t0, now, v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(F.ptr[rank] + t) : "memory");
        if (v >= epoch) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t0 > timeout_ns) {
            if (abort_flag) *(volatile int*)abort_flag = 1 + t;   // what the apply kernels of this device look at
            if (timed_out) {                       // the caller's copy, device or pinned host memory: a plain system-scope store
                *(volatile int*)timed_out = 1 + t;
                __threadfence_system();
            }
            break;
        }
        __nanosleep(200);
    }
}

// ---- exact synchronous step on a SHARDED table, ROUTED: every table access is local to the owner ---------------
// k_qlearn_emit_owned looks s and s' up wherever they live: ~2.4 small NVLink requests per env, and the rate of small
// remote requests (6.7 G/s per GPU) is what bounds that step.  Here nothing but coalesced lists crosses NVLink:
//   k_routed_request  every env takes (slot, row) of its state from the answers of the last step, chooses, steps, and
//                     appends the key of s' (and of the fresh board after a game over) to its list for the GPU that OWNS
//                     the key's home slot; counts its record per (warp, owner)                 [local writes]
//   k_routed_scan     the counts become places: every rank's records for one owner in ascending env order
//   k_routed_offsets  (after the barrier) where this rank's records start in every owner's sort input = the records of
//                     the lower ranks; the total for this GPU goes to the host (the only synchronisation of the step)
//   k_routed_lookup   the owner pulls the lists written for it (coalesced peer reads), finds-or-inserts every key in
//                     its OWN shard and pushes {slot, max Q} to the same place of the requester's answer buffer
//   k_routed_records  the requester builds r + gamma max Q(s') and pushes {slot * 4 + action, target} straight into the
//                     sort input of the owner of s, at its place: the input is then in ascending GLOBAL env order
//   (stable radix sort on the slot bits + k_segment_apply: the single-GPU deterministic apply, on the owner's share)
//   k_routed_rows     the owner pushes the rows as they are AFTER the apply for every request: what the next step's
//                     choose_action reads
// with a flag barrier after request, lookup, records and rows.  A request handle = owner << 28 | place in the list.
struct RoutedLocal {                 // the requester's side
    u64* req_out[G2048_MAX_PEERS];            // keys for owner d                                   (local)
    u64* push_rec[G2048_MAX_PEERS];           // owner d's sort input: (slot * 4 + action) << 32 | TD target bits  (peer memory, written)
    const uint2* reply1[G2048_MAX_PEERS];     // {slot, max Q bits} from owner d, same places as req_out[d]   (local)
    const float4* reply2[G2048_MAX_PEERS];    // rows after the apply from owner d                  (local)
    unsigned long long* req_count;            // [world]
    unsigned long long* rec_count;            // [world] records for owner d (k_routed_scan)
    const u32* off;                           // [world] my first place in owner d's sort input (k_routed_offsets)
    u32* sk;                                  // per env: slot * 4 + action of its record
    u32* req1;                                // per env: handle of the request for s'
    u32* cur;                                 // per env: handle of the request that answers for the state it sits in
    float* reward;                            // per env
    u32* meta;                                // per env: owner of s (255: the state has no slot, no record) | done << 8
    u32* chunk;                               // [world][warps]: records of a warp's 32 envs per owner (k_routed_request)
    u32* place;                               // [world][warps]: their first place among this rank's records for that owner (k_routed_scan)
    long long n_warps;                        // row length of `chunk` and `place`
    int world;
    u32 owner_shift;                          // owner(key) = (mix64(key) >> owner_shift) & (world - 1)
};
struct RoutedServe {                 // the owner's side
    const u64* req[G2048_MAX_PEERS];                        // rank r's keys for this GPU          (peer memory, read)
    const unsigned long long* req_count[G2048_MAX_PEERS];   // how many                            (peer memory, read)
    uint2* reply1[G2048_MAX_PEERS];                          // this GPU's part of r's answer buffer (peer memory, written)
    float4* reply2[G2048_MAX_PEERS];
    u32* saved_slot[G2048_MAX_PEERS];                        // local: slot of every request, kept for k_routed_rows
    unsigned long long* count_cache;                        // local [world]: the counts k_routed_lookup read
    int world;
};
constexpr u32 kHandleBits = 28;
constexpr int kInlineRun = 8;      // sorted runs of more records than this are applied by a warp (k_long_run_apply)
// the lanes of a warp that append to the same list take consecutive places with one atomicAdd (all 32 lanes call this;
// list < 0: nothing to append); returns the place
__device__ __forceinline__ u32 warp_append(int list, unsigned long long* counts) {
    const int lane = threadIdx.x & 31;
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, list);
    const int leader = __ffs(peers) - 1;
    unsigned long long base = 0;
    if (lane == leader && list >= 0) base = atomicAdd(&counts[list], (unsigned long long)__popc(peers));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    return (u32)base + (u32)__popc(peers & ((1u << lane) - 1u));
}
// SMEM_LUT (big batches): one block of 1024 threads per SM with the row LUT staged in shared memory, as k_env_step
// (with the LUT in global memory the kernel needs 76 registers and runs latency-bound at 768 threads per SM)
template <int FLAVOUR, bool PRIME, bool SMEM_LUT>
__global__ void __launch_bounds__(SMEM_LUT ? kRolloutThreads : 256, SMEM_LUT ? 1 : 3)
k_routed_request(Tables T, u64* boards, u64* aux, int* score, const __grid_constant__ RoutedLocal R, long long n,
                 u64 eps_thresh, u64 seed, u64 t, u64 id_base, long long* counters) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ u64* req_out[G2048_MAX_PEERS];
    __shared__ const uint2* reply1[G2048_MAX_PEERS];
    __shared__ const float4* reply2[G2048_MAX_PEERS];
    if (threadIdx.x < G2048_MAX_PEERS) {
        req_out[threadIdx.x] = R.req_out[threadIdx.x];
        reply1[threadIdx.x] = R.reply1[threadIdx.x];
        reply2[threadIdx.x] = R.reply2[threadIdx.x];
    }
    __syncthreads();
    Lut L = SMEM_LUT ? stage_lut(T, smem) : global_lut(T);
    Counters c;
    const int lane = threadIdx.x & 31;
    const u32 owner_mask = (u32)R.world - 1u;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = blockIdx.x * (long long)blockDim.x + (threadIdx.x & ~31); i0 < n; i0 += stride) {   // warp-uniform
        const long long i = i0 + lane;
        int owner1 = -1, owner2 = -1, owner_s = -1;
        u64 key1 = 0, key2 = 0;
        if (i < n) {
            if (PRIME) {
                key1 = boards[i];
            } else {
                Env e;
                env_load(e, boards[i], (FLAVOUR == G2048_FLAVOUR_PENALTY && aux) ? aux[i] : G2048_AUX_INIT, score ? score[i] : 0);
                const u64 id = id_base + (u64)i;
                const u32 cur = R.cur[i], os = cur >> kHandleBits, pos = cur & ((1u << kHandleBits) - 1u);
                const u32 slot = reply1[os][pos].x;
                const float4 row = reply2[os][pos];
                Draw4 x = philox(seed, id, t, G2048_STREAM_STEP);
                const int a = choose_action(row, x, eps_thresh);
                StepOut o;
                philox_step<FLAVOUR>(e, a, x, seed, id, t, L, T, o);
                c.add(o);
                key1 = e.board;                       // s' (== s after an invalid move: the owner answers for s again)
                R.sk[i] = (slot << 2) | (u32)a;
                if (slot != kNoSlot) owner_s = (int)os;
                R.reward[i] = (float)o.reward;
                R.meta[i] = (slot != kNoSlot ? os : 255u) | (o.done ? 256u : 0u);
                if (o.done) {
                    philox_autoreset(e, seed, id, t);
                    key2 = e.board;
                    owner2 = (int)((u32)(mix64(key2) >> R.owner_shift) & owner_mask);
                }
                boards[i] = e.board;
                if (FLAVOUR == G2048_FLAVOUR_PENALTY && aux) aux[i] = env_to_aux(e);
                if (score) score[i] = e.score;
            }
            owner1 = (int)((u32)(mix64(key1) >> R.owner_shift) & owner_mask);
        }
        const u32 p1 = warp_append(owner1, R.req_count);
        if (owner1 >= 0) req_out[owner1][p1] = key1;
        u32 h1 = ((u32)owner1 << kHandleBits) | p1, h2 = h1;
        if (!PRIME) {
            // how many records this warp's 32 envs will send to every owner: k_routed_scan turns the counts into places,
            // so that k_routed_records writes every owner's input in ascending env order without an atomic
            u32 mine = 0;
            for (int j = 0; j < R.world; ++j) {
                const u32 cnt = (u32)__popc(__ballot_sync(0xFFFFFFFFu, owner_s == j));
                if (lane == j) mine = cnt;
            }
            if (lane < R.world) R.chunk[(long long)lane * R.n_warps + (i0 >> 5)] = mine;
            if (__any_sync(0xFFFFFFFFu, owner2 >= 0)) {          // a game ended in this warp (rare): the fresh board
                const u32 p2 = warp_append(owner2, R.req_count);
                if (owner2 >= 0) {
                    req_out[owner2][p2] = key2;
                    h2 = ((u32)owner2 << kHandleBits) | p2;
                }
            }
        }
        if (i < n) {
            R.req1[i] = h1;
            R.cur[i] = h2;
        }
    }
    flush_counters(c, counters);
}
// chunk[j][w] = records warp w sends to owner j  ->  place[j][w] = exclusive prefix over the warps, the total to count[j].
// Block (t, j) scans tile t (1024 warps) of owner j's row; the tile's first place = the sum of everything before it,
// which the block adds up itself (the rows are a few hundred KB and sit in L2; no atomics, nothing to clear).
__global__ void __launch_bounds__(1024) k_routed_scan(const u32* chunk, u32* place, long long n_warps, long long row, int tiles,
                                                      unsigned long long* count) {
    __shared__ u32 warp_tot[32];
    __shared__ u32 base_sh;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, t = blockIdx.x, j = blockIdx.y;
    const u32* col = chunk + (long long)j * row;
    u32 before = 0;
    const long long stop = (long long)t * 1024 < n_warps ? (long long)t * 1024 : n_warps;
#pragma unroll 8
    for (long long c = threadIdx.x; c < stop; c += 1024) before += col[c];   // (independent loads: eight in flight per thread)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) before += __shfl_xor_sync(0xFFFFFFFFu, before, d);
    if (lane == 0) warp_tot[warp] = before;
    __syncthreads();
    if (warp == 0) {
        u32 z = warp_tot[lane];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) z += __shfl_xor_sync(0xFFFFFFFFu, z, d);
        if (lane == 0) base_sh = z;
    }
    __syncthreads();
    const u32 base = base_sh;
    const long long c = (long long)t * 1024 + threadIdx.x;
    const u32 mine = c < n_warps ? col[c] : 0u;
    u32 x = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const u32 y = __shfl_up_sync(0xFFFFFFFFu, x, d); if (lane >= d) x += y; }
    __syncthreads();
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
        const u32 w = warp_tot[lane];
        u32 z = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const u32 y = __shfl_up_sync(0xFFFFFFFFu, z, d); if (lane >= d) z += y; }
        warp_tot[lane] = z - w;
        if (lane == 31 && t == tiles - 1) count[j] = (unsigned long long)(base + z);
    }
    __syncthreads();
    if (c < n_warps) place[(long long)j * row + c] = base + x - mine + warp_tot[warp];
}
// after the barrier: every rank's record counts (peer memory) -> where my records start in every owner's sort input
// (= the records of the lower ranks for that owner), and how many records this GPU will receive (to the host)
struct PeerWords { const u64* ptr[G2048_MAX_PEERS]; };
__global__ void __launch_bounds__(256) k_routed_offsets(PeerWords counts, int world, int me, u32* off, u64* host_total) {
    __shared__ u64 m[G2048_MAX_PEERS][G2048_MAX_PEERS];
    const int r = threadIdx.x / G2048_MAX_PEERS, j = threadIdx.x % G2048_MAX_PEERS;
    if (r < world && j < world) {
        u64 v;
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(counts.ptr[r] + j) : "memory");
        m[r][j] = v;
    }
    __syncthreads();
    if ((int)threadIdx.x < world) {
        u64 s = 0;
        for (int q = 0; q < me; ++q) s += m[q][threadIdx.x];
        off[threadIdx.x] = (u32)s;
    }
    if (threadIdx.x == 0) {
        u64 s = 0;
        for (int q = 0; q < world; ++q) s += m[q][me];
        *host_total = s;
    }
}
// prefix ends of the `world` request lists in shared memory; returns the total
__device__ __forceinline__ long long routed_ends(const unsigned long long* cnt, int world, long long* end) {
    if (threadIdx.x == 0) {
        long long s = 0;
        for (int j = 0; j < world; ++j) { s += (long long)cnt[j]; end[j] = s; }
    }
    __syncthreads();
    return end[world - 1];
}
__global__ void __launch_bounds__(256)
k_routed_lookup(Slot* shard, u64 mask, const __grid_constant__ RoutedServe V, long long* counters, const int* abort_flag) {
    __shared__ unsigned long long cnt[G2048_MAX_PEERS];
    __shared__ long long end[G2048_MAX_PEERS];
    if (abort_flag && *(const volatile int*)abort_flag != 0) return;   // a peer never reached the barrier: its list is not valid
    if ((int)threadIdx.x < V.world) {
        unsigned long long v;
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(V.req_count[threadIdx.x]) : "memory");
        cnt[threadIdx.x] = v;
        if (blockIdx.x == 0) V.count_cache[threadIdx.x] = v;
    }
    __syncthreads();
    const long long m = routed_ends(cnt, V.world, end);
    Counters c;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        int j = 0;
        while (i >= end[j]) ++j;
        const long long local = i - (j ? end[j - 1] : 0);
        u64 key;
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(key) : "l"(V.req[j] + local) : "memory");
        float4 row;
        const u32 slot = table_find<true>(shard, mask, key, row, c.inserts);
        c.dropped += (slot == kNoSlot);
        V.saved_slot[j][local] = slot;
        V.reply1[j][local] = make_uint2(slot, __float_as_uint(max4(row)));
    }
    flush_counters(c, counters);
}
__global__ void __launch_bounds__(256)
k_routed_rows(const Slot* shard, const __grid_constant__ RoutedServe V, const int* abort_flag) {
    __shared__ unsigned long long cnt[G2048_MAX_PEERS];
    __shared__ long long end[G2048_MAX_PEERS];
    if (abort_flag && *(const volatile int*)abort_flag != 0) return;
    if ((int)threadIdx.x < V.world) cnt[threadIdx.x] = V.count_cache[threadIdx.x];
    __syncthreads();
    const long long m = routed_ends(cnt, V.world, end);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        int j = 0;
        while (i >= end[j]) ++j;
        const long long local = i - (j ? end[j - 1] : 0);
        const u32 slot = V.saved_slot[j][local];
        float4 row = make_float4(0.f, 0.f, 0.f, 0.f);
        if (slot != kNoSlot) {
            u64 k;
            load_slot<false>(shard + slot, k, row);
        }
        V.reply2[j][local] = row;
    }
}
__global__ void __launch_bounds__(256)
k_routed_records(const __grid_constant__ RoutedLocal R, long long n, float gamma) {
    __shared__ u64* push_rec[G2048_MAX_PEERS];
    __shared__ const uint2* reply1[G2048_MAX_PEERS];
    __shared__ u32 off[G2048_MAX_PEERS];
    if (threadIdx.x < G2048_MAX_PEERS) {
        push_rec[threadIdx.x] = R.push_rec[threadIdx.x];
        reply1[threadIdx.x] = R.reply1[threadIdx.x];
        off[threadIdx.x] = (int)threadIdx.x < R.world ? R.off[threadIdx.x] : 0u;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = blockIdx.x * (long long)blockDim.x + (threadIdx.x & ~31); i0 < n; i0 += stride) {   // warp-uniform
        const long long i = i0 + lane;
        int owner = -1;
        u64 rec = 0;
        if (i < n) {
            const u32 h = R.req1[i], meta = R.meta[i];
            const uint2 ans = reply1[h >> kHandleBits][h & ((1u << kHandleBits) - 1u)];
            const float target = td_target(gamma, R.reward[i], __uint_as_float(ans.y), (meta & 256u) != 0);
            if ((meta & 255u) != 255u) owner = (int)(meta & 255u);
            rec = ((u64)R.sk[i] << 32) | (u64)__float_as_uint(target);
        }
        // place = my first place in the owner's input + the warp's first place among my records for it (k_routed_scan)
        //         + the rank among the warp's lanes for that owner: ascending env order.  One 8-byte store per record.
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, owner);
        if (owner >= 0)
            push_rec[owner][off[owner] + R.place[(long long)owner * R.n_warps + (i0 >> 5)] + (u32)__popc(peers & ((1u << lane) - 1u))] = rec;
    }
}
// k_segment_apply / k_long_run_apply on packed records (key << 32 | target bits), sorted by key, equal keys in env order
__global__ void __launch_bounds__(256)
k_segment_apply_packed(Slot* tab, const u64* rec, float lr, long long n, u64* worklist, const int* abort_flag) {
    if (abort_flag && *(const volatile int*)abort_flag != 0) return;   // a peer never reached the barrier: the records are not valid
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const u32 k = (u32)(rec[i] >> 32);
        if (i > 0 && (u32)(rec[i - 1] >> 32) == k) continue;
        long long j = i + 1;
        while (j - i <= kInlineRun && j < n && (u32)(rec[j] >> 32) == k) ++j;
        if (j - i > kInlineRun) {
            const u64 w = atomicAdd((unsigned long long*)&worklist[0], 1ull);
            worklist[1 + w] = (u64)i;
            continue;
        }
        float* qp = &tab[k >> 2].q[k & 3];
        float q = *qp;
        for (long long t = i; t < j; ++t) q = td_apply(q, lr, __uint_as_float((u32)rec[t]));
        *qp = q;
    }
}
__global__ void __launch_bounds__(256)
k_long_run_apply_packed(Slot* tab, const u64* rec, float lr, long long n, const u64* worklist, const int* abort_flag) {
    if (abort_flag && *(const volatile int*)abort_flag != 0) return;
    const int lane = threadIdx.x & 31;
    const u64 n_warps = (u64)gridDim.x * (blockDim.x >> 5);
    const u64 count = worklist[0];
    for (u64 w = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < count; w += n_warps) {
        long long j = (long long)worklist[1 + w];
        const u32 k = (u32)(rec[j] >> 32);
        float* qp = &tab[k >> 2].q[k & 3];
        float q = *qp;
        u64 r = (j + lane < n) ? rec[j + lane] : ((u64)~k << 32);
        for (;;) {
            const unsigned m = __ballot_sync(0xFFFFFFFFu, (u32)(r >> 32) == k);
            const int cnt = (m == 0xFFFFFFFFu) ? 32 : __ffs(~m) - 1;   // records of this run in the chunk (a prefix)
            u64 rn = (u64)~k << 32;
            if (cnt == 32 && j + 32 + lane < n) rn = rec[j + 32 + lane];
            const float tt = __uint_as_float((u32)r);
            if (cnt == 32) {
#pragma unroll
                for (int l = 0; l < 32; ++l) q = td_apply(q, lr, __shfl_sync(0xFFFFFFFFu, tt, l));
            } else {
                for (int l = 0; l < cnt; ++l) q = td_apply(q, lr, __shfl_sync(0xFFFFFFFFu, tt, l));
                break;
            }
            j += 32; r = rn;
        }
        if (lane == 0) *qp = q;
    }
}

__global__ void __launch_bounds__(256)
k_apply_atomic(Slot* tab, const u64* sortkey, const float* target, float lr, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        u64 k = sortkey[i];
        if (k == ~0ull) continue;
        float* q = &tab[k >> 2].q[k & 3];
        q_update_atomic(q, __ldcg(q), lr, target[i]);
    }
}
// sorted (stable) records: every run of equal keys is applied in order, q <- q + lr (target - q) record after
// record (the reference's update, main.py:43, in sequence).  The head of a run of up to kInlineRun records applies
// it itself; longer runs (early-game states shared by thousands of envs) are queued for k_long_run_apply, because
// one thread walking a long run pays a full load latency per record (measured: 16 ms for the first step after a
// reset of 8 M envs).  worklist[0] = number of queued runs (zeroed by the caller), worklist[1 + w] = index of the head.
__global__ void __launch_bounds__(256)
k_segment_apply(Slot* tab, const u64* sortkey, const float* target, float lr, long long n, u64* worklist, int kshift,
                const int* abort_flag = nullptr) {
    if (abort_flag && *(const volatile int*)abort_flag != 0) return;   // a peer never reached the barrier: its records are not valid
    // kshift: low bits of the sort key that only order the records of a run (0 here; the global env index in the
    // owner-computes form); the all-ones key ("no slot") is checked before shifting
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        u64 raw = sortkey[i];
        u64 k = raw >> kshift;
        if (raw == ~0ull || (i > 0 && (sortkey[i - 1] >> kshift) == k)) continue;
        long long j = i + 1;
        while (j - i <= kInlineRun && j < n && (sortkey[j] >> kshift) == k) ++j;
        if (j - i > kInlineRun) {
            u64 w = atomicAdd((unsigned long long*)&worklist[0], 1ull);
            worklist[1 + w] = (u64)i;
            continue;
        }
        float* qp = &tab[k >> 2].q[k & 3];
        float q = *qp;
        for (long long t = i; t < j; ++t) q = td_apply(q, lr, target[t]);
        *qp = q;
    }
}
// One warp per long run: the lanes fetch 32 records at a time (coalesced, the next chunk in flight while the
// current one is consumed) and every lane runs the same sequential chain over the shuffled targets -- 3 dependent
// float operations per record, the order of the records untouched, so the result is bit-identical to the serial walk.
__global__ void __launch_bounds__(256)
k_long_run_apply(Slot* tab, const u64* sortkey, const float* target, float lr, long long n, const u64* worklist,
                 int kshift, const int* abort_flag = nullptr) {
    if (abort_flag && *(const volatile int*)abort_flag != 0) return;
    const int lane = threadIdx.x & 31;
    const u64 n_warps = (u64)gridDim.x * (blockDim.x >> 5);
    const u64 count = worklist[0];
    for (u64 w = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < count; w += n_warps) {
        long long j = (long long)worklist[1 + w];
        const u64 k = sortkey[j] >> kshift;
        float* qp = &tab[k >> 2].q[k & 3];
        float q = *qp;
        u64 kk = (j + lane < n) ? (sortkey[j + lane] >> kshift) : ~k;
        float tt = (j + lane < n) ? target[j + lane] : 0.f;
        for (;;) {
            unsigned m = __ballot_sync(0xFFFFFFFFu, kk == k);
            int cnt = (m == 0xFFFFFFFFu) ? 32 : __ffs(~m) - 1;   // records of this run in the chunk (a prefix)
            u64 kn = ~k;
            float tn = 0.f;
            if (cnt == 32 && j + 32 + lane < n) { kn = sortkey[j + 32 + lane] >> kshift; tn = target[j + 32 + lane]; }
            if (cnt == 32) {
#pragma unroll
                for (int l = 0; l < 32; ++l) q = td_apply(q, lr, __shfl_sync(0xFFFFFFFFu, tt, l));
            } else {
                for (int l = 0; l < cnt; ++l) q = td_apply(q, lr, __shfl_sync(0xFFFFFFFFu, tt, l));
                break;
            }
            j += 32; kk = kn; tt = tn;
        }
        if (lane == 0) *qp = q;
    }
}

// ---- grouping update records by address without a sort ------------------------------------------------------------
// Both the deferred updates of a fused rollout and the records of the exact synchronous step need "all the records of
// one Q value in one thread's hands", not a total order: the records are counted into 2^bits hash buckets of their
// address (one atomicAdd each, which also gives the record its place in its bucket), the bucket sizes are prefix-summed
// and the records scattered bucket by bucket, packed as {address, target bits | order << 32}.  A bucket then holds a few
// records of different addresses, or the hundreds that pile up on the action a popular start state takes.  Inside a
// group the records are applied by ascending `order` (the place in the list / the global env index): exactly so in the
// synchronous step (k_group_apply_exact), and for groups of up to kInlineBucket records in the asynchronous rollout.
struct Buckets {
    u32* boff;      // [2^bits] records per bucket, then (k_bucket_scan) offset inside its block of 1024 buckets
    u32* sums;      // [blocks] offset of every block of 1024 buckets, [blocks] = total, [blocks + 1] = long buckets queued
    u32* work;      // [2^bits] queue of the buckets with more than kInlineBucket records
    int bits;       // 10 .. 22
    __host__ __device__ u32 count() const { return 1u << bits; }
    __host__ __device__ u32 blocks() const { return 1u << (bits - 10); }
};
constexpr u32 kInlineBucket = 24;     // buckets of more records than this are applied by a warp
constexpr u32 kFoldBucket = 64;        // asynchronous rollout: ... and from this size on the warp folds them segment-wise
__device__ __forceinline__ u32 bucket_of(u64 key, int bits) { return (u32)((key * 0x9E3779B97F4A7C15ull) >> (64 - bits)); }
__device__ __forceinline__ void bucket_range(const Buckets& B, u32 b, u32& begin, u32& end) {
    begin = B.sums[b >> 10] + B.boff[b];
    end = (b + 1 == B.count()) ? B.sums[B.blocks()] : B.sums[(b + 1) >> 10] + B.boff[b + 1];
}
__global__ void __launch_bounds__(256)
k_defer_count(const ulonglong2* rec, const unsigned long long* count, unsigned long long cap, u32* pos, Buckets B) {
    const unsigned long long m = *count < cap ? *count : cap;
    for (unsigned long long j = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; j < m; j += (unsigned long long)gridDim.x * blockDim.x) {
        const u64 k = rec[j].x;
        if (k != ~0ull) pos[j] = atomicAdd(&B.boff[bucket_of(k, B.bits)], 1u);
    }
}
// exclusive prefix sum over the bucket sizes: every block scans 1024 of them and queues its long buckets, then one block
// scans the block totals
__global__ void __launch_bounds__(1024) k_bucket_scan(Buckets B) {
    __shared__ u32 warp_tot[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 idx = blockIdx.x * 1024u + threadIdx.x;
    const u32 v = B.boff[idx];
    if (v > kInlineBucket) B.work[atomicAdd(&B.sums[B.blocks() + 1], 1u)] = idx;
    u32 x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { u32 y = __shfl_up_sync(0xFFFFFFFFu, x, d); if (lane >= d) x += y; }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
        u32 w = warp_tot[lane], z = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { u32 y = __shfl_up_sync(0xFFFFFFFFu, z, d); if (lane >= d) z += y; }
        warp_tot[lane] = z - w;
        if (lane == 31) B.sums[blockIdx.x] = z;        // this block's total
    }
    __syncthreads();
    B.boff[idx] = x - v + warp_tot[warp];              // exclusive, within the block
}
__global__ void __launch_bounds__(1024) k_bucket_scan_sums(Buckets B) {   // block totals -> exclusive offsets, grand total behind them
    __shared__ u32 warp_tot[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 nb = B.blocks(), per = (nb + 1023) / 1024, lo = threadIdx.x * per;
    u32 mine = 0;
    for (u32 i = lo; i < lo + per && i < nb; ++i) mine += B.sums[i];
    u32 x = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { u32 y = __shfl_up_sync(0xFFFFFFFFu, x, d); if (lane >= d) x += y; }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
        u32 w = warp_tot[lane], z = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { u32 y = __shfl_up_sync(0xFFFFFFFFu, z, d); if (lane >= d) z += y; }
        warp_tot[lane] = z - w;
        if (lane == 31) B.sums[nb] = z;                // grand total
    }
    __syncthreads();
    u32 run = x - mine + warp_tot[warp];
    for (u32 i = lo; i < lo + per && i < nb; ++i) { const u32 v = B.sums[i]; B.sums[i] = run; run += v; }
}
__global__ void __launch_bounds__(256)
k_defer_scatter(const ulonglong2* rec, const unsigned long long* count, unsigned long long cap, const u32* pos, Buckets B,
                ulonglong2* out) {
    const unsigned long long m = *count < cap ? *count : cap;
    for (unsigned long long j = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; j < m; j += (unsigned long long)gridDim.x * blockDim.x) {
        const ulonglong2 r = rec[j];
        if (r.x == ~0ull) continue;
        const u32 b = bucket_of(r.x, B.bits);
        // one 16-byte store per record: {address, target bits | its place in the list << 32} (the list index = the order
        // in which its warp appended it)
        out[B.sums[b >> 10] + B.boff[b] + pos[j]] = make_ulonglong2(r.x, (r.y & 0xFFFFFFFFull) | ((u64)j << 32));
    }
}
// q <- q + lr (target - q) record after record; the value is written back with a compare-and-swap from what was read,
// so that a rollout running on another stream cannot be overwritten (it rarely is: start again).  One kernel, two kinds
// of blocks that run side by side: the first `long_blocks` blocks take the long buckets (more than kInlineBucket records:
// a popular start state's action collects thousands per launch; queued by k_bucket_scan), one WARP per bucket; the other
// blocks take one bucket per THREAD and skip the long ones.
template <class TAB>
__device__ __forceinline__ u32 load_q_bits(const float* p) {   // coherent at L2 / at the owner GPU
    u32 v;
    if (TAB::kSysLoad) asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
template <class TAB>
__global__ void __launch_bounds__(256)
k_defer_apply(const __grid_constant__ TAB table, ulonglong2* rec, Buckets B, float lr, int long_blocks) {
    __shared__ Slot* shard_base[G2048_MAX_PEERS];
    const auto tab = table.view(shard_base);
    using VIEW = decltype(tab);
    if ((int)blockIdx.x >= long_blocks) {
        const u32 b = (blockIdx.x - long_blocks) * blockDim.x + threadIdx.x;
        if (b >= B.count()) return;
        u32 begin, end;
        bucket_range(B, b, begin, end);
        if (end - begin > kInlineBucket) return;
        for (u32 r = begin; r < end; ++r) {
            const u64 k = rec[r].x;
            if (k == ~0ull) continue;                  // applied together with an earlier record of the same address
            float* qp = &tab.at(k >> 2)->q[k & 3];
            u32 seen = load_q_bits<VIEW>(qp);
            for (;;) {
                // the records of this address in the order of the list (an env's own updates of one value keep their
                // order: a warp appends in time order): repeatedly the smallest list index above the last one
                float q = __uint_as_float(seen);
                long long last = -1;
                for (;;) {
                    u32 best = ~0u;
                    float best_target = 0.f;
                    for (u32 t = r; t < end; ++t) {
                        const ulonglong2 x = rec[t];
                        const u32 ix = (u32)(x.y >> 32);
                        if (x.x == k && (long long)ix > last && ix < best) { best = ix; best_target = __uint_as_float((u32)x.y); }
                    }
                    if (best == ~0u) break;
                    q = td_apply(q, lr, best_target);
                    last = (long long)best;
                }
                const u32 old = cas32<VIEW::kSys>(reinterpret_cast<u32*>(qp), seen, __float_as_uint(q));
                if (old == seen) break;
                seen = old;
            }
            for (u32 t = r + 1; t < end; ++t)
                if (rec[t].x == k) rec[t].x = ~0ull;
        }
        return;
    }
    // for every address of the bucket (in order of first appearance) the lanes fetch the bucket 32 records at a time and
    // fold the targets into the chain in record order: 3 dependent float operations per record, the sequential update
    const int lane = threadIdx.x & 31;
    const u32 n_warps = (u32)long_blocks * (blockDim.x >> 5), n_work = B.sums[B.blocks() + 1];
    for (u32 w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < n_work; w += n_warps) {
        u32 begin, end;
        bucket_range(B, B.work[w], begin, end);
        u32 r = begin;
        for (;;) {                                     // warp-uniform
            // the next record that has not been applied yet, 32 at a time
            u64 k = ~0ull;
            while (r < end) {
                const u32 t = r + lane;
                const u64 kt = t < end ? rec[t].x : ~0ull;
                const unsigned m = __ballot_sync(0xFFFFFFFFu, kt != ~0ull);
                if (m) {
                    const int l = __ffs((int)m) - 1;
                    k = __shfl_sync(0xFFFFFFFFu, kt, l);
                    r += (u32)l;
                    break;
                }
                r += 32;
            }
            if (r >= end) break;
            float* qp = &tab.at(k >> 2)->q[k & 3];
            u32 seen = load_q_bits<VIEW>(qp);
            if (end - r >= kFoldBucket) {
                // Thousands of records on one value (the action a popular start state takes): a chain of that length
                // would be the whole phase.  Every lane folds the records r + lane, r + lane + 32, ... into its own map
                // q -> A q + B (B = the exact chain started from 0, A = (1 - lr)^count), then the 32 maps are composed
                // lane after lane: one sequential order of all the updates, evaluated segment-wise (equal to the plain
                // chain up to float32 rounding of the composition).
                const float keep = __fsub_rn(1.0f, lr);
                float A = 1.0f, B = 0.0f;
                for (u32 t = r + lane; t < end; t += 32) {
                    const ulonglong2 x = rec[t];
                    if (x.x == k) { B = td_apply(B, lr, __uint_as_float((u32)x.y)); A = __fmul_rn(A, keep); }
                }
                for (;;) {
                    float q = __uint_as_float(seen);
#pragma unroll
                    for (int l = 0; l < 32; ++l)
                        q = __fadd_rn(__fmul_rn(__shfl_sync(0xFFFFFFFFu, A, l), q), __shfl_sync(0xFFFFFFFFu, B, l));
                    u32 old = 0;
                    if (lane == 0) old = cas32<VIEW::kSys>(reinterpret_cast<u32*>(qp), seen, __float_as_uint(q));
                    old = __shfl_sync(0xFFFFFFFFu, old, 0);
                    if (old == seen) break;
                    seen = old;
                }
            } else
            for (;;) {
                float q = __uint_as_float(seen);
                u32 t = r + lane;
                ulonglong2 x = t < end ? rec[t] : make_ulonglong2(~k, 0ull);
                bool mine = x.x == k;
                float tt = __uint_as_float((u32)x.y);
                for (u32 base = r; base < end; base += 32) {
                    const u32 tn = base + 32 + lane;           // the next chunk is in flight while this one is folded
                    x = tn < end ? rec[tn] : make_ulonglong2(~k, 0ull);
                    const bool mine_n = x.x == k;
                    const float tt_n = __uint_as_float((u32)x.y);
                    unsigned m = __ballot_sync(0xFFFFFFFFu, mine);
                    if (m == 0xFFFFFFFFu) {
#pragma unroll
                        for (int l = 0; l < 32; ++l) q = td_apply(q, lr, __shfl_sync(0xFFFFFFFFu, tt, l));
                    } else {
                        while (m) {
                            const int l = __ffs((int)m) - 1;
                            m &= m - 1;
                            q = td_apply(q, lr, __shfl_sync(0xFFFFFFFFu, tt, l));
                        }
                    }
                    mine = mine_n;
                    tt = tt_n;
                }
                u32 old = 0;
                if (lane == 0) old = cas32<VIEW::kSys>(reinterpret_cast<u32*>(qp), seen, __float_as_uint(q));
                old = __shfl_sync(0xFFFFFFFFu, old, 0);
                if (old == seen) break;
                seen = old;
            }
            for (u32 t = r + lane; t < end; t += 32)
                if (rec[t].x == k) rec[t].x = ~0ull;
            __syncwarp();
            r += 1;
        }
    }
}

template <bool INSERT>
__global__ void __launch_bounds__(256)
k_q_lookup(Slot* tab, u64 mask, const u64* keys, long long n, float4* rows, uint8_t* found) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float4 row;
        u32 ins = 0;
        u32 slot = table_find<INSERT>(tab, mask, keys[i], row, ins);
        rows[i] = row;
        if (found) found[i] = (slot != kNoSlot) && !ins;
    }
}
template <bool INSERT, class TAB>
__global__ void __launch_bounds__(256)
k_q_lookup_tab(const __grid_constant__ TAB table, const u64* keys, long long n, float4* rows, uint8_t* found) {
    __shared__ Slot* shard_base[G2048_MAX_PEERS];
    const auto tab = table.view(shard_base);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float4 row;
        u32 ins = 0;
        u32 slot = table_find<INSERT>(tab, keys[i], row, ins);
        rows[i] = row;
        if (found) found[i] = (slot != kNoSlot) && !ins;
    }
}
__global__ void __launch_bounds__(256)
k_choose_action(Slot* tab, u64 mask, const u64* boards, uint8_t* actions, long long n, u64 eps_thresh, u64 seed, u64 t,
                u64 id_base) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float4 row;
        u32 ins = 0;
        table_find<true>(tab, mask, boards[i], row, ins);
        Draw4 x = philox(seed, id_base + (u64)i, t, G2048_STREAM_STEP);
        actions[i] = (uint8_t)choose_action(row, x, eps_thresh);
    }
}
__global__ void __launch_bounds__(256) k_q_size(const Slot* tab, u64 capacity, long long* count) {
    long long c = 0;
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < capacity; i += (u64)gridDim.x * blockDim.x)
        c += __ldcg(&tab[i].key) != 0;
    c = warp_sum(c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd((unsigned long long*)count, (unsigned long long)c);
}
// occupancy statistics: out[0] = states, out[1] = sum over states of the number of probes that precede their slot in
// their key's probe sequence, out[2] = the largest such number.  A lookup of a stored state costs 1 + that many probes.
__global__ void __launch_bounds__(256) k_q_probe_stats(const Slot* tab, u64 capacity, long long* out) {
    long long cnt = 0, sum = 0, mx = 0;
    const u64 mask = capacity - 1;
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < capacity; i += (u64)gridDim.x * blockDim.x) {
        u64 k = __ldcg(&tab[i].key);
        if (k) {
            // probes before this slot in its key's sequence (pair by pair, the home slot's side first: next_probe)
            const u64 home = mix64(k) & mask;
            long long d = (long long)(2 * (((i >> 1) - (home >> 1)) & (mask >> 1)) + ((i ^ home) & 1));
            cnt += 1; sum += d; mx = max(mx, d);
        }
    }
    cnt = warp_sum(cnt);
    sum = warp_sum(sum);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, s));
    if ((threadIdx.x & 31) == 0 && cnt) {
        atomicAdd((unsigned long long*)out, (unsigned long long)cnt);
        atomicAdd((unsigned long long*)(out + 1), (unsigned long long)sum);
        atomicMax(out + 2, mx);
    }
}
__global__ void __launch_bounds__(256)
k_q_export(const Slot* tab, u64 capacity, u64* keys, float4* rows, long long max_out, long long* count) {
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < capacity; i += (u64)gridDim.x * blockDim.x) {
        u64 k;
        float4 q;
        load_slot(tab + i, k, q);
        if (k) {
            long long j = (long long)atomicAdd((unsigned long long*)count, 1ull);
            if (j < max_out) { keys[j] = k; rows[j] = q; }
        }
    }
}

__global__ void __launch_bounds__(256)
k_move_trial(Tables T, const u64* in, const uint8_t* actions, u64* out, uint8_t* moved, int* move_score, long long n) {
    Lut L = global_lut(T);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        Moved m = do_move(in[i], actions[i] & 3, L);
        if (out) out[i] = m.board;
        if (moved) moved[i] = m.moved;
        if (move_score) move_score[i] = m.score;
    }
}
__global__ void __launch_bounds__(256) k_legal_mask(const u64* boards, uint8_t* out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = (uint8_t)legal_mask(boards[i]);
}
// one thread per cell: coalesced 8-byte tile reads, the 16 lanes of a board OR their nibbles together
__global__ void __launch_bounds__(256) k_pack(const long long* tiles, u64* boards, long long n, long long* bad) {
    long long total = n * 16;
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < ((total + 31) & ~31ll);
         j += (long long)gridDim.x * blockDim.x) {
        long long v = j < total ? tiles[j] : 0;
        int lvl = 0, is_bad = 0;
        if (v != 0) {
            lvl = 63 - __clzll(v);
            if (v < 0 || (1ll << lvl) != v || lvl < 1 || lvl > 15) { is_bad = 1; lvl = 0; }
        }
        u64 nib = (u64)lvl << (4 * (j & 15));
#pragma unroll
        for (int s = 8; s > 0; s >>= 1) nib |= __shfl_xor_sync(0xFFFFFFFFu, nib, s);
        if ((j & 15) == 0 && j < total) boards[j >> 4] = nib;
        if (is_bad && bad) atomicAdd((unsigned long long*)bad, 1ull);
    }
}
__global__ void __launch_bounds__(256) k_unpack(const u64* boards, long long* tiles, long long n) {
    long long total = n * 16;
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < total; j += (long long)gridDim.x * blockDim.x) {
        int lvl = (int)((boards[j >> 4] >> (4 * (j & 15))) & 15);
        tiles[j] = lvl ? (1ll << lvl) : 0;
    }
}
// One-hot pieces: group g of a board = level plane (g >> 2), row (g & 3): four consecutive cells
__device__ __forceinline__ float4 onehot_f32(u64 b, int g) {
    u32 row = (u32)(b >> (16 * (g & 3))) & 0xFFFFu, lvl = (u32)(g >> 2);
    return make_float4((row & 15) == lvl ? 1.f : 0.f, ((row >> 4) & 15) == lvl ? 1.f : 0.f,
                       ((row >> 8) & 15) == lvl ? 1.f : 0.f, ((row >> 12) & 15) == lvl ? 1.f : 0.f);
}
__device__ __forceinline__ uint2 onehot_bf16(u64 b, int g) {   // bf16 1.0 = 0x3F80
    u32 row = (u32)(b >> (16 * (g & 3))) & 0xFFFFu, lvl = (u32)(g >> 2);
    return make_uint2(((row & 15) == lvl ? 0x3F80u : 0u) | (((row >> 4) & 15) == lvl ? 0x3F800000u : 0u),
                      (((row >> 8) & 15) == lvl ? 0x3F80u : 0u) | (((row >> 12) & 15) == lvl ? 0x3F800000u : 0u));
}
// encode_state: out[n][16][4][4]; one thread writes 4 consecutive cells (one row of one level plane)
__global__ void __launch_bounds__(256) k_onehot_f32(const u64* boards, float4* out, long long n) {
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n * 64; j += (long long)gridDim.x * blockDim.x)
        out[j] = onehot_f32(boards[j >> 6], (int)(j & 63));
}
__global__ void __launch_bounds__(256) k_onehot_bf16(const u64* boards, uint2* out, long long n) {
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n * 64; j += (long long)gridDim.x * blockDim.x)
        out[j] = onehot_bf16(boards[j >> 6], (int)(j & 63));
}
// act / act_ripetitive (Dqn8TestNOPERCNN.py:312-336): lm == 0 -> act (uniform over 4 / plain argmax), else uniform over
// the legal moves (np.random.choice) / first maximum among the legal moves in ascending action order
__device__ __forceinline__ int select_action(const float4& q, u32 lm, const Draw4& x, u64 eps_thresh) {
    bool explore = (u64)x.x2 < eps_thresh;
    if (lm == 0) return explore ? (int)(x.x3 >> 30) : argmax4(q);
    if (explore) {
        int j = (int)__umulhi(x.x3, (u32)__popc(lm));
        u32 m = lm;
        for (int s = 0; s < j; ++s) m &= m - 1;
        return __ffs((int)m) - 1;
    }
    int act = -1;
    float best = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if ((lm >> c) & 1u) {
            float v = q_at(q, c);
            if (act < 0 || v > best) { best = v; act = c; }
        }
    return act;
}
__global__ void __launch_bounds__(256)
k_select_action(const float4* qv, const uint8_t* legal, uint8_t* actions, long long n, u64 eps_thresh, u64 seed, u64 t,
                u64 id_base) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        Draw4 x = philox(seed, id_base + (u64)i, t, G2048_STREAM_STEP);
        actions[i] = (uint8_t)select_action(qv[i], legal ? (legal[i] & 15u) : 0u, x, eps_thresh);
    }
}

// The env side of one DQN driver step (mainDQL_CNN_step2.py:163-237) fused: select -> step -> terminal bonus -> reset ->
// legal mask -> one-hot.  One thread per env for the game logic; the one-hot planes of the block's 256 boards are
// then written by the whole block, 16 bytes per thread and store instruction, fully coalesced (1 KB per board).
template <bool BF16>
__global__ void __launch_bounds__(256)
k_dqn_env_step(Tables T, u64* boards, int* score, const float4* qv, const uint8_t* legal_in, uint8_t* actions, u64* state_out,
               u64* next_out, float* reward, uint8_t* done_out, uint8_t* legal_out, void* onehot, long long n,
               u64 eps_thresh, u32 opts, u64 seed, u64 t, u64 reset_idx, u64 id_base) {
    __shared__ u64 sb[256];
    Lut L = global_lut(T);
    long long block0 = (long long)blockIdx.x * 256;
    long long i = block0 + threadIdx.x;
    u64 cont = 0;
    if (i < n) {
        u64 id = id_base + (u64)i;
        Env e;
        env_load(e, boards[i], G2048_AUX_INIT, score ? score[i] : 0);
        Draw4 x = philox(seed, id, t, G2048_STREAM_STEP);
        int a = select_action(qv[i], legal_in ? (legal_in[i] & 15u) : 0u, x, eps_thresh);
        u64 s0 = e.board;
        StepOut o;
        philox_step<G2048_FLAVOUR_NOPENALTY>(e, a, x, seed, id, t, L, T, o);
        float r = (float)o.reward;
        if (o.done && (opts & G2048_DQN_TERMINAL_BONUS)) {   // mainDQL_CNN_step2.py:202-213
            u64 b = e.board;
            int big = __popcll((b >> 3) & ((b >> 2) | (b >> 1)) & kNib1);   // cells holding 1024 or more
            if (o.maxlvl >= 11) r += 100.f;
            else if (o.maxlvl >= 10 && big >= 2) r += 50.f;
        }
        if (actions) actions[i] = (uint8_t)a;
        if (state_out) state_out[i] = s0;
        if (next_out) next_out[i] = e.board;
        if (reward) reward[i] = r;
        if (done_out) done_out[i] = o.done;
        if (o.done && (opts & G2048_DQN_AUTO_RESET)) {
            Draw4 y = philox(seed, id, reset_idx, G2048_STREAM_RESET);
            e.board = fresh_board<false>(y.x0, y.x1, y.x2, y.x3);
            e.score = 0;
        }
        boards[i] = e.board;
        if (score) score[i] = e.score;
        if (legal_out) legal_out[i] = (uint8_t)legal_mask(e.board);
        cont = e.board;
    }
    sb[threadIdx.x] = cont;
    __syncthreads();
    if (!onehot) return;
    long long nb = n - block0 < 256 ? n - block0 : 256;   // boards of this block
    for (int g = threadIdx.x; g < nb * 64; g += 256) {
        u64 b = sb[g >> 6];
        if (BF16) reinterpret_cast<uint2*>(onehot)[block0 * 64 + g] = onehot_bf16(b, g & 63);
        else reinterpret_cast<float4*>(onehot)[block0 * 64 + g] = onehot_f32(b, g & 63);
    }
}

}  // namespace

// ============================================================================================ C ABI
G2048_API int g2048_version(void) { return G2048_VERSION; }
G2048_API const char* g2048_last_error(void) { return g_err; }
G2048_API int g2048_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
G2048_API void* g2048_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { fail((int)cudaGetLastError(), "cudaHostAlloc"); return nullptr; }
    return p;
}
G2048_API void g2048_host_free(void* p) { if (p) cudaFreeHost(p); }

G2048_API void g2048_host_tables(uint16_t* row, uint8_t* merged, uint32_t* mscore, double* reward_valid,
                                 double* reward_invalid, double* pen) {
    std::vector<uint16_t> r;
    std::vector<uint8_t> m;
    std::vector<uint32_t> ms;
    std::vector<double> v, iv, p;
    build_row_tables(r, m, ms);
    build_reward_tables(v, iv, p);
    for (unsigned i = 0; i < 65536; ++i) {   // undo the bank swizzle
        unsigned s = lut_index2(i) & 0xFFFFu;
        if (row) row[i] = r[s];
        if (merged) merged[i] = m[s];
    }
    if (mscore) memcpy(mscore, ms.data(), ms.size() * sizeof(uint32_t));
    if (reward_valid) memcpy(reward_valid, v.data(), v.size() * sizeof(double));
    if (reward_invalid) memcpy(reward_invalid, iv.data(), iv.size() * sizeof(double));
    if (pen) memcpy(pen, p.data(), p.size() * sizeof(double));
}

G2048_API int g2048_init(int device) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (device < 0 || device >= kMaxDevices) return fail(G2048_ERR_ARG, "g2048_init: bad device index");
    CK(cudaSetDevice(device));
    DeviceState& d = g_dev[device];
    if (d.ready) return 0;
    std::vector<uint16_t> row;
    std::vector<uint8_t> merged;
    std::vector<uint32_t> mscore;
    std::vector<double> valid, invalid, pen;
    build_row_tables(row, merged, mscore);
    build_reward_tables(valid, invalid, pen);
    void *lut = nullptr, *rv = nullptr, *ri = nullptr, *pn = nullptr;
    CK(cudaMalloc(&lut, kLutBytes));
    CK(cudaMalloc(&rv, valid.size() * sizeof(double)));
    CK(cudaMalloc(&ri, invalid.size() * sizeof(double)));
    CK(cudaMalloc(&pn, pen.size() * sizeof(double)));
    CK(cudaMemcpy(lut, row.data(), kLutRowBytes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy((char*)lut + kLutRowBytes, merged.data(), kLutMergedBytes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy((char*)lut + kLutRowBytes + kLutMergedBytes, mscore.data(), 256 * sizeof(uint32_t), cudaMemcpyHostToDevice));
    {   // hot reward entries, staged into shared memory together with the LUT
        std::vector<double> hot(kHotDoubles, 0.0);
        for (int i = 0; i < 512; ++i) hot[kHotInvalid + i] = invalid[i];
        for (int i = 0; i < 32; ++i) hot[kHotPen + i] = pen[i];
        for (int lvl = 0; lvl < 16; ++lvl)
            for (int d = 0; d < 2; ++d)
                for (int s4 = 0; s4 < 64; ++s4) hot[kHotValid + (lvl * 2 + d) * 64 + s4] = valid[(lvl * 16 + d) * 256 + s4];
        CK(cudaMemcpy((char*)lut + kLutCoreBytes, hot.data(), hot.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    CK(cudaMemcpy(rv, valid.data(), valid.size() * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ri, invalid.data(), invalid.size() * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(pn, pen.data(), pen.size() * sizeof(double), cudaMemcpyHostToDevice));
    d.lut = lut;
    d.tables = Tables{(const uint16_t*)lut, (const uint8_t*)lut + kLutRowBytes,
                      (const uint32_t*)((const char*)lut + kLutRowBytes + kLutMergedBytes), (const double*)rv,
                      (const double*)ri, (const double*)pn};
    CK(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, device));
    CK(cudaMalloc(&d.queue, kQueueSlots * sizeof(unsigned long long)));
    CK(cudaMalloc(&d.abort_flag, sizeof(int)));
    CK(cudaMemset(d.abort_flag, 0, sizeof(int)));
    CK(cudaFuncSetAttribute(k_rollout_random<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLutBytes));
    CK(cudaFuncSetAttribute(k_rollout_random<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLutBytes));
    CK(cudaFuncSetAttribute(k_env_step<0, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLutBytes));
    CK(cudaFuncSetAttribute(k_env_step<0, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLutBytes));
    CK(cudaFuncSetAttribute(k_env_step<1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLutBytes));
    CK(cudaFuncSetAttribute(k_env_step<1, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLutBytes));
    CK(cudaFuncSetAttribute(k_routed_request<0, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLutBytes));
    CK(cudaFuncSetAttribute(k_routed_request<1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLutBytes));
    CK(cudaFuncSetAttribute(k_rollout_qlearn<0, true, LocalTable>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLutBytes));
    CK(cudaFuncSetAttribute(k_rollout_qlearn<1, true, LocalTable>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLutBytes));
    CK(cudaFuncSetAttribute(k_rollout_qlearn<0, true, ShardedTable>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLutBytes));
    CK(cudaFuncSetAttribute(k_rollout_qlearn<1, true, ShardedTable>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLutBytes));
    d.ready = true;
    return 0;
}

#define DEVSTATE()                         \
    DeviceState* D = nullptr;              \
    {                                      \
        int rc_ = current_device_state(&D); \
        if (rc_) return rc_;               \
    }
#define LAUNCH_CHECK(what)                                   \
    do {                                                     \
        cudaError_t e_ = cudaGetLastError();                 \
        if (e_ != cudaSuccess) return fail((int)e_, what);   \
    } while (0)

G2048_API int g2048_env_reset(uint64_t* boards, int32_t* score, const uint8_t* mask, const uint8_t* replay_draws,
                              int64_t n, uint64_t seed, uint64_t episode_idx, uint64_t env_id_base, void* stream) {
    DEVSTATE();
    if (n < 0 || (n && !boards)) return fail(G2048_ERR_ARG, "g2048_env_reset: bad arguments");
    if (n == 0) return 0;
    int g = grid_for(n, 256, D->sm_count);
    if (replay_draws)
        k_env_reset<true><<<g, 256, 0, S(stream)>>>((u64*)boards, score, mask, replay_draws, n, seed, episode_idx, env_id_base);
    else
        k_env_reset<false><<<g, 256, 0, S(stream)>>>((u64*)boards, score, mask, nullptr, n, seed, episode_idx, env_id_base);
    LAUNCH_CHECK("k_env_reset");
    return 0;
}

G2048_API int g2048_env_step(uint64_t* boards, uint64_t* aux, int32_t* score, const uint8_t* actions,
                             const uint8_t* replay_draws, double* reward_f64, float* reward_f32, uint8_t* flags,
                             uint8_t* maxlvl, int32_t* move_score, int64_t n, int flavour, uint64_t seed,
                             uint64_t step_idx, uint64_t env_id_base, void* stream) {
    DEVSTATE();
    if (n < 0 || (n && (!boards || !actions)) || (flavour != 0 && flavour != 1))
        return fail(G2048_ERR_ARG, "g2048_env_step: bad arguments");
    if (n == 0) return 0;
    const bool big = n >= kEnvStepSmemLutMinEnvs;   // enough work to amortise the 213 KB staging copy per SM
    int g = big ? D->sm_count : grid_for(n, 256, D->sm_count);
    int blk = big ? kRolloutThreads : 256;
    size_t smem = big ? kLutBytes : 0;
#define STEP2(F, R, SM)                                                                                              \
    k_env_step<F, R, SM><<<g, blk, smem, S(stream)>>>(D->tables, (u64*)boards, (u64*)aux, score, actions, replay_draws, \
                                                      reward_f64, reward_f32, flags, maxlvl, move_score, n, seed,     \
                                                      step_idx, env_id_base)
#define STEP(F, R) do { if (big) STEP2(F, R, true); else STEP2(F, R, false); } while (0)
    if (flavour == 0) { if (replay_draws) STEP(0, true); else STEP(0, false); }
    else { if (replay_draws) STEP(1, true); else STEP(1, false); }
#undef STEP
#undef STEP2
    LAUNCH_CHECK("k_env_step");
    return 0;
}

G2048_API int g2048_move_trial(const uint64_t* boards_in, const uint8_t* actions, uint64_t* out_boards, uint8_t* moved,
                               int32_t* move_score, int64_t n, void* stream) {
    DEVSTATE();
    if (n < 0 || (n && (!boards_in || !actions))) return fail(G2048_ERR_ARG, "g2048_move_trial: bad arguments");
    if (n == 0) return 0;
    k_move_trial<<<grid_for(n, 256, D->sm_count), 256, 0, S(stream)>>>(D->tables, (const u64*)boards_in, actions,
                                                                        (u64*)out_boards, moved, move_score, n);
    LAUNCH_CHECK("k_move_trial");
    return 0;
}
G2048_API int g2048_legal_mask(const uint64_t* boards, uint8_t* out, int64_t n, void* stream) {
    DEVSTATE();
    if (n < 0 || (n && (!boards || !out))) return fail(G2048_ERR_ARG, "g2048_legal_mask: bad arguments");
    if (n == 0) return 0;
    k_legal_mask<<<grid_for(n, 256, D->sm_count), 256, 0, S(stream)>>>((const u64*)boards, out, n);
    LAUNCH_CHECK("k_legal_mask");
    return 0;
}
G2048_API int g2048_pack_i64(const int64_t* tiles, uint64_t* boards, int64_t n, int64_t* bad_count, void* stream) {
    DEVSTATE();
    if (n < 0 || (n && (!tiles || !boards))) return fail(G2048_ERR_ARG, "g2048_pack_i64: bad arguments");
    if (n == 0) return 0;
    k_pack<<<grid_for(n * 16, 256, D->sm_count), 256, 0, S(stream)>>>((const long long*)tiles, (u64*)boards, n,
                                                                       (long long*)bad_count);
    LAUNCH_CHECK("k_pack");
    return 0;
}
G2048_API int g2048_unpack_i64(const uint64_t* boards, int64_t* tiles, int64_t n, void* stream) {
    DEVSTATE();
    if (n < 0 || (n && (!tiles || !boards))) return fail(G2048_ERR_ARG, "g2048_unpack_i64: bad arguments");
    if (n == 0) return 0;
    k_unpack<<<grid_for(n * 16, 256, D->sm_count), 256, 0, S(stream)>>>((const u64*)boards, (long long*)tiles, n);
    LAUNCH_CHECK("k_unpack");
    return 0;
}
G2048_API int g2048_encode_onehot(const uint64_t* boards, void* out, int64_t n, int dtype, void* stream) {
    DEVSTATE();
    if (n < 0 || (n && (!boards || !out)) || (dtype != G2048_DTYPE_F32 && dtype != G2048_DTYPE_BF16))
        return fail(G2048_ERR_ARG, "g2048_encode_onehot: bad arguments");
    if (n == 0) return 0;
    int g = grid_for(n * 64, 256, D->sm_count, 16);
    if (dtype == G2048_DTYPE_F32) k_onehot_f32<<<g, 256, 0, S(stream)>>>((const u64*)boards, (float4*)out, n);
    else k_onehot_bf16<<<g, 256, 0, S(stream)>>>((const u64*)boards, (uint2*)out, n);
    LAUNCH_CHECK("k_onehot");
    return 0;
}
G2048_API int g2048_select_action(const float* qvalues, const uint8_t* legal_mask, uint8_t* actions, int64_t n,
                                  double eps, uint64_t seed, uint64_t step_idx, uint64_t env_id_base, void* stream) {
    DEVSTATE();
    if (n < 0 || (n && (!qvalues || !actions))) return fail(G2048_ERR_ARG, "g2048_select_action: bad arguments");
    if (n == 0) return 0;
    k_select_action<<<grid_for(n, 256, D->sm_count), 256, 0, S(stream)>>>((const float4*)qvalues, legal_mask, actions, n,
                                                                           eps_threshold(eps), seed, step_idx, env_id_base);
    LAUNCH_CHECK("k_select_action");
    return 0;
}

G2048_API int g2048_dqn_env_step(uint64_t* boards, int32_t* score, const float* qvalues, const uint8_t* legal_in,
                                 uint8_t* actions, uint64_t* state_out, uint64_t* next_state_out, float* reward,
                                 uint8_t* done, uint8_t* legal_out, void* onehot_out, int dtype, int64_t n, double eps,
                                 uint32_t opts, uint64_t seed, uint64_t step_idx, uint64_t reset_idx,
                                 uint64_t env_id_base, void* stream) {
    DEVSTATE();
    if (n < 0 || (n && (!boards || !qvalues)) || (dtype != G2048_DTYPE_F32 && dtype != G2048_DTYPE_BF16))
        return fail(G2048_ERR_ARG, "g2048_dqn_env_step: bad arguments");
    if (n == 0) return 0;
    int g = (int)((n + 255) / 256);
#define LAUNCH_DQN(B)                                                                                                   \
    k_dqn_env_step<B><<<g, 256, 0, S(stream)>>>(D->tables, (u64*)boards, score, (const float4*)qvalues, legal_in, actions,  \
                                                (u64*)state_out, (u64*)next_state_out, reward, done, legal_out, onehot_out, \
                                                n, eps_threshold(eps), opts, seed, step_idx, reset_idx, env_id_base)
    if (dtype == G2048_DTYPE_BF16) LAUNCH_DQN(true); else LAUNCH_DQN(false);
#undef LAUNCH_DQN
    LAUNCH_CHECK("k_dqn_env_step");
    return 0;
}

// fused rollouts: one persistent CTA per SM with the LUT in shared memory once there is enough work to
// amortise the 213 KB staging copy; small batches read the LUT through L1 instead.
static inline void rollout_geometry(const DeviceState* D, int64_t n, int& grid, int& block, int& smem_lut, size_t& smem,
                                    int big_block = kRolloutThreads) {
    if (n >= 16384) {
        block = big_block;
        grid = (int)((n + block - 1) / block);
        if (grid > D->sm_count) grid = D->sm_count;
        smem_lut = 1;
        smem = kLutBytes;
    } else {
        block = kSmallRolloutThreads;
        grid = (int)((n + block - 1) / block);
        if (grid < 1) grid = 1;
        smem_lut = 0;
        smem = 0;
    }
}

G2048_API int g2048_rollout_random(uint64_t* boards, uint64_t* aux, int32_t* score, int64_t n, int64_t k_steps,
                                   int flavour, uint64_t seed, uint64_t step_base, uint64_t env_id_base,
                                   int64_t* counters, void* stream) {
    DEVSTATE();
    if (n < 0 || k_steps < 0 || (n && !boards) || (flavour != 0 && flavour != 1))
        return fail(G2048_ERR_ARG, "g2048_rollout_random: bad arguments");
    if (n == 0 || k_steps == 0) return 0;
    int grid, block, smem_lut;
    size_t smem;
    rollout_geometry(D, n, grid, block, smem_lut, smem);
#define LAUNCH_RR(F, SM)                                                                                                \
    k_rollout_random<F, SM><<<grid, block, smem, S(stream)>>>(D->tables, (u64*)boards, (u64*)aux, score, n, k_steps, seed, \
                                                              step_base, env_id_base, (long long*)counters)
    if (flavour == 0) { if (smem_lut) LAUNCH_RR(0, true); else LAUNCH_RR(0, false); }
    else { if (smem_lut) LAUNCH_RR(1, true); else LAUNCH_RR(1, false); }
#undef LAUNCH_RR
    LAUNCH_CHECK("k_rollout_random");
    return 0;
}

namespace {
struct Scratch;
size_t scratch_bytes(int64_t n);
int carve(void* scratch, size_t bytes, int64_t n, Scratch& s);
int apply_records(DeviceState* D, Slot* tab, uint64_t capacity, Scratch& s, int64_t n, float lr, int mode, cudaStream_t st,
                  int kshift = 0, const int* abort_flag = nullptr);

struct DeferBuffers;
template <class TAB>
int launch_rollout_qlearn(DeviceState* D, const TAB& tab, uint64_t* boards, uint64_t* aux, int32_t* score, int64_t n,
                          int64_t k_steps, int flavour, float lr, float gamma, double eps, uint64_t seed,
                          uint64_t step_base, uint64_t env_id_base, int64_t* counters, void* stream,
                          const DeferBuffers* shared_list = nullptr);
// shards[j] = device pointer to slots_per_shard slots (local or peer memory); n_shards and slots_per_shard powers of two
int make_sharded(const void* const* shards, int n_shards, uint64_t slots_per_shard, ShardedTable& t, const char* who) {
    if (!shards || n_shards < 1 || n_shards > G2048_MAX_PEERS || (n_shards & (n_shards - 1)) || !pow2(slots_per_shard) ||
        slots_per_shard * (uint64_t)n_shards > (1ull << 31))
        return fail(G2048_ERR_ARG, who);
    t = ShardedTable{};
    for (int j = 0; j < n_shards; ++j) {
        if (!shards[j] || ((uintptr_t)shards[j] & 31)) return fail(G2048_ERR_ARG, who);
        t.base[j] = (Slot*)shards[j];
    }
    t.mask = slots_per_shard * (uint64_t)n_shards - 1;
    t.low = slots_per_shard - 1;
    t.shift = 0;
    while ((1ull << t.shift) < slots_per_shard) ++t.shift;
    return 0;
}
}  // namespace

G2048_API int g2048_rollout_qlearn(uint64_t* boards, uint64_t* aux, int32_t* score, void* table, uint64_t capacity,
                                   int64_t n, int64_t k_steps, int flavour, float lr, float gamma, double eps,
                                   uint64_t seed, uint64_t step_base, uint64_t env_id_base, int64_t* counters,
                                   void* stream) {
    DEVSTATE();
    if (n < 0 || k_steps < 0 || (n && !boards) || !table || !pow2(capacity) || capacity > (1ull << 31) ||
        (flavour != 0 && flavour != 1))
        return fail(G2048_ERR_ARG, "g2048_rollout_qlearn: bad arguments");
    if (n == 0 || k_steps == 0) return 0;
    return launch_rollout_qlearn(D, LocalTable{(Slot*)table, capacity - 1}, boards, aux, score, n, k_steps, flavour, lr,
                                 gamma, eps, seed, step_base, env_id_base, counters, stream);
}

G2048_API int g2048_rollout_qlearn_sharded(uint64_t* boards, uint64_t* aux, int32_t* score, const void* const* shards,
                                           int n_shards, uint64_t slots_per_shard, int64_t n, int64_t k_steps,
                                           int flavour, float lr, float gamma, double eps, uint64_t seed,
                                           uint64_t step_base, uint64_t env_id_base, int64_t* counters, void* stream) {
    DEVSTATE();
    if (n < 0 || k_steps < 0 || (n && !boards) || (flavour != 0 && flavour != 1))
        return fail(G2048_ERR_ARG, "g2048_rollout_qlearn_sharded: bad arguments");
    ShardedTable t;
    int rc = make_sharded(shards, n_shards, slots_per_shard, t, "g2048_rollout_qlearn_sharded: bad shard list");
    if (rc) return rc;
    if (n == 0 || k_steps == 0) return 0;
    return launch_rollout_qlearn(D, t, boards, aux, score, n, k_steps, flavour, lr, gamma, eps, seed, step_base,
                                 env_id_base, counters, stream);
}

G2048_API int g2048_qtable_lookup_sharded(const void* const* shards, int n_shards, uint64_t slots_per_shard,
                                          const uint64_t* keys, int64_t n, float* rows, uint8_t* found, int insert,
                                          void* stream) {
    DEVSTATE();
    if (n < 0 || (n && (!keys || !rows))) return fail(G2048_ERR_ARG, "g2048_qtable_lookup_sharded: bad arguments");
    ShardedTable t;
    int rc = make_sharded(shards, n_shards, slots_per_shard, t, "g2048_qtable_lookup_sharded: bad shard list");
    if (rc) return rc;
    if (n == 0) return 0;
    int g = grid_for(n, 256, D->sm_count);
    if (insert) k_q_lookup_tab<true, ShardedTable><<<g, 256, 0, S(stream)>>>(t, (const u64*)keys, n, (float4*)rows, found);
    else k_q_lookup_tab<false, ShardedTable><<<g, 256, 0, S(stream)>>>(t, (const u64*)keys, n, (float4*)rows, found);
    LAUNCH_CHECK("k_q_lookup_tab");
    return 0;
}

// ---- scratch layout for the synchronous update: key_in | key_out | val_in | val_out | cub temp
namespace {
inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
struct Scratch {
    u64 *key_in, *key_out;
    float *val_in, *val_out;
    void* cub_temp;
    size_t cub_bytes;
};
size_t cub_temp_bytes(int64_t n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const u64*)nullptr, (u64*)nullptr, (const float*)nullptr,
                                    (float*)nullptr, (int64_t)n, 0, 64, (cudaStream_t)0);
    return bytes;
}
size_t scratch_bytes(int64_t n) {
    size_t m = (size_t)(n > 0 ? n : 1);
    return 2 * align256(m * 8) + 2 * align256(m * 4) + align256(cub_temp_bytes(n)) + 256;
}
int carve(void* scratch, size_t bytes, int64_t n, Scratch& s) {
    if (!scratch || bytes < scratch_bytes(n)) return fail(G2048_ERR_NOMEM, "scratch buffer missing or smaller than g2048_qlearn_scratch_bytes(n)");
    size_t m = (size_t)(n > 0 ? n : 1);
    char* p = (char*)(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
    s.key_in = (u64*)p; p += align256(m * 8);
    s.key_out = (u64*)p; p += align256(m * 8);
    s.val_in = (float*)p; p += align256(m * 4);
    s.val_out = (float*)p; p += align256(m * 4);
    s.cub_temp = p;
    s.cub_bytes = cub_temp_bytes(n);
    return 0;
}
// apply the records in s.key_in / s.val_in
int apply_records(DeviceState* D, Slot* tab, uint64_t capacity, Scratch& s, int64_t n, float lr, int mode, cudaStream_t st,
                  int kshift, const int* abort_flag) {
    int g = grid_for(n, 256, D->sm_count);
    if (mode == G2048_MODE_ATOMIC) {
        k_apply_atomic<<<g, 256, 0, st>>>(tab, s.key_in, s.val_in, lr, n);
        LAUNCH_CHECK("k_apply_atomic");
        return 0;
    }
    // Valid keys are slot * 4 + action < 4 * capacity; the all-ones "no slot" key is also all ones in the low
    // log2(capacity) + 3 bits, so sorting just those bits still puts it last (fewer radix passes than 64 bits).
    // The sort is stable: equal keys keep ascending record (= env) order.
    int end_bit = 3;
    while ((1ull << (end_bit - 3)) < capacity) ++end_bit;
    end_bit += kshift;
    if (end_bit > 64) end_bit = 64;
    CK(cub::DeviceRadixSort::SortPairs(s.cub_temp, s.cub_bytes, (const u64*)s.key_in, s.key_out, (const float*)s.val_in,
                                       s.val_out, (int64_t)n, 0, end_bit, st));
    u64* worklist = s.key_in;   // the sort's input is dead now: reuse it for the queue of long runs (< n / 8 entries)
    CK(cudaMemsetAsync(worklist, 0, sizeof(u64), st));
    k_segment_apply<<<g, 256, 0, st>>>(tab, s.key_out, s.val_out, lr, n, worklist, kshift, abort_flag);
    LAUNCH_CHECK("k_segment_apply");
    if (n > kInlineRun) {
        k_long_run_apply<<<D->sm_count * 4, 256, 0, st>>>(tab, s.key_out, s.val_out, lr, n, worklist, kshift, abort_flag);
        LAUNCH_CHECK("k_long_run_apply");
    }
    return 0;
}
}  // namespace


namespace {
// buckets for grouping about `records` records: four per bucket on average (2^10 .. 2^22 buckets)
inline int bucket_bits_for(int64_t records) {
    int bits = 10;
    while (bits < 22 && (1ll << (bits + 2)) < records) ++bits;
    return bits;
}
inline size_t bucket_bytes(int bits) { return 2 * align256(((size_t)1 << bits) * 4) + align256((((size_t)1 << (bits - 10)) + 2) * 4); }
// carve boff | work | sums out of `p` and clear what must start at zero
int make_buckets(char* p, int bits, cudaStream_t st, Buckets& B) {
    B.bits = bits;
    B.boff = (u32*)p; p += align256(((size_t)1 << bits) * 4);
    B.work = (u32*)p; p += align256(((size_t)1 << bits) * 4);
    B.sums = (u32*)p;
    CK(cudaMemsetAsync(B.boff, 0, ((size_t)1 << bits) * sizeof(u32), st));
    CK(cudaMemsetAsync(B.sums, 0, (((size_t)1 << (bits - 10)) + 2) * sizeof(u32), st));
    return 0;
}
// a zeroed list of `cap` deferred updates and the buffers that group them, from the device's ring
struct DeferBuffers {
    unsigned long long* count;
    ulonglong2 *rec, *rec_out;
    u32* pos;
    Buckets buckets;
    int64_t cap;
};
int deferred_list(DeviceState* D, int64_t cap, cudaStream_t st, DeferBuffers& B) {
    DeviceState::DeferSet* set;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        set = &D->defer[D->defer_next++ % 4];
    }
    const size_t m = (size_t)cap;
    const int bits = bucket_bits_for(cap / 2);      // the list is about half used
    const size_t need = 256 + 2 * align256(m * 16) + align256(m * 4) + bucket_bytes(bits);
    if (set->bytes < need) {
        CK(cudaDeviceSynchronize());               // growing: nobody may still be using the old buffer
        if (set->buf) CK(cudaFree(set->buf));
        set->buf = nullptr;
        set->bytes = 0;
        CK(cudaMalloc(&set->buf, need));
        set->bytes = need;
    }
    char* p = (char*)set->buf;
    B.count = (unsigned long long*)p; p += 256;
    B.rec = (ulonglong2*)p; p += align256(m * 16);
    B.rec_out = (ulonglong2*)p; p += align256(m * 16);
    B.pos = (u32*)p; p += align256(m * 4);
    B.cap = cap;
    CK(cudaMemsetAsync(B.count, 0, sizeof(unsigned long long), st));
    return make_buckets(p, bits, st, B.buckets);
}
// room for one lost race in four env steps (measured: one in 18 on a local table, one in 9 on a table shared by 8 GPUs,
// plus a quarter of slack in the warps' reservations; only the first launches after a common reset, where every env sits
// on one of 480 boards, overflow into the in-place loop)
int64_t deferred_capacity(int64_t n, int64_t k_steps) {
    const long double want = (long double)n * (long double)k_steps / 4;
    int64_t cap = 1 << 16;
    while (cap < (1 << 23) && (long double)cap < want) cap <<= 1;
    return cap;
}
template <class TAB>
int apply_deferred(DeviceState* D, const TAB& tab, const DeferBuffers& B, float lr, cudaStream_t st) {
    const int g = grid_for(B.cap, 256, D->sm_count);
    k_defer_count<<<g, 256, 0, st>>>(B.rec, B.count, (unsigned long long)B.cap, B.pos, B.buckets);
    k_bucket_scan<<<B.buckets.blocks(), 1024, 0, st>>>(B.buckets);
    k_bucket_scan_sums<<<1, 1024, 0, st>>>(B.buckets);
    k_defer_scatter<<<g, 256, 0, st>>>(B.rec, B.count, (unsigned long long)B.cap, B.pos, B.buckets, B.rec_out);
    const int long_blocks = D->sm_count * 4;
    k_defer_apply<TAB><<<long_blocks + B.buckets.count() / 256, 256, 0, st>>>(tab, B.rec_out, B.buckets, lr, long_blocks);
    LAUNCH_CHECK("apply_deferred");
    return 0;
}
template <class TAB>
int launch_rollout_qlearn(DeviceState* D, const TAB& tab, uint64_t* boards, uint64_t* aux, int32_t* score, int64_t n,
                          int64_t k_steps, int flavour, float lr, float gamma, double eps, uint64_t seed,
                          uint64_t step_base, uint64_t env_id_base, int64_t* counters, void* stream,
                          const DeferBuffers* shared_list) {
    int grid, block, smem_lut;
    size_t smem;
    rollout_geometry(D, n, grid, block, smem_lut, smem, QlearnThreads<TAB>::value);
    // the warps take their envs from a queue: its head is a counter zeroed in stream order before the launch
    unsigned long long* queue;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        queue = D->queue + (D->queue_next++ % kQueueSlots);
    }
    CK(cudaMemsetAsync(queue, 0, sizeof(unsigned long long), S(stream)));
    // big launches: lost races go to a list that is applied after the rollout (see the kernel)
    // (shared_list: the caller runs several launches into one list and applies it itself)
    Deferred defer{};
    DeferBuffers db{};
    int64_t cap = 0;
    if (shared_list) {
        defer = Deferred{shared_list->rec, shared_list->count, (unsigned long long)shared_list->cap};
    } else if (n >= kDeferMinEnvs) {
        cap = deferred_capacity(n, k_steps);
        int rc = deferred_list(D, cap, S(stream), db);
        if (rc) return rc;
        defer = Deferred{db.rec, db.count, (unsigned long long)cap};
    }
#define LAUNCH_RQ(F, SM)                                                                                              \
    k_rollout_qlearn<F, SM, TAB><<<grid, block, smem, S(stream)>>>(D->tables, (u64*)boards, (u64*)aux, score, tab, n,   \
                                                                   k_steps, lr, gamma, eps_threshold(eps), seed,       \
                                                                   step_base, env_id_base, (long long*)counters, queue, defer)
    if (flavour == 0) { if (smem_lut) LAUNCH_RQ(0, true); else LAUNCH_RQ(0, false); }
    else { if (smem_lut) LAUNCH_RQ(1, true); else LAUNCH_RQ(1, false); }
#undef LAUNCH_RQ
    LAUNCH_CHECK("k_rollout_qlearn");
    if (cap) return apply_deferred(D, tab, db, lr, S(stream));
    return 0;
}
}  // namespace

G2048_API size_t g2048_qlearn_scratch_bytes(int64_t n) { return scratch_bytes(n); }

G2048_API int g2048_qlearn_step(uint64_t* boards, uint64_t* aux, int32_t* score, void* table, uint64_t capacity,
                                int64_t n, int flavour, float lr, float gamma, double eps, int mode, int apply,
                                uint64_t seed, uint64_t step_idx, uint64_t env_id_base, int64_t* counters,
                                uint64_t* rec_key, uint8_t* rec_action, float* rec_target, void* scratch,
                                size_t scratch_bytes_, void* stream) {
    DEVSTATE();
    if (n < 0 || (n && !boards) || !table || !pow2(capacity) || capacity > (1ull << 31) ||
        (flavour != 0 && flavour != 1) || (mode != 0 && mode != 1))
        return fail(G2048_ERR_ARG, "g2048_qlearn_step: bad arguments");
    if (n == 0) return 0;
    Scratch s{};
    if (apply) {
        int rc = carve(scratch, scratch_bytes_, n, s);
        if (rc) return rc;
    }
    int g = grid_for(n, 256, D->sm_count);
#define PHASE_A(F)                                                                                                    \
    k_qlearn_phase_a<F><<<g, 256, 0, S(stream)>>>(D->tables, (u64*)boards, (u64*)aux, score, (Slot*)table, capacity - 1, \
                                                  n, lr, gamma, eps_threshold(eps), seed, step_idx, env_id_base,      \
                                                  (long long*)counters, s.key_in, s.val_in, (u64*)rec_key, rec_action, \
                                                  rec_target, nullptr)
    if (flavour == 0) PHASE_A(0); else PHASE_A(1);
#undef PHASE_A
    LAUNCH_CHECK("k_qlearn_phase_a");
    if (apply) return apply_records(D, (Slot*)table, capacity, s, n, lr, mode, S(stream));
    return 0;
}

// ---- synchronous step, exchange form: emit packed records / apply record lists (local or peer memory)
G2048_API int g2048_qlearn_emit(uint64_t* boards, uint64_t* aux, int32_t* score, void* table, uint64_t capacity,
                                int64_t n, int flavour, float gamma, double eps, uint64_t seed, uint64_t step_idx,
                                uint64_t env_id_base, int64_t* counters, g2048_record* records, void* stream) {
    DEVSTATE();
    if (n < 0 || (n && (!boards || !records)) || !table || !pow2(capacity) || capacity > (1ull << 31) ||
        (flavour != 0 && flavour != 1) || ((uintptr_t)records & 15))
        return fail(G2048_ERR_ARG, "g2048_qlearn_emit: bad arguments");
    if (n == 0) return 0;
    int g = grid_for(n, 256, D->sm_count);
#define EMIT(F)                                                                                                        \
    k_qlearn_phase_a<F><<<g, 256, 0, S(stream)>>>(D->tables, (u64*)boards, (u64*)aux, score, (Slot*)table, capacity - 1, \
                                                  n, 0.f, gamma, eps_threshold(eps), seed, step_idx, env_id_base,      \
                                                  (long long*)counters, nullptr, nullptr, nullptr, nullptr, nullptr,   \
                                                  (ulonglong2*)records)
    if (flavour == 0) EMIT(0); else EMIT(1);
#undef EMIT
    LAUNCH_CHECK("k_qlearn_phase_a");
    return 0;
}

G2048_API int g2048_qtable_apply_records(void* table, uint64_t capacity, const g2048_record* const* lists,
                                         const int64_t* counts, int n_lists, float lr, int mode, void* scratch,
                                         size_t scratch_bytes_, void* stream) {
    DEVSTATE();
    if (!table || !pow2(capacity) || capacity > (1ull << 31) || !lists || !counts || n_lists < 1 ||
        n_lists > G2048_MAX_PEERS || (mode != 0 && mode != 1))
        return fail(G2048_ERR_ARG, "g2048_qtable_apply_records: bad arguments");
    RecordLists R{};
    long long n = 0;
    for (int j = 0; j < n_lists; ++j) {
        if (counts[j] < 0 || (counts[j] && !lists[j]) || ((uintptr_t)lists[j] & 15))
            return fail(G2048_ERR_ARG, "g2048_qtable_apply_records: bad record list");
        R.ptr[j] = (const ulonglong2*)lists[j];
        n += counts[j];
        R.end[j] = n;
    }
    R.n_lists = n_lists;
    if (n == 0) return 0;
    Scratch sc{};
    int rc = carve(scratch, scratch_bytes_, n, sc);
    if (rc) return rc;
    k_peer_records_to_sortkeys<<<grid_for(n, 256, D->sm_count), 256, 0, S(stream)>>>((Slot*)table, capacity - 1, R, n,
                                                                                      sc.key_in, sc.val_in, D->abort_flag);
    LAUNCH_CHECK("k_peer_records_to_sortkeys");
    return apply_records(D, (Slot*)table, capacity, sc, n, lr, mode, S(stream));
}

G2048_API int g2048_qlearn_emit_owned(uint64_t* boards, uint64_t* aux, int32_t* score, const void* const* shards,
                                      int n_shards, uint64_t slots_per_shard, int64_t n, int flavour, float gamma,
                                      double eps, uint64_t seed, uint64_t step_idx, uint64_t env_id_base,
                                      uint64_t record_index_base, int idx_bits, int64_t* counters,
                                      g2048_record* const* owner_lists, uint64_t* owner_counts, uint32_t* carry_slot,
                                      float* carry_row, int use_carry, void* stream) {
    DEVSTATE();
    if (n < 0 || (n && !boards) || (flavour != 0 && flavour != 1) || !owner_lists || !owner_counts || idx_bits < 1 ||
        idx_bits > 40 || (!carry_slot != !carry_row) || (use_carry && !carry_slot) || ((uintptr_t)carry_row & 15))
        return fail(G2048_ERR_ARG, "g2048_qlearn_emit_owned: bad arguments");
    ShardedTable t;
    int rc = make_sharded(shards, n_shards, slots_per_shard, t, "g2048_qlearn_emit_owned: bad shard list");
    if (rc) return rc;
    if ((int)t.shift + 2 + idx_bits > 64 || ((record_index_base + (uint64_t)n - 1) >> idx_bits) != 0)
        return fail(G2048_ERR_ARG, "g2048_qlearn_emit_owned: record index does not fit idx_bits");
    OwnedLists out{};
    for (int j = 0; j < n_shards; ++j) {
        if (!owner_lists[j] || ((uintptr_t)owner_lists[j] & 15)) return fail(G2048_ERR_ARG, "g2048_qlearn_emit_owned: bad owner list");
        out.list[j] = (ulonglong2*)owner_lists[j];
    }
    out.count = (unsigned long long*)owner_counts;
    out.idx_bits = idx_bits;
    if (n == 0) return 0;
    int g = grid_for(n, 256, D->sm_count);
#define EMIT(F)                                                                                                       \
    k_qlearn_emit_owned<F><<<g, 256, 0, S(stream)>>>(D->tables, (u64*)boards, (u64*)aux, score, t, n, gamma,           \
                                                     eps_threshold(eps), seed, step_idx, env_id_base, record_index_base, \
                                                     (long long*)counters, out, carry_slot, (float4*)carry_row, use_carry)
    if (flavour == 0) EMIT(0); else EMIT(1);
#undef EMIT
    LAUNCH_CHECK("k_qlearn_emit_owned");
    return 0;
}

G2048_API int g2048_qtable_apply_owned(void* shard, uint64_t slots_per_shard, const g2048_record* const* lists,
                                       const int64_t* counts, int n_lists, int idx_bits, float lr, void* scratch,
                                       size_t scratch_bytes_, void* stream) {
    DEVSTATE();
    if (!shard || !pow2(slots_per_shard) || !lists || !counts || n_lists < 1 || n_lists > G2048_MAX_PEERS || idx_bits < 1 ||
        idx_bits > 40)
        return fail(G2048_ERR_ARG, "g2048_qtable_apply_owned: bad arguments");
    RecordLists R{};
    long long n = 0;
    for (int j = 0; j < n_lists; ++j) {
        if (counts[j] < 0 || (counts[j] && !lists[j]) || ((uintptr_t)lists[j] & 15))
            return fail(G2048_ERR_ARG, "g2048_qtable_apply_owned: bad record list");
        R.ptr[j] = (const ulonglong2*)lists[j];
        n += counts[j];
        R.end[j] = n;
    }
    R.n_lists = n_lists;
    if (n == 0) return 0;
    Scratch sc{};
    int rc = carve(scratch, scratch_bytes_, n, sc);
    if (rc) return rc;
    k_gather_owned<<<grid_for(n, 256, D->sm_count), 256, 0, S(stream)>>>(R, n, sc.key_in, sc.val_in, D->abort_flag);
    LAUNCH_CHECK("k_gather_owned");
    return apply_records(D, (Slot*)shard, slots_per_shard, sc, n, lr, G2048_MODE_DETERMINISTIC, S(stream), idx_bits);
}

G2048_API int g2048_peer_read_u64(const uint64_t* const* src, int n, uint64_t* host_out, void* stream) {
    if (!src || !host_out || n < 0) return fail(G2048_ERR_ARG, "g2048_peer_read_u64: bad arguments");
    for (int j = 0; j < n; ++j) {
        if (!src[j]) return fail(G2048_ERR_ARG, "g2048_peer_read_u64: null pointer");
        CK(cudaMemcpyAsync(host_out + j, src[j], sizeof(uint64_t), cudaMemcpyDeviceToHost, S(stream)));
    }
    CK(cudaStreamSynchronize(S(stream)));
    return 0;
}
G2048_API int g2048_peer_memset(void* dev_ptr, int value, size_t bytes, void* stream) {
    if (!dev_ptr) return fail(G2048_ERR_ARG, "g2048_peer_memset: bad arguments");
    CK(cudaMemsetAsync(dev_ptr, value, bytes, S(stream)));
    return 0;
}

// ---- NVLink peer memory between the per-GPU processes of one box (CUDA IPC)
G2048_API int g2048_peer_alloc(size_t bytes, void** dev_ptr, void* ipc_handle_out) {
    if (!dev_ptr || !ipc_handle_out || bytes == 0) return fail(G2048_ERR_ARG, "g2048_peer_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == G2048_IPC_HANDLE_BYTES, "IPC handle size");
    void* p = nullptr;
    CK(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail((int)e, "g2048_peer_alloc: cudaIpcGetMemHandle");
    }
    memcpy(ipc_handle_out, &h, sizeof h);
    *dev_ptr = p;
    return 0;
}
G2048_API int g2048_peer_open(const void* ipc_handle, void** dev_ptr) {
    if (!ipc_handle || !dev_ptr) return fail(G2048_ERR_ARG, "g2048_peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof h);
    CK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
G2048_API int g2048_peer_close(void* dev_ptr) {
    if (dev_ptr) CK(cudaIpcCloseMemHandle(dev_ptr));
    return 0;
}
G2048_API int g2048_peer_free(void* dev_ptr) {
    if (dev_ptr) CK(cudaFree(dev_ptr));
    return 0;
}
G2048_API int g2048_peer_barrier(uint64_t* const* flags, int rank, int world, uint64_t epoch, uint64_t timeout_ns,
                                 int* timed_out, void* stream) {
    if (!flags || world < 1 || world > G2048_MAX_PEERS || rank < 0 || rank >= world)
        return fail(G2048_ERR_ARG, "g2048_peer_barrier: bad arguments");
    PeerFlags F{};
    for (int j = 0; j < world; ++j) {
        if (!flags[j]) return fail(G2048_ERR_ARG, "g2048_peer_barrier: null flag pointer");
        F.ptr[j] = (u64*)flags[j];
    }
    int dev = 0;
    CK(cudaGetDevice(&dev));
    int* abort_flag = (dev >= 0 && dev < kMaxDevices && g_dev[dev].ready) ? g_dev[dev].abort_flag : nullptr;
    k_peer_barrier<<<1, 32, 0, S(stream)>>>(F, rank, world, epoch, timeout_ns ? timeout_ns : 5000000000ull, timed_out, abort_flag);
    LAUNCH_CHECK("k_peer_barrier");
    return 0;
}

// ---- exact synchronous step on a sharded table, routed (see k_routed_request)
// Layout of the buffer every rank shares with its peers (CUDA IPC), C = cap (the largest env count of any rank):
//   [0, 128) barrier flags | [256, 384) request counts per owner | [512, 640) record counts per owner | from 1024:
//   req_out[world][2C] u64 | reply1[world][2C] 8 B | reply2[world][2C] 16 B | sort input: records[world * C] u64
namespace {
constexpr size_t kRoutedHead = 1024, kRoutedReqCount = 256, kRoutedRecCount = 512;
struct RoutedLayout {
    size_t req, reply1, reply2, key_in, req_stride, r1_stride, r2_stride, total;
};
RoutedLayout routed_layout(int world, int64_t cap) {
    RoutedLayout l{};
    const size_t c = (size_t)(cap > 0 ? cap : 1);
    l.req_stride = align256(2 * c * 8);
    l.r1_stride = align256(2 * c * 8);
    l.r2_stride = align256(2 * c * 16);
    l.req = kRoutedHead;
    l.reply1 = l.req + (size_t)world * l.req_stride;
    l.reply2 = l.reply1 + (size_t)world * l.r1_stride;
    l.key_in = l.reply2 + (size_t)world * l.r2_stride;
    l.total = l.key_in + align256((size_t)world * c * 8);
    return l;
}
}  // namespace
struct g2048_routed {
    int device = 0, rank = 0, world = 1;
    int64_t cap = 0, n_total = 0;
    Slot* shard = nullptr;
    uint64_t slots = 0;
    RoutedLocal L{};
    RoutedServe V{};
    PeerFlags F{};
    PeerWords rec_counts{};          // rank r's record counts (all owners)
    u64* key_in = nullptr;           // this GPU's sort input, in the shared buffer: the peers push their records into it
    u32* off = nullptr;              // local [world]
    char* local = nullptr;           // per-env state of the requester side + saved slots of the owner side
    void* sort_buf = nullptr;        // sort output + CUB's temporary storage
    size_t sort_bytes = 0;
    u64* host_total = nullptr;       // pinned: [0] records for this GPU in the current step, [1] = the barrier's time-out flag (int)
    int* timed_out = nullptr;
    u64 epoch = 0;
    bool primed = false;
    // G2048_ROUTED_PROFILE=1: device time of every phase of the step (CUDA events), printed by g2048_routed_destroy
    bool profile = false;
    static constexpr int kPhases = 10;
    cudaEvent_t ev[kPhases + 1] = {};
    double phase_ms[kPhases] = {};
    long long profiled_steps = 0;
};
namespace {
int routed_barrier(g2048_routed* r, DeviceState* D, cudaStream_t st) {
    r->epoch += 1;
    k_peer_barrier<<<1, 32, 0, st>>>(r->F, r->rank, r->world, r->epoch, 5000000000ull, r->timed_out, D->abort_flag);
    LAUNCH_CHECK("k_peer_barrier");
    return 0;
}
int routed_check(g2048_routed* r) {
    if (*(volatile int*)r->timed_out) return fail(G2048_ERR_PEER, "g2048_routed: a peer did not reach the barrier (the ranks are no longer in step)");
    return 0;
}
}  // namespace

G2048_API size_t g2048_routed_buffer_bytes(int world, int64_t cap) {
    if (world < 1 || world > G2048_MAX_PEERS || cap < 1) return 0;
    return routed_layout(world, cap).total;
}
G2048_API g2048_routed* g2048_routed_create(int rank, int world, int64_t cap, int64_t n_total, void* const* peer_buffers,
                                            void* shard, uint64_t slots_per_shard) {
    DeviceState* D = nullptr;
    if (current_device_state(&D)) return nullptr;
    if (world < 1 || world > G2048_MAX_PEERS || (world & (world - 1)) || rank < 0 || rank >= world || cap < 1 ||
        2 * cap >= (1ll << kHandleBits) || (int64_t)world * cap >= (1ll << 32) || n_total < 1 || !peer_buffers || !shard ||
        !pow2(slots_per_shard) || slots_per_shard > (1ull << 30)) {   // slot * 4 + action must fit 32 bits of a record
        fail(G2048_ERR_ARG, "g2048_routed_create: bad arguments");
        return nullptr;
    }
    for (int j = 0; j < world; ++j)
        if (!peer_buffers[j] || ((uintptr_t)peer_buffers[j] & 255)) { fail(G2048_ERR_ARG, "g2048_routed_create: bad peer buffer"); return nullptr; }
    g2048_routed* r = new g2048_routed();
    cudaGetDevice(&r->device);
    r->rank = rank; r->world = world; r->cap = cap; r->n_total = n_total;
    r->shard = (Slot*)shard; r->slots = slots_per_shard;
    int slot_bits = 0;
    while ((1ull << slot_bits) < slots_per_shard) ++slot_bits;
    const RoutedLayout lay = routed_layout(world, cap);
    char* mine = (char*)peer_buffers[rank];
    const size_t c = (size_t)cap;
    const size_t n_warps = (c + 31) / 32;
    const size_t local_bytes = 5 * align256(c * 4) + 2 * align256(n_warps * world * 4) + (size_t)world * align256(2 * c * 4) + 512;
    if (cudaMalloc(&r->local, local_bytes) != cudaSuccess || cudaMemset(r->local, 0, local_bytes) != cudaSuccess ||
        cudaHostAlloc(&r->host_total, 2 * sizeof(u64), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
        fail((int)cudaGetLastError(), "g2048_routed_create: allocation");
        if (r->local) cudaFree(r->local);
        delete r;
        return nullptr;
    }
    r->host_total[0] = r->host_total[1] = 0;
    r->timed_out = (int*)(r->host_total + 1);
    char* p = r->local;
    r->L.sk = (u32*)p; p += align256(c * 4);
    r->L.req1 = (u32*)p; p += align256(c * 4);
    r->L.cur = (u32*)p; p += align256(c * 4);
    r->L.reward = (float*)p; p += align256(c * 4);
    r->L.meta = (u32*)p; p += align256(c * 4);
    r->L.chunk = (u32*)p; p += align256(n_warps * world * 4);
    r->L.place = (u32*)p; p += align256(n_warps * world * 4);
    r->L.n_warps = (long long)n_warps;
    for (int j = 0; j < world; ++j) { r->V.saved_slot[j] = (u32*)p; p += align256(2 * c * 4); }
    r->V.count_cache = (unsigned long long*)p; p += 256;
    r->off = (u32*)p;
    r->L.off = r->off;
    r->L.world = world; r->L.owner_shift = (u32)slot_bits;
    r->L.req_count = (unsigned long long*)(mine + kRoutedReqCount);
    r->L.rec_count = (unsigned long long*)(mine + kRoutedRecCount);
    r->key_in = (u64*)(mine + lay.key_in);
    r->V.world = world;
    for (int j = 0; j < world; ++j) {
        char* peer = (char*)peer_buffers[j];
        r->L.req_out[j] = (u64*)(mine + lay.req + (size_t)j * lay.req_stride);
        r->L.reply1[j] = (const uint2*)(mine + lay.reply1 + (size_t)j * lay.r1_stride);
        r->L.reply2[j] = (const float4*)(mine + lay.reply2 + (size_t)j * lay.r2_stride);
        r->L.push_rec[j] = (u64*)(peer + lay.key_in);
        r->V.req[j] = (const u64*)(peer + lay.req + (size_t)rank * lay.req_stride);
        r->V.req_count[j] = (const unsigned long long*)(peer + kRoutedReqCount) + rank;
        r->V.reply1[j] = (uint2*)(peer + lay.reply1 + (size_t)rank * lay.r1_stride);
        r->V.reply2[j] = (float4*)(peer + lay.reply2 + (size_t)rank * lay.r2_stride);
        r->F.ptr[j] = (u64*)peer;
        r->rec_counts.ptr[j] = (const u64*)(peer + kRoutedRecCount);
    }
    const char* prof = getenv("G2048_ROUTED_PROFILE");
    r->profile = prof && prof[0] == '1';
    if (r->profile)
        for (auto& e : r->ev) cudaEventCreate(&e);
    return r;
}
G2048_API void g2048_routed_destroy(g2048_routed* r) {
    if (!r) return;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(r->device);
    if (r->profile && r->profiled_steps) {
        static const char* names[g2048_routed::kPhases] = {"request + scan", "barrier 1 + offsets", "host (sync, launch)", "lookup", "barrier 2",
                                                           "records", "barrier 3", "sort + apply", "rows", "barrier 4"};
        fprintf(stderr, "g2048_routed rank %d: %lld steps, ms per step:", r->rank, r->profiled_steps);
        for (int i = 0; i < g2048_routed::kPhases; ++i) fprintf(stderr, " %s %.3f |", names[i], r->phase_ms[i] / (double)r->profiled_steps);
        fprintf(stderr, "\n");
        for (auto& e : r->ev) cudaEventDestroy(e);
    }
    if (r->local) cudaFree(r->local);
    if (r->sort_buf) cudaFree(r->sort_buf);
    if (r->host_total) cudaFreeHost(r->host_total);
    cudaSetDevice(prev);
    delete r;
}
G2048_API int g2048_routed_prime(g2048_routed* r, const uint64_t* boards, int64_t n, void* stream) {
    DEVSTATE();
    if (!r || n < 0 || n > r->cap || (n && !boards)) return fail(G2048_ERR_ARG, "g2048_routed_prime: bad arguments");
    cudaStream_t st = S(stream);
    int rc;
    CK(cudaMemsetAsync(r->L.req_count, 0, 128, st));
    CK(cudaMemsetAsync(r->L.rec_count, 0, 128, st));
    if ((rc = routed_barrier(r, D, st))) return rc;            // nobody still reads the counts of an earlier use
    if (n) {
        k_routed_request<0, true, false><<<grid_for(n, 256, D->sm_count), 256, 0, st>>>(D->tables, (u64*)boards, nullptr, nullptr, r->L, n, 0, 0,
                                                                                  0, 0, nullptr);
        LAUNCH_CHECK("k_routed_request");
    }
    if ((rc = routed_barrier(r, D, st))) return rc;
    const int g = grid_for(2 * r->cap, 256, D->sm_count);
    k_routed_lookup<<<g, 256, 0, st>>>(r->shard, r->slots - 1, r->V, nullptr, D->abort_flag);
    LAUNCH_CHECK("k_routed_lookup");
    if ((rc = routed_barrier(r, D, st))) return rc;
    CK(cudaMemsetAsync(r->L.req_count, 0, 128, st));
    k_routed_rows<<<g, 256, 0, st>>>(r->shard, r->V, D->abort_flag);
    LAUNCH_CHECK("k_routed_rows");
    if ((rc = routed_barrier(r, D, st))) return rc;
    CK(cudaStreamSynchronize(st));
    r->primed = true;
    return routed_check(r);
}
G2048_API int g2048_routed_step(g2048_routed* r, uint64_t* boards, uint64_t* aux, int32_t* score, int64_t n, int flavour,
                                float lr, float gamma, double eps, uint64_t seed, uint64_t step_idx, uint64_t env_id_base,
                                int64_t* counters, int64_t* applied, void* stream) {
    DEVSTATE();
    if (!r || n < 0 || n > r->cap || (n && !boards) || (flavour != 0 && flavour != 1))
        return fail(G2048_ERR_ARG, "g2048_routed_step: bad arguments");
    if (!r->primed) return fail(G2048_ERR_ARG, "g2048_routed_step: call g2048_routed_prime first");
    if (env_id_base + (uint64_t)n > (uint64_t)r->n_total)
        return fail(G2048_ERR_ARG, "g2048_routed_step: env_id_base + n exceeds the total the exchange was created for");
    cudaStream_t st = S(stream);
    int rc;
    if ((rc = routed_check(r))) return rc;
    const int ge = grid_for(n, 256, D->sm_count), gs = grid_for(2 * r->cap, 256, D->sm_count);
#define MARK(i) do { if (r->profile) cudaEventRecord(r->ev[i], st); } while (0)
    MARK(0);
    if (n) {
        const bool big = n >= kEnvStepSmemLutMinEnvs;   // enough work to amortise the 213 KB staging copy per SM
#define REQ(F, SM) k_routed_request<F, false, SM><<<big ? D->sm_count : ge, big ? kRolloutThreads : 256, big ? kLutBytes : 0, st>>>( \
        D->tables, (u64*)boards, (u64*)aux, score, r->L, n, eps_threshold(eps), seed, step_idx, env_id_base, (long long*)counters)
        if (flavour == 0) { if (big) REQ(0, true); else REQ(0, false); }
        else { if (big) REQ(1, true); else REQ(1, false); }
#undef REQ
        LAUNCH_CHECK("k_routed_request");
    }
    // (the peers read the record counts of the step before ahead of its last barrier)
    const long long warps = (n + 31) / 32;
    const int tiles = (int)((warps + 1023) / 1024) > 0 ? (int)((warps + 1023) / 1024) : 1;
    k_routed_scan<<<dim3((unsigned)tiles, (unsigned)r->world), 1024, 0, st>>>(r->L.chunk, r->L.place, warps, r->L.n_warps, tiles, r->L.rec_count);
    LAUNCH_CHECK("k_routed_scan");
    MARK(1);
    if ((rc = routed_barrier(r, D, st))) return rc;            // every rank's requests and record counts are written
    k_routed_offsets<<<1, 256, 0, st>>>(r->rec_counts, r->world, r->rank, r->off, r->host_total);
    LAUNCH_CHECK("k_routed_offsets");
    MARK(2);
    // the one synchronisation of the step: the sort needs the record count on the host.  It comes this early so that
    // everything below is enqueued while the lookup kernel runs -- the device never waits for a launch.
    CK(cudaStreamSynchronize(st));
    if ((rc = routed_check(r))) return rc;
    const long long total = (long long)r->host_total[0];
    if (total > (long long)r->world * r->cap) return fail(G2048_ERR_ARG, "g2048_routed_step: more records than envs");
    // the records are sorted as 64-bit KEYS on the bits of (slot * 4 + action) only: the target rides in the low half, and
    // the sort is stable, so equal (slot, action) keep their ascending env order
    int slot_bits = 0;
    while ((1ull << slot_bits) < r->slots) ++slot_bits;
    u64 *rec_out = nullptr, *worklist = nullptr;
    void* temp = nullptr;
    size_t temp_bytes = 0;
    if (total > 0) {
        const size_t m = (size_t)total;
        CK(cub::DeviceRadixSort::SortKeys(nullptr, temp_bytes, (const u64*)nullptr, (u64*)nullptr, (int64_t)total, 32, 34 + slot_bits, st));
        const size_t need = align256(m * 8) + align256((m / kInlineRun + 2) * 8) + align256(temp_bytes) + 256;
        if (r->sort_bytes < need) {
            if (r->sort_buf) CK(cudaFree(r->sort_buf));
            r->sort_buf = nullptr;
            r->sort_bytes = 0;
            const size_t want = need + need / 4;
            CK(cudaMalloc(&r->sort_buf, want));
            r->sort_bytes = want;
        }
        char* p = (char*)r->sort_buf;
        rec_out = (u64*)p; p += align256(m * 8);
        worklist = (u64*)p; p += align256((m / kInlineRun + 2) * 8);
        temp = p;
    }
    MARK(3);
    k_routed_lookup<<<gs, 256, 0, st>>>(r->shard, r->slots - 1, r->V, (long long*)counters, D->abort_flag);
    LAUNCH_CHECK("k_routed_lookup");
    MARK(4);
    if ((rc = routed_barrier(r, D, st))) return rc;            // the answers are there; every owner has read my requests
    CK(cudaMemsetAsync(r->L.req_count, 0, 128, st));
    MARK(5);
    if (n) {
        k_routed_records<<<ge, 256, 0, st>>>(r->L, n, gamma);
        LAUNCH_CHECK("k_routed_records");
    }
    MARK(6);
    if ((rc = routed_barrier(r, D, st))) return rc;            // every rank's records are in their owner's sort input
    MARK(7);
    if (total > 0) {
        CK(cub::DeviceRadixSort::SortKeys(temp, temp_bytes, (const u64*)r->key_in, rec_out, (int64_t)total, 32, 34 + slot_bits, st));
        CK(cudaMemsetAsync(worklist, 0, sizeof(u64), st));
        k_segment_apply_packed<<<grid_for(total, 256, D->sm_count), 256, 0, st>>>(r->shard, rec_out, lr, total, worklist, D->abort_flag);
        k_long_run_apply_packed<<<D->sm_count * 4, 256, 0, st>>>(r->shard, rec_out, lr, total, worklist, D->abort_flag);
        LAUNCH_CHECK("k_segment_apply_packed");
    }
    MARK(8);
    k_routed_rows<<<gs, 256, 0, st>>>(r->shard, r->V, D->abort_flag);
    LAUNCH_CHECK("k_routed_rows");
    MARK(9);
    if ((rc = routed_barrier(r, D, st))) return rc;            // the rows for the next step are there
    MARK(10);
#undef MARK
    if (applied) *applied = total;
    if (r->profile && total > 0) {
        CK(cudaStreamSynchronize(st));
        for (int i = 0; i < g2048_routed::kPhases; ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, r->ev[i], r->ev[i + 1]);
            r->phase_ms[i] += ms;
        }
        r->profiled_steps += 1;
    }
    return 0;
}

G2048_API size_t g2048_qtable_bytes(uint64_t capacity) { return (size_t)capacity * sizeof(Slot); }
G2048_API int g2048_qtable_clear(void* table, uint64_t capacity, void* stream) {
    if (!table || !pow2(capacity)) return fail(G2048_ERR_ARG, "g2048_qtable_clear: bad arguments");
    CK(cudaMemsetAsync(table, 0, (size_t)capacity * sizeof(Slot), S(stream)));
    return 0;
}
G2048_API int g2048_qtable_lookup(void* table, uint64_t capacity, const uint64_t* keys, int64_t n, float* rows,
                                  uint8_t* found, int insert, void* stream) {
    DEVSTATE();
    if (n < 0 || !table || !pow2(capacity) || (n && (!keys || !rows))) return fail(G2048_ERR_ARG, "g2048_qtable_lookup: bad arguments");
    if (n == 0) return 0;
    int g = grid_for(n, 256, D->sm_count);
    if (insert) k_q_lookup<true><<<g, 256, 0, S(stream)>>>((Slot*)table, capacity - 1, (const u64*)keys, n, (float4*)rows, found);
    else k_q_lookup<false><<<g, 256, 0, S(stream)>>>((Slot*)table, capacity - 1, (const u64*)keys, n, (float4*)rows, found);
    LAUNCH_CHECK("k_q_lookup");
    return 0;
}
G2048_API int g2048_choose_action(void* table, uint64_t capacity, const uint64_t* boards, uint8_t* actions, int64_t n,
                                  double eps, uint64_t seed, uint64_t step_idx, uint64_t env_id_base, void* stream) {
    DEVSTATE();
    if (n < 0 || !table || !pow2(capacity) || (n && (!boards || !actions))) return fail(G2048_ERR_ARG, "g2048_choose_action: bad arguments");
    if (n == 0) return 0;
    k_choose_action<<<grid_for(n, 256, D->sm_count), 256, 0, S(stream)>>>((Slot*)table, capacity - 1, (const u64*)boards,
                                                                           actions, n, eps_threshold(eps), seed, step_idx,
                                                                           env_id_base);
    LAUNCH_CHECK("k_choose_action");
    return 0;
}
G2048_API int g2048_qtable_update(void* table, uint64_t capacity, const uint64_t* s, const uint8_t* a, const float* r,
                                  const uint64_t* s2, const uint8_t* done, int64_t n, float lr, float gamma, int mode,
                                  void* scratch, size_t scratch_bytes_, void* stream) {
    DEVSTATE();
    if (n < 0 || !table || !pow2(capacity) || capacity > (1ull << 31) || (n && (!s || !a || !r || !s2 || !done)) ||
        (mode != 0 && mode != 1))
        return fail(G2048_ERR_ARG, "g2048_qtable_update: bad arguments");
    if (n == 0) return 0;
    Scratch sc{};
    int rc = carve(scratch, scratch_bytes_, n, sc);
    if (rc) return rc;
    k_q_update_phase_a<<<grid_for(n, 256, D->sm_count), 256, 0, S(stream)>>>((Slot*)table, capacity - 1, (const u64*)s, a, r,
                                                                              (const u64*)s2, done, n, gamma, sc.key_in,
                                                                              sc.val_in);
    LAUNCH_CHECK("k_q_update_phase_a");
    return apply_records(D, (Slot*)table, capacity, sc, n, lr, mode, S(stream));
}
G2048_API int g2048_qtable_apply_targets(void* table, uint64_t capacity, const uint64_t* keys, const uint8_t* a,
                                         const float* target, int64_t n, float lr, int mode, void* scratch,
                                         size_t scratch_bytes_, void* stream) {
    DEVSTATE();
    if (n < 0 || !table || !pow2(capacity) || capacity > (1ull << 31) || (n && (!keys || !a || !target)) ||
        (mode != 0 && mode != 1))
        return fail(G2048_ERR_ARG, "g2048_qtable_apply_targets: bad arguments");
    if (n == 0) return 0;
    Scratch sc{};
    int rc = carve(scratch, scratch_bytes_, n, sc);
    if (rc) return rc;
    k_keys_to_records<<<grid_for(n, 256, D->sm_count), 256, 0, S(stream)>>>((Slot*)table, capacity - 1, (const u64*)keys, a,
                                                                             n, sc.key_in);
    LAUNCH_CHECK("k_keys_to_records");
    CK(cudaMemcpyAsync(sc.val_in, target, (size_t)n * sizeof(float), cudaMemcpyDefault, S(stream)));
    return apply_records(D, (Slot*)table, capacity, sc, n, lr, mode, S(stream));
}
G2048_API int g2048_qtable_size(const void* table, uint64_t capacity, int64_t* count, void* stream) {
    DEVSTATE();
    if (!table || !pow2(capacity) || !count) return fail(G2048_ERR_ARG, "g2048_qtable_size: bad arguments");
    CK(cudaMemsetAsync(count, 0, sizeof(int64_t), S(stream)));
    k_q_size<<<grid_for((int64_t)capacity, 256, D->sm_count, 16), 256, 0, S(stream)>>>((const Slot*)table, capacity,
                                                                                        (long long*)count);
    LAUNCH_CHECK("k_q_size");
    return 0;
}
G2048_API int g2048_qtable_probe_stats(const void* table, uint64_t capacity, int64_t* stats, void* stream) {
    DEVSTATE();
    if (!table || !pow2(capacity) || !stats) return fail(G2048_ERR_ARG, "g2048_qtable_probe_stats: bad arguments");
    CK(cudaMemsetAsync(stats, 0, 3 * sizeof(int64_t), S(stream)));
    k_q_probe_stats<<<grid_for((int64_t)capacity, 256, D->sm_count, 16), 256, 0, S(stream)>>>((const Slot*)table, capacity,
                                                                                               (long long*)stats);
    LAUNCH_CHECK("k_q_probe_stats");
    return 0;
}
G2048_API int g2048_qtable_export(const void* table, uint64_t capacity, uint64_t* keys, float* rows, int64_t max_out,
                                  int64_t* count, void* stream) {
    DEVSTATE();
    if (!table || !pow2(capacity) || !count || max_out < 0 || (max_out && (!keys || !rows)))
        return fail(G2048_ERR_ARG, "g2048_qtable_export: bad arguments");
    k_q_export<<<grid_for((int64_t)capacity, 256, D->sm_count, 16), 256, 0, S(stream)>>>((const Slot*)table, capacity,
                                                                                          (u64*)keys, (float4*)rows, max_out,
                                                                                          (long long*)count);
    LAUNCH_CHECK("k_q_export");
    return 0;
}

// ============================================================================================ host-buffer API
struct g2048_ctx {
    int device = 0;
    int64_t max_envs = 0;
    uint64_t capacity = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t pipe[3] = {nullptr, nullptr, nullptr};   // chunked rollouts: copies of one chunk overlap the kernel of another
    cudaEvent_t ev_start = nullptr, ev_done[3] = {nullptr, nullptr, nullptr}, ev_kernel[3] = {nullptr, nullptr, nullptr};
    void* table = nullptr;
    // device staging, sized for max_envs
    u64 *boards = nullptr, *aux = nullptr, *keys2 = nullptr;
    int *score = nullptr, *move_score = nullptr;
    uint8_t *bytes_a = nullptr, *bytes_b = nullptr, *flags = nullptr, *maxlvl = nullptr, *draws = nullptr;
    double* rew64 = nullptr;
    float *rows = nullptr, *rew32 = nullptr;
    long long* counters = nullptr;  // G2048_N_COUNTERS + 1 (export/size count)
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    // Contexts for a handful of envs (the N = 1 drop-in adapters) stage through ONE block of pinned host memory that
    // the kernels read and write directly (mapped under unified addressing): a call then costs its kernel and one
    // stream synchronisation instead of about ten small copies from pageable memory.
    void* arena = nullptr;
    size_t arena_bytes = 0;
};
constexpr int64_t kTinyCtxEnvs = 64;

G2048_API g2048_ctx* g2048_ctx_create(int device, int64_t max_envs, uint64_t table_capacity) {
    if (max_envs <= 0 || (table_capacity && !pow2(table_capacity))) { fail(G2048_ERR_ARG, "g2048_ctx_create: bad arguments"); return nullptr; }
    if (g2048_init(device)) return nullptr;
    g2048_ctx* c = new g2048_ctx();
    c->device = device; c->max_envs = max_envs; c->capacity = table_capacity;
    size_t m = (size_t)max_envs;
    c->scratch_bytes = scratch_bytes(max_envs);
    bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->ev_start, cudaEventDisableTiming) == cudaSuccess;
    for (int j = 0; ok && j < 3; ++j)
        ok = cudaStreamCreateWithFlags(&c->pipe[j], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->ev_done[j], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->ev_kernel[j], cudaEventDisableTiming) == cudaSuccess;
    if (ok && max_envs <= kTinyCtxEnvs) {
        const size_t slot = align256(m * 16);            // every staging buffer fits one slot (rows: 16 B per env)
        c->arena_bytes = 13 * slot;
        ok = cudaHostAlloc(&c->arena, c->arena_bytes, cudaHostAllocDefault) == cudaSuccess;
        if (ok) {
            char* p = (char*)c->arena;
            memset(p, 0, c->arena_bytes);
            auto take = [&]() { char* q = p; p += slot; return (void*)q; };
            c->boards = (u64*)take(); c->aux = (u64*)take(); c->keys2 = (u64*)take();
            c->score = (int*)take(); c->move_score = (int*)take();
            c->bytes_a = (uint8_t*)take(); c->bytes_b = (uint8_t*)take(); c->flags = (uint8_t*)take();
            c->maxlvl = (uint8_t*)take(); c->draws = (uint8_t*)take();
            c->rew64 = (double*)take(); c->rows = (float*)take(); c->rew32 = (float*)take();
        }
    } else {
        ok = ok &&
              cudaMalloc(&c->boards, m * 8) == cudaSuccess && cudaMalloc(&c->aux, m * 8) == cudaSuccess &&
              cudaMalloc(&c->keys2, m * 8) == cudaSuccess && cudaMalloc(&c->score, m * 4) == cudaSuccess &&
              cudaMalloc(&c->move_score, m * 4) == cudaSuccess && cudaMalloc(&c->bytes_a, m) == cudaSuccess &&
              cudaMalloc(&c->bytes_b, m) == cudaSuccess && cudaMalloc(&c->flags, m) == cudaSuccess &&
              cudaMalloc(&c->maxlvl, m) == cudaSuccess && cudaMalloc(&c->draws, m * 4) == cudaSuccess &&
              cudaMalloc(&c->rew64, m * 8) == cudaSuccess && cudaMalloc(&c->rows, m * 16) == cudaSuccess &&
              cudaMalloc(&c->rew32, m * 4) == cudaSuccess;
    }
    ok = ok && cudaMalloc(&c->counters, (G2048_N_COUNTERS + 1) * sizeof(long long)) == cudaSuccess &&
         cudaMalloc(&c->scratch, c->scratch_bytes) == cudaSuccess;
    if (ok && table_capacity) {
        ok = cudaMalloc(&c->table, g2048_qtable_bytes(table_capacity)) == cudaSuccess &&
             cudaMemsetAsync(c->table, 0, g2048_qtable_bytes(table_capacity), c->stream) == cudaSuccess &&
             cudaStreamSynchronize(c->stream) == cudaSuccess;
    }
    if (!ok) {
        fail((int)cudaGetLastError(), "g2048_ctx_create: allocation failed");
        g2048_ctx_destroy(c);
        return nullptr;
    }
    return c;
}
G2048_API void g2048_ctx_destroy(g2048_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    void* staging[] = {c->boards, c->aux, c->keys2, c->score, c->move_score, c->bytes_a, c->bytes_b, c->flags, c->maxlvl,
                       c->draws, c->rew64, c->rows, c->rew32};
    if (c->arena) cudaFreeHost(c->arena);
    else
        for (void* p : staging)
            if (p) cudaFree(p);
    void* ptrs[] = {c->counters, c->scratch, c->table};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (c->stream) cudaStreamDestroy(c->stream);
    for (int j = 0; j < 3; ++j) {
        if (c->pipe[j]) cudaStreamDestroy(c->pipe[j]);
        if (c->ev_done[j]) cudaEventDestroy(c->ev_done[j]);
        if (c->ev_kernel[j]) cudaEventDestroy(c->ev_kernel[j]);
    }
    if (c->ev_start) cudaEventDestroy(c->ev_start);
    delete c;
}
G2048_API void* g2048_ctx_table(g2048_ctx* c) { return c ? c->table : nullptr; }
G2048_API uint64_t g2048_ctx_table_capacity(g2048_ctx* c) { return c ? c->capacity : 0; }
G2048_API void* g2048_ctx_stream(g2048_ctx* c) { return c ? (void*)c->stream : nullptr; }

#define CTX_ENTER(need_table)                                                                    \
    if (!c) return fail(G2048_ERR_ARG, "null context");                                          \
    if (n < 0 || n > c->max_envs) return fail(G2048_ERR_ARG, "n exceeds the context's max_envs"); \
    if ((need_table) && !c->table) return fail(G2048_ERR_ARG, "context has no Q-table");         \
    CK(cudaSetDevice(c->device));                                                                \
    cudaStream_t st = c->stream;                                                                 \
    (void)st
static inline bool in_arena(const g2048_ctx* c, const void* p) {
    return c->arena && (const char*)p >= (const char*)c->arena && (const char*)p < (const char*)c->arena + c->arena_bytes;
}
// host -> staging: a plain memcpy into the pinned arena (nothing is in flight: every call ends with a synchronise)
static inline int ctx_h2d(g2048_ctx* c, void* dst, const void* src, size_t bytes, cudaStream_t st) {
    if (!src || !bytes) return 0;
    if (in_arena(c, dst)) { memcpy(dst, src, bytes); return 0; }
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    return 0;
}
// staging -> host: from the arena once the kernels have finished, else an asynchronous copy
static inline int ctx_d2h(g2048_ctx* c, void* dst, const void* src, size_t bytes, cudaStream_t st) {
    if (!dst || !bytes) return 0;
    if (in_arena(c, src)) { CK(cudaStreamSynchronize(st)); memcpy(dst, src, bytes); return 0; }
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
    return 0;
}
#define H2D(dst, src, bytes) RC(ctx_h2d(c, dst, src, (size_t)(bytes), st))
#define D2H(dst, src, bytes) RC(ctx_d2h(c, dst, src, (size_t)(bytes), st))
#define RC(expr)              \
    do {                      \
        int rc_ = (expr);     \
        if (rc_) return rc_;  \
    } while (0)

G2048_API int g2048_ctx_env_reset(g2048_ctx* c, uint64_t* boards, int32_t* score, const uint8_t* mask,
                                  const uint8_t* replay_draws, int64_t n, uint64_t seed, uint64_t episode_idx,
                                  uint64_t env_id_base) {
    CTX_ENTER(false);
    if (!boards) return fail(G2048_ERR_ARG, "g2048_ctx_env_reset: boards is null");
    if (mask) { H2D(c->boards, boards, n * 8); H2D(c->score, score, n * 4); H2D(c->bytes_a, mask, n); }
    H2D(c->draws, replay_draws, n * 4);
    RC(g2048_env_reset((uint64_t*)c->boards, score ? c->score : nullptr, mask ? c->bytes_a : nullptr,
                       replay_draws ? c->draws : nullptr, n, seed, episode_idx, env_id_base, st));
    D2H(boards, c->boards, n * 8);
    D2H(score, c->score, n * 4);
    CK(cudaStreamSynchronize(st));
    return 0;
}
G2048_API int g2048_ctx_env_step(g2048_ctx* c, uint64_t* boards, uint64_t* aux, int32_t* score, const uint8_t* actions,
                                 const uint8_t* replay_draws, double* reward_f64, uint8_t* flags, uint8_t* maxlvl,
                                 int32_t* move_score, int64_t n, int flavour, uint64_t seed, uint64_t step_idx,
                                 uint64_t env_id_base) {
    CTX_ENTER(false);
    if (!boards || !actions) return fail(G2048_ERR_ARG, "g2048_ctx_env_step: boards/actions is null");
    H2D(c->boards, boards, n * 8);
    H2D(c->aux, aux, n * 8);
    H2D(c->score, score, n * 4);
    H2D(c->bytes_a, actions, n);
    H2D(c->draws, replay_draws, n * 4);
    RC(g2048_env_step((uint64_t*)c->boards, aux ? (uint64_t*)c->aux : nullptr, score ? c->score : nullptr, c->bytes_a,
                      replay_draws ? c->draws : nullptr, reward_f64 ? c->rew64 : nullptr, nullptr,
                      flags ? c->flags : nullptr, maxlvl ? c->maxlvl : nullptr, move_score ? c->move_score : nullptr, n,
                      flavour, seed, step_idx, env_id_base, st));
    D2H(boards, c->boards, n * 8);
    D2H(aux, c->aux, n * 8);
    D2H(score, c->score, n * 4);
    D2H(reward_f64, c->rew64, n * 8);
    D2H(flags, c->flags, n);
    D2H(maxlvl, c->maxlvl, n);
    D2H(move_score, c->move_score, n * 4);
    CK(cudaStreamSynchronize(st));
    return 0;
}
G2048_API int g2048_ctx_legal_mask(g2048_ctx* c, const uint64_t* boards, uint8_t* out, int64_t n) {
    CTX_ENTER(false);
    if (n && (!boards || !out)) return fail(G2048_ERR_ARG, "g2048_ctx_legal_mask: null pointer");
    H2D(c->boards, boards, n * 8);
    RC(g2048_legal_mask((const uint64_t*)c->boards, c->flags, n, st));
    D2H(out, c->flags, n);
    CK(cudaStreamSynchronize(st));
    return 0;
}
G2048_API int g2048_ctx_move_trial(g2048_ctx* c, const uint64_t* boards_in, const uint8_t* actions, uint64_t* out_boards,
                                   uint8_t* moved, int32_t* move_score, int64_t n) {
    CTX_ENTER(false);
    if (n && (!boards_in || !actions)) return fail(G2048_ERR_ARG, "g2048_ctx_move_trial: null pointer");
    H2D(c->boards, boards_in, n * 8);
    H2D(c->bytes_a, actions, n);
    RC(g2048_move_trial((const uint64_t*)c->boards, c->bytes_a, (uint64_t*)c->keys2, c->flags, c->move_score, n, st));
    D2H(out_boards, c->keys2, n * 8);
    D2H(moved, c->flags, n);
    D2H(move_score, c->move_score, n * 4);
    CK(cudaStreamSynchronize(st));
    return 0;
}
static int ctx_rollout(g2048_ctx* c, uint64_t* boards, uint64_t* aux, int32_t* score, int64_t n, int64_t k_steps,
                       int flavour, float lr, float gamma, double eps, uint64_t seed, uint64_t step_base,
                       uint64_t env_id_base, int64_t* counters, bool qlearn) {
    CTX_ENTER(qlearn);
    if (!boards) return fail(G2048_ERR_ARG, "rollout: boards is null");
    CK(cudaMemsetAsync(c->counters, 0, G2048_N_COUNTERS * sizeof(long long), st));
    // Large batches go through in chunks on three streams, so that the host<->device copies of one chunk overlap
    // the kernel of another (PCIe is full duplex: H2D, D2H and compute all run at once).  Env ids are global, so
    // the result does not depend on the chunking.  The rollout kernels run one env per thread and round with
    // T = SMs x 1024 threads resident, so every chunk but the last is a multiple of T (no chunk adds a partly
    // filled round); the first and the last chunk are the small ones, because their copies are the exposed ones.
    DeviceState* D = nullptr;
    RC(current_device_state(&D));
    const int64_t T = (int64_t)D->sm_count * kRolloutThreads;
    const int64_t unit = T * (n / (6 * T) > 1 ? n / (6 * T) : 1);
    const bool pipelined = n >= (1 << 18) && n > unit;
    // the chunks of a Q-learning rollout share one list of deferred updates, applied once all their kernels are done
    // (while the last chunk's results are still on their way back to the host)
    DeferBuffers db{};
    const bool shared_list = qlearn && pipelined && n >= kDeferMinEnvs;
    if (shared_list) RC(deferred_list(D, deferred_capacity(n, k_steps), st, db));
    CK(cudaEventRecord(c->ev_start, st));
    int64_t lo = 0;
    for (int j = 0; lo < n; ++j) {
        int64_t rem = n - lo, m = rem;
        if (pipelined) m = j == 0 ? unit : rem >= 3 * unit ? 2 * unit : rem > unit ? unit : rem;
        cudaStream_t ps = pipelined ? c->pipe[j % 3] : st;
        if (pipelined) CK(cudaStreamWaitEvent(ps, c->ev_start, 0));
        CK(cudaMemcpyAsync(c->boards + lo, boards + lo, (size_t)m * 8, cudaMemcpyDefault, ps));
        if (aux) CK(cudaMemcpyAsync(c->aux + lo, aux + lo, (size_t)m * 8, cudaMemcpyDefault, ps));
        if (score) CK(cudaMemcpyAsync(c->score + lo, score + lo, (size_t)m * 4, cudaMemcpyDefault, ps));
        uint64_t* dbo = (uint64_t*)c->boards + lo;
        uint64_t* da = aux ? (uint64_t*)c->aux + lo : nullptr;
        int32_t* ds = score ? c->score + lo : nullptr;
        if (qlearn)
            RC(launch_rollout_qlearn(D, LocalTable{(Slot*)c->table, c->capacity - 1}, dbo, da, ds, m, k_steps, flavour, lr,
                                     gamma, eps, seed, step_base, env_id_base + (uint64_t)lo, (int64_t*)c->counters, ps,
                                     shared_list ? &db : nullptr));
        else
            RC(g2048_rollout_random(dbo, da, ds, m, k_steps, flavour, seed, step_base, env_id_base + (uint64_t)lo,
                                    (int64_t*)c->counters, ps));
        if (shared_list) CK(cudaEventRecord(c->ev_kernel[j % 3], ps));
        CK(cudaMemcpyAsync(boards + lo, c->boards + lo, (size_t)m * 8, cudaMemcpyDefault, ps));
        if (aux) CK(cudaMemcpyAsync(aux + lo, c->aux + lo, (size_t)m * 8, cudaMemcpyDefault, ps));
        if (score) CK(cudaMemcpyAsync(score + lo, c->score + lo, (size_t)m * 4, cudaMemcpyDefault, ps));
        lo += m;
    }
    if (shared_list) {
        for (int j = 0; j < 3; ++j) CK(cudaStreamWaitEvent(st, c->ev_kernel[j], 0));
        RC(apply_deferred(D, LocalTable{(Slot*)c->table, c->capacity - 1}, db, lr, st));
    }
    if (pipelined)
        for (int j = 0; j < 3; ++j) {
            CK(cudaEventRecord(c->ev_done[j], c->pipe[j]));
            CK(cudaStreamWaitEvent(st, c->ev_done[j], 0));
        }
    D2H(counters, c->counters, G2048_N_COUNTERS * sizeof(long long));
    CK(cudaStreamSynchronize(st));
    return 0;
}
G2048_API int g2048_ctx_rollout_random(g2048_ctx* c, uint64_t* boards, uint64_t* aux, int32_t* score, int64_t n,
                                       int64_t k_steps, int flavour, uint64_t seed, uint64_t step_base,
                                       uint64_t env_id_base, int64_t* counters) {
    return ctx_rollout(c, boards, aux, score, n, k_steps, flavour, 0.f, 0.f, 0.0, seed, step_base, env_id_base, counters, false);
}
G2048_API int g2048_ctx_rollout_qlearn(g2048_ctx* c, uint64_t* boards, uint64_t* aux, int32_t* score, int64_t n,
                                       int64_t k_steps, int flavour, float lr, float gamma, double eps, uint64_t seed,
                                       uint64_t step_base, uint64_t env_id_base, int64_t* counters) {
    return ctx_rollout(c, boards, aux, score, n, k_steps, flavour, lr, gamma, eps, seed, step_base, env_id_base, counters, true);
}
G2048_API int g2048_ctx_qtable_lookup(g2048_ctx* c, const uint64_t* keys, int64_t n, float* rows, uint8_t* found,
                                      int insert) {
    CTX_ENTER(true);
    if (n && (!keys || !rows)) return fail(G2048_ERR_ARG, "g2048_ctx_qtable_lookup: null pointer");
    H2D(c->boards, keys, n * 8);
    RC(g2048_qtable_lookup(c->table, c->capacity, (const uint64_t*)c->boards, n, c->rows, found ? c->flags : nullptr,
                           insert, st));
    D2H(rows, c->rows, n * 16);
    D2H(found, c->flags, n);
    CK(cudaStreamSynchronize(st));
    return 0;
}
G2048_API int g2048_ctx_choose_action(g2048_ctx* c, const uint64_t* boards, uint8_t* actions, int64_t n, double eps,
                                      uint64_t seed, uint64_t step_idx, uint64_t env_id_base) {
    CTX_ENTER(true);
    if (n && (!boards || !actions)) return fail(G2048_ERR_ARG, "g2048_ctx_choose_action: null pointer");
    H2D(c->boards, boards, n * 8);
    RC(g2048_choose_action(c->table, c->capacity, (const uint64_t*)c->boards, c->bytes_a, n, eps, seed, step_idx,
                           env_id_base, st));
    D2H(actions, c->bytes_a, n);
    CK(cudaStreamSynchronize(st));
    return 0;
}
G2048_API int g2048_ctx_qtable_update(g2048_ctx* c, const uint64_t* s, const uint8_t* a, const float* r,
                                      const uint64_t* s2, const uint8_t* done, int64_t n, float lr, float gamma,
                                      int mode) {
    CTX_ENTER(true);
    if (n && (!s || !a || !r || !s2 || !done)) return fail(G2048_ERR_ARG, "g2048_ctx_qtable_update: null pointer");
    H2D(c->boards, s, n * 8);
    H2D(c->keys2, s2, n * 8);
    H2D(c->bytes_a, a, n);
    H2D(c->bytes_b, done, n);
    H2D(c->rew32, r, n * 4);
    RC(g2048_qtable_update(c->table, c->capacity, (const uint64_t*)c->boards, c->bytes_a, c->rew32,
                           (const uint64_t*)c->keys2, c->bytes_b, n, lr, gamma, mode, c->scratch, c->scratch_bytes, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}
G2048_API int64_t g2048_ctx_qtable_size(g2048_ctx* c) {
    if (!c || !c->table) { fail(G2048_ERR_ARG, "context has no Q-table"); return -1; }
    if (cudaSetDevice(c->device) != cudaSuccess) return -1;
    long long* cnt = c->counters + G2048_N_COUNTERS;
    if (g2048_qtable_size(c->table, c->capacity, (int64_t*)cnt, c->stream)) return -1;
    long long h = 0;
    if (cudaMemcpyAsync(&h, cnt, sizeof h, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess) { fail((int)cudaGetLastError(), "g2048_ctx_qtable_size"); return -1; }
    return h;
}
G2048_API int64_t g2048_ctx_qtable_export(g2048_ctx* c, uint64_t* keys, float* rows, int64_t max_out) {
    if (!c || !c->table || max_out < 0) { fail(G2048_ERR_ARG, "g2048_ctx_qtable_export: bad arguments"); return -1; }
    if (cudaSetDevice(c->device) != cudaSuccess) return -1;
    long long* cnt = c->counters + G2048_N_COUNTERS;
    u64* dk = nullptr;
    float* dr = nullptr;
    long long h = -1;
    size_t m = (size_t)(max_out > 0 ? max_out : 1);
    if (cudaMalloc(&dk, m * 8) != cudaSuccess || cudaMalloc(&dr, m * 16) != cudaSuccess) {
        fail((int)cudaGetLastError(), "g2048_ctx_qtable_export: cudaMalloc");
    } else if (cudaMemsetAsync(cnt, 0, sizeof(long long), c->stream) == cudaSuccess &&
               g2048_qtable_export(c->table, c->capacity, (uint64_t*)dk, dr, max_out, (int64_t*)cnt, c->stream) == 0 &&
               cudaMemcpyAsync(&h, cnt, sizeof h, cudaMemcpyDeviceToHost, c->stream) == cudaSuccess &&
               cudaStreamSynchronize(c->stream) == cudaSuccess) {
        long long w = h < max_out ? h : max_out;
        if (w > 0 && (cudaMemcpy(keys, dk, (size_t)w * 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
                      cudaMemcpy(rows, dr, (size_t)w * 16, cudaMemcpyDeviceToHost) != cudaSuccess)) {
            fail((int)cudaGetLastError(), "g2048_ctx_qtable_export: copy");
            h = -1;
        }
    } else {
        fail((int)cudaGetLastError(), "g2048_ctx_qtable_export");
        h = -1;
    }
    if (dk) cudaFree(dk);
    if (dr) cudaFree(dr);
    return h;
}
G2048_API int g2048_ctx_qtable_clear(g2048_ctx* c) {
    if (!c || !c->table) return fail(G2048_ERR_ARG, "context has no Q-table");
    CK(cudaSetDevice(c->device));
    RC(g2048_qtable_clear(c->table, c->capacity, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}
