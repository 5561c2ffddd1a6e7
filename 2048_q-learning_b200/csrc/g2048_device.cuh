// g2048_device.cuh -- device-side building blocks of the batched 2048 env and the
// HBM hash Q-table (sm_100a).  Board = uint64, cell (r,c) = nibble 4r+c holding
// log2(tile), 0 = empty; row r = bits 16r..16r+15, "left" = towards nibble 0.
//
// Reference semantics (paths relative to the reference root):
//   QLearningBase/environment/Game2048_env.py        penalty-flavour env
//   Deep_QLearning/environment/Game2048_nopenalty_env.py + mainDQL_CNN_step2.py:163-237
//   QLearningBase/Agent/main.py:34-43                choose_action / update_q_value
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/g2048.h"

namespace g2048 {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr u64 kNib1 = 0x1111111111111111ull;   // LSB of every nibble
constexpr u64 kColsLeft3 = 0x0111011101110111ull;  // cells with a right-hand neighbour in the same row
constexpr u64 kRowsTop3 = 0x0000111111111111ull;   // cells with a neighbour below
constexpr u32 kIs4Thresh = 0xE6666666u;        // floor(0.9 * 2^32): `random() < 0.9 -> 2`, Game2048_env.py:20
constexpr int kPenSat = 25;                    // stall penalty table saturates at -10 from index 25 on
constexpr u32 kNoSlot = 0xFFFFFFFFu;
constexpr int kMaxProbe = 256;

// Host-built tables resident in HBM/L2 (built once per device by g2048_init).
struct Tables {
    const uint16_t* lut_row;      // [65536] row moved left, stored at the swizzled index lut_index(row)
    const uint8_t* lut_merged;    // [65536] the (at most two) merged levels, hi nibble >= lo nibble (same index)
    const uint32_t* lut_mscore;   // [256] merged byte -> move score (bits 0-23) | hi level << 24
    const double* rew_valid;      // [16 lvl][16 d][256 score/4] normalised reward of a valid move
    const double* rew_invalid;    // [2 game_over][16 lvl][16 d]
    const double* pen;            // [32] stall penalty sequence, pen[0] = -1
};

// Row LUT view (shared-memory copy in the fused kernels, global copy otherwise).
struct Lut {
    const uint16_t* row;
    const uint8_t* merged;
    const uint32_t* mscore;
    // hot part of the reward tables staged next to the LUT (NULL on the global-LUT path):
    // [rew_invalid 512][pen 32][rew_valid for d < 2 and score < 256: 16 levels x 2 x 64]
    const double* rew_hot;
};
constexpr int kHotInvalid = 0, kHotPen = 512, kHotValid = 544, kHotDoubles = 544 + 16 * 2 * 64;
// Bank swizzle of the LUT index, applied to two packed rows at once.  Real boards hold small levels, so the
// plain index puts most lanes of a warp into ~8 of the 32 banks (measured 6.9 wavefronts per LDS); XOR-folding
// bits 7-13 into bits 1-6 spreads them (2.7 wavefronts on rollout boards) and stays a bijection per row.
constexpr u32 kSwzMask = 0x007E007Eu;
__host__ __device__ __forceinline__ u32 lut_index2(u32 two_rows) {
    return two_rows ^ (((two_rows >> 6) ^ (two_rows >> 7)) & kSwzMask);
}

// ------------------------------------------------------------------ Philox4x32-10
// counter = (env id lo, env id hi, step lo, stream<<24 | step hi), key = seed
struct Draw4 { u32 x0, x1, x2, x3; };
__device__ __forceinline__ Draw4 philox(u64 seed, u64 env_id, u64 step, u32 stream) {
    u32 c0 = (u32)env_id, c1 = (u32)(env_id >> 32), c2 = (u32)step;
    u32 c3 = (stream << 24) | ((u32)(step >> 32) & 0x00FFFFFFu);
    u32 k0 = (u32)seed, k1 = (u32)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        u32 h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        u32 h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ k0; c1 = l1;
        c2 = h0 ^ c3 ^ k1; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return Draw4{c0, c1, c2, c3};
}

// ------------------------------------------------------------------ board bit tricks
__device__ __forceinline__ u64 nzmask(u64 b) {  // bit 4i set iff nibble i != 0
    u64 t = b | (b >> 1);
    t |= t >> 2;
    return t & kNib1;
}
__device__ __forceinline__ u64 is15mask(u64 b) {  // bit 4i set iff nibble i == 15
    u64 t = b & (b >> 1);
    t &= t >> 2;
    return t & kNib1;
}
// Direction handling on the two 32-bit halves (lo = rows 0,1; hi = rows 2,3).  Both stages are involutions
// and are made conditional through their PRMT selectors / delta-swap masks, so a warp whose lanes move in four
// different directions runs one instruction stream:
//   T (transpose)    byte permute [r0.lo r2.lo r1.lo r3.lo | r0.hi r2.hi r1.hi r3.hi], then swap the high nibble
//                    of bytes 0,1 with the low nibble of bytes 2,3 (delta swap, shift 12)
//   R (reverse rows) swap the bytes of each row, then the nibbles of each byte (delta swap, shift 4)
struct MoveCtl { u32 selTlo, selThi, maskT, selR, maskR; };
__device__ __forceinline__ MoveCtl move_ctl(int a) {  // 0 left | 1 up = T | 2 right = R | 3 down = T then R
    MoveCtl c;
    bool t = a & 1, r = a & 2;
    c.selTlo = t ? 0x6240u : 0x3210u;
    c.selThi = t ? 0x7351u : 0x7654u;
    c.maskT = t ? 0x0000F0F0u : 0u;
    c.selR = r ? 0x2301u : 0x3210u;
    c.maskR = r ? 0x0F0F0F0Fu : 0u;
    return c;
}
__device__ __forceinline__ u32 delta_swap(u32 x, u32 mask, int shift) {
    u32 t = ((x >> shift) ^ x) & mask;
    return x ^ t ^ (t << shift);
}
__device__ __forceinline__ void stage_T(u32& lo, u32& hi, const MoveCtl& c) {
    u32 l = __byte_perm(lo, hi, c.selTlo), h = __byte_perm(lo, hi, c.selThi);
    lo = delta_swap(l, c.maskT, 12);
    hi = delta_swap(h, c.maskT, 12);
}
__device__ __forceinline__ void stage_R(u32& lo, u32& hi, const MoveCtl& c) {
    lo = delta_swap(__byte_perm(lo, 0, c.selR), c.maskR, 4);
    hi = delta_swap(__byte_perm(hi, 0, c.selR), c.maskR, 4);
}
__device__ __forceinline__ int max_level(u64 b) {  // largest nibble
    u32 lo = (u32)b, hi = (u32)(b >> 32);
    u32 v0 = lo & 0x0F0F0F0Fu, v1 = (lo >> 4) & 0x0F0F0F0Fu, v2 = hi & 0x0F0F0F0Fu, v3 = (hi >> 4) & 0x0F0F0F0Fu;
    // bytewise max of values < 128: bit 7 of ((a|0x80..) - b) is set iff a >= b
    auto bmax = [](u32 a, u32 b) -> u32 {
        u32 ge = (((a | 0x80808080u) - b) >> 7) & 0x01010101u;
        u32 m = ge * 0xFFu;
        return (a & m) | (b & ~m);
    };
    u32 m = bmax(bmax(v0, v1), bmax(v2, v3));
    m = bmax(m, m >> 16);
    m = bmax(m, m >> 8);
    return (int)(m & 0xFFu);
}

// Result of one move: the moved board, whether anything moved, the move score and the largest merged level.
struct Moved {
    u64 board;
    int score;     // score += merged value (Game2048_env.py:35)
    int hi_level;  // largest level created by a merge (0 = none)
    bool moved;
};
// Game2048.move without the spawn (Game2048_env.py:51-60): canonicalise the direction (T, R), move towards
// nibble 0 of every row through the row LUT (move_left, :22-46), undo (R, T).
__device__ __forceinline__ Moved do_move(u64 b, int a, const Lut& L) {
    MoveCtl c = move_ctl(a);
    u32 lo = (u32)b, hi = (u32)(b >> 32);
    stage_T(lo, hi, c);
    stage_R(lo, hi, c);
    u32 plo = lut_index2(lo), phi = lut_index2(hi);
    u32 i0 = plo & 0xFFFFu, i1 = plo >> 16, i2 = phi & 0xFFFFu, i3 = phi >> 16;
    u32 r0 = L.row[i0], r1 = L.row[i1], r2 = L.row[i2], r3 = L.row[i3];
    u32 e0 = L.mscore[L.merged[i0]], e1 = L.mscore[L.merged[i1]], e2 = L.mscore[L.merged[i2]], e3 = L.mscore[L.merged[i3]];
    u32 nlo = r0 | (r1 << 16), nhi = r2 | (r3 << 16);
    Moved m;
    m.moved = (nlo != lo) | (nhi != hi);
    m.score = (int)((e0 + e1 + e2 + e3) & 0x00FFFFFFu);
    m.hi_level = (int)(max(max(e0, e1), max(e2, e3)) >> 24);
    stage_R(nlo, nhi, c);
    stage_T(nlo, nhi, c);
    m.board = ((u64)nhi << 32) | nlo;
    return m;
}

// position (bit offset, multiple of 4) of the k-th set bit of E (bits only at nibble LSBs),
// counted from nibble 0 = row-major order of np.where (add_number, Game2048_env.py:17-19)
__device__ __forceinline__ int kth_empty_pos(u64 E, int k) {
    u32 w = (u32)E;
    int base = 0, c = __popc(w);
    if (k >= c) { k -= c; w = (u32)(E >> 32); base = 32; }
    c = __popc(w & 0xFFFFu);
    if (k >= c) { k -= c; w >>= 16; base += 16; }
    c = __popc(w & 0xFFu);
    if (k >= c) { k -= c; w >>= 8; base += 8; }
    if (k >= (int)(w & 1u)) base += 4;
    return base;
}
// add_number (Game2048_env.py:16-20).  REPLAY: d0 = recorded cell index k, d1 = recorded is-4 flag;
// otherwise d0/d1 are 32-bit uniform draws: k = floor(d0 * n_empty / 2^32), 4 iff d1 >= 0.9 * 2^32.
// lvl_max is raised to the spawned level (1 or 2) when it is larger.
template <bool REPLAY>
__device__ __forceinline__ u64 spawn(u64 b, u32 d0, u32 d1, int& lvl_max, int* empties_left = nullptr) {
    u64 E = ~nzmask(b) & kNib1;
    int ne = __popcll(E);
    if (empties_left) *empties_left = ne > 0 ? ne - 1 : 0;
    if (ne == 0) return b;
    int k = REPLAY ? (int)d0 : (int)__umulhi(d0, (u32)ne);
    if (k >= ne) k = ne - 1;  // malformed replay input: stay in range
    int lvl = (REPLAY ? (d1 != 0) : (d1 >= kIs4Thresh)) ? 2 : 1;
    lvl_max = max(lvl_max, lvl);
    return b | ((u64)lvl << kth_empty_pos(E, k));
}
// Game2048.__init__ (Game2048_env.py:11-14): empty board + two spawns
template <bool REPLAY>
__device__ __forceinline__ u64 fresh_board(u32 d0, u32 d1, u32 d2, u32 d3) {
    int ka = REPLAY ? (int)(d0 & 15u) : (int)__umulhi(d0, 16u);
    int kb = REPLAY ? (int)d2 : (int)__umulhi(d2, 15u);
    if (kb > 14) kb = 14;
    bool fa = REPLAY ? (d1 != 0) : (d1 >= kIs4Thresh);
    bool fb = REPLAY ? (d3 != 0) : (d3 >= kIs4Thresh);
    int pb = kb < ka ? kb : kb + 1;
    return ((u64)(fa ? 2 : 1) << (4 * ka)) | ((u64)(fb ? 2 : 1) << (4 * pb));
}

// pairs of equal neighbours that can merge (level 15 cannot: 65536 is unrepresentable)
__device__ __forceinline__ void equal_pairs(u64 b, u64 F, u64& eqH, u64& eqV) {
    u64 ok = F & ~is15mask(b);
    eqH = ~nzmask(b ^ (b >> 4)) & kColsLeft3 & ok;
    eqV = ~nzmask(b ^ (b >> 16)) & kRowsTop3 & ok;
}
// is_game_over for a board KNOWN to be full: no two equal neighbours.  x = b ^ (b >> 4) has a zero nibble where two
// horizontal neighbours are equal (column 3 pairs with the next row: forced non-zero), same with >> 16 vertically;
// "has a zero nibble" is the exact (x - 0x11..1) & ~x & 0x88..8 test.  Level-15 pairs cannot merge (they are the
// only equal neighbours that do not count): boards holding a 15 take the general path.
__device__ __forceinline__ bool is_dead_full(u64 b);
__device__ __forceinline__ bool has_zero_nibble(u64 x) { return ((x - kNib1) & ~x & (kNib1 << 3)) != 0; }

// is_game_over as a predicate (Game2048_env.py:65-75): full and nothing merges
__device__ __forceinline__ bool is_dead(u64 b) {
    u64 F = nzmask(b);
    if (F != kNib1) return false;
    u64 eqH, eqV;
    equal_pairs(b, F, eqH, eqV);
    return (eqH | eqV) == 0;
}
__device__ __forceinline__ bool is_dead_full(u64 b) {
    if (is15mask(b)) return is_dead(b);
    u64 h = (b ^ (b >> 4)) | 0xF000F000F000F000ull;
    u64 v = (b ^ (b >> 16)) | 0xFFFF000000000000ull;
    return !has_zero_nibble(h) && !has_zero_nibble(v);
}
// bit a set iff move(a, trial=True) would move (mainDQL_CNN_step2.py:169-174)
__device__ __forceinline__ u32 legal_mask(u64 b) {
    u64 F = nzmask(b), E = ~F & kNib1, eqH, eqV;
    equal_pairs(b, F, eqH, eqV);
    u32 left = ((E & (F >> 4) & kColsLeft3) | eqH) != 0;
    u32 right = ((F & (E >> 4) & kColsLeft3) | eqH) != 0;
    u32 up = ((E & (F >> 16) & kRowsTop3) | eqV) != 0;
    u32 down = ((F & (E >> 16) & kRowsTop3) | eqV) != 0;
    return left | (up << 1) | (right << 2) | (down << 3);
}

// ------------------------------------------------------------------ env state in registers
struct Env {
    u64 board;
    u32 cons_count;   // consecutive_count (saturating)
    u32 small;        // prev_level | cons_action << 8 | pen_idx << 16   (aux low word)
    int score;        // env.score
    int maxlvl;       // largest level on `board`, carried across steps (tiles only grow within an episode)
};
__device__ __forceinline__ void env_load(Env& e, u64 board, u64 aux, int score) {
    e.board = board; e.small = (u32)aux; e.cons_count = (u32)(aux >> 32); e.score = score;
    e.maxlvl = max_level(board);
}
__device__ __forceinline__ u64 env_to_aux(const Env& e) { return ((u64)e.cons_count << 32) | e.small; }

struct StepOut {
    double reward;
    int move_score;
    int maxlvl;
    bool valid, game_over, done;
};

// calculate_reward + update_and_normalize (Game2048_env.py:136-184, 197-205) through host-built
// tables: the normalised reward of an invalid move depends only on (game_over, level, d) and that of
// a valid move on (level, d, score) with score a multiple of 4; score >= 1024 gives exactly 10.
// d = max(level - prev_level, 0) is the progress step (previous_max is raised even when the move was
// invalid and the bonus discarded, :148-150).
__device__ __forceinline__ double shaped_reward(int score, bool valid, bool game_over, int lvl, int& prev_level,
                                                const Tables& T, const Lut& L) {
    int d = lvl > prev_level ? lvl - prev_level : 0;
    if (lvl > prev_level) prev_level = lvl;
    if (!valid) {
        int i = (game_over ? 256 : 0) + lvl * 16 + d;
        return L.rew_hot ? L.rew_hot[kHotInvalid + i] : __ldg(T.rew_invalid + i);
    }
    if (score >= 1024) return 10.0;
    if (L.rew_hot && d < 2 && score < 256) return L.rew_hot[kHotValid + (lvl * 2 + d) * 64 + (score >> 2)];
    return __ldg(T.rew_valid + ((lvl * 16 + d) * 256 + (score >> 2)));
}

// Game2048_env.step, penalty flavour (Game2048_env.py:97-129).  d0,d1: spawn draws of the move.
template <bool REPLAY>
__device__ __forceinline__ void penalty_step(Env& e, int a, u32 d0, u32 d1, const Lut& L, const Tables& T, StepOut& o) {
    Moved m = do_move(e.board, a, L);                            // :98
    bool valid = m.moved;
    int ms = m.score, lvl = max(e.maxlvl, m.hi_level);           // np.max(board) :100, incrementally
    u64 b1 = m.board;
    int empties = 1;                                             // cells still empty after the spawn
    if (valid) b1 = spawn<REPLAY>(b1, d0, d1, lvl, &empties);    // :61-62
    else empties = (nzmask(b1) != kNib1);                        // unchanged board: only "full or not" matters
    bool game_over = empties == 0 && is_dead_full(b1);           // :99
    if (lvl < 1) lvl = 1;                                        // max(2, max_number) :141
    e.board = b1;
    e.maxlvl = lvl;
    e.score += ms;                                               // :104
    int prev_level = (int)(e.small & 0xFFu), cons_action = (int)((e.small >> 8) & 0xFFu);
    int pen_idx = (int)((e.small >> 16) & 0xFFu);
    double reward = shaped_reward(ms, valid, game_over, lvl, prev_level, T, L);  // :107
    if (a == cons_action) {                                      // :110-115
        if (e.cons_count != 0xFFFFFFFFu) e.cons_count += 1;
    } else {
        cons_action = a;
        e.cons_count = 1;
        pen_idx = 0;
    }
    bool done = !valid && game_over;                             // :117-118
    if (e.cons_count > 10) {                                     // :121-127
        if (e.cons_count > 100) done = true;
        if (pen_idx < kPenSat) pen_idx += 1;
        reward = __dadd_rn(reward, L.rew_hot ? L.rew_hot[kHotPen + pen_idx] : __ldg(T.pen + pen_idx));
    }
    e.small = (u32)prev_level | ((u32)cons_action << 8) | ((u32)pen_idx << 16);
    o.reward = reward; o.move_score = ms; o.maxlvl = lvl;
    o.valid = valid; o.game_over = game_over; o.done = done;
}

// Game2048_env.step, nopenalty flavour (Game2048_nopenalty_env.py:106-138) under the caller protocol
// of mainDQL_CNN_step2.py:163-237: e.board is the committed board S on entry and the returned
// moved_board M on exit.  (q0,q1): draws of the spawn inside is_game_over when S is full and some
// action a' is legal -- the returned board is then the result of the FIRST legal a', not of the
// agent's action (full-board quirk).
template <bool REPLAY>
__device__ __forceinline__ void nopenalty_step(Env& e, int a, u32 d0, u32 d1, u32 q0, u32 q1, const Lut& L, StepOut& o) {
    u64 S = e.board;
    Moved m = do_move(S, a, L);                                  // :53-66
    bool valid = m.moved, game_over = false;
    int ms = m.score, lvl = e.maxlvl;
    u64 M = S;
    if (nzmask(S) != kNib1) {                                    // S has an empty cell: is_game_over is False (:69)
        if (valid) { lvl = max(lvl, m.hi_level); M = spawn<REPLAY>(m.board, d0, d1, lvl); }
    } else {                                                     // :70-78, evaluated on S; the agent's own result is discarded
        u32 lm = legal_mask(S);
        if (lm == 0) {
            game_over = true;
        } else {
            int a2 = __ffs((int)lm) - 1;
            Moved m2 = (a2 == a) ? m : do_move(S, a2, L);
            lvl = max(lvl, m2.hi_level);
            M = spawn<REPLAY>(m2.board, q0, q1, lvl);
        }
    }
    e.board = M;                                                 // np.max(moved_board) :108
    e.maxlvl = lvl;
    e.score += ms;                                               // :111
    o.reward = (!valid && !game_over) ? -10.0 : (double)ms;      // :122-128
    o.move_score = ms; o.maxlvl = lvl;
    o.valid = valid; o.game_over = game_over; o.done = game_over; // :117-118
}

// one env step with Philox draws x (STREAM_STEP).  The nopenalty full-board spawn reuses x0, x1: in that case the
// spawn of the agent's own move is discarded, so those two draws are otherwise unused and the joint distribution
// is the reference's (which draws fresh numbers there).
template <int FLAVOUR>
__device__ __forceinline__ void philox_step(Env& e, int a, const Draw4& x, u64 seed, u64 env_id, u64 t, const Lut& L,
                                            const Tables& T, StepOut& o) {
    if (FLAVOUR == G2048_FLAVOUR_PENALTY) penalty_step<false>(e, a, x.x0, x.x1, L, T, o);
    else nopenalty_step<false>(e, a, x.x0, x.x1, x.x0, x.x1, L, o);
}
__device__ __forceinline__ void philox_autoreset(Env& e, u64 seed, u64 env_id, u64 t) {
    Draw4 y = philox(seed, env_id, t, G2048_STREAM_AUTORESET);
    e.board = fresh_board<false>(y.x0, y.x1, y.x2, y.x3);
    e.maxlvl = ((y.x1 >= kIs4Thresh) || (y.x3 >= kIs4Thresh)) ? 2 : 1;
    e.score = 0;   // reset() zeroes env.score only (Game2048_env.py:187-191); aux state survives
}

// ------------------------------------------------------------------ HBM hash Q-table
// 32-byte slots (one DRAM sector) = two 16-byte halves:  W0 = {key, q[0], q[1]}   W1 = {q[2], q[3], meta (unused)}.
// key 0 = empty (an all-empty board never occurs).  Open addressing, linear probing, capacity a power of two.
// The key shares W0 with two Q values so that a NEW state and its first update can be written by ONE 128-bit
// compare-and-swap (see k_rollout_qlearn): on B200 every write-type request to the table costs about as much as the
// line fill of the lookup itself, whatever it writes (tools/membench5.cu), so one request less per new state is worth
// a third of the table time.
struct __align__(32) Slot {
    u64 key;
    float q[4];
    u64 meta;
};
static_assert(sizeof(Slot) == G2048_QTABLE_SLOT_BYTES, "slot size");
static_assert(offsetof(Slot, q) == 8 && offsetof(Slot, meta) == 24, "slot layout");

__device__ __forceinline__ u64 mix64(u64 x) {  // splitmix64 finaliser
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}
// Where the slots live.  LocalTable: one array in this GPU's HBM.  ShardedTable: the global slot range cut into
// 2^k equal shards, shard j in the HBM of GPU j and mapped into every process of the box (CUDA IPC over NVLink 5 /
// NVSwitch): a slot is addressed the same way from every GPU, reads are plain (system-scope) loads and updates are
// system-scope atomics executed at the owner's L2, so all GPUs learn ONE table.  Shard = the TOP bits of the slot
// index, so a probe sequence stays on one GPU.
struct LocalTable {
    Slot* base;
    u64 mask;
    static constexpr bool kSys = false;
    static constexpr bool kSysLoad = false;
    static constexpr bool kMerge128 = true;    // insert + first update as one 128-bit CAS (k_rollout_qlearn)
    __device__ __forceinline__ Slot* at(u64 h) const { return base + h; }
    __device__ __forceinline__ LocalTable view(Slot**) const { return *this; }
};
// device-side view of a sharded table: the shard pointers sit in shared memory (a per-lane index into the kernel
// parameters would be a constant-bank load, which the hardware serialises per distinct index)
template <bool SYS_LOAD, bool SYS_ATOM>
struct ShardedView {
    Slot* const* base;
    u64 mask, low;
    u32 shift;
    static constexpr bool kSys = SYS_ATOM;
    static constexpr bool kSysLoad = SYS_LOAD;
    static constexpr bool kMerge128 = true;    // (also over NVLink: measured 5x faster than a 64-bit insert with the value deferred)
    __device__ __forceinline__ Slot* at(u64 h) const { return base[h >> shift] + (h & low); }
};
template <bool SYS_LOAD, bool SYS_ATOM>
struct ShardedTableT {
    Slot* base[G2048_MAX_PEERS];
    u64 mask;       // global capacity - 1
    u64 low;        // slots per shard - 1
    u32 shift;      // log2(slots per shard)
    // call from all threads of the block, once, at kernel start; `sm` = G2048_MAX_PEERS pointers of shared memory
    __device__ __forceinline__ ShardedView<SYS_LOAD, SYS_ATOM> view(Slot** sm) const {
        if (threadIdx.x < G2048_MAX_PEERS) sm[threadIdx.x] = base[threadIdx.x];
        __syncthreads();
        return ShardedView<SYS_LOAD, SYS_ATOM>{sm, mask, low, shift};
    }
};
using ShardedTable = ShardedTableT<true, true>;
template <bool SYS> __device__ __forceinline__ u64 cas64(u64* p, u64 cmp, u64 val) {
    return SYS ? atomicCAS_system((unsigned long long*)p, cmp, val) : atomicCAS((unsigned long long*)p, cmp, val);
}
template <bool SYS> __device__ __forceinline__ u32 cas32(u32* p, u32 cmp, u32 val) {
    return SYS ? atomicCAS_system(p, cmp, val) : atomicCAS(p, cmp, val);
}
// one 256-bit load of a whole slot, coherent at L2 (L1 is bypassed: other SMs -- or, system scope, other GPUs --
// update rows with atomics).  .L2::64B: a miss fills 64 bytes instead of the default whole 128-byte line (ncu:
// 2.0 instead of 3.98 DRAM sectors per random load, same request rate -- tools/membench6.cu).
template <bool SYS = false>
__device__ __forceinline__ void load_slot(const Slot* s, u64& key, float4& q) {
    u64 k, q01, q23, m;
    if (SYS) asm volatile("ld.relaxed.sys.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(k), "=l"(q01), "=l"(q23), "=l"(m) : "l"(s));
#ifdef G2048_EXP_CGLOAD
    else asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(k), "=l"(q01), "=l"(q23), "=l"(m) : "l"(s));
#else
    else asm volatile("ld.relaxed.gpu.global.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(k), "=l"(q01), "=l"(q23), "=l"(m) : "l"(s));
#endif
    key = k;
    q.x = __uint_as_float((u32)q01); q.y = __uint_as_float((u32)(q01 >> 32));
    q.z = __uint_as_float((u32)q23); q.w = __uint_as_float((u32)(q23 >> 32));
}
// 128-bit compare-and-swap on the first half of a slot {key, q[0] | q[1] << 32}; returns what was there
template <bool SYS>
__device__ __forceinline__ void cas_w0(Slot* s, u64 cmp_key, u64 cmp_q01, u64 new_key, u64 new_q01, u64& old_key, u64& old_q01) {
    if (SYS)
        asm volatile("{\n\t.reg .b128 c, n, d;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 n, {%4, %5};\n\t"
                     "atom.relaxed.sys.global.cas.b128 d, [%6], c, n;\n\tmov.b128 {%0, %1}, d;\n\t}"
                     : "=l"(old_key), "=l"(old_q01) : "l"(cmp_key), "l"(cmp_q01), "l"(new_key), "l"(new_q01), "l"(s) : "memory");
    else
        asm volatile("{\n\t.reg .b128 c, n, d;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 n, {%4, %5};\n\t"
                     "atom.relaxed.gpu.global.cas.b128 d, [%6], c, n;\n\tmov.b128 {%0, %1}, d;\n\t}"
                     : "=l"(old_key), "=l"(old_q01) : "l"(cmp_key), "l"(cmp_q01), "l"(new_key), "l"(new_q01), "l"(s) : "memory");
}
// The probe sequence: the home slot, then its partner in the same 64 bytes (h ^ 1: a lookup's L2 fill brings both, so
// the second probe is an L2 hit instead of a second miss half of the time), then the next pair, starting on the home
// slot's side again.  p = index of the probe just made.
__device__ __forceinline__ u64 next_probe(u64 h, int p, u64 mask) { return (p & 1) ? (((h ^ 1) + 2) & mask) : (h ^ 1); }
// Probing by single 32-byte slots, pair by pair.  (A two-slot 64-byte bucket loaded as a pair was measured SLOWER --
// 10.7 vs 14.2 G steps/s: what saturates is the number of L2-miss sector requests, ~36 G/s on B200 whatever the
// fetch granularity, tools/membench.cu -- so every extra sector costs, even an adjacent one.)
// defaultdict semantics (main.py:16): reading a state creates its zero row.  Returns the slot index
// (kNoSlot if the probe limit is hit: the state is then treated as a zero row and not updated).
template <bool INSERT, class TAB>
__device__ __forceinline__ u32 table_find(const TAB& tab, u64 key, float4& q, u32& inserted) {
    u64 h = mix64(key) & tab.mask;
    for (int p = 0; p < kMaxProbe; h = next_probe(h, p, tab.mask), ++p) {
        u64 k;
        Slot* sp = tab.at(h);
        load_slot<TAB::kSysLoad>(sp, k, q);
        if (k == key) return (u32)h;
        if (k == 0) {
            if (!INSERT) break;
            u64 old = cas64<TAB::kSys>(&sp->key, 0ull, key);
            if (old == 0) { inserted += 1; q = make_float4(0.f, 0.f, 0.f, 0.f); return (u32)h; }
            if (old == key) { load_slot<TAB::kSysLoad>(sp, k, q); return (u32)h; }
        }
    }
    q = make_float4(0.f, 0.f, 0.f, 0.f);
    return kNoSlot;
}
template <bool INSERT>
__device__ __forceinline__ u32 table_find(Slot* tab, u64 mask, u64 key, float4& q, u32& inserted) {
    return table_find<INSERT>(LocalTable{tab, mask}, key, q, inserted);
}
// Lookup WITHOUT insert for the fused rollout: returns the slot of `key` and its row, or -- fresh = true -- the
// empty slot that ends its probe sequence with a zero row; nothing is written.  The fused kernel inserts a fresh
// state together with its first update (one 128-bit CAS on {key, q[0], q[1]}), or on its own when no update follows.
template <class TAB>
__device__ __forceinline__ u32 table_probe(const TAB& tab, u64 key, float4& q, bool& fresh, u32& dropped) {
    u64 h = mix64(key) & tab.mask, k = 1;
    int p = 0;
    for (; p < kMaxProbe; h = next_probe(h, p, tab.mask), ++p) {
        load_slot<TAB::kSysLoad>(tab.at(h), k, q);
        if (k == key || k == 0) break;
    }
    fresh = false;
    if (k == key) return (u32)h;
    q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p == kMaxProbe) { dropped += 1; return kNoSlot; }
    fresh = true;
    return (u32)h;
}
// Insert `key` (zero row) at `slot`, the empty slot a table_probe returned; if another state took the slot meanwhile,
// find-or-insert from the home slot again.  Returns the slot that holds the key now (kNoSlot: table full).
template <class TAB>
__device__ __forceinline__ u32 insert_at(const TAB& tab, u32 slot, u64 key, u32& inserted, u32& dropped) {
    u64 old = cas64<TAB::kSys>(&tab.at(slot)->key, 0ull, key);
    if (old == 0) { inserted += 1; return slot; }
    if (old == key) return slot;
    float4 r;
    u32 s = table_find<true>(tab, key, r, inserted);
    dropped += (s == kNoSlot);
    return s;
}
__device__ __forceinline__ float q_at(const float4& q, int a) { return a == 0 ? q.x : a == 1 ? q.y : a == 2 ? q.z : q.w; }
__device__ __forceinline__ void q_set(float4& q, int a, float v) {
    if (a == 0) q.x = v; else if (a == 1) q.y = v; else if (a == 2) q.z = v; else q.w = v;
}
// np.argmax: first maximum (main.py:38, :41)
__device__ __forceinline__ int argmax4(const float4& q) {
    int b = 0; float v = q.x;
    if (q.y > v) { v = q.y; b = 1; }
    if (q.z > v) { v = q.z; b = 2; }
    if (q.w > v) { b = 3; }
    return b;
}
__device__ __forceinline__ float max4(const float4& q) { return q_at(q, argmax4(q)); }
// choose_action (main.py:34-38): explore iff x2 < floor(eps * 2^32), random action = x3 >> 30
__device__ __forceinline__ int choose_action(const float4& q, const Draw4& x, u64 eps_thresh) {
    return ((u64)x.x2 < eps_thresh) ? (int)(x.x3 >> 30) : argmax4(q);
}
// update_q_value (main.py:40-43) in float32, every operation rounded on its own (-fmad=false):
//   target = r + (done ? 0 : gamma * best_next)     q = q + lr * (target - q)
__device__ __forceinline__ float td_target(float gamma, float r, float best_next, bool done) {
    float g = __fmul_rn(gamma, best_next);
    return __fadd_rn(r, done ? 0.0f : g);
}
__device__ __forceinline__ float td_apply(float q, float lr, float target) {
    return __fadd_rn(q, __fmul_rn(lr, __fsub_rn(target, q)));
}
// 16-byte exchange record of one transition: {state key, action | float bits of the TD target << 32}
__device__ __forceinline__ ulonglong2 pack_record(u64 key, int a, float target) {
    return make_ulonglong2(key, (u64)(a & 3) | ((u64)__float_as_uint(target) << 32));
}
// Atomic read-modify-write of one Q value: q <- q + lr * (target - q) as a CAS loop, so concurrent
// updates of the same (state, action) compose like the reference's sequential loop (a contraction
// towards the targets) instead of summing stale deltas, which diverges once the number of
// simultaneous updaters exceeds 2 / lr.  `guess` is the caller's last view of the value.
template <bool SYS = false>
__device__ __forceinline__ float q_update_atomic(float* addr, float guess, float lr, float target) {
    u32 assumed = __float_as_uint(guess);
    while (true) {
        float nq = td_apply(__uint_as_float(assumed), lr, target);
        u32 old = cas32<SYS>(reinterpret_cast<u32*>(addr), assumed, __float_as_uint(nq));
        if (old == assumed) return nq;
        assumed = old;
    }
}

// ------------------------------------------------------------------ counters
// Per-thread partial sums (32-bit where a launch cannot overflow them), reduced per warp and added to the
// caller's int64 counters with one atomic per warp and counter.
struct Counters {
    u32 steps = 0, valid = 0, episodes = 0, inserts = 0, dropped = 0, lost = 0, retried = 0;
    int maxlvl = 0;
    long long score = 0, reward_fx = 0;
    __device__ __forceinline__ void add(const StepOut& o) {
        steps += 1; valid += o.valid; episodes += o.done; score += o.move_score;
        maxlvl = max(o.maxlvl, maxlvl);
        reward_fx += (long long)(o.reward * 1048576.0);
    }
};
__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, s);
    return v;
}
__device__ __forceinline__ void flush_counters(const Counters& c, long long* out) {
    if (!out) return;
    long long v[9] = {c.steps, c.valid, c.episodes, c.score, c.reward_fx, c.inserts, c.dropped, c.lost, c.retried};
    const int idx[9] = {G2048_C_STEPS, G2048_C_VALID, G2048_C_EPISODES, G2048_C_SCORE, G2048_C_REWARD_FX,
                        G2048_C_INSERTS, G2048_C_DROPPED, G2048_C_LOST, G2048_C_RETRIED};
    int mx = c.maxlvl;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, s));
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        long long s = warp_sum(v[i]);
        if ((threadIdx.x & 31) == 0 && s) atomicAdd((unsigned long long*)(out + idx[i]), (unsigned long long)s);
    }
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(out + G2048_C_MAXLVL, (long long)mx);
}

}  // namespace g2048
