/*
 * g2048_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the hot path of Rocco9999/2048_Q-Learning:
 * batched 2048 env reset()/step(), epsilon-greedy choose_action and the tabular
 * Q-update.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` leg may load this file's library; the product
 * (2048_q-learning_b200/) never does.
 *
 * Parity pin: the reference has no tests or golden vectors of its own
 * (SURVEY.md section 4), so this oracle is pinned against outputs of the
 * reference itself, imported unmodified in the build container by
 * oracle/make_golden.py and committed under tests/golden/ (see
 * tests/test_oracle_golden.py).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference root).  The board is restated on 16 explicit cells
 * (cell[r][c] = log2(tile), 0 = empty) -- deliberately NOT the bit tricks the
 * CUDA kernels use, so that the two implementations are independent.
 *
 * Packed interchange format: board = uint64, cell (r,c) is
 * nibble 4r+c, value log2(tile).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ----- flavours / flag bits: same numeric values as include/g2048.h ------- */
enum { FLAVOUR_PENALTY = 0, FLAVOUR_NOPENALTY = 1 };
enum { FLAG_VALID = 1, FLAG_GAME_OVER = 2, FLAG_DONE = 4 };
#define AUX_INIT 0x000000000000FF01ull /* prev_level=1, cons_action=None(0xFF), pen_idx=0, cons_count=0 */
#define PEN_SAT 25 /* stall_penalty(k) == -10 for every k >= 25 */

/* ------------------------------------------------------------------------- */
/* board <-> cells                                                           */
/* ------------------------------------------------------------------------- */
static void unpack(uint64_t b, int cell[4][4]) {
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) cell[r][c] = (int)((b >> (4 * (4 * r + c))) & 0xF);
}
static uint64_t pack(int cell[4][4]) {
    uint64_t b = 0;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) b |= (uint64_t)(cell[r][c] & 0xF) << (4 * (4 * r + c));
    return b;
}

/* np.rot90(board): counter-clockwise quarter turn (Game2048_env.py:48-49).
 * new[i][j] = old[j][3-i]. */
static void rot90(int cell[4][4]) {
    int t[4][4];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) t[i][j] = cell[j][3 - i];
    memcpy(cell, t, sizeof t);
}

/* One row of move_left (Game2048_env.py:25-41): drop zeros, merge equal
 * neighbours once scanning left to right (skip flag :28-37), pad with zeros
 * (:41).  Levels: merging two tiles of level t gives t+1 and scores 2^(t+1)
 * (:34-35).  merged[] receives the (at most two) merged levels, 0 = none.
 * Two level-15 tiles cannot merge in the nibble format (65536 is
 * unrepresentable): they stay unmerged -- the one documented divergence from
 * the unbounded-int reference (DESIGN.md "nibble range"). */
static void row_left(const int in[4], int out[4], int merged[2]) {
    int nz[4], n = 0, m = 0, k = 0;
    out[0] = out[1] = out[2] = out[3] = 0;
    merged[0] = merged[1] = 0;
    for (int c = 0; c < 4; ++c)
        if (in[c]) nz[n++] = in[c];
    for (int i = 0; i < n; ++i) {
        if (i + 1 < n && nz[i] == nz[i + 1] && nz[i] < 15) {
            out[m++] = nz[i] + 1;
            merged[k++] = nz[i] + 1;
            ++i; /* skip the partner */
        } else {
            out[m++] = nz[i];
        }
    }
}

/* move_left (Game2048_env.py:22-46): all four rows; score += merged tile
 * value (:35); moved if any merge or any row changed (:38, :42-43). */
static int move_left(int cell[4][4], int64_t *score) {
    int moved = 0;
    for (int r = 0; r < 4; ++r) {
        int out[4], merged[2];
        row_left(cell[r], out, merged);
        for (int k = 0; k < 2; ++k)
            if (merged[k]) { *score += (int64_t)1 << merged[k]; moved = 1; }
        for (int c = 0; c < 4; ++c) {
            if (cell[r][c] != out[c]) moved = 1;
            cell[r][c] = out[c];
        }
    }
    return moved;
}

/* Game2048.move without the spawn (Game2048_env.py:51-60): rotate CCW `action`
 * times, move_left, rotate back (-action % 4) times.  0=left 1=up 2=right 3=down. */
static int move_nospawn(uint64_t *b, int action, int64_t *score) {
    int cell[4][4];
    unpack(*b, cell);
    for (int i = 0; i < action; ++i) rot90(cell);
    int moved = move_left(cell, score);
    for (int i = 0; i < ((4 - action) % 4); ++i) rot90(cell);
    *b = pack(cell);
    return moved;
}

static int count_empty(uint64_t b) {
    int n = 0;
    for (int i = 0; i < 16; ++i) n += ((b >> (4 * i)) & 0xF) == 0;
    return n;
}

/* add_number (Game2048_env.py:16-20): the k-th empty cell in row-major order
 * (np.where order) receives 2 (level 1) or 4 (level 2).  Returns 0 when the
 * board is full (no draw consumed, :18). */
static int spawn(uint64_t *b, int k, int is4) {
    int seen = 0;
    for (int i = 0; i < 16; ++i) {
        if (((*b >> (4 * i)) & 0xF) == 0) {
            if (seen == k) {
                *b |= (uint64_t)(is4 ? 2 : 1) << (4 * i);
                return 1;
            }
            ++seen;
        }
    }
    return 0;
}

static int max_level(uint64_t b) {
    int m = 0;
    for (int i = 0; i < 16; ++i) {
        int v = (int)((b >> (4 * i)) & 0xF);
        if (v > m) m = v;
    }
    return m;
}

/* is_game_over as a pure predicate (Game2048_env.py:65-75): no empty cell and
 * no action moves anything.  (The reference's phantom spawn + board restore
 * :70-74 has no effect on the board, only on the RNG stream.) */
static int dead(uint64_t b) {
    if (count_empty(b)) return 0;
    for (int a = 0; a < 4; ++a) {
        uint64_t t = b;
        int64_t s = 0;
        if (move_nospawn(&t, a, &s)) return 0;
    }
    return 1;
}

/* legal-move probe, mainDQL_CNN_step2.py:169-174: bit a set iff move(a, trial=True) moves. */
static int legal_mask(uint64_t b) {
    int m = 0;
    for (int a = 0; a < 4; ++a) {
        uint64_t t = b;
        int64_t s = 0;
        if (move_nospawn(&t, a, &s)) m |= 1 << a;
    }
    return m;
}

/* ------------------------------------------------------------------------- */
/* Philox4x32-10 (Salmon et al. 2011) -- the framework's counter-based RNG.  */
/* Not part of the reference (which uses the global MT19937); restated here  */
/* independently so that GPU rollouts can be bit-compared with CPU rollouts. */
/* counter = (env_id lo, env_id hi, step lo, (stream<<24)|step hi), key=seed */
/* ------------------------------------------------------------------------- */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
enum { STREAM_STEP = 0, STREAM_RESET = 1, STREAM_QUIRK = 2, STREAM_AUTORESET = 3 };
static void draws(uint64_t seed, uint64_t env_id, uint64_t step, uint32_t stream, uint32_t x[4]) {
    x[0] = (uint32_t)env_id;
    x[1] = (uint32_t)(env_id >> 32);
    x[2] = (uint32_t)step;
    x[3] = (stream << 24) | ((uint32_t)(step >> 32) & 0x00FFFFFFu);
    philox4x32_10(x, (uint32_t)seed, (uint32_t)(seed >> 32));
}
ORC_API void orc_philox(uint64_t seed, uint64_t env_id, uint64_t step, uint32_t stream, uint32_t *out4) {
    draws(seed, env_id, step, stream, out4);
}
/* spawn draw mapping: k = floor(x * n / 2^32); is4 iff x' >= floor(0.9 * 2^32)
 * (the reference's `np.random.random() < 0.9 -> 2`, Game2048_env.py:20). */
#define IS4_THRESH 0xE6666666u
static int draw_k(uint32_t x, int n) { return (int)(((uint64_t)x * (uint64_t)n) >> 32); }

/* ------------------------------------------------------------------------- */
/* reward shaping, penalty flavour                                           */
/* ------------------------------------------------------------------------- */
/* update_and_normalize (Game2048_env.py:197-205) */
static double normalize(double reward) {
    if (reward >= 0) return fmin(log2(reward + 1), 10);
    return -fmin(log2(fabs(reward - 1)), 10);
}

/* calculate_reward (Game2048_env.py:136-184) on levels.  max_number = 2^lvl,
 * previous_max = 2^(*prev_level).  Float expression order follows the
 * reference statement by statement. */
static double calculate_reward(int64_t score, int valid, int game_over, int lvl, int *prev_level) {
    double reward = 0;
    if (lvl < 1) lvl = 1;                        /* max(2, max_number) :141 */
    double current_level = (double)lvl;          /* log2(max_number) :144, exact */
    double bonus_progress = 0;
    if (lvl > *prev_level) {                     /* :148-150 */
        bonus_progress = (current_level - (double)*prev_level) * pow(current_level, 1.2);
        *prev_level = lvl;
    }
    if (!valid) {
        if (game_over) {
            if (lvl == 9 || lvl == 10 || lvl == 11)          /* max in [512,1024,2048] :156 */
                reward = bonus_progress + pow(current_level, 1.2);
            else
                reward -= log2((double)(((int64_t)1 << lvl) + 1)); /* :160 */
        } else {
            reward -= 0.1 * current_level;       /* :164 */
        }
    } else {
        reward = (double)score;                  /* :168 */
        if (bonus_progress > 0) reward += bonus_progress;     /* :171-172 */
        else if (bonus_progress == 0) reward += current_level * 0.05; /* :173-174 */
        if (lvl >= 9) reward += pow(current_level, 1.2) * 2;  /* :176-177 */
    }
    return normalize(reward);                    /* :181 */
}

/* stall penalty sequence (Game2048_env.py:124-125): p <- max(p * 1.1, -10), p0 = -1 */
static double stall_penalty(int idx) {
    double p = -1;
    for (int i = 0; i < idx; ++i) {
        double q = p * 1.1;
        p = q > -10 ? q : -10;
    }
    return p;
}

/* ------------------------------------------------------------------------- */
/* env steps                                                                 */
/* ------------------------------------------------------------------------- */
typedef struct {
    uint64_t board;
    uint64_t aux;   /* prev_level | cons_action<<8 | pen_idx<<16 | cons_count<<32 */
    int32_t score;  /* env.score: sum of merge scores this episode */
} env_t;

typedef struct {
    double reward;
    int32_t move_score;
    uint8_t flags; /* FLAG_VALID | FLAG_GAME_OVER | FLAG_DONE | legal_mask(new board)<<4 */
    uint8_t maxlvl;
} out_t;

/* Game2048_env.step, penalty flavour (Game2048_env.py:97-129).
 * d = {k, is4} spawn draws for the agent's move (used iff the move is valid). */
static int g_want_legal = 1; /* rollouts do not need the legal-move nibble of flags */
static void penalty_step(env_t *e, int action, int k, int is4, out_t *o) {
    int prev_level = (int)(e->aux & 0xFF);
    int cons_action = (int)((e->aux >> 8) & 0xFF);
    int pen_idx = (int)((e->aux >> 16) & 0xFF);
    uint32_t cons_count = (uint32_t)(e->aux >> 32);

    int64_t score = 0;
    int valid = move_nospawn(&e->board, action, &score);      /* :98 -> :51-60 */
    if (valid) spawn(&e->board, k, is4);                       /* :61-62 */
    int game_over = dead(e->board);                            /* :99 */
    int lvl = max_level(e->board);                             /* :100 */
    e->score += (int32_t)score;                                /* :104 */
    double reward = calculate_reward(score, valid, game_over, lvl, &prev_level); /* :107 */

    if (action == cons_action) {                               /* :110-115 */
        if (cons_count != 0xFFFFFFFFu) cons_count += 1;
    } else {
        cons_action = action;
        cons_count = 1;
        pen_idx = 0;
    }
    int done = (!valid && game_over);                          /* :117-118 */
    if (cons_count > 10) {                                     /* :121-127 */
        if (cons_count > 100) done = 1;
        if (pen_idx < PEN_SAT) pen_idx += 1;
        reward += stall_penalty(pen_idx);
    }
    e->aux = (uint64_t)prev_level | ((uint64_t)cons_action << 8) | ((uint64_t)pen_idx << 16) |
             ((uint64_t)cons_count << 32);
    o->reward = reward;
    o->move_score = (int32_t)score;
    o->flags = (uint8_t)((valid ? FLAG_VALID : 0) | (game_over ? FLAG_GAME_OVER : 0) |
                         (done ? FLAG_DONE : 0) | (g_want_legal ? legal_mask(e->board) << 4 : 0));
    o->maxlvl = (uint8_t)lvl;
}

/* Game2048_env.step, nopenalty flavour (Game2048_nopenalty_env.py:106-120) under
 * the caller protocol of mainDQL_CNN_step2.py:163-237 (the caller commits the
 * returned board as the next env.game.board).  e->board = committed board S;
 * on return e->board = moved_board M.  (k1,f1) = spawn draws of the agent's
 * move; (k2,f2) = draws of the spawn inside is_game_over when S is full and
 * some action a' is legal (the full-board quirk). */
static void nopenalty_step(env_t *e, int action, int k1, int f1, int k2, int f2, out_t *o) {
    uint64_t S = e->board, M = S;
    int64_t score = 0;
    int valid = move_nospawn(&M, action, &score);              /* :53-66 */
    if (valid) spawn(&M, k1, f1);
    int game_over = 0;                                         /* :68-78 on S */
    if (count_empty(S) == 0) {
        game_over = 1;
        for (int a = 0; a < 4; ++a) {
            uint64_t T = S;
            int64_t s2 = 0;
            int moved = move_nospawn(&T, a, &s2);
            M = T;                                             /* moved_board is overwritten by every probe */
            if (moved) {
                spawn(&M, k2, f2);
                game_over = 0;
                break;
            }
        }
    }
    int lvl = max_level(M);                                    /* :108 */
    e->score += (int32_t)score;                                /* :111 */
    double reward = (!valid && !game_over) ? -10.0 : (double)score; /* :122-128 */
    int done = game_over;                                      /* :117-118 */
    e->board = M;
    o->reward = reward;
    o->move_score = (int32_t)score;
    o->flags = (uint8_t)((valid ? FLAG_VALID : 0) | (game_over ? FLAG_GAME_OVER : 0) |
                         (done ? FLAG_DONE : 0) | (g_want_legal ? legal_mask(M) << 4 : 0));
    o->maxlvl = (uint8_t)lvl;
}

/* Game2048.__init__ (Game2048_env.py:11-14): empty board + two spawns.
 * reset() (:187-191) zeroes env.score; the penalty flavour's previous_max and
 * consecutive-action state survive (they are only set in __init__ :84-95). */
static void reset_env(env_t *e, int ka, int fa, int kb, int fb) {
    e->board = 0;
    spawn(&e->board, ka, fa);
    spawn(&e->board, kb, fb);
    e->score = 0;
}

/* ------------------------------------------------------------------------- */
/* batched entry points (array-in / array-out, same shapes as the C-ABI)     */
/* ------------------------------------------------------------------------- */

/* One env step with Philox draws: x = draws(STREAM_STEP) supplies the spawn of
 * the agent's move (x[0] -> cell, x[1] -> 2/4); the nopenalty full-board quirk
 * spawn reuses the same two draws.  The spawn cell index is
 * floor(x * n_empty / 2^32) with n_empty counted on the moved board. */
static void philox_env_step(env_t *e, int a, int flavour, uint64_t seed, uint64_t env_id, uint64_t t,
                            const uint32_t x[4], out_t *o) {
    (void)seed; (void)env_id; (void)t;
    int k1, f1, k2 = 0, f2 = 0;
    uint64_t tb = e->board; int64_t s = 0;
    move_nospawn(&tb, a, &s);
    int ne = count_empty(tb);
    k1 = ne ? draw_k(x[0], ne) : 0;
    f1 = x[1] >= IS4_THRESH;
    if (flavour == FLAVOUR_NOPENALTY && count_empty(e->board) == 0) {
        /* full-board quirk spawn: the spawn of the agent's own move is discarded in this case, so its two draws
         * x[0], x[1] are reused here (same joint distribution as the reference's fresh draws) */
        for (int a2 = 0; a2 < 4; ++a2) {
            uint64_t t2 = e->board; int64_t s2 = 0;
            if (move_nospawn(&t2, a2, &s2)) {
                int ne2 = count_empty(t2);
                k2 = ne2 ? draw_k(x[0], ne2) : 0;
                f2 = x[1] >= IS4_THRESH;
                break;
            }
        }
    }
    if (flavour == FLAVOUR_PENALTY) penalty_step(e, a, k1, f1, o);
    else nopenalty_step(e, a, k1, f1, k2, f2, o);
}
static void philox_reset(env_t *e, uint64_t seed, uint64_t env_id, uint64_t idx, uint32_t stream) {
    uint32_t x[4];
    draws(seed, env_id, idx, stream, x);
    reset_env(e, draw_k(x[0], 16), x[1] >= IS4_THRESH, draw_k(x[2], 15), x[3] >= IS4_THRESH);
}

/* replay_draws: NULL (Philox) or uint8[n][4] = {k1, is4_1, k2, is4_2}. */
ORC_API void orc_env_step(uint64_t *boards, uint64_t *aux, int32_t *score, const uint8_t *actions,
                          const uint8_t *replay_draws, double *reward, uint8_t *flags, uint8_t *maxlvl,
                          int32_t *move_score, int64_t n, int flavour, uint64_t seed, uint64_t step_idx,
                          uint64_t env_id_base) {
    g_want_legal = flags != NULL;
    for (int64_t i = 0; i < n; ++i) {
        env_t e = {boards[i], aux ? aux[i] : AUX_INIT, score ? score[i] : 0};
        out_t o;
        int a = actions[i] & 3;
        if (replay_draws) {
            const uint8_t *d = replay_draws + 4 * i;
            if (flavour == FLAVOUR_PENALTY) penalty_step(&e, a, d[0], d[1], &o);
            else nopenalty_step(&e, a, d[0], d[1], d[2], d[3], &o);
        } else {
            uint32_t x[4];
            draws(seed, env_id_base + (uint64_t)i, step_idx, STREAM_STEP, x);
            philox_env_step(&e, a, flavour, seed, env_id_base + (uint64_t)i, step_idx, x, &o);
        }
        boards[i] = e.board;
        if (aux) aux[i] = e.aux;
        if (score) score[i] = e.score;
        if (reward) reward[i] = o.reward;
        if (flags) flags[i] = o.flags;
        if (maxlvl) maxlvl[i] = o.maxlvl;
        if (move_score) move_score[i] = o.move_score;
    }
}

/* mask: NULL = all envs.  replay_draws: NULL (Philox, stream RESET keyed by
 * (env id, episode_idx)) or uint8[n][4] = {ka, is4_a, kb, is4_b}. */
ORC_API void orc_env_reset(uint64_t *boards, int32_t *score, const uint8_t *mask, const uint8_t *replay_draws,
                           int64_t n, uint64_t seed, uint64_t episode_idx, uint64_t env_id_base) {
    for (int64_t i = 0; i < n; ++i) {
        if (mask && !mask[i]) continue;
        env_t e = {0, 0, 0};
        if (replay_draws) {
            reset_env(&e, replay_draws[4 * i], replay_draws[4 * i + 1], replay_draws[4 * i + 2],
                      replay_draws[4 * i + 3]);
        } else {
            philox_reset(&e, seed, env_id_base + (uint64_t)i, episode_idx, STREAM_RESET);
        }
        boards[i] = e.board;
        if (score) score[i] = 0;
    }
}

ORC_API void orc_legal_mask(const uint64_t *boards, uint8_t *out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) out[i] = (uint8_t)legal_mask(boards[i]);
}
ORC_API void orc_dead(const uint64_t *boards, uint8_t *out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) out[i] = (uint8_t)dead(boards[i]);
}
/* out_moved[n], out_score[n], boards updated in place, no spawn */
ORC_API void orc_move(uint64_t *boards, const uint8_t *actions, uint8_t *out_moved, int32_t *out_score, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
        int64_t s = 0;
        out_moved[i] = (uint8_t)move_nospawn(&boards[i], actions[i] & 3, &s);
        out_score[i] = (int32_t)s;
    }
}
/* the 65,536-entry row table the CUDA side must reproduce: result row and the
 * two merged levels packed (hi nibble >= lo nibble, 0 = none) */
ORC_API void orc_row_table(uint16_t *result, uint8_t *merged) {
    for (int row = 0; row < 65536; ++row) {
        int in[4], out[4], mg[2];
        for (int c = 0; c < 4; ++c) in[c] = (row >> (4 * c)) & 0xF;
        row_left(in, out, mg);
        int o = 0;
        for (int c = 0; c < 4; ++c) o |= out[c] << (4 * c);
        result[row] = (uint16_t)o;
        int hi = mg[0] > mg[1] ? mg[0] : mg[1], lo = mg[0] > mg[1] ? mg[1] : mg[0];
        merged[row] = (uint8_t)((hi << 4) | lo);
    }
}
ORC_API double orc_calculate_reward(int64_t score, int valid, int game_over, int lvl, int *prev_level) {
    return calculate_reward(score, valid, game_over, lvl, prev_level);
}
ORC_API double orc_stall_penalty(int idx) { return stall_penalty(idx); }

/* raw tile values (np.int64[n][16], the reference's board format) <-> packed */
ORC_API int64_t orc_pack_i64(const int64_t *tiles, uint64_t *boards, int64_t n) {
    int64_t bad = 0;
    for (int64_t i = 0; i < n; ++i) {
        uint64_t b = 0;
        for (int j = 0; j < 16; ++j) {
            int64_t v = tiles[16 * i + j];
            int lvl = 0;
            if (v != 0) {
                while (lvl < 16 && ((int64_t)1 << lvl) != v) ++lvl;
                if (lvl < 1 || lvl > 15) { ++bad; lvl = 0; }
            }
            b |= (uint64_t)lvl << (4 * j);
        }
        boards[i] = b;
    }
    return bad;
}
ORC_API void orc_unpack_i64(const uint64_t *boards, int64_t *tiles, int64_t n) {
    for (int64_t i = 0; i < n; ++i)
        for (int j = 0; j < 16; ++j) {
            int lvl = (int)((boards[i] >> (4 * j)) & 0xF);
            tiles[16 * i + j] = lvl ? ((int64_t)1 << lvl) : 0;
        }
}

/* DQNAgent.encode_state (Dqn8TestNOPERCNN.py:271-277): log2 per cell (0 -> 0),
 * one-hot depth 16, laid out [batch, level, row, col] float32. */
ORC_API void orc_encode_onehot(const uint64_t *boards, float *out, int64_t n) {
    memset(out, 0, (size_t)n * 256 * sizeof(float));
    for (int64_t i = 0; i < n; ++i)
        for (int j = 0; j < 16; ++j) {
            int lvl = (int)((boards[i] >> (4 * j)) & 0xF);
            out[i * 256 + lvl * 16 + j] = 1.0f;
        }
}

/* ------------------------------------------------------------------------- */
/* Q-table (CPU): open addressing, key = packed board, rows of 4 values.     */
/* The reference's defaultdict(lambda: np.zeros(4)) (main.py:16) inserts a    */
/* zero row on every read; lookups here insert too, so sizes are comparable. */
/* ------------------------------------------------------------------------- */
typedef struct {
    uint64_t *keys;
    void *rows;     /* double[cap][4] or float[cap][4] */
    uint64_t cap;   /* power of two */
    uint64_t size;
    int f32;
} qtab_t;

static uint64_t mix64(uint64_t x) { /* splitmix64 finaliser */
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}
ORC_API qtab_t *orc_qtab_new(uint64_t cap, int f32) {
    qtab_t *t = (qtab_t *)calloc(1, sizeof *t);
    t->cap = cap; t->f32 = f32;
    t->keys = (uint64_t *)calloc(cap, sizeof(uint64_t));
    t->rows = calloc(cap, 4 * (f32 ? sizeof(float) : sizeof(double)));
    return t;
}
ORC_API void orc_qtab_free(qtab_t *t) { if (t) { free(t->keys); free(t->rows); free(t); } }
ORC_API uint64_t orc_qtab_size(const qtab_t *t) { return t->size; }
static int64_t qtab_slot(qtab_t *t, uint64_t key, int insert) {
    uint64_t m = t->cap - 1, h = mix64(key) & m;
    for (uint64_t p = 0; p < t->cap; ++p, h = (h + 1) & m) {
        if (t->keys[h] == key) return (int64_t)h;
        if (t->keys[h] == 0) {
            if (!insert || t->size * 10 >= t->cap * 9) return -1;
            t->keys[h] = key; t->size++;
            return (int64_t)h;
        }
    }
    return -1;
}
ORC_API int64_t orc_qtab_export(const qtab_t *t, uint64_t *keys, double *rows, int64_t max_out) {
    int64_t m = 0;
    for (uint64_t h = 0; h < t->cap; ++h)
        if (t->keys[h]) {
            if (m < max_out) {
                keys[m] = t->keys[h];
                for (int a = 0; a < 4; ++a)
                    rows[4 * m + a] = t->f32 ? (double)((float *)t->rows)[4 * h + a] : ((double *)t->rows)[4 * h + a];
            }
            ++m;
        }
    return m;
}
ORC_API int orc_qtab_get(qtab_t *t, uint64_t key, double *row4) {
    int64_t h = qtab_slot(t, key, 0);
    for (int a = 0; a < 4; ++a)
        row4[a] = h < 0 ? 0.0 : (t->f32 ? (double)((float *)t->rows)[4 * h + a] : ((double *)t->rows)[4 * h + a]);
    return h >= 0;
}

/* np.argmax: first maximum (main.py:38, :41) */
static int argmax4d(const double *q) {
    int b = 0;
    for (int a = 1; a < 4; ++a) if (q[a] > q[b]) b = a;
    return b;
}
static int argmax4f(const float *q) {
    int b = 0;
    for (int a = 1; a < 4; ++a) if (q[a] > q[b]) b = a;
    return b;
}

/* QLearningAgent.update_q_value (main.py:40-43), strictly sequential, float64:
 * the reference's own order.  n transitions applied one after another. */
ORC_API void orc_q_update_seq_f64(qtab_t *t, const uint64_t *s, const uint8_t *a, const double *r,
                                  const uint64_t *s2, const uint8_t *done, int64_t n, double lr, double gamma) {
    double *rows = (double *)t->rows;
    for (int64_t i = 0; i < n; ++i) {
        int64_t h2 = qtab_slot(t, s2[i], 1);
        int best = argmax4d(&rows[4 * h2]);
        double target = r[i] + (gamma * rows[4 * h2 + best] * (double)(1 - (done[i] ? 1 : 0)));
        int64_t h = qtab_slot(t, s[i], 1);
        rows[4 * h + a[i]] += lr * (target - rows[4 * h + a[i]]);
    }
}

/* choose_action (main.py:34-38) followed by update_q_value (main.py:40-43) per
 * transition, float64: explore[i] says whether random.random() < epsilon held
 * (then the action is the recorded randint), else the action is the first-max
 * argmax of the current row.  actions_out lets a test check the agent's
 * greedy choices against the reference's. */
ORC_API void orc_q_replay_agent_f64(qtab_t *t, const uint64_t *s, const uint8_t *explore, const uint8_t *rand_a,
                                    const double *r, const uint64_t *s2, const uint8_t *done, uint8_t *actions_out,
                                    int64_t n, double lr, double gamma) {
    double *rows = (double *)t->rows;
    for (int64_t i = 0; i < n; ++i) {
        int64_t h = qtab_slot(t, s[i], 1);
        int a = explore[i] ? rand_a[i] : argmax4d(&rows[4 * h]);
        actions_out[i] = (uint8_t)a;
        int64_t h2 = qtab_slot(t, s2[i], 1);
        int best = argmax4d(&rows[4 * h2]);
        double target = r[i] + (gamma * rows[4 * h2 + best] * (double)(1 - (done[i] ? 1 : 0)));
        rows[4 * h + a] += lr * (target - rows[4 * h + a]);
    }
}

/* float32 arithmetic helpers: every operation rounds to float32 on its own
 * (no FMA contraction, no excess precision) -- the CUDA side is compiled with
 * -fmad=false and must produce the same bits.
 * update_q_value (main.py:40-43) split in two:
 *   target = r + (done ? 0 : gamma * best_next)           (:41-42)
 *   q      = q + lr * (target - q)                         (:43) */
static float f32_target(float gamma, float r, float best_next, int done) {
    volatile float g = gamma * best_next;
    volatile float gn = done ? 0.0f : g;
    volatile float target = r + gn;
    return target;
}
static float f32_apply(float q, float lr, float target) {
    volatile float diff = target - q;
    volatile float d = lr * diff;
    volatile float nq = q + d;
    return nq;
}
typedef struct { uint64_t key; int64_t idx; } sortrec_t;
static int sortrec_cmp(const void *pa, const void *pb) {
    const sortrec_t *a = (const sortrec_t *)pa, *b = (const sortrec_t *)pb;
    if (a->key != b->key) return a->key < b->key ? -1 : 1;
    return a->idx < b->idx ? -1 : (a->idx > b->idx);
}
/* Transitions that hit the same (state, action) in one batch are applied one after
 * another in ascending i -- q = q + lr * (target_i - q) -- exactly what the
 * reference's sequential loop does to Q[s][a]; only the bootstrap values inside
 * target_i come from the snapshot.  (Summing the deltas instead would multiply the
 * step size by the number of duplicates and diverge once that exceeds 2 / lr.) */
static void apply_targets_sorted(float *rows, const int64_t *slot, const uint8_t *a, const float *target, float lr,
                                 int64_t n) {
    sortrec_t *rec = (sortrec_t *)malloc((size_t)(n ? n : 1) * sizeof *rec);
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i)
        if (slot[i] >= 0) { rec[m].key = (uint64_t)slot[i] * 4 + a[i]; rec[m].idx = i; ++m; }
    qsort(rec, (size_t)m, sizeof *rec, sortrec_cmp);
    for (int64_t i = 0; i < m; ++i) rows[rec[i].key] = f32_apply(rows[rec[i].key], lr, target[rec[i].idx]);
    free(rec);
}

/* Batched synchronous update, float32 rows:
 * every transition bootstraps on the snapshot at batch start,
 *   target_i = r_i + (done_i ? 0 : gamma * max_a Q_snap[s2_i][a])
 * then each (s, a) receives its targets in ascending i: q = q + lr * (target_i - q).
 * N = 1 is exactly update_q_value (main.py:40-43) in float32. */
ORC_API void orc_q_update_batch_f32(qtab_t *t, const uint64_t *s, const uint8_t *a, const float *r,
                                    const uint64_t *s2, const uint8_t *done, int64_t n, float lr, float gamma) {
    float *rows = (float *)t->rows;
    int64_t *slot = (int64_t *)malloc((size_t)(n ? n : 1) * sizeof(int64_t));
    float *target = (float *)malloc((size_t)(n ? n : 1) * sizeof(float));
    for (int64_t i = 0; i < n; ++i) { /* inserts create zero rows only: they change no value */
        qtab_slot(t, s2[i], 1);
        slot[i] = qtab_slot(t, s[i], 1);
    }
    for (int64_t i = 0; i < n; ++i) {
        int64_t h2 = qtab_slot(t, s2[i], 0);
        float best = h2 < 0 ? 0.0f : rows[4 * h2 + argmax4f(&rows[4 * h2])];
        target[i] = f32_target(gamma, r[i], best, done[i] != 0);
    }
    apply_targets_sorted(rows, slot, a, target, lr, n);
    free(target); free(slot);
}

/* (key, action, target) records, e.g. received from other ranks: look the key up
 * (insert if absent) and apply as above. */
ORC_API void orc_q_apply_targets_f32(qtab_t *t, const uint64_t *keys, const uint8_t *a, const float *target, float lr,
                                     int64_t n) {
    int64_t *slot = (int64_t *)malloc((size_t)(n ? n : 1) * sizeof(int64_t));
    for (int64_t i = 0; i < n; ++i) slot[i] = qtab_slot(t, keys[i], 1);
    apply_targets_sorted((float *)t->rows, slot, a, target, lr, n);
    free(slot);
}

/* ------------------------------------------------------------------------- */
/* fused rollouts (the framework's batched drivers of the hot loop,          */
/* main.py:80-101 run for N envs at once)                                    */
/* ------------------------------------------------------------------------- */
/* counters: [0] env steps, [1] valid moves, [2] finished episodes, [3] sum of
 * move scores, [4] max level seen, [5] sum of trunc(reward * 2^20) (an
 * order-independent integer checksum of the float64 rewards), [6] table
 * inserts, [7] dropped inserts (table full), [8] lost updates (always 0 here:
 * the sequential oracle never races). */
enum { C_STEPS, C_VALID, C_EPISODES, C_SCORE, C_MAXLVL, C_REWARD_FX, C_INSERTS, C_DROPPED, C_LOST, C_N };
static void count_step(int64_t *c, const out_t *o) {
    c[C_STEPS] += 1;
    c[C_VALID] += (o->flags & FLAG_VALID) != 0;
    c[C_EPISODES] += (o->flags & FLAG_DONE) != 0;
    c[C_SCORE] += o->move_score;
    if (o->maxlvl > c[C_MAXLVL]) c[C_MAXLVL] = o->maxlvl;
    c[C_REWARD_FX] += (int64_t)(o->reward * 1048576.0);
}

/* uniform-random policy: action = x[3] >> 30 of the step's Philox draw; on done
 * the env is reset in place with draws(STREAM_AUTORESET) of the same step. */
ORC_API void orc_rollout_random(uint64_t *boards, uint64_t *aux, int32_t *score, int64_t n, int64_t k_steps,
                                int flavour, uint64_t seed, uint64_t step_base, uint64_t env_id_base,
                                int64_t *counters) {
    g_want_legal = 0;
    for (int64_t i = 0; i < n; ++i) {
        env_t e = {boards[i], aux ? aux[i] : AUX_INIT, score ? score[i] : 0};
        uint64_t id = env_id_base + (uint64_t)i;
        for (int64_t k = 0; k < k_steps; ++k) {
            uint64_t t = step_base + (uint64_t)k;
            uint32_t x[4];
            out_t o;
            draws(seed, id, t, STREAM_STEP, x);
            philox_env_step(&e, (int)(x[3] >> 30), flavour, seed, id, t, x, &o);
            count_step(counters, &o);
            if (o.flags & FLAG_DONE) philox_reset(&e, seed, id, t, STREAM_AUTORESET);
        }
        boards[i] = e.board;
        if (aux) aux[i] = e.aux;
        if (score) score[i] = e.score;
    }
}

static int64_t qtab_slot_counted(qtab_t *t, uint64_t key, int64_t *counters) {
    uint64_t before = t->size;
    int64_t h = qtab_slot(t, key, 1);
    if (h < 0) counters[C_DROPPED] += 1;
    else if (t->size != before) counters[C_INSERTS] += 1;
    return h;
}

/* choose_action (main.py:34-38) with Philox draws: explore iff x[2] < eps_thresh
 * (eps_thresh = floor(eps * 2^32), as a 64-bit compare so eps = 1 always
 * explores), random action = x[3] >> 30, else first-max argmax of the row. */
static int choose_action_f32(const float *row, const uint32_t x[4], uint64_t eps_thresh) {
    if ((uint64_t)x[2] < eps_thresh) return (int)(x[3] >> 30);
    return argmax4f(row);
}
ORC_API void orc_choose_action(qtab_t *t, const uint64_t *boards, int64_t n, uint64_t eps_thresh, uint64_t seed,
                               uint64_t step_idx, uint64_t env_id_base, uint8_t *actions) {
    static const float zero[4] = {0, 0, 0, 0};
    for (int64_t i = 0; i < n; ++i) {
        uint32_t x[4];
        draws(seed, env_id_base + (uint64_t)i, step_idx, STREAM_STEP, x);
        int64_t h = qtab_slot(t, boards[i], 1);
        actions[i] = (uint8_t)choose_action_f32(h < 0 ? zero : (float *)t->rows + 4 * h, x, eps_thresh);
    }
}

/* Sequential-semantics epsilon-greedy Q-learning rollout (float32 table):
 * time-major, env-minor; each env's transition is applied to the table
 * immediately (the reference loop main.py:91-101 for N = 1).  The GPU's
 * asynchronous atomic mode is bit-comparable with this only at N = 1. */
ORC_API void orc_rollout_qlearn_seq(uint64_t *boards, uint64_t *aux, int32_t *score, qtab_t *t, int64_t n,
                                    int64_t k_steps, int flavour, float lr, float gamma, uint64_t eps_thresh,
                                    uint64_t seed, uint64_t step_base, uint64_t env_id_base, int64_t *counters) {
    float *rows = (float *)t->rows;
    static const float zero[4] = {0, 0, 0, 0};
    g_want_legal = 0;
    for (int64_t k = 0; k < k_steps; ++k) {
        uint64_t step = step_base + (uint64_t)k;
        for (int64_t i = 0; i < n; ++i) {
            env_t e = {boards[i], aux ? aux[i] : AUX_INIT, score ? score[i] : 0};
            uint64_t id = env_id_base + (uint64_t)i;
            uint32_t x[4];
            out_t o;
            draws(seed, id, step, STREAM_STEP, x);
            int64_t h = qtab_slot_counted(t, e.board, counters);
            int a = choose_action_f32(h < 0 ? zero : rows + 4 * h, x, eps_thresh);
            philox_env_step(&e, a, flavour, seed, id, step, x, &o);
            count_step(counters, &o);
            int64_t h2 = qtab_slot_counted(t, e.board, counters);
            float best = h2 < 0 ? 0.0f : rows[4 * h2 + argmax4f(rows + 4 * h2)];
            if (h >= 0)
                rows[4 * h + a] = f32_apply(rows[4 * h + a], lr,
                                            f32_target(gamma, (float)o.reward, best, (o.flags & FLAG_DONE) != 0));
            if (o.flags & FLAG_DONE) {
                philox_reset(&e, seed, id, step, STREAM_AUTORESET);
                qtab_slot_counted(t, e.board, counters); /* state = env.reset() is read at once (main.py:81-82, :92) */
            }
            boards[i] = e.board;
            if (aux) aux[i] = e.aux;
            if (score) score[i] = e.score;
        }
    }
}

/* One synchronous batched Q-learning step over n envs:
 * all envs choose from and bootstrap on the table snapshot at step start;
 * the targets are applied afterwards, each (state, action) receiving its
 * targets in ascending env order.  Bit-comparable with the GPU's deterministic mode for
 * any n and any sharding.  Optionally exports the transition records. */
ORC_API void orc_qlearn_step_sync(uint64_t *boards, uint64_t *aux, int32_t *score, qtab_t *t, int64_t n,
                                  int flavour, float lr, float gamma, uint64_t eps_thresh, uint64_t seed,
                                  uint64_t step, uint64_t env_id_base, int64_t *counters,
                                  uint64_t *rec_key, uint8_t *rec_action, float *rec_target, int apply) {
    float *rows = (float *)t->rows;
    static const float zero[4] = {0, 0, 0, 0};
    g_want_legal = 0;
    int64_t *slot = (int64_t *)malloc((size_t)(n ? n : 1) * sizeof(int64_t));
    uint8_t *act = (uint8_t *)malloc((size_t)(n ? n : 1));
    float *target = (float *)malloc((size_t)(n ? n : 1) * sizeof(float));
    for (int64_t i = 0; i < n; ++i) {
        env_t e = {boards[i], aux ? aux[i] : AUX_INIT, score ? score[i] : 0};
        uint64_t id = env_id_base + (uint64_t)i;
        uint32_t x[4];
        out_t o;
        draws(seed, id, step, STREAM_STEP, x);
        int64_t h = qtab_slot_counted(t, e.board, counters);
        int a = choose_action_f32(h < 0 ? zero : rows + 4 * h, x, eps_thresh);
        if (rec_key) rec_key[i] = e.board;
        philox_env_step(&e, a, flavour, seed, id, step, x, &o);
        count_step(counters, &o);
        int64_t h2 = qtab_slot_counted(t, e.board, counters);
        float best = h2 < 0 ? 0.0f : rows[4 * h2 + argmax4f(rows + 4 * h2)];
        slot[i] = h; act[i] = (uint8_t)a;
        target[i] = f32_target(gamma, (float)o.reward, best, (o.flags & FLAG_DONE) != 0);
        if (rec_action) rec_action[i] = (uint8_t)a;
        if (rec_target) rec_target[i] = target[i];
        if (o.flags & FLAG_DONE) philox_reset(&e, seed, id, step, STREAM_AUTORESET);
        boards[i] = e.board;
        if (aux) aux[i] = e.aux;
        if (score) score[i] = e.score;
    }
    /* apply = 0: emit the records only (the cross-rank exchange applies them) */
    if (apply) apply_targets_sorted(rows, slot, act, target, lr, n);
    free(target); free(act); free(slot);
}

/* ------------------------------------------------------------------------- */
/* multi-threaded drivers: the CPU baseline bench.py times (all host cores). */
/* Env shards are independent; the Q-learning variant gives every thread its  */
/* own table (replicas), i.e. the most generous CPU figure.                  */
/* ------------------------------------------------------------------------- */
typedef struct {
    uint64_t *boards, *aux; int32_t *score; int64_t n, k_steps; int flavour; float lr, gamma;
    uint64_t eps_thresh, seed, step_base, env_id_base, table_cap; int qlearn; int64_t counters[C_N];
    qtab_t **tables;   /* optional: one persistent table per thread (tables[w]) instead of a fresh one per call */
    int worker;
} mt_job_t;
static void *mt_worker(void *arg) {
    mt_job_t *j = (mt_job_t *)arg;
    if (j->qlearn) {
        qtab_t *t = j->tables ? j->tables[j->worker] : orc_qtab_new(j->table_cap, 1);
        orc_rollout_qlearn_seq(j->boards, j->aux, j->score, t, j->n, j->k_steps, j->flavour, j->lr, j->gamma,
                               j->eps_thresh, j->seed, j->step_base, j->env_id_base, j->counters);
        if (!j->tables) orc_qtab_free(t);
    } else {
        orc_rollout_random(j->boards, j->aux, j->score, j->n, j->k_steps, j->flavour, j->seed, j->step_base,
                           j->env_id_base, j->counters);
    }
    return NULL;
}
static void mt_run(mt_job_t proto, int64_t n, int threads, int64_t *counters) {
    if (threads < 1) threads = 1;
    mt_job_t *jobs = (mt_job_t *)calloc((size_t)threads, sizeof *jobs);
    pthread_t *tid = (pthread_t *)calloc((size_t)threads, sizeof *tid);
    for (int w = 0; w < threads; ++w) {
        int64_t lo = n * w / threads, hi = n * (w + 1) / threads;
        jobs[w] = proto;
        jobs[w].boards = proto.boards + lo;
        jobs[w].aux = proto.aux ? proto.aux + lo : NULL;
        jobs[w].score = proto.score ? proto.score + lo : NULL;
        jobs[w].n = hi - lo;
        jobs[w].env_id_base = proto.env_id_base + (uint64_t)lo;
        jobs[w].worker = w;
        memset(jobs[w].counters, 0, sizeof jobs[w].counters);
        pthread_create(&tid[w], NULL, mt_worker, &jobs[w]);
    }
    for (int w = 0; w < threads; ++w) {
        pthread_join(tid[w], NULL);
        for (int c = 0; c < C_N; ++c) {
            if (c == C_MAXLVL) { if (jobs[w].counters[c] > counters[c]) counters[c] = jobs[w].counters[c]; }
            else counters[c] += jobs[w].counters[c];
        }
    }
    free(tid); free(jobs);
}
ORC_API void orc_rollout_random_mt(uint64_t *boards, uint64_t *aux, int32_t *score, int64_t n, int64_t k_steps,
                                   int flavour, uint64_t seed, uint64_t step_base, uint64_t env_id_base,
                                   int64_t *counters, int threads) {
    mt_job_t p = {boards, aux, score, n, k_steps, flavour, 0, 0, 0, seed, step_base, env_id_base, 0, 0, {0}, NULL, 0};
    mt_run(p, n, threads, counters);
}
ORC_API void orc_rollout_qlearn_mt(uint64_t *boards, uint64_t *aux, int32_t *score, int64_t n, int64_t k_steps,
                                   int flavour, float lr, float gamma, uint64_t eps_thresh, uint64_t seed,
                                   uint64_t step_base, uint64_t env_id_base, int64_t *counters, uint64_t table_cap,
                                   int threads) {
    mt_job_t p = {boards, aux, score, n, k_steps, flavour, lr, gamma, eps_thresh, seed, step_base, env_id_base,
                  table_cap, 1, {0}, NULL, 0};
    mt_run(p, n, threads, counters);
}
/* the same with one persistent table per thread (tables[0 .. threads)), kept across calls: the training run bench.py
 * times as the CPU baseline (thread w plays envs [n w / threads, n (w + 1) / threads) against tables[w]) */
ORC_API void orc_rollout_qlearn_mt_tables(uint64_t *boards, uint64_t *aux, int32_t *score, int64_t n, int64_t k_steps,
                                          int flavour, float lr, float gamma, uint64_t eps_thresh, uint64_t seed,
                                          uint64_t step_base, uint64_t env_id_base, int64_t *counters,
                                          qtab_t **tables, int threads) {
    mt_job_t p = {boards, aux, score, n, k_steps, flavour, lr, gamma, eps_thresh, seed, step_base, env_id_base,
                  0, 1, {0}, tables, 0};
    mt_run(p, n, threads, counters);
}
