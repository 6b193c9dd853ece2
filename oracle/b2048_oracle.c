/*
 * b2048_oracle.c -- CPU ORACLE for the 2048NN playout path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is a plain-C restatement of the reference's board engine and fixed-order
 * playout / root-search algorithms (fqjin/2048NN: board.py, mcts.py).  It exists so that
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs have
 * something to check the CUDA path against and to time on the host.  Nothing in the
 * product package (2048nn_b200/) may import, link or call it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function below against
 * golden vectors that oracle/make_golden.py produced by importing and running the
 * UNMODIFIED Python reference in the build container (row tables, move vectors,
 * generate_tile, injected-spawn play_fixed games, mcts_fixed* results), plus the
 * known answers of SURVEY.md Appendix B.
 *
 * Reference lines restated are cited per function as board.py:L / mcts.py:L.
 *
 * Spawn randomness: the reference draws from CPython's MT19937 `randrange`, which cannot be
 * reproduced per game on a GPU.  Both sides therefore use the injected-spawn protocol of
 * every game owns a counter-based stream of 32-bit words and a draw with
 * bound n is (word * n) >> 32.  The stream definition (orc_rng_*) is part of the boundary
 * spec (include/b2048.h), restated here independently of the CUDA source.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>
#include <unistd.h>

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint16_t u16;
typedef uint8_t u8;

/* ------------------------------------------------------------------------------------ */
/* Row tables: merge_table / merge_table_rev  (board.py:104-148)                          */
/* ------------------------------------------------------------------------------------ */
static u16 g_final[2][65536];
static u32 g_score[2][65536];
static u8 g_moved[2][65536];
static int g_ready = 0;

/* board.py:104-132  move_row(x, rev): slide+merge one 4-nibble row toward index 0
 * (or index 3 when rev).  A tile merges at most once; merging exponent t gives t+1 and
 * adds 2^(t+1) to the score; 15+15 wraps to an empty cell but still scores 65536. */
static void row_slide(u32 x, int rev, u16 *final_out, u32 *score_out, u8 *moved_out) {
    int cell[4], out[4] = {0, 0, 0, 0};
    int n_out = 0, pending = 0;
    u32 score = 0;
    for (int i = 0; i < 4; i++) cell[i] = (x >> (4 * (rev ? 3 - i : i))) & 0xF;
    for (int i = 0; i < 4; i++) {
        int t = cell[i];
        if (!t) continue;
        if (pending == t) {
            int nt = t + 1;
            score += 1u << nt;
            out[n_out++] = (nt == 16) ? 0 : nt;
            pending = 0;
        } else {
            if (pending) out[n_out++] = pending;
            pending = t;
        }
    }
    if (pending) out[n_out++] = pending;
    /* Note: an annihilated pair appended a 0 above, exactly as the reference's list does. */
    u32 f = 0;
    for (int i = 0; i < 4; i++) f |= (u32)out[i] << (4 * (rev ? 3 - i : i));
    *final_out = (u16)f;
    *score_out = score;
    *moved_out = (f != x);
}

void orc_init(void) {
    if (g_ready) return;
    for (u32 r = 0; r < 65536; r++) {
        row_slide(r, 0, &g_final[0][r], &g_score[0][r], &g_moved[0][r]);
        row_slide(r, 1, &g_final[1][r], &g_score[1][r], &g_moved[1][r]);
    }
    g_ready = 1;
}

/* copies the six 65536-entry tables out (fwd then rev) */
void orc_tables(u16 *fin_fwd, u16 *fin_rev, u32 *sc_fwd, u32 *sc_rev, u8 *mv_fwd, u8 *mv_rev) {
    orc_init();
    memcpy(fin_fwd, g_final[0], sizeof g_final[0]);
    memcpy(fin_rev, g_final[1], sizeof g_final[1]);
    memcpy(sc_fwd, g_score[0], sizeof g_score[0]);
    memcpy(sc_rev, g_score[1], sizeof g_score[1]);
    memcpy(mv_fwd, g_moved[0], sizeof g_moved[0]);
    memcpy(mv_rev, g_moved[1], sizeof g_moved[1]);
}

/* ------------------------------------------------------------------------------------ */
/* Bit tricks (board.py:23-65)                                                           */
/* ------------------------------------------------------------------------------------ */
/* board.py:23-33 get_tiles: 16 nibbles, least significant first */
void orc_get_tiles(u64 x, u8 *tiles16) {
    for (int k = 0; k < 16; k++) tiles16[k] = (x >> (4 * k)) & 0xF;
}

/* board.py:46-54 transpose of the 4x4 nibble matrix (two mask/shift stages) */
u64 orc_transpose(u64 x) {
    u64 keep = x & 0xF0F00F0FF0F00F0FULL;
    u64 up = x & 0x0000F0F00000F0F0ULL;
    u64 dn = x & 0x0F0F00000F0F0000ULL;
    u64 a = keep | (up << 12) | (dn >> 12);
    u64 keep2 = a & 0xFF00FF0000FF00FFULL;
    u64 hi = a & 0x00FF00FF00000000ULL;
    u64 lo = a & 0x00000000FF00FF00ULL;
    return keep2 | (hi >> 24) | (lo << 24);
}

/* board.py:57-65 countzero: number of empty nibbles MOD 16 (16 empties -> 0) */
int orc_countzero(u64 x) {
    x |= (x >> 2) & 0x3333333333333333ULL;
    x |= (x >> 1);
    x = ~x & 0x1111111111111111ULL;
    x += x >> 32;
    x += x >> 16;
    x += x >> 8;
    x += x >> 4;
    return (int)(x & 0xF);
}

/* ------------------------------------------------------------------------------------ */
/* move (board.py:151-189)                                                               */
/* ------------------------------------------------------------------------------------ */
void orc_move(u64 x, int direction, u64 *board_out, u32 *score_out, u8 *moved_out) {
    orc_init();
    int t = (direction & 2) ? 1 : 0;     /* board.py:169-172: bit 1 -> reverse table */
    int vertical = direction & 1;        /* board.py:174: bit 0 -> transpose before/after */
    u64 src = vertical ? orc_transpose(x) : x;
    u64 fin = 0;
    u32 score = 0;
    u8 moved = 0;
    for (int sh = 0; sh < 64; sh += 16) {
        u32 r = (u32)(src >> sh) & 0xFFFF;
        fin |= (u64)g_final[t][r] << sh;
        score += g_score[t][r];
        moved |= g_moved[t][r];
    }
    *board_out = vertical ? orc_transpose(fin) : fin;
    *score_out = score;
    *moved_out = moved;
}

void orc_move_batch(const u64 *x, const u8 *dirs, int dir_scalar, u64 *b, u32 *s, u8 *m, long n) {
    orc_init();
    for (long i = 0; i < n; i++) orc_move(x[i], dirs ? dirs[i] : dir_scalar, &b[i], &s[i], &m[i]);
}

/* ------------------------------------------------------------------------------------ */
/* Injected spawn stream (boundary spec; include/b2048.h "RNG protocol")                 */
/* ------------------------------------------------------------------------------------ */
#define ORC_GOLDEN 0x9E3779B97F4A7C15ULL
#define ORC_INIT_DRAWS 8 /* 64-bit draws (=16 words) reserved for generate_init_tiles */

static inline u64 mix64(u64 z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
u64 orc_rng_key(u64 seed, u64 game_id) {
    return mix64(mix64(seed + ORC_GOLDEN) ^ (game_id * 0xD1342543DE82EF95ULL + 1ULL));
}
u64 orc_rng_draw(u64 key, u64 i) { return mix64(key + (i + 1ULL) * ORC_GOLDEN); }
/* word w of the game's 32-bit stream: even words are the high half of draw w/2 */
u32 orc_rng_word(u64 key, u64 w) {
    u64 d = orc_rng_draw(key, w >> 1);
    return (w & 1) ? (u32)d : (u32)(d >> 32);
}
static inline u32 bounded(u32 word, u32 n) { return (u32)(((u64)word * n) >> 32); }

/* board.py:85-101 generate_tile with the two randrange results replaced by words:
 * position = randrange(countzero) first, then randrange(10) != 0 -> 2-tile else 4-tile.
 * Returns 0 and sets *err when countzero(board) == 0 (the reference raises ValueError
 * from randrange(0): full board, or the 16-empties wrap). */
u64 orc_generate_tile(u64 board, u32 w_pos, u32 w_ten, int *err) {
    int nz = orc_countzero(board);
    if (nz == 0) {
        if (err) *err = 1;
        return 0;
    }
    int position = (int)bounded(w_pos, (u32)nz);
    u64 scan = board, tile = 1;
    for (;;) {
        if ((scan & 0xF) == 0) {
            if (position == 0) break;
            position--;
        }
        scan >>= 4;
        tile <<= 4;
    }
    return bounded(w_ten, 10) ? (board | tile) : (board | (tile << 1));
}

/* board.py:68-82 generate_init_tiles, words consumed in the reference's call order:
 * pos1, pos2 (redrawn while equal), tile1, tile2.  The stream reserves 16 words for this;
 * if 13 redraws all collide (p = 16^-13) pos2 falls back to (pos1+1)%16 on both sides. */
u64 orc_init_board(u64 key) {
    u64 w = 0;
    u32 pos1 = bounded(orc_rng_word(key, w++), 16);
    u32 pos2 = bounded(orc_rng_word(key, w++), 16);
    while (pos1 == pos2) {
        if (w >= 2 * ORC_INIT_DRAWS - 2) {
            pos2 = (pos1 + 1) & 15;
            break;
        }
        pos2 = bounded(orc_rng_word(key, w++), 16);
    }
    u64 t1 = bounded(orc_rng_word(key, w++), 10) ? (1ULL << (4 * pos1)) : (1ULL << (4 * pos1 + 1));
    u64 t2 = bounded(orc_rng_word(key, w++), 10) ? (1ULL << (4 * pos2)) : (1ULL << (4 * pos2 + 1));
    return t1 | t2;
}

/* k-th spawn of a game (k counts accepted moves from 0): draw ORC_INIT_DRAWS + k */
static inline u64 spawn_k(u64 board, u64 key, u64 k, int *err) {
    u64 d = orc_rng_draw(key, ORC_INIT_DRAWS + k);
    return orc_generate_tile(board, (u32)(d >> 32), (u32)d, err);
}
u64 orc_spawn(u64 board, u64 key, u64 k, int *err) { return spawn_k(board, key, k, err); }

/* ------------------------------------------------------------------------------------ */
/* One game with a fixed priority order (board.py:192-225 play_fixed; the per-game body  */
/* of BoardArray.move_batch board.py:246-277 iterated to the end)                        */
/* ------------------------------------------------------------------------------------ */
/* start == 0 means "no board given" (board.py:205 `if not board`), i.e. fresh board.
 * spawn0: index of the first spawn draw to use (0 for a fresh game). */
int orc_play_order(u64 start, const u8 *order4, u64 key, u64 spawn0, u64 *final_out, u64 *score_out,
                   u32 *count_out) {
    orc_init();
    u64 board = start ? start : orc_init_board(key);
    u64 score = 0;
    u32 count = 0;
    int err = 0;
    for (;;) {
        int advanced = 0;
        for (int a = 0; a < 4; a++) {
            u64 f;
            u32 s;
            u8 m;
            orc_move(board, order4[a], &f, &s, &m);
            if (m) {
                board = spawn_k(f, key, spawn0 + count, &err);
                if (err) return 1;
                score += s;
                count++;
                advanced = 1;
                break;
            }
        }
        if (!advanced) break;
    }
    *final_out = board;
    *score_out = score;
    *count_out = count;
    return 0;
}

static const u8 LUDR[4] = {0, 1, 3, 2}; /* board.py:214, board.py:310, mcts.py:26 */

/* board.py:306-315 play_fixed_batch(number): `number` fresh LUDR games to the end.
 * Output is indexed by game id (the reference returns scores in death order; compare as
 * multisets -- SURVEY.md A.7).  starts may be NULL (fresh boards). */
typedef struct {
    long n;
    const u64 *starts;
    u64 seed, gid_base;
    u64 *finals, *scores;
    u32 *counts;
    long next; /* shared chunk cursor (atomic) */
    int bad;
} batch_job_t;

static void *batch_worker(void *arg) {
    batch_job_t *job = (batch_job_t *)arg;
    const long chunk = 256;
    for (;;) {
        long lo = __atomic_fetch_add(&job->next, chunk, __ATOMIC_RELAXED);
        if (lo >= job->n) break;
        long hi = lo + chunk < job->n ? lo + chunk : job->n;
        for (long g = lo; g < hi; g++) {
            u64 key = orc_rng_key(job->seed, job->gid_base + (u64)g);
            u64 f, s;
            u32 c;
            if (orc_play_order(job->starts ? job->starts[g] : 0, LUDR, key, 0, &f, &s, &c))
                __atomic_store_n(&job->bad, 1, __ATOMIC_RELAXED);
            if (job->finals) job->finals[g] = f;
            if (job->scores) job->scores[g] = s;
            if (job->counts) job->counts[g] = c;
        }
    }
    return NULL;
}

int orc_play_fixed_batch(long n, const u64 *starts, u64 seed, u64 gid_base, u64 *finals, u64 *scores,
                         u32 *counts, int nthreads) {
    orc_init();
    batch_job_t job = {n, starts, seed, gid_base, finals, scores, counts, 0, 0};
    if (nthreads <= 1) {
        batch_worker(&job);
        return job.bad;
    }
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < nthreads; t++)
        if (pthread_create(&th[started], NULL, batch_worker, &job) == 0) started++;
    if (!started) batch_worker(&job);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    return job.bad;
}

/* ------------------------------------------------------------------------------------ */
/* Root search (mcts.py:4-103).  Line j of root move i of evaluation `eval_id` owns the   */
/* stream game_id = eval_id*4*number... see orc_line_gid; its first spawn (mcts.py:23     */
/* generate_tile(b)) is spawn 0 and the playout continues with spawns 1,2,...             */
/* ------------------------------------------------------------------------------------ */
u64 orc_line_gid(u64 eval_id, int root_move, u64 line, u64 number) {
    return (eval_id * 4ULL + (u64)root_move) * number + line;
}

/* per-line raw results for one origin: legal[4], root_score[4], and for every (i, j):
 * playout score (BoardArray scores start at 0: board.py:244) and playout length L
 * (accepted moves after the first spawn). */
int orc_root_lines(u64 origin, u64 number, u64 seed, u64 eval_id, const u8 *order4, u8 *legal4,
                   u32 *root_score4, u64 *line_score, u32 *line_moves, u64 *line_final) {
    orc_init();
    for (int i = 0; i < 4; i++) {
        u64 b;
        u32 s;
        u8 m;
        orc_move(origin, i, &b, &s, &m); /* mcts.py:21 */
        legal4[i] = m;
        root_score4[i] = s;
        for (u64 j = 0; j < number; j++) {
            u64 idx = (u64)i * number + j;
            if (!m) {
                line_score[idx] = 0;
                line_moves[idx] = 0;
                if (line_final) line_final[idx] = 0;
                continue;
            }
            u64 key = orc_rng_key(seed, orc_line_gid(eval_id, i, j, number));
            int err = 0;
            u64 start = spawn_k(b, key, 0, &err); /* mcts.py:23 */
            if (err) return 1;
            u64 f, sc;
            u32 c;
            if (orc_play_order(start, order4 ? order4 : LUDR, key, 1, &f, &sc, &c)) return 1;
            line_score[idx] = sc;
            line_moves[idx] = c;
            if (line_final) line_final[idx] = f;
        }
    }
    return 0;
}

/* mcts.py:4-34 mcts_fixed: mean(log10(playout_score + s + 1)) per legal root move, -1 if illegal.
 * NumPy's mean of a float64 array uses pairwise summation; for number <= 8 lanes... we
 * return the per-line log10 terms too so the Python side can apply np.mean exactly. */
int orc_mcts_fixed(u64 origin, u64 number, u64 seed, u64 eval_id, double *result4, double *terms) {
    u8 legal[4];
    u32 rs[4];
    u64 *sc = (u64 *)malloc(sizeof(u64) * 4 * number);
    u32 *mv = (u32 *)malloc(sizeof(u32) * 4 * number);
    int rc = orc_root_lines(origin, number, seed, eval_id, LUDR, legal, rs, sc, mv, NULL);
    for (int i = 0; i < 4 && !rc; i++) {
        if (!legal[i]) {
            result4[i] = -1.0; /* mcts.py:33 */
            continue;
        }
        double acc = 0.0;
        for (u64 j = 0; j < number; j++) {
            double t = log10((double)(sc[i * number + j] + rs[i] + 1)); /* mcts.py:31 */
            if (terms) terms[i * number + j] = t;
            acc += t;
        }
        result4[i] = acc / (double)number;
    }
    free(sc);
    free(mv);
    return rc;
}

/* mcts.py:73-103 mcts_fixed_min: 1-based step of the first death = 1 + min_j L_j; illegal -> 0 */
int orc_mcts_fixed_min(u64 origin, u64 number, u64 seed, u64 eval_id, int *result4) {
    u8 legal[4];
    u32 rs[4];
    u64 *sc = (u64 *)malloc(sizeof(u64) * 4 * number);
    u32 *mv = (u32 *)malloc(sizeof(u32) * 4 * number);
    int rc = orc_root_lines(origin, number, seed, eval_id, LUDR, legal, rs, sc, mv, NULL);
    for (int i = 0; i < 4 && !rc; i++) {
        if (!legal[i]) {
            result4[i] = 0; /* mcts.py:102 */
            continue;
        }
        u32 mn = 0xFFFFFFFFu;
        for (u64 j = 0; j < number; j++)
            if (mv[i * number + j] < mn) mn = mv[i * number + j];
        result4[i] = (int)mn + 1;
    }
    free(sc);
    free(mv);
    return rc;
}

static int cmp_u32(const void *a, const void *b) {
    u32 x = *(const u32 *)a, y = *(const u32 *)b;
    return (x > y) - (x < y);
}
/* mcts.py:37-70 mcts_fixed_moves: 1-based step at which cumulative deaths >= number//2.
 * A line of length L dies in step L+1.  With h = number//2: h == 0 -> the first step that has
 * any death (same as _min); else the h-th smallest L, plus 1.  illegal -> -1. */
int orc_mcts_fixed_moves(u64 origin, u64 number, u64 seed, u64 eval_id, int *result4) {
    u8 legal[4];
    u32 rs[4];
    u64 *sc = (u64 *)malloc(sizeof(u64) * 4 * number);
    u32 *mv = (u32 *)malloc(sizeof(u32) * 4 * number);
    int rc = orc_root_lines(origin, number, seed, eval_id, LUDR, legal, rs, sc, mv, NULL);
    for (int i = 0; i < 4 && !rc; i++) {
        if (!legal[i]) {
            result4[i] = -1; /* mcts.py:69 */
            continue;
        }
        qsort(mv + i * number, number, sizeof(u32), cmp_u32);
        u64 h = number / 2;
        u64 k = h ? h - 1 : 0;
        result4[i] = (int)mv[i * number + k] + 1;
    }
    free(sc);
    free(mv);
    return rc;
}

/* ------------------------------------------------------------------------------------ */
/* One BoardArray.move_batch step with per-game priority rows (board.py:246-277), the     */
/* building block of the NN-policy loops (mcts_nn.py:24-39, eval_nn.py:28-47).           */
/* alive[g] in/out; spawn_idx[g] is advanced for games that moved.                        */
/* ------------------------------------------------------------------------------------ */
int orc_policy_step(long n, u64 *boards, u64 *scores, u32 *spawn_idx, const u64 *keys, const u8 *prio /* n x 4 */,
                    u8 *alive) {
    orc_init();
    int deaths = 0;
    for (long g = 0; g < n; g++) {
        if (!alive[g]) continue;
        int advanced = 0;
        for (int a = 0; a < 4; a++) {
            u64 f;
            u32 s;
            u8 m;
            orc_move(boards[g], prio[4 * g + a], &f, &s, &m);
            if (m) {
                int err = 0;
                boards[g] = spawn_k(f, keys[g], spawn_idx[g], &err);
                if (err) return -1;
                scores[g] += s;
                spawn_idx[g]++;
                advanced = 1;
                break;
            }
        }
        if (!advanced) {
            alive[g] = 0;
            deaths++;
        }
    }
    return deaths;
}

/* legal-move mask of a board: bit d set when move(board, d) moves */
int orc_legal_mask(u64 board) {
    int mask = 0;
    for (int d = 0; d < 4; d++) {
        u64 f;
        u32 s;
        u8 m;
        orc_move(board, d, &f, &s, &m);
        if (m) mask |= 1 << d;
    }
    return mask;
}

int orc_num_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
