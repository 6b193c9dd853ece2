// playout_nn.cuh -- K4: persistent NN-policy playout kernels (mcts_nn.py:4-43 mcts_nn,
// mcts_nn.py:46-94 play_nn, eval_nn.py:4-95 eval_nn / eval_nn_min).
//
// A CTA owns a small set of game slots and loops {refill empty slots from the global work
// counter; one policy forward for all its boards; per game: first legal move in descending
// logit order (mcts_nn.py:36-37 argsort + board.py:264-269), spawn}, entirely on the device:
// no host round trip per move.  Work items are "lines" (root-search) or whole games.
//
// Min-move reduction with early stop (mcts_nn.py:37-39, eval_nn.py:91-93): every game belongs to
// a group; a group's result is min_j L_j.  A game stops as soon as its move count reaches the
// group's current minimum, because it can no longer lower it -- the minimum itself is
// independent of scheduling, so results are deterministic.
#pragma once
#include "convnet_fp32.cuh"
#include "convnet_tc.cuh"
#include "engine.cuh"

namespace b2048 {

// ---- device-resident main games -------------------------------------------------------------------------
// One MainGame per main game of the launch, in caller-provided workspace (HBM, read/written through L2 only).
// A main step = root moves of `board`, `number` policy lines per legal root move (work items in the MainQueue ring),
// min accepted moves per root move (gmin, shared with the lines' early stop), then the line that retires last
// ("closer": pending hits 0) picks argmax(1 + gmin), moves, spawns, records and queues the next step's lines.
// Main games advance independently: no grid-wide step barrier, a free slot takes the next queued line of ANY game.
struct MainGame {
    u64 board;       // current main board
    u64 key;         // main stream (seed, game id): start board and main spawns
    long long score; // merge score of the main moves made
    u64 rboard[4];   // boards after the four root moves of the current step
    u32 rscore[4];
    int step;        // main moves made
    int pending;     // lines of the current root evaluation queued or running
    int legal;       // bit i: root move i moves
    int over;
    int pad[10];
};
static_assert(sizeof(MainGame) == 128, "MainGame layout");
struct MainQueue {
    unsigned long long head, tail;  // consumed / produced line counts (ring index = count & (cap-1))
    int games_alive;
    int pad[11];
};
static_assert(sizeof(MainQueue) == 64, "MainQueue layout");

struct NNPlayParams {
    Tables tb;
    const void *blob;   // folded weights: fp32 image (convnet_fp32.cuh) or fp16 tensor-core image (convnet_tc.cuh)
    long n;             // work items
    // whole-game mode
    const u64 *start;  // start boards (NULL / 0 entry -> fresh board)
    // root mode
    int root_mode;
    const u64 *origins;     // [G] root boards; root move r = 4*g + i is recomputed by the line that starts from it
    const uint8_t *rlegal;  // [G*4]
    int number, line_begin, line_count;
    u64 seed, gid_base;
    u32 spawn0;
    // grouping for the min reduction: group = item / group_size (whole-game mode) or root index
    int group_size;      // 0 = no grouping (play every game to the end)
    int *group_min;      // [groups], caller-initialised to INT32_MAX
    // per-item outputs (may be NULL)
    u64 *out_final;
    u32 *out_score;
    u32 *out_moves;
    unsigned long long *counter;
    unsigned long long *move_counter;  // total accepted moves (for throughput accounting), may be NULL
    int single_set;                    // tensor-core kernel: use only tile-set 0 of every CTA (small batches)
    // population mode (evolve.py:30-36: the same evaluation for every network of a population): CTA c plays for
    // model c % n_models with that model's weights at blob + model*blob_stride and work cursor counter[model];
    // `n` is then the number of games PER MODEL and outputs are indexed model*n + game
    int n_models;          // <= 1: single model
    long blob_stride;      // bytes between consecutive models' weight images
    u64 gid_model_stride;  // game id offset per model (0: every model sees the same spawn streams)
    int *err_flag;
    // device-resident main games (selfplay.py:62-115 with mcts_nn.py:4-43 as the search): see MainGame below
    int main_mode;
    MainGame *mg;               // [G] main-game states
    MainQueue *mgq;             // line queue header
    u32 *mg_items;              // ring of queued lines, mg_cap entries (power of two), 0 = empty
    u32 mg_cap;
    u64 mg_g0, mg_total;        // first main-game id of this launch, games of the whole session (eval-id stride)
    int mg_max_steps;
    u64 *mg_out_boards;         // [G][max_steps] board before each main move (may be NULL)
    float *mg_out_results;      // [G][max_steps][4] root values the move was chosen from (may be NULL)
    u64 *mg_out_final;          // [G]
    long long *mg_out_score;    // [G]
    int *mg_out_steps;          // [G] main moves made
    int *mg_out_status;         // [G] 0 game over, 1 stopped at max_steps
};

struct GameSlot {
    u64 board, key;
    u32 score, count;
    long slot;   // output index (main mode: local main-game index)
    int group;   // -1 = none
    int live;
    int closing; // main mode: this line was the last one of its main game's step -> its lane advances the game
};

// fetch + initialise one work item into `g`; returns false when the queue is empty.
__device__ __forceinline__ bool nn_fetch(const NNPlayParams &p, GameSlot &g) {
    const int model = p.n_models > 1 ? (int)(blockIdx.x % (unsigned)p.n_models) : 0;
    for (;;) {
        long item = (long)atomicAdd(p.counter + model, 1ULL);
        if (item >= p.n) return false;
        g.score = 0;
        g.count = 0;
        g.group = -1;
        g.closing = 0;
        if (p.root_mode) {
            long r = item / p.line_count;
            int j = p.line_begin + (int)(item - r * p.line_count);
            g.slot = r * p.number + j;
            if (!p.rlegal[r]) {
                if (p.out_final) p.out_final[g.slot] = 0;
                if (p.out_score) p.out_score[g.slot] = 0;
                if (p.out_moves) p.out_moves[g.slot] = 0;
                continue;
            }
            g.key = rng_key(p.seed, (p.gid_base + (u64)r) * (u64)p.number + (u64)j);
            u32 rs;
            u64 rb = move_exact(p.origins[r >> 2], (int)(r & 3), p.tb, rs);  // mcts_nn.py:19
            if (countzero(rb) == 0) {
                *p.err_flag = 1;
                continue;
            }
            u32 four;
            g.board = spawn_tile(rb, rng_draw(g.key, B2048_INIT_DRAWS), four);  // mcts_nn.py:22
            if (p.group_min) g.group = (int)r;
        } else {
            g.slot = (long)model * p.n + item;
            g.key = rng_key(p.seed, p.gid_base + (u64)model * p.gid_model_stride + (u64)item);
            u64 s0 = p.start ? p.start[item] : 0ULL;
            g.board = s0 ? s0 : init_board(g.key);
            if (p.group_min && p.group_size > 0) g.group = (int)(g.slot / p.group_size);
        }
        g.live = 1;
        return true;
    }
}

__device__ __forceinline__ void nn_finish(const NNPlayParams &p, GameSlot &g) {
    g.live = 0;
    if (p.main_mode) {
        // the group's minimum must be visible before this line is counted as retired (the closer reads it)
        atomicMin(p.group_min + g.group, (int)g.count);
        __threadfence();
        g.closing = atomicSub(&p.mg[g.slot].pending, 1) == 1;
        return;
    }
    if (p.out_final) p.out_final[g.slot] = g.board;
    if (p.out_score) p.out_score[g.slot] = g.score;
    if (p.out_moves) p.out_moves[g.slot] = g.count;
    if (g.group >= 0) atomicMin(p.group_min + g.group, (int)g.count);
}

// One policy move for a live game given its 4 logits: directions in descending logit order
// (stable: ties go to the lower direction index), first legal one is played.
__device__ __forceinline__ void nn_move(const NNPlayParams &p, GameSlot &g, const float *z, u32 &moves_done) {
    // early stop: cannot lower the group's minimum any more
    if (g.group >= 0 && (int)g.count >= *((volatile int *)(p.group_min + g.group))) {
        nn_finish(p, g);
        return;
    }
    int order[4] = {0, 1, 2, 3};
#pragma unroll
    for (int a = 1; a < 4; a++)  // insertion sort, descending, stable
#pragma unroll
        for (int c = a; c > 0; c--)
            if (z[order[c]] > z[order[c - 1]]) {
                int t = order[c];
                order[c] = order[c - 1];
                order[c - 1] = t;
            }
    bool moved = false;
#pragma unroll 1
    for (int a = 0; a < 4 && !moved; a++) {
        u32 s;
        u64 f = move_exact(g.board, order[a], p.tb, s);
        if (f != g.board) {
            if (countzero(f) == 0) {
                *p.err_flag = 1;
                nn_finish(p, g);
                return;
            }
            u32 four;
            g.board = spawn_tile(f, rng_draw(g.key, B2048_INIT_DRAWS + p.spawn0 + g.count), four);
            g.score += s;
            g.count++;
            moves_done++;
            moved = true;
        }
    }
    if (!moved) nn_finish(p, g);
}

// Candidate moves of a slot, computed while the network runs (PlayWork::prefetch): for each direction the
// board AFTER the move and its spawn (the spawn draw depends only on the game's key and move count), the
// merge score, a legality mask, and the group's current minimum (slightly stale is fine: a larger bound
// only delays the early stop, the minimum itself is unaffected).  After the forward the move choice is a
// handful of ALU instructions.
struct SlotCand {
    u64 next[4];
    u32 score[4];
    int gmin;
    u32 legal;  // bit d: direction d moves; bit 8+d: the move empties the board (randrange(0) in the reference)
    u32 pad[2];
};
static_assert(sizeof(SlotCand) == 64, "SlotCand must be 64 bytes");

__device__ __forceinline__ void nn_move_cand(const NNPlayParams &p, GameSlot &g, const float *z, const SlotCand &c,
                                             u32 &moves_done) {
    if (g.group >= 0 && (int)g.count >= c.gmin) {
        nn_finish(p, g);
        return;
    }
    // highest logit among the legal directions (ties -> lower index, as the stable descending argsort)
    int best = -1;
    float bz = 0.f;
#pragma unroll
    for (int d = 0; d < 4; d++) {
        const bool legal = (c.legal >> d) & 1u;
        if (legal && (best < 0 || z[d] > bz)) best = d, bz = z[d];
    }
    if (best < 0) {
        nn_finish(p, g);
        return;
    }
    if ((c.legal >> (8 + best)) & 1u) {
        *p.err_flag = 1;
        nn_finish(p, g);
        return;
    }
    u64 f = c.next[0];
    u32 s = c.score[0];
#pragma unroll
    for (int d = 1; d < 4; d++)
        if (best == d) f = c.next[d], s = c.score[d];
    g.board = f;
    g.score += s;
    g.count++;
    moves_done++;
}


// ---------------------------------------------------------------------------------------------------------
// Device-resident main games: queue + main-move logic.  Everything below reads and writes the workspace with
// volatile / atomic accesses (L2): the states are handed between SMs many times per launch.
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T ld_vol(const T *p) { return *reinterpret_cast<const volatile T *>(p); }
template <typename T>
__device__ __forceinline__ void st_vol(T *p, T v) { *reinterpret_cast<volatile T *>(p) = v; }

__device__ __forceinline__ void mg_finalize(const NNPlayParams &p, long gl, u64 board, long long score, int step, int status) {
    p.mg_out_final[gl] = board;
    p.mg_out_score[gl] = score;
    p.mg_out_steps[gl] = step;
    p.mg_out_status[gl] = status;
    st_vol(&p.mg[gl].over, 1);
    __threadfence();
    atomicSub(&p.mgq->games_alive, 1);
}

// Root moves of the game's current board (mcts_nn.py:18-19): fills rboard / rscore / legal, resets the four minima
// and `pending`.  Returns the legal mask; 0 = no root move moves -> mcts_nn returns [0,0,0,0], argmax is move 0, it
// does not move: the main game is over (selfplay.py:95-98) and has been finalized here.
__device__ __forceinline__ int mg_publish(const NNPlayParams &p, long gl, u64 board, long long score, int step) {
    MainGame *s = p.mg + gl;
    int mask = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        u32 sc;
        const u64 f = move_exact(board, i, p.tb, sc);
        if (f != board) {
            if (countzero(f) == 0) {  // generate_tile on an emptied board: ValueError in the reference (board.py:87)
                *p.err_flag = 1;
                mask = 0;
                break;
            }
            mask |= 1 << i;
        }
        st_vol(&s->rboard[i], f);
        st_vol(&s->rscore[i], sc);
        st_vol(p.group_min + gl * 4 + i, 0x7fffffff);
    }
    if (!mask) {
        mg_finalize(p, gl, board, score, step, 0);
        return 0;
    }
    st_vol(&s->legal, mask);
    st_vol(&s->pending, __popc(mask) * p.number);
    return mask;
}

// The closer's part of a main step (selfplay.py:93-105): values 1 + min (0 where illegal), first maximum, move,
// spawn number `step` of the main stream, record.  Returns the next step's legal mask (0: the game has ended).
__device__ __forceinline__ int mg_advance(const NNPlayParams &p, long gl) {
    MainGame *s = p.mg + gl;
    __threadfence();  // the retired lines' minima (released before their `pending` decrement)
    const int legal = ld_vol(&s->legal);
    const int step = ld_vol(&s->step);
    const u64 board = ld_vol(&s->board);
    float val[4];
    int best = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        val[i] = ((legal >> i) & 1) ? (float)(ld_vol(p.group_min + gl * 4 + i) + 1) : 0.f;
        if (val[i] > val[best]) best = i;  // np.argmax: first maximum
    }
    if (p.mg_out_boards) p.mg_out_boards[gl * p.mg_max_steps + step] = board;
    if (p.mg_out_results)
        reinterpret_cast<float4 *>(p.mg_out_results)[gl * p.mg_max_steps + step] = make_float4(val[0], val[1], val[2], val[3]);
    const long long score = ld_vol(&s->score) + (long long)ld_vol(&s->rscore[best]);
    u32 four;
    const u64 nb = spawn_tile(ld_vol(&s->rboard[best]), rng_draw(ld_vol(&s->key), B2048_INIT_DRAWS + (u64)step), four);
    st_vol(&s->board, nb);
    st_vol(&s->score, score);
    st_vol(&s->step, step + 1);
    if (step + 1 >= p.mg_max_steps) {
        mg_finalize(p, gl, nb, score, step + 1, 1);
        return 0;
    }
    return mg_publish(p, gl, nb, score, step + 1);
}

// Queue the lines of game gl's current step: `number` per legal root move, ring slots [base, base + n).
// Called by `nlanes` cooperating lanes (lane = 0..nlanes-1) after the game's state was written by one of them.
__device__ __forceinline__ void mg_push(const NNPlayParams &p, long gl, int mask, unsigned long long base, int lane, int nlanes) {
    const int n = __popc(mask) * p.number;
    __threadfence();  // state before items
    for (int k = lane; k < n; k += nlanes) {
        int ii = k / p.number;
        const int j = k - ii * p.number;
        int i = 0;
        for (int m = mask;; i++, m >>= 1)
            if (m & 1) {
                if (ii == 0) break;
                ii--;
            }
        const u32 code = (u32)(gl * 4 * p.number + i * p.number + j) + 1u;
        st_vol(p.mg_items + ((base + (unsigned long long)k) & (unsigned long long)(p.mg_cap - 1)), code);
    }
}

// Start the queued line `code - 1` in slot g.
__device__ __forceinline__ void mg_start_line(const NNPlayParams &p, GameSlot &g, u32 code) {
    code -= 1u;
    const u32 per = 4u * (u32)p.number;
    const long gl = code / per;
    const u32 rem = code - (u32)gl * per;
    const int i = (int)(rem / (u32)p.number), j = (int)(rem - (u32)i * (u32)p.number);
    const MainGame *s = p.mg + gl;
    const u64 eval_id = ((u64)ld_vol(&s->step) + 1ULL) * p.mg_total + p.mg_g0 + (u64)gl;
    g.key = rng_key(p.seed, (eval_id * 4ULL + (u64)i) * (u64)p.number + (u64)j);
    u32 four;
    g.board = spawn_tile(ld_vol(&s->rboard[i]), rng_draw(g.key, B2048_INIT_DRAWS), four);  // mcts_nn.py:22
    g.score = 0;
    g.count = 0;
    g.slot = gl;
    g.group = (int)gl * 4 + i;
    g.closing = 0;
    g.live = 1;
}

// Warp-level service routine of the main-game mode, called by all 32 lanes of a slot-owning warp after the move
// phase: (1) lanes whose line closed a main step advance that game and queue its next lines (one closer at a time,
// the warp writes the items together), (2) free slots claim queued lines (one CAS per warp), (3) a warp with nothing
// live waits while main games are still running elsewhere (`spin`; the tensor-core kernel passes false and idles the
// tile-set through its RUN_IDLE protocol instead, because its MMA warps serve two sets).  `has_slot`: this lane owns a slot.
__device__ __noinline__ void mg_service(const NNPlayParams &p, GameSlot &g, bool has_slot, bool &queue_empty, int lane, bool spin) {
    MainQueue *q = p.mgq;
    for (;;) {
        unsigned cm = __ballot_sync(0xffffffffu, has_slot && g.closing);
        while (cm) {
            const int l = __ffs(cm) - 1;
            cm &= cm - 1;
            const long gl = __shfl_sync(0xffffffffu, g.slot, l);
            int mask = 0;
            unsigned long long base = 0;
            if (lane == l) {
                mask = mg_advance(p, gl);
                if (mask) base = atomicAdd(&q->tail, (unsigned long long)(__popc(mask) * p.number));
                g.closing = 0;
            }
            mask = __shfl_sync(0xffffffffu, mask, l);
            base = __shfl_sync(0xffffffffu, base, l);
            __syncwarp();
            if (mask) mg_push(p, gl, mask, base, lane, 32);
        }
        const bool want = has_slot && !g.live && !queue_empty;
        const unsigned need = __ballot_sync(0xffffffffu, want);
        if (need) {
            int take = 0;
            unsigned long long base = 0;
            bool all_over = false;
            if (lane == 0) {
                const int cnt = __popc(need);
                for (;;) {
                    const unsigned long long h = ld_vol(&q->head);
                    const long long avail = (long long)(ld_vol(&q->tail) - h);
                    if (avail <= 0) {
                        all_over = ld_vol(&q->games_alive) <= 0;
                        break;
                    }
                    take = avail < cnt ? (int)avail : cnt;
                    if (atomicCAS(&q->head, h, h + (unsigned long long)take) == h) {
                        base = h;
                        break;
                    }
                    take = 0;
                }
            }
            take = __shfl_sync(0xffffffffu, take, 0);
            base = __shfl_sync(0xffffffffu, base, 0);
            all_over = __shfl_sync(0xffffffffu, (int)all_over, 0) != 0;
            const int rank = __popc(need & ((1u << lane) - 1u));
            if (want && rank < take) {
                u32 *slot = p.mg_items + ((base + (unsigned long long)rank) & (unsigned long long)(p.mg_cap - 1));
                u32 code;
                while ((code = atomicExch(slot, 0u)) == 0u) __nanosleep(64);  // reserved by a producer, being written
                __threadfence();
                mg_start_line(p, g, code);
            }
            if (all_over && want) queue_empty = true;  // every main game has ended: nothing will be queued again
        }
        const unsigned live = __ballot_sync(0xffffffffu, has_slot && g.live);
        const unsigned open = __ballot_sync(0xffffffffu, has_slot && !queue_empty);
        if (live || !open || !spin) return;
        __nanosleep(512);  // all slots free, main games still running on other SMs: poll the queue
    }
}

// One thread per main game: main stream key, start board, first root evaluation queued.
__global__ void mg_init_kernel(const NNPlayParams p, const u64 *starts, long G) {
    const long gl = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gl >= G) return;
    MainGame *s = p.mg + gl;
    const u64 key = rng_key(p.seed, p.mg_g0 + (u64)gl);
    const u64 s0 = starts ? starts[gl] : 0ULL;
    const u64 board = s0 ? s0 : init_board(key);  // selfplay.py:83 `if not board`
    s->key = key;
    s->board = board;
    s->score = 0;
    s->step = 0;
    s->over = 0;
    atomicAdd(&p.mgq->games_alive, 1);
    int mask = p.mg_max_steps > 0 ? mg_publish(p, gl, board, 0, 0) : 0;
    if (p.mg_max_steps <= 0) mg_finalize(p, gl, board, 0, 0, 1);
    if (mask) {
        const unsigned long long base = atomicAdd(&p.mgq->tail, (unsigned long long)(__popc(mask) * p.number));
        mg_push(p, gl, mask, base, 0, 1);
    }
}

// Exact-mode kernel: 8 game slots per CTA, fp32 CUDA-core forward.
__global__ void __launch_bounds__(512) playout_nn_fp32_kernel(const NNPlayParams p) {
    extern __shared__ __align__(16) float smem[];
    float *act0 = smem, *act1 = smem + 8 * 64 * 16, *s_head = act1 + 8 * 64 * 16, *s_logits = s_head + 8 * 64;
    u64 *s_boards = reinterpret_cast<u64 *>(s_logits + 32);
    __shared__ int s_live;
    GameSlot g;
    g.live = 0;
    g.closing = 0;
    g.slot = 0;
    g.board = 0;
    u32 moves_done = 0;
    bool queue_empty = false;
    const int t = threadIdx.x;
    const float *blob = reinterpret_cast<const float *>(
        reinterpret_cast<const uint8_t *>(p.blob) + (p.n_models > 1 ? (size_t)(blockIdx.x % (unsigned)p.n_models) * p.blob_stride : 0));
    for (;;) {
        if (t == 0) s_live = 0;
        __syncthreads();
        if (p.main_mode) {
            if (t < 32) mg_service(p, g, t < CN_BOARDS_PER_CTA, queue_empty, t, true);
        } else if (t < CN_BOARDS_PER_CTA) {
            if (!g.live && !queue_empty) queue_empty = !nn_fetch(p, g);
        }
        if (t < CN_BOARDS_PER_CTA) {
            s_boards[t] = g.live ? g.board : 0ULL;
            if (g.live) atomicOr(&s_live, 1);
        }
        __syncthreads();
        if (!s_live) break;
        convnet_fp32_cta(blob, s_boards, CN_BOARDS_PER_CTA, act0, act1, s_head, s_logits);
        if (t < CN_BOARDS_PER_CTA && g.live) nn_move(p, g, s_logits + 4 * t, moves_done);
        __syncthreads();
    }
    if (t < CN_BOARDS_PER_CTA && p.move_counter && moves_done) atomicAdd(p.move_counter, (unsigned long long)moves_done);
}

// Tensor-core kernel (tc::tc_persistent_kernel<PlayWork>): 2 x 32 game slots per CTA, one slot per lane
// of each worker group's owner warp: consume the forward's logits (move choice, spawn, death), refill
// from the global work counter, hand the next 32 boards to the network.  The engine phase of one set
// overlaps the other set's forward.
// Candidate moves of `board` whose next spawn is draw number `draw_idx` of stream `key` (see SlotCand).
__device__ __forceinline__ void fill_cand(const NNPlayParams &p, SlotCand *c, u64 board, u64 key, u32 draw_idx, int gmin) {
    const u64 draw = rng_draw(key, B2048_INIT_DRAWS + p.spawn0 + draw_idx);
    u32 legal = 0;
#pragma unroll
    for (int d = 0; d < 4; d++) {
        u32 s;
        u64 f = move_exact(board, d, p.tb, s);
        if (f != board) {
            legal |= 1u << d;
            if (countzero(f) == 0) {
                legal |= 1u << (8 + d);
            } else {
                u32 four;
                f = spawn_tile(f, draw, four);
            }
        }
        c->next[d] = f;
        c->score[d] = s;
    }
    c->legal = legal;
    c->gmin = gmin;
}

// highest logit among the legal directions (ties -> lower index, as the stable descending argsort); -1: no legal move
__device__ __forceinline__ int best_legal(const float *z, u32 legal) {
    int best = -1;
    float bz = 0.f;
#pragma unroll
    for (int d = 0; d < 4; d++)
        if (((legal >> d) & 1u) && (best < 0 || z[d] > bz)) best = d, bz = z[d];
    return best;
}

// Two-ply speculation (latency regime: far fewer lines than game slots, e.g. BASELINE config 3 = 200 lines on a GPU with
// 4,736 slots).  A line owns SPEC_GROUP = 5 consecutive slots of a tile: its current board and the four boards that follow from
// it (move d + the spawn that is already determined by the line's stream).  One forward evaluates all five, so after it the line
// makes its move AND the move of the board it arrives at: two policy moves per forward, i.e. half the dependent forwards per
// line, for 5x the (otherwise idle) tensor work.  The boards evaluated are the boards the one-ply kernel would evaluate and a
// board's logits do not depend on its slot, so the moves -- and every result -- are identical.
constexpr int SPEC_GROUP = 5, SPEC_LINES = 6;  // 6 lines x 5 slots = 30 of a tile's 32 slots
__device__ __forceinline__ void nn_move_spec(const NNPlayParams &p, GameSlot &g, const float *logits, const SlotCand *cands,
                                             int lead, u32 &moves_done) {
    const SlotCand &c0 = cands[lead];
    if (g.group >= 0 && (int)g.count >= c0.gmin) {
        nn_finish(p, g);
        return;
    }
    const int b0 = best_legal(logits + 4 * lead, c0.legal);
    if (b0 < 0) {
        nn_finish(p, g);
        return;
    }
    if ((c0.legal >> (8 + b0)) & 1u) {
        *p.err_flag = 1;
        nn_finish(p, g);
        return;
    }
    g.board = c0.next[b0];
    g.score += c0.score[b0];
    g.count++;
    moves_done++;
    // second ply: the board just reached is the one slot lead + 1 + b0 evaluated in the same forward
    if (g.group >= 0 && (int)g.count >= c0.gmin) {
        nn_finish(p, g);
        return;
    }
    const int cl = lead + 1 + b0;
    const SlotCand &c1 = cands[cl];
    const int b1 = best_legal(logits + 4 * cl, c1.legal);
    if (b1 < 0) {
        nn_finish(p, g);
        return;
    }
    if ((c1.legal >> (8 + b1)) & 1u) {
        *p.err_flag = 1;
        nn_finish(p, g);
        return;
    }
    g.board = c1.next[b1];
    g.score += c1.score[b1];
    g.count++;
    moves_done++;
}

// Three-ply speculation (even fewer lines: at most 2 per SM, e.g. BASELINE config 3 = 200 lines on 148 SMs).  A line owns
// SPEC3_GROUP = 16 slots of a tile: slot 0 its current board, slots 1..4 the boards after move d (+ the spawn its stream has
// already fixed), slots 5..15 the first 11 boards -- in (d, e) order -- that follow from a legal second move e.  One forward
// evaluates all of them, so the line plays up to THREE policy moves per forward; when the pair it plays is not among the 11
// evaluated ones it plays two.  `pairs` = bit 4d+e set when the board after moves d, e exists.
constexpr int SPEC3_GROUP = 16, SPEC3_LINES = 2, SPEC3_GRAND = SPEC3_GROUP - 5;
__device__ __forceinline__ void nn_move_spec3(const NNPlayParams &p, GameSlot &g, const float *logits, const SlotCand *cands,
                                              int lead, u32 pairs, u32 &moves_done) {
    const int gmin = cands[lead].gmin;
    int slot = lead;
#pragma unroll 1
    for (int ply = 0; ply < 3; ply++) {
        if (g.group >= 0 && (int)g.count >= gmin) {
            nn_finish(p, g);
            return;
        }
        const SlotCand &c = cands[slot];
        const int b = best_legal(logits + 4 * slot, c.legal);
        if (b < 0) {
            nn_finish(p, g);
            return;
        }
        if ((c.legal >> (8 + b)) & 1u) {
            *p.err_flag = 1;
            nn_finish(p, g);
            return;
        }
        g.board = c.next[b];
        g.score += c.score[b];
        g.count++;
        moves_done++;
        if (ply == 0) {
            slot = lead + 1 + b;  // the board just reached was evaluated by child slot b
        } else if (ply == 1) {
            const int idx = 4 * (slot - lead - 1) + b;
            const int rank = __popc(pairs & ((1u << idx) - 1u));
            if (!((pairs >> idx) & 1u) || rank >= SPEC3_GRAND) return;  // that grandchild was not evaluated: two plies this time
            slot = lead + 5 + rank;
        }
    }
}

// MAIN = true: the device-resident main-game mode (slots are fed by the line queue, mg_service), a separate kernel
// instantiation so that the plain playout kernel carries none of its code.  SPEC = 1: two-ply speculation, SPEC = 2:
// three-ply speculation (above), both for launches that are small from the start.
// SPEC = 3 (ADAPTIVE, not with MAIN): a tile-set starts with one line per slot; once the launch's work queue is empty and
// at most 6 (2) of its 32 lines are still running, the survivors are moved to the leader slots of the two-ply (three-ply)
// layout and the free slots evaluate their successors: the tail of a launch -- a few long lines holding an SM each --
// then runs at two to three policy moves per forward.  The boards evaluated are the boards the plain kernel evaluates,
// so every result is identical.
template <bool MAIN, int SPEC = 0>
struct PlayWorkT {
    static_assert(!(MAIN && SPEC == 3), "main-game launches have no tail: their slots are fed by the line queue");
    typedef NNPlayParams Params;
    struct State {
        GameSlot g;
        u32 moves_done;
        bool queue_empty;
        // speculation layouts: a successor slot is never live; it keeps the board it evaluates, the line's stream and the
        // index of the spawn that follows a move from that board in its own g.board / g.key / g.count
        u32 pairs;  // three-ply layout: which (first move, second move) successors of the line's board exist (nn_move_spec3)
        int mode;   // layout of this tile-set: 0 = one line per slot, 1 = two-ply groups, 2 = three-ply groups
    };
    __device__ static size_t blob_offset(const Params &p) {
        return p.n_models > 1 ? (size_t)(blockIdx.x % (unsigned)p.n_models) * (size_t)p.blob_stride : 0;
    }
    __device__ static int mode_of(const State &st) { return SPEC == 3 ? st.mode : SPEC; }
    __device__ static bool is_leader(int mode, int lane) {
        if (mode == 2) return (lane & (SPEC3_GROUP - 1)) == 0;
        return mode == 0 || (lane < SPEC_GROUP * SPEC_LINES && lane % SPEC_GROUP == 0);
    }
    __device__ static void init(State &st, const Params &p, int set) {
        st.g.live = 0;
        st.g.closing = 0;
        st.g.slot = 0;
        st.g.board = 0;
        st.g.key = 0;
        st.g.count = 0;
        st.g.group = -1;
        st.g.score = 0;
        st.moves_done = 0;
        st.pairs = 0;
        st.mode = SPEC == 3 ? 0 : SPEC;
        // latency mode (few work items): one tile-set per CTA so a step is not interleaved with a second set
        st.queue_empty = (set > 0) && p.single_set;
        if (!is_leader(st.mode, (int)(threadIdx.x & 31))) st.queue_empty = true;  // successor slots never take work items
    }
    __device__ static void prefetch(const Params &p, State &st, uint8_t *scratch, int lane) {
        SlotCand *c = reinterpret_cast<SlotCand *>(scratch) + lane;
        const int mode = mode_of(st);
        const bool lead = is_leader(mode, lane);
        if (lead ? !st.g.live : !st.g.board) return;  // (successor slots exist in the speculation layouts only)
        fill_cand(p, c, st.g.board, st.g.key, st.g.count,
                  (lead && st.g.group >= 0) ? *((volatile int *)(p.group_min + st.g.group)) : 0x7fffffff);
    }
    // ADAPTIVE: move the k-th live line of the set to lane k * stride (the leader slots of the next layout)
    __device__ static void compact(State &st, unsigned live, int stride, int lane) {
        const int k = lane / stride;
        const bool target = (lane % stride == 0) && k < __popc(live);
        const int src = target ? (int)__fns(live, 0, k + 1) : lane;
        GameSlot &g = st.g;
        g.board = __shfl_sync(0xffffffffu, g.board, src);
        g.key = __shfl_sync(0xffffffffu, g.key, src);
        g.score = __shfl_sync(0xffffffffu, g.score, src);
        g.count = __shfl_sync(0xffffffffu, g.count, src);
        g.slot = __shfl_sync(0xffffffffu, g.slot, src);
        g.group = __shfl_sync(0xffffffffu, g.group, src);
        g.live = target ? 1 : 0;
        g.closing = 0;
        st.queue_empty = true;
    }
    __device__ static int step(const Params &p, State &st, bool first, const float *logits, unsigned long long *s_boards,
                               int lane, long &dbg_base, uint8_t *scratch) {
        dbg_base = 0;
        int mode = mode_of(st);
        const bool lead = is_leader(mode, lane);
        if (!first && st.g.live) {
            if (mode == 2) nn_move_spec3(p, st.g, logits, reinterpret_cast<const SlotCand *>(scratch), lane, st.pairs, st.moves_done);
            else if (mode == 1) nn_move_spec(p, st.g, logits, reinterpret_cast<const SlotCand *>(scratch), lane, st.moves_done);
            else nn_move_cand(p, st.g, logits + 4 * lane, reinterpret_cast<const SlotCand *>(scratch)[lane], st.moves_done);
        }
        if (MAIN) mg_service(p, st.g, lead, st.queue_empty, lane, false);
        else if (lead && !st.g.live && !st.queue_empty) st.queue_empty = !nn_fetch(p, st.g);
        if (SPEC == 3 && mode < 2) {
            // the queue is empty (every free slot has seen it) and few lines are left: regroup them so that the free slots
            // evaluate their successors
            const unsigned live = __ballot_sync(0xffffffffu, st.g.live);
            const unsigned open = __ballot_sync(0xffffffffu, !st.g.live && !st.queue_empty);
            const int n = __popc(live);
            if (!open && n > 0 && n < 32) {
                const int want = n <= SPEC3_LINES ? 2 : (n <= SPEC_LINES ? 1 : 0);
                if (want > mode) {
                    compact(st, live, want == 2 ? SPEC3_GROUP : SPEC_GROUP, lane);
                    st.mode = mode = want;
                }
            }
        }
        if (mode == 2) {
            // lane 16 l + 4 d + e of line l computes the board after moves d, e (with the two spawns the line's stream has
            // already fixed); the legal ones are packed, in (d, e) order, into the line's 11 grandchild slots
            const int sub = lane & (SPEC3_GROUP - 1), src = lane & SPEC3_GROUP;
            const u64 lb = __shfl_sync(0xffffffffu, st.g.board, src), lk = __shfl_sync(0xffffffffu, st.g.key, src);
            const u32 lc = __shfl_sync(0xffffffffu, st.g.count, src);
            const int ll = __shfl_sync(0xffffffffu, st.g.live, src);
            u64 child = 0, grand = 0;
            if (ll) {
                u32 sc, four;
                const u64 f = move_exact(lb, sub >> 2, p.tb, sc);
                if (f != lb && countzero(f) != 0) {
                    child = spawn_tile(f, rng_draw(lk, B2048_INIT_DRAWS + p.spawn0 + lc), four);  // == SlotCand.next[d] of the leader
                    const u64 f2 = move_exact(child, sub & 3, p.tb, sc);
                    if (f2 != child && countzero(f2) != 0)
                        grand = spawn_tile(f2, rng_draw(lk, B2048_INIT_DRAWS + p.spawn0 + lc + 1u), four);
                }
            }
            const u32 pairs = (__ballot_sync(0xffffffffu, grand != 0) >> src) & 0xFFFFu;
            // slot `sub` of the line: 0 = the line's board, 1..4 = child d = sub - 1 (held by lane 4 d), 5..15 = the
            // (sub - 5)-th existing grandchild
            const u32 nth = sub >= 5 ? __fns(pairs, 0, sub - 4) : 0xFFFFFFFFu;
            const u64 cb = __shfl_sync(0xffffffffu, child, src + ((sub >= 1 && sub <= 4) ? 4 * (sub - 1) : 0));
            const u64 gb = __shfl_sync(0xffffffffu, grand, src + (nth < 16u ? (int)nth : 0));
            u64 mine = 0;
            u32 draw = lc;
            if (sub == 0) mine = ll ? lb : 0ULL;
            else if (sub <= 4) mine = cb, draw = lc + 1u;
            else if (nth < 16u) mine = gb, draw = lc + 2u;
            st.pairs = pairs;
            if (sub) st.g.board = mine, st.g.key = lk, st.g.count = draw;
            s_boards[lane] = mine;
        } else if (mode == 1) {
            // the line's leader hands its board / stream / move count to its four successor slots
            const bool lead1 = is_leader(1, lane);
            const int sub = lane % SPEC_GROUP, src = (lane < SPEC_GROUP * SPEC_LINES) ? lane - sub : 0;
            const u64 lb = __shfl_sync(0xffffffffu, st.g.board, src), lk = __shfl_sync(0xffffffffu, st.g.key, src);
            const u32 lc = __shfl_sync(0xffffffffu, st.g.count, src);
            const int ll = __shfl_sync(0xffffffffu, st.g.live, src);
            u64 mine = 0;
            if (lead1) {
                mine = st.g.live ? st.g.board : 0ULL;
            } else if (lane < SPEC_GROUP * SPEC_LINES && ll) {
                u32 sc;
                const u64 f = move_exact(lb, sub - 1, p.tb, sc);
                if (f != lb && countzero(f) != 0) {
                    u32 four;
                    mine = spawn_tile(f, rng_draw(lk, B2048_INIT_DRAWS + p.spawn0 + lc), four);  // == SlotCand.next[sub - 1] of the leader
                }
            }
            if (!lead1) st.g.board = mine, st.g.key = lk, st.g.count = lc + 1u;
            s_boards[lane] = mine;
        } else {
            s_boards[lane] = st.g.live ? st.g.board : 0ULL;
        }
        const unsigned any = __ballot_sync(0xffffffffu, st.g.live);
        if (MAIN && !any && __ballot_sync(0xffffffffu, !st.queue_empty))
            return tc::RUN_IDLE;  // no line here now, main games still running: lines may be queued any moment
        if (!any && p.move_counter) {
            u32 m = st.moves_done;
            for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
            if (lane == 0 && m) atomicAdd(p.move_counter, (unsigned long long)m);
        }
        return any ? tc::RUN_FWD : tc::RUN_STOP;
    }
};
typedef PlayWorkT<false> PlayWork;
typedef PlayWorkT<true> MainWork;
typedef PlayWorkT<false, 1> PlaySpecWork;
typedef PlayWorkT<true, 1> MainSpecWork;
typedef PlayWorkT<false, 2> PlaySpec3Work;
typedef PlayWorkT<true, 2> MainSpec3Work;
typedef PlayWorkT<false, 3> PlayAdaptWork;

}  // namespace b2048
