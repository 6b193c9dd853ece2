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

struct NNPlayParams {
    Tables tb;
    const void *blob;   // folded weights: fp32 image (convnet_fp32.cuh) or fp16 tensor-core image (convnet_tc.cuh)
    long n;             // work items
    // whole-game mode
    const u64 *start;  // start boards (NULL / 0 entry -> fresh board)
    // root mode
    int root_mode;
    const u64 *rboard;      // [G*4] boards after the root move
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
};

struct GameSlot {
    u64 board, key;
    u32 score, count;
    long slot;   // output index
    int group;   // -1 = none
    int live;
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
            u64 rb = p.rboard[r];
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
    if (p.out_final) p.out_final[g.slot] = g.board;
    if (p.out_score) p.out_score[g.slot] = g.score;
    if (p.out_moves) p.out_moves[g.slot] = g.count;
    if (g.group >= 0) atomicMin(p.group_min + g.group, (int)g.count);
    g.live = 0;
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

// Exact-mode kernel: 8 game slots per CTA, fp32 CUDA-core forward.
__global__ void __launch_bounds__(512) playout_nn_fp32_kernel(const NNPlayParams p) {
    extern __shared__ __align__(16) float smem[];
    float *act0 = smem, *act1 = smem + 8 * 64 * 16, *s_head = act1 + 8 * 64 * 16, *s_logits = s_head + 8 * 64;
    u64 *s_boards = reinterpret_cast<u64 *>(s_logits + 32);
    __shared__ int s_live;
    GameSlot g;
    g.live = 0;
    g.board = 0;
    u32 moves_done = 0;
    bool queue_empty = false;
    const int t = threadIdx.x;
    const float *blob = reinterpret_cast<const float *>(
        reinterpret_cast<const uint8_t *>(p.blob) + (p.n_models > 1 ? (size_t)(blockIdx.x % (unsigned)p.n_models) * p.blob_stride : 0));
    for (;;) {
        if (t == 0) s_live = 0;
        __syncthreads();
        if (t < CN_BOARDS_PER_CTA) {
            if (!g.live && !queue_empty) queue_empty = !nn_fetch(p, g);
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
struct PlayWork {
    typedef NNPlayParams Params;
    struct State {
        GameSlot g;
        u32 moves_done;
        bool queue_empty;
    };
    __device__ static size_t blob_offset(const Params &p) {
        return p.n_models > 1 ? (size_t)(blockIdx.x % (unsigned)p.n_models) * (size_t)p.blob_stride : 0;
    }
    __device__ static void init(State &st, const Params &p, int set) {
        st.g.live = 0;
        st.g.board = 0;
        st.moves_done = 0;
        // latency mode (few work items): one tile-set per CTA so a step is not interleaved with a second set
        st.queue_empty = (set > 0) && p.single_set;
    }
    __device__ static void prefetch(const Params &p, State &st, uint8_t *scratch, int lane) {
        if (!st.g.live) return;
        SlotCand *c = reinterpret_cast<SlotCand *>(scratch) + lane;
        const u64 draw = rng_draw(st.g.key, B2048_INIT_DRAWS + p.spawn0 + st.g.count);  // this move's spawn
        u32 legal = 0;
#pragma unroll
        for (int d = 0; d < 4; d++) {
            u32 s;
            u64 f = move_exact(st.g.board, d, p.tb, s);
            if (f != st.g.board) {
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
        c->gmin = (st.g.group >= 0) ? *((volatile int *)(p.group_min + st.g.group)) : 0x7fffffff;
    }
    __device__ static bool step(const Params &p, State &st, bool first, const float *logits, unsigned long long *s_boards,
                                int lane, long &dbg_base, uint8_t *scratch) {
        dbg_base = 0;
        if (!first && st.g.live)
            nn_move_cand(p, st.g, logits + 4 * lane, reinterpret_cast<const SlotCand *>(scratch)[lane], st.moves_done);
        if (!st.g.live && !st.queue_empty) st.queue_empty = !nn_fetch(p, st.g);
        s_boards[lane] = st.g.live ? st.g.board : 0ULL;
        const unsigned any = __ballot_sync(0xffffffffu, st.g.live);
        if (!any && p.move_counter) {
            u32 m = st.moves_done;
            for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
            if (lane == 0 && m) atomicAdd(p.move_counter, (unsigned long long)m);
        }
        return any != 0;
    }
};

}  // namespace b2048
