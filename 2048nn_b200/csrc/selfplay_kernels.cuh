// selfplay_kernels.cuh -- whole fixed-order root-search MAIN GAMES on the device (mcts.py:106-148 play_mcts_fixed,
// selfplay.py:9-59 selfplay_fixed): one CTA plays one main game from the first board to the end without a host
// round trip per main move.  Per main move: the 4 root moves (mcts.py:21), times x 4 x number fixed-order lines to
// the end (one line per thread, looping), the reduction of the chosen mode, argmax, move + spawn, record.
//
// Why a separate kernel from K2: a root evaluation of `number`=10 is 40 short lines -- a grid-wide persistent launch
// plus a host decision per main move costs ~440 us of latency for ~30 us of work.  Here the decision stays in the
// CTA; lines read the row tables through L1 (__ldg) instead of a 192 KB shared-memory copy, so dozens of
// games share an SM.
//
// Streams (identical to the host loops in 2048nn_b200/mcts.py and selfplay.py, so the games are the same games):
//   main game with id m: start board and main spawns from stream (seed, m), spawn k = main moves made so far;
//   root evaluation t of `times` at main step s: eval_id = (s*times + t + 1) * g_total + m; its line j of root move i
//   is game (eval_id*4 + i)*number + j, first spawn k = 0 (b2048.h, "Root-search line").
#pragma once
#include "engine.cuh"

namespace b2048 {

#define SELFPLAY_MAX_TIMES 64

struct SelfplayParams {
    Tables tb;
    const u64 *starts;  // [G] or NULL; 0 entry -> fresh board
    long G;
    u64 g0, g_total;
    int number, times, mode, max_steps;
    u64 seed;
    u64 *out_boards;      // [G][max_steps] board before each main move
    float *out_results;   // [G][max_steps][4] root values the move was chosen from
    u64 *out_final;       // [G]
    long long *out_score; // [G]
    int *out_steps;       // [G] main moves made
    int *out_status;      // [G] 0 = game over, 1 = stopped at max_steps
    unsigned long long *move_counter;
    int *err_flag;
};

// one fixed-order (L,U,D,R) line from `b` to its end (board.py:192-225 semantics on injected spawns).  Legality first
// (8 byte lookups: rows of the board and of its transpose), then ONE table slide in the chosen direction -- the same
// scheme as K2 variant 1, with the tables read through L1 instead of shared memory.
template <bool NEED_SCORE>
__device__ __forceinline__ void play_line(const Tables &tb, u64 b, u64 key, u32 &score, u32 &count, int *err_flag) {
    score = 0, count = 0;
    const uint8_t *__restrict__ L = tb.legal;
    const uint16_t *__restrict__ T = tb.fin_fwd;
    for (;;) {
        const u64 bt = transpose(b);
        const u32 lo = lo32(b), hi = hi32(b), tlo = lo32(bt), thi = hi32(bt);
        const u32 hm = __ldg(L + (lo & 0xFFFFu)) | __ldg(L + (lo >> 16)) | __ldg(L + (hi & 0xFFFFu)) | __ldg(L + (hi >> 16));
        const u32 vm = __ldg(L + (tlo & 0xFFFFu)) | __ldg(L + (tlo >> 16)) | __ldg(L + (thi & 0xFFFFu)) | __ldg(L + (thi >> 16));
        if (!(hm | vm)) return;
        const u32 d = (hm & 1u) ? 0u : (vm & 1u) ? 1u : (vm & 2u) ? 3u : 2u;  // first legal of L, U, D, R
        u64 x = (d & 1u) ? bt : b;
        const bool rev = (d & 2u) != 0;
        const u64 xm = mirror(x);
        x = rev ? xm : x;
        const u32 xl = lo32(x), xh = hi32(x);
        const u32 r0 = xl & 0xFFFFu, r1 = xl >> 16, r2 = xh & 0xFFFFu, r3 = xh >> 16;
        u64 y = pack64(__ldg(T + r0) | ((u32)__ldg(T + r1) << 16), __ldg(T + r2) | ((u32)__ldg(T + r3) << 16));
        if (NEED_SCORE) score += __ldg(tb.score + r0) + __ldg(tb.score + r1) + __ldg(tb.score + r2) + __ldg(tb.score + r3);
        const u64 ym = mirror(y);
        y = rev ? ym : y;
        const u64 yt = transpose(y);
        y = (d & 1u) ? yt : y;
        if (countzero(y) == 0) {
            *err_flag = 1;
            return;
        }
        u32 four;
        b = spawn_tile(y, rng_draw(key, B2048_INIT_DRAWS + 1 + count), four);
        count++;
    }
}

__global__ void selfplay_fixed_kernel(const SelfplayParams p) {
    extern __shared__ u32 s_lm[];  // mode 1: [4][number] moves of every line
    __shared__ u64 s_board, s_rboard[4];
    __shared__ u32 s_rscore[4];
    __shared__ int s_legal[4], s_min[SELFPLAY_MAX_TIMES][4], s_done;
    __shared__ unsigned long long s_logsum[4];
    __shared__ double s_val[4];
    const int tid = threadIdx.x;
    const long g = blockIdx.x;
    const u64 gid_main = p.g0 + (u64)g;
    const u64 key_main = rng_key(p.seed, gid_main);
    if (tid == 0) {
        const u64 s0 = p.starts ? p.starts[g] : 0ULL;
        s_board = s0 ? s0 : init_board(key_main);
        s_done = 0;
    }
    long long score = 0;  // thread 0
    int step = 0;
    unsigned long long moves_done = 0;
    const int per_eval = 4 * p.number, total = p.times * per_eval;
    for (;;) {
        __syncthreads();
        const u64 board = s_board;
        if (tid < 4) {
            u32 s;
            const u64 f = move_exact(board, tid, p.tb, s);
            s_legal[tid] = f != board;
            s_rscore[tid] = s;
            s_rboard[tid] = f;
            s_logsum[tid] = 0ULL;
            s_val[tid] = 0.0;
        }
        for (int k = tid; k < p.times * 4; k += blockDim.x) s_min[k >> 2][k & 3] = 0x7fffffff;
        __syncthreads();
        for (int it = tid; it < total; it += blockDim.x) {
            const int t = it / per_eval, rem = it - t * per_eval, i = rem / p.number, j = rem - i * p.number;
            if (!s_legal[i]) continue;
            const u64 rb = s_rboard[i];
            if (countzero(rb) == 0) {
                *p.err_flag = 1;
                continue;
            }
            const u64 eval_id = ((u64)step * (u64)p.times + (u64)t + 1ULL) * p.g_total + gid_main;
            const u64 key = rng_key(p.seed, (eval_id * 4ULL + (u64)i) * (u64)p.number + (u64)j);
            u32 four, sc, cnt;
            const u64 first = spawn_tile(rb, rng_draw(key, B2048_INIT_DRAWS), four);
            if (p.mode == B2048_MODE_MEANLOG) play_line<true>(p.tb, first, key, sc, cnt, p.err_flag);
            else play_line<false>(p.tb, first, key, sc, cnt, p.err_flag);
            moves_done += cnt;
            if (p.mode == B2048_MODE_MEANLOG) {
                const double lg = log10((double)sc + (double)s_rscore[i] + 1.0);  // mcts.py:31
                atomicAdd(&s_logsum[i], (unsigned long long)__double2ll_rn(ldexp(lg, 40)));
            } else if (p.mode == B2048_MODE_MEDIAN) {
                s_lm[i * p.number + j] = cnt;
            } else {
                atomicMin(&s_min[t][i], (int)cnt);
            }
        }
        __syncthreads();
        if (p.mode == B2048_MODE_MEDIAN) {
            // mcts.py:37-70: 1-based step at which cumulative deaths reach number//2 = the (h-1)-th smallest move
            // count + 1; rank by (value, index) so exactly one thread owns each rank
            const int h = p.number / 2, target = h ? h - 1 : 0;
            for (int it = tid; it < per_eval; it += blockDim.x) {
                const int i = it / p.number, j = it - i * p.number;
                if (!s_legal[i]) continue;
                const u32 mine = s_lm[i * p.number + j];
                int rank = 0;
                for (int k = 0; k < p.number; k++) {
                    const u32 o = s_lm[i * p.number + k];
                    rank += (o < mine) || (o == mine && k < j);
                }
                if (rank == target) s_val[i] = (double)(mine + 1u);
            }
            __syncthreads();
        }
        if (tid == 0) {
            double val[4];
            for (int i = 0; i < 4; i++) {
                if (p.mode == B2048_MODE_MEANLOG) {
                    val[i] = s_legal[i] ? ldexp((double)s_logsum[i], -40) / (double)p.number : -1.0;
                } else if (p.mode == B2048_MODE_MEDIAN) {
                    val[i] = s_legal[i] ? s_val[i] : -1.0;
                } else {  // mean over `times` of mcts_fixed_min = 1 + min moves (selfplay.py:33-38); 0 if illegal
                    double acc = 0.0;
                    for (int t = 0; t < p.times; t++) acc += s_legal[i] ? (double)(s_min[t][i] + 1) : 0.0;
                    val[i] = acc / (double)p.times;
                }
            }
            int best = 0;
            for (int i = 1; i < 4; i++)
                if (val[i] > val[best]) best = i;  // np.argmax: first maximum
            bool over = !s_legal[best];  // mcts.py:138, selfplay.py:40-42: the chosen move does not move
            if (!over && countzero(s_rboard[best]) == 0) {
                *p.err_flag = 1;
                over = true;
            }
            if (!over && step >= p.max_steps) {
                p.out_status[g] = 1;
                over = true;
            } else if (over) {
                p.out_status[g] = 0;
            }
            if (over) {
                p.out_final[g] = board;
                p.out_score[g] = score;
                p.out_steps[g] = step;
                s_done = 1;
            } else {
                if (p.out_boards) p.out_boards[g * p.max_steps + step] = board;
                if (p.out_results)
                    reinterpret_cast<float4 *>(p.out_results)[g * p.max_steps + step] =
                        make_float4((float)val[0], (float)val[1], (float)val[2], (float)val[3]);
                u32 four;
                s_board = spawn_tile(s_rboard[best], rng_draw(key_main, B2048_INIT_DRAWS + (u64)step), four);
                score += s_rscore[best];
            }
        }
        __syncthreads();
        if (s_done) break;
        step++;
    }
    if (p.move_counter) {
        for (int o = 16; o > 0; o >>= 1) moves_done += __shfl_xor_sync(0xffffffffu, moves_done, o);
        if ((tid & 31) == 0 && moves_done) atomicAdd(p.move_counter, moves_done);
    }
}

}  // namespace b2048
