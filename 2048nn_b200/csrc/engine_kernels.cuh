// engine_kernels.cuh -- K1 (stateless move/spawn/board ops) and K2 (persistent fixed-order
// playout + root-search reductions) for sm_100a.  Included by b2048_capi.cu.
#pragma once
#include "engine.cuh"

namespace b2048 {

// ---------------------------------------------------------------------------------------
// Table build on the device (replaces board.py:134-148).  One thread per row.
// ---------------------------------------------------------------------------------------
__global__ void build_tables_kernel(uint16_t *fin_fwd, uint16_t *fin_rev, uint32_t *score, uint8_t *moved_fwd,
                                    uint8_t *moved_rev, uint8_t *legal, uint2 *rowstat, uint32_t *pair, int *mismatch) {
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= 65536) return;
    u32 ff, sf, fr, sr;
    row_slide(r, false, ff, sf);
    row_slide(r, true, fr, sr);
    fin_fwd[r] = (uint16_t)ff;
    fin_rev[r] = (uint16_t)fr;
    score[r] = sf;
    moved_fwd[r] = ff != r;
    moved_rev[r] = fr != r;
    legal[r] = (uint8_t)((ff != r) | ((fr != r) << 1));
    rowstat[r] = make_uint2(potential((u64)r), tile_sum((u64)r));
    if (r < K3_ROWS) pair[r] = (ff ^ r) | ((fr ^ r) << 16);  // both slides of a row as XOR deltas in one word (K2 variant 2)
    // invariants the playout kernel relies on: equal scores, mirrored finals
    u32 mf, ms;
    row_slide(mirror32(r) & 0xFFFFu, false, mf, ms);
    if (sf != sr || (mirror32(mf) & 0xFFFFu) != fr) atomicAdd(mismatch, 1);
}

// ---------------------------------------------------------------------------------------
// K1 kernels
// ---------------------------------------------------------------------------------------
__global__ void move_batch_kernel(Tables tb, const u64 *__restrict__ boards, const uint8_t *__restrict__ dirs,
                                  int dir_scalar, u64 *__restrict__ out_b, u32 *__restrict__ out_s,
                                  uint8_t *__restrict__ out_m, long n) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long stride = (long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        u64 x = boards[i];
        int d = dirs ? dirs[i] : dir_scalar;
        u32 s;
        u64 f = move_exact(x, d, tb, s);
        if (out_b) out_b[i] = f;
        if (out_s) out_s[i] = s;
        if (out_m) out_m[i] = f != x;
    }
}

__global__ void board_ops_kernel(Tables tb, const u64 *__restrict__ boards, u64 *__restrict__ out_t,
                                 uint8_t *__restrict__ out_cz, uint8_t *__restrict__ out_legal, long n) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long stride = (long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        u64 x = boards[i];
        if (out_t) out_t[i] = transpose(x);
        if (out_cz) out_cz[i] = (uint8_t)countzero(x);
        if (out_legal) {
            int mask = 0;
            for (int d = 0; d < 4; d++) {
                u32 s;
                mask |= (move_exact(x, d, tb, s) != x) << d;
            }
            out_legal[i] = (uint8_t)mask;
        }
    }
}

__global__ void init_boards_kernel(u64 seed, u64 gid_base, u64 *__restrict__ out, long n) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long stride = (long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = init_board(rng_key(seed, gid_base + (u64)i));
}

__global__ void spawn_batch_kernel(const u64 *__restrict__ boards, u64 seed, const u64 *__restrict__ gids,
                                   u64 gid_base, const u32 *__restrict__ k, u32 k_scalar, u64 *__restrict__ out,
                                   int *err_flag, long n) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long stride = (long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        u64 b = boards[i];
        if (countzero(b) == 0) {  // randrange(0): ValueError in the reference (board.py:87)
            out[i] = 0;
            if (err_flag) *err_flag = 1;
            continue;
        }
        u64 key = rng_key(seed, gids ? gids[i] : gid_base + (u64)i);
        u32 four;
        out[i] = spawn_tile(b, rng_draw(key, B2048_INIT_DRAWS + (k ? k[i] : k_scalar)), four);
    }
}

__global__ void rng_draws_kernel(u64 seed, u64 gid_base, u64 first, int count, u64 *__restrict__ out, long n) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long stride = (long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        u64 key = rng_key(seed, gid_base + (u64)i);
        for (int j = 0; j < count; j++) out[i * count + j] = rng_draw(key, first + j);
    }
}

// One BoardArray.move_batch step (board.py:246-277): per game, the first legal direction of
// its own priority row, then generate_tile.  Games that cannot move keep their board.
__global__ void policy_step_kernel(Tables tb, const u64 *__restrict__ boards, const uint8_t *__restrict__ prio,
                                   u64 seed, const u64 *__restrict__ gids, u64 gid_base, const u32 *__restrict__ k,
                                   u64 *__restrict__ out_b, u32 *__restrict__ out_gain, uint8_t *__restrict__ out_moved,
                                   int *err_flag, long n) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long stride = (long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        u64 b = boards[i];
        u32 pr = *reinterpret_cast<const u32 *>(prio + 4 * i);
        u64 nb = b;
        u32 gain = 0;
        bool moved = false;
#pragma unroll 1
        for (int a = 0; a < 4 && !moved; a++) {
            nb = move_exact(b, (pr >> (8 * a)) & 3, tb, gain);
            moved = nb != b;
        }
        if (moved) {
            if (countzero(nb) == 0) {
                *err_flag = 1;
                nb = 0;
            } else {
                u64 key = rng_key(seed, gids ? gids[i] : gid_base + (u64)i);
                u32 four;
                nb = spawn_tile(nb, rng_draw(key, B2048_INIT_DRAWS + k[i]), four);
            }
        } else {
            gain = 0;
        }
        out_b[i] = nb;
        out_gain[i] = gain;
        out_moved[i] = moved;
    }
}

// ---------------------------------------------------------------------------------------
// Training-record side of the path: board encodings and soft targets of
// game_dataset.py:6-48.  Pure HBM streaming: 8 B in, 64 B (exponents) or 1,024 B (one-hot) out per board.
// ---------------------------------------------------------------------------------------
// mode 0: exponents float32[N][16] (game_dataset.py:15-21); mode 1: one-hot float32[N][16 ch][16 pos]
// (game_dataset.py:46-48, mcts_nn.py:25-35).  One float4 per thread, consecutive threads write consecutive
// 16-byte pieces (coalesced 512 B per warp).
__global__ void encode_boards_kernel(const u64 *__restrict__ boards, int mode, float4 *__restrict__ out, long n) {
    const long per = mode ? 64 : 4;  // float4 per board
    const long total = n * per;
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long stride = (long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const long b = i / per;
        const int q = (int)(i - b * per);
        const u64 x = __ldg(boards + b);
        float4 v;
        if (mode) {
            const int c = q >> 2, p0 = (q & 3) * 4;  // channel, first of 4 positions
            const u32 nib = (u32)(x >> (4 * p0)) & 0xFFFFu;
            v.x = ((nib & 0xF) == (u32)c) ? 1.f : 0.f;
            v.y = (((nib >> 4) & 0xF) == (u32)c) ? 1.f : 0.f;
            v.z = (((nib >> 8) & 0xF) == (u32)c) ? 1.f : 0.f;
            v.w = (((nib >> 12) & 0xF) == (u32)c) ? 1.f : 0.f;
        } else {
            const u32 nib = (u32)(x >> (16 * q)) & 0xFFFFu;
            v.x = (float)(nib & 0xF);
            v.y = (float)((nib >> 4) & 0xF);
            v.z = (float)((nib >> 8) & 0xF);
            v.w = (float)((nib >> 12) & 0xF);
        }
        out[i] = v;
    }
}

// game_dataset.py:25-28: softmax(results / rowmax * soft) over the 4 root-move values of a record
__global__ void soft_targets_kernel(const float4 *__restrict__ results, float soft, float4 *__restrict__ out, long n) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long stride = (long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float4 r = results[i];
        const float mx = fmaxf(fmaxf(r.x, r.y), fmaxf(r.z, r.w));
        r.x = r.x / mx * soft, r.y = r.y / mx * soft, r.z = r.z / mx * soft, r.w = r.w / mx * soft;
        const float m2 = fmaxf(fmaxf(r.x, r.y), fmaxf(r.z, r.w));
        const float e0 = expf(r.x - m2), e1 = expf(r.y - m2), e2 = expf(r.z - m2), e3 = expf(r.w - m2);
        const float inv = 1.f / (e0 + e1 + e2 + e3);
        out[i] = make_float4(e0 * inv, e1 * inv, e2 * inv, e3 * inv);
    }
}

// Distribution summary of finished games (notebooks/distribution.py:11-14: [max(get_tiles(board)), score, moves]
// per game, then histograms / mean log score in the notebooks): per-game max tile exponent, and accumulated into
// `acc` (integer atomics only -> the result does not depend on scheduling):
//   acc[0..15]  games by max tile exponent      acc[16] sum of round(log10(score+1) * 2^32)  (2^32: 2^29 games fit)
//   acc[17]     sum of moves                    acc[18] max score            acc[19] games
__global__ void __launch_bounds__(256) game_summary_kernel(const u64 *__restrict__ finals, const u32 *__restrict__ score,
                                                           const u32 *__restrict__ moves, uint8_t *__restrict__ out_maxtile,
                                                           unsigned long long *__restrict__ acc, long n) {
    __shared__ unsigned long long s_acc[20];
    if (threadIdx.x < 20) s_acc[threadIdx.x] = 0ULL;
    __syncthreads();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long stride = (long)gridDim.x * blockDim.x;
    unsigned long long logsum = 0, movsum = 0, games = 0;
    u32 best = 0;
    for (; i < n; i += stride) {
        u64 x = finals[i];
        // max nibble: fold 16 -> 1 with byte/nibble-wise maxima
        u32 m = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const u32 t = (u32)(x >> (4 * k)) & 0xFu;
            m = t > m ? t : m;
        }
        if (out_maxtile) out_maxtile[i] = (uint8_t)m;
        atomicAdd(&s_acc[m], 1ULL);
        const u32 sc = score[i];
        logsum += (unsigned long long)__double2ll_rn(ldexp(log10((double)sc + 1.0), 32));
        movsum += moves ? moves[i] : 0u;
        best = sc > best ? sc : best;
        games++;
    }
    for (int o = 16; o > 0; o >>= 1) {
        logsum += __shfl_xor_sync(0xffffffffu, logsum, o);
        movsum += __shfl_xor_sync(0xffffffffu, movsum, o);
        games += __shfl_xor_sync(0xffffffffu, games, o);
        const u32 ob = __shfl_xor_sync(0xffffffffu, best, o);
        best = ob > best ? ob : best;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_acc[16], logsum);
        atomicAdd(&s_acc[17], movsum);
        atomicMax(&s_acc[18], (unsigned long long)best);
        atomicAdd(&s_acc[19], games);
    }
    __syncthreads();
    if (threadIdx.x < 20 && s_acc[threadIdx.x]) {
        if (threadIdx.x == 18) atomicMax(acc + 18, s_acc[18]);
        else atomicAdd(acc + threadIdx.x, s_acc[threadIdx.x]);
    }
}

// root moves of G origins: legal flag and merge score (mcts.py:21); the lines recompute the moved board themselves
__global__ void root_moves_kernel(Tables tb, const u64 *__restrict__ origins, long G, uint8_t *__restrict__ legal,
                                  u32 *__restrict__ rscore) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G * 4) return;
    u64 x = origins[i >> 2];
    u32 s;
    u64 f = move_exact(x, (int)(i & 3), tb, s);
    legal[i] = f != x;
    rscore[i] = s;
}

// ---------------------------------------------------------------------------------------
// K2: persistent fixed-order playout kernel.
//
// grid = #SMs, block = 1024 threads, one game per thread, 128 KB forward row table in shared
// memory.  Each lane runs a flat loop {refill if needed; one move attempt cascade; spawn};
// finished lanes refill from a global work counter (warp-aggregated atomic) so a warp never
// waits for its longest game.  The L,U,D,R cascade is specialised: Up/Down work on the
// transposed board (computed only when Left fails), Right/Down reuse the forward table through
// mirror().  Scores come from the potential() closed form while no 32768+32768 annihilation is
// possible; otherwise the game is finished on the exact table path.
// ---------------------------------------------------------------------------------------
struct PlayParams {
    Tables tb;
    // work definition
    long n;                 // number of games (work items)
    const u64 *start;       // FRESH/START mode: start boards (NULL or 0 entry -> fresh board)
    const u64 *origins;     // ROOT mode: [G] root boards (line r = 4*g + i starts from move(origins[g], i))
    const uint8_t *rlegal;  // ROOT mode: [G*4]
    const u32 *rscore;      // ROOT mode: [G*4]
    int root_mode;
    int number, line_begin, line_count;  // ROOT mode: lines [line_begin, line_begin+line_count) of `number`
    u64 seed, gid_base;                  // ROOT mode: gid_base = eval_id_base*4
    u32 spawn0;
    uint8_t order[4];
    // outputs
    u64 *out_final;
    u32 *out_score;
    u32 *out_moves;
    long long *out_logsum;
    int *out_count;
    int *out_minmov;
    // scratch
    unsigned long long *counter;
    int *err_flag;
};

// generic direction (any order): uniform code, no table for the reverse direction
template <typename TablePtr>
__device__ __forceinline__ u64 move_dir(u64 b, int d, TablePtr T) {
    u64 x = (d & 1) ? transpose(b) : b;
    x = (d & 2) ? slide_right(x, T) : slide_left(x, T);
    return (d & 1) ? transpose(x) : x;
}

// Finish a game on the exact path (global tables, per-move scores): used when a 32768+32768
// annihilation is possible, i.e. never in practice but required for bit-exactness.
struct ExactResult {
    u64 b;
    u32 count, score;
    int err;
};
__device__ __noinline__ ExactResult finish_exact(const PlayParams &p, u64 b, u64 key, u32 count, u32 score) {
    ExactResult r;
    r.err = 0;
    for (;;) {
        bool adv = false;
#pragma unroll 1
        for (int a = 0; a < 4; a++) {
            u32 s;
            u64 f = move_exact(b, p.order[a], p.tb, s);
            if (f != b) {
                if (countzero(f) == 0) {  // total annihilation -> randrange(0) in the reference
                    r.err = 1;
                    adv = false;
                    break;
                }
                u32 four;
                b = spawn_tile(f, rng_draw(key, B2048_INIT_DRAWS + p.spawn0 + count), four);
                score += s;
                count++;
                adv = true;
                break;
            }
        }
        if (!adv) break;
    }
    r.b = b;
    r.count = count;
    r.score = score;
    return r;
}

template <bool LUDR>
__global__ void __launch_bounds__(1024, 1) playout_fixed_kernel(const PlayParams p) {
    extern __shared__ __align__(16) uint16_t sT[];
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.tb.fin_fwd);
        uint4 *dst = reinterpret_cast<uint4 *>(sT);
        for (int i = threadIdx.x; i < 65536 * 2 / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();

    const unsigned lane = threadIdx.x & 31;
    const long total_threads = (long)gridDim.x * blockDim.x;
    long item = (long)blockIdx.x * blockDim.x + threadIdx.x;  // first game: static assignment
    bool first = true;

    bool have = false;
    u64 b = 0, key = 0;
    u32 count = 0, n4 = 0, pot0 = 0, limit = 0, rs = 0;
    long slot = 0;  // output index of the current game

    for (;;) {
        if (!have) {
            // ---- refill: lanes here may be any subset of the warp --------------------------
            if (!first) {
                unsigned m = __activemask();
                int leader = __ffs(m) - 1;
                unsigned long long base = 0;
                if ((int)lane == leader) base = atomicAdd(p.counter, (unsigned long long)__popc(m));
                base = __shfl_sync(m, base, leader);
                item = total_threads + (long)base + __popc(m & ((1u << lane) - 1));
            }
            first = false;
            if (item >= p.n) break;
            slot = item;
            count = 0;
            n4 = 0;
            rs = 0;
            bool skip = false;
            if (p.root_mode) {
                long r = item / p.line_count;  // root index g*4+i
                int j = p.line_begin + (int)(item - r * p.line_count);
                slot = r * p.number + j;
                if (!p.rlegal[r]) {
                    skip = true;
                } else {
                    key = rng_key(p.seed, (p.gid_base + (u64)r) * (u64)p.number + (u64)j);
                    u32 rs_;
                            u64 rb = move_exact(p.origins[r >> 2], (int)(r & 3), p.tb, rs_);  // mcts.py:21
                    rs = p.rscore[r];
                    if (countzero(rb) == 0) {
                        *p.err_flag = 1;
                        skip = true;
                    } else {
                        u32 four;
                        b = spawn_tile(rb, rng_draw(key, B2048_INIT_DRAWS), four);  // mcts.py:23
                    }
                }
            } else {
                key = rng_key(p.seed, p.gid_base + (u64)item);
                u64 s0 = p.start ? p.start[item] : 0ULL;
                b = s0 ? s0 : init_board(key);  // board.py:205 `if not board`
            }
            if (skip) {
                if (p.out_final) p.out_final[slot] = 0;
                if (p.out_score) p.out_score[slot] = 0;
                if (p.out_moves) p.out_moves[slot] = 0;
                continue;
            }
            pot0 = potential(b);
            u32 ts = tile_sum(b);
            limit = ts >= 65536u ? 0u : ((65536u - ts + 3u) >> 2);  // moves that cannot annihilate
            have = true;
        }

        // ---- one move: first legal direction of the priority order ----------------------
        u64 nb;
        bool moved;
        if (LUDR) {
            nb = slide_left(b, sT);  // Left
            moved = nb != b;
            if (!moved) {
                u64 bt = transpose(b);
                u64 nt = slide_left(bt, sT);  // Up
                if (nt == bt) nt = slide_right(bt, sT);  // Down
                if (nt != bt) {
                    nb = transpose(nt);
                    moved = true;
                } else {
                    nb = slide_right(b, sT);  // Right
                    moved = nb != b;
                }
            }
        } else {
            moved = false;
            nb = b;
#pragma unroll 1
            for (int a = 0; a < 4 && !moved; a++) {
                nb = move_dir(b, p.order[a], sT);
                moved = nb != b;
            }
        }

        if (moved && count < limit) {
            u32 four;
            b = spawn_tile(nb, rng_draw(key, B2048_INIT_DRAWS + p.spawn0 + count), four);
            n4 += four;
            count++;
            continue;
        }

        // ---- game over (or hand-over to the exact path) ----------------------------------
        u32 score = potential(b) - pot0 - 4u * n4;
        if (moved) {  // count >= limit: annihilation possible from here on
            ExactResult r = finish_exact(p, b, key, count, score);
            b = r.b;
            count = r.count;
            score = r.score;
            if (r.err) *p.err_flag = 1;
        }
        if (p.out_final) p.out_final[slot] = b;
        if (p.out_score) p.out_score[slot] = score;
        if (p.out_moves) p.out_moves[slot] = count;
        if (p.root_mode) {
            long r = slot / p.number;
            if (p.out_logsum) {
                double t = log10((double)score + (double)rs + 1.0);  // mcts.py:31
                long long q = __double2ll_rn(ldexp(t, 40));
                atomicAdd(reinterpret_cast<unsigned long long *>(p.out_logsum + r), (unsigned long long)q);
            }
            if (p.out_count) atomicAdd(p.out_count + r, 1);
            if (p.out_minmov) atomicMin(p.out_minmov + r, (int)count);
        }
        have = false;
    }
}

// ---------------------------------------------------------------------------------------
// K2 variant 2: legality first, then ONE uniform slide per move (no L,U,D,R attempt cascade).
//
// The cascade kernel above executes the Up/Down/Right attempts for a whole warp whenever any lane
// needs them, with 27% / 7% / 2% of the lanes active (ncu: 14.8 of 32 threads per instruction).
// Here every lane does the same work every step: 8 byte lookups in a 64 KB row-legality table (4 rows
// of the board, 4 rows of its transpose) give the legal-move mask, the first legal direction of the
// priority order is selected branch-free, and a single table slide is applied through
// select / mirror / transpose steps that are predicated by the chosen direction.
// Shared memory: 128 KB forward table + 64 KB legality table.
// ---------------------------------------------------------------------------------------
#define K2V2_SMEM_BYTES (131072 + 65536)
#ifndef K2_REFILL_MIN
#define K2_REFILL_MIN 3
#endif

template <bool LUDR>
__global__ void __launch_bounds__(1024, 1) playout_fixed_kernel2(const PlayParams p) {
    extern __shared__ __align__(16) uint16_t sT[];
    uint8_t *sL = reinterpret_cast<uint8_t *>(sT) + 131072;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.tb.fin_fwd);
        uint4 *dst = reinterpret_cast<uint4 *>(sT);
        for (int i = threadIdx.x; i < 131072 / 16; i += blockDim.x) dst[i] = __ldg(src + i);
        const uint4 *srcl = reinterpret_cast<const uint4 *>(p.tb.legal);
        uint4 *dstl = reinterpret_cast<uint4 *>(sL);
        for (int i = threadIdx.x; i < 65536 / 16; i += blockDim.x) dstl[i] = __ldg(srcl + i);
    }
    __syncthreads();

    const unsigned lane = threadIdx.x & 31;
    const long total_threads = (long)gridDim.x * blockDim.x;
    long item = (long)blockIdx.x * blockDim.x + threadIdx.x;  // first game: static assignment
    bool first = true, exhausted = false;
    bool have = false;
    u64 b = 0, key = 0;
    u32 count = 0, n4 = 0, pot0 = 0, limit = 0, rs = 0;
    long slot = 0;
    const u32 o0 = p.order[0], o1 = p.order[1], o2 = p.order[2], o3 = p.order[3];

    // All 32 lanes stay in the loop until the warp has no game left.  Starting a game costs a few
    // hundred instructions for the whole warp, so dead lanes wait until K2_REFILL_MIN of them can
    // be refilled together (ncu: per-lane refill left only 22 of 32 lanes active per instruction).
    for (;;) {
        unsigned alive_mask = __ballot_sync(0xffffffffu, have);
        if (__popc(~alive_mask) >= K2_REFILL_MIN || alive_mask == 0u) {
            if (!have && !exhausted) {
                if (!first) {
                    unsigned m = __activemask();
                    int leader = __ffs(m) - 1;
                    unsigned long long base = 0;
                    if ((int)lane == leader) base = atomicAdd(p.counter, (unsigned long long)__popc(m));
                    base = __shfl_sync(m, base, leader);
                    item = total_threads + (long)base + __popc(m & ((1u << lane) - 1));
                }
                first = false;
                if (item >= p.n) {
                    exhausted = true;
                } else {
                    slot = item;
                    count = 0;
                    n4 = 0;
                    rs = 0;
                    bool skip = false;
                    if (p.root_mode) {
                        long r = item / p.line_count;
                        int j = p.line_begin + (int)(item - r * p.line_count);
                        slot = r * p.number + j;
                        if (!p.rlegal[r]) {
                            skip = true;
                        } else {
                            key = rng_key(p.seed, (p.gid_base + (u64)r) * (u64)p.number + (u64)j);
                            u32 rs_;
                            u64 rb = move_exact(p.origins[r >> 2], (int)(r & 3), p.tb, rs_);  // mcts.py:21
                            rs = p.rscore[r];
                            if (countzero(rb) == 0) {
                                *p.err_flag = 1;
                                skip = true;
                            } else {
                                u32 four;
                                b = spawn_tile(rb, rng_draw(key, B2048_INIT_DRAWS), four);
                            }
                        }
                    } else {
                        key = rng_key(p.seed, p.gid_base + (u64)item);
                        u64 s0 = p.start ? p.start[item] : 0ULL;
                        b = s0 ? s0 : init_board(key);
                    }
                    if (skip) {
                        if (p.out_final) p.out_final[slot] = 0;
                        if (p.out_score) p.out_score[slot] = 0;
                        if (p.out_moves) p.out_moves[slot] = 0;
                    } else {
                        const uint2 st = board_stat(b, p.tb);
                        pot0 = st.x;
                        limit = st.y >= 65536u ? 0u : ((65536u - st.y + 3u) >> 2);  // moves that cannot annihilate
                        have = true;
                    }
                }
            }
            alive_mask = __ballot_sync(0xffffffffu, have);
            if (alive_mask == 0u && __all_sync(0xffffffffu, exhausted)) break;
        }
        if (!have) continue;

        // ---- legality of all four directions ---------------------------------------------------------
        const u64 bt = transpose(b);
        const u32 lo = lo32(b), hi = hi32(b), tlo = lo32(bt), thi = hi32(bt);
        const u32 hm = sL[lo & 0xFFFFu] | sL[lo >> 16] | sL[hi & 0xFFFFu] | sL[hi >> 16];      // bit0 L, bit1 R
        const u32 vm = sL[tlo & 0xFFFFu] | sL[tlo >> 16] | sL[thi & 0xFFFFu] | sL[thi >> 16];  // bit0 U, bit1 D
        const bool alive = (hm | vm) != 0;
        if (alive && count < limit) {
            u32 d;
            if (LUDR) {
                d = (hm & 1u) ? 0u : (vm & 1u) ? 1u : (vm & 2u) ? 3u : 2u;
            } else {
                const u32 mask = (hm & 1u) | ((vm & 1u) << 1) | ((hm & 2u) << 1) | ((vm & 2u) << 2);
                d = ((mask >> o0) & 1u) ? o0 : ((mask >> o1) & 1u) ? o1 : ((mask >> o2) & 1u) ? o2 : o3;
            }
            // ---- one slide, direction folded into data movement ------------------------------------
            u64 x = (d & 1u) ? bt : b;
            const bool rev = (d & 2u) != 0;
            u64 xm = mirror(x);
            x = rev ? xm : x;
            u64 y = slide_left(x, sT);
            u64 ym = mirror(y);
            y = rev ? ym : y;
            u64 yt = transpose(y);
            y = (d & 1u) ? yt : y;
            u32 four;
            b = spawn_tile(y, rng_draw(key, B2048_INIT_DRAWS + p.spawn0 + count), four);
            n4 += four;
            count++;
            continue;
        }

        // ---- game over (or hand-over to the exact path) ----------------------------------------------
        u32 score = board_stat(b, p.tb).x - pot0 - 4u * n4;
        if (alive) {
            ExactResult r = finish_exact(p, b, key, count, score);
            b = r.b;
            count = r.count;
            score = r.score;
            if (r.err) *p.err_flag = 1;
        }
        if (p.out_final) p.out_final[slot] = b;
        if (p.out_score) p.out_score[slot] = score;
        if (p.out_moves) p.out_moves[slot] = count;
        if (p.root_mode) {
            long r = slot / p.number;
            if (p.out_logsum) {
                double t = log10((double)score + (double)rs + 1.0);
                long long q = __double2ll_rn(ldexp(t, 40));
                atomicAdd(reinterpret_cast<unsigned long long *>(p.out_logsum + r), (unsigned long long)q);
            }
            if (p.out_count) atomicAdd(p.out_count + r, 1);
            if (p.out_minmov) atomicMin(p.out_minmov + r, (int)count);
        }
        have = false;
    }
}

// ---------------------------------------------------------------------------------------
// K2 variant 2 (default): the pair table.
//
// ncu on variant 1 (profiles/r02_k2_playout_fixed_ncu.txt): ALU pipe 88 % busy, shared-memory wavefronts 86 % of peak,
// 174 executed thread-instructions per move -- the kernel is bound by the NUMBER of ALU-pipe instructions and of
// shared-memory lookups per move.  This variant removes both kinds of work instead of rescheduling it:
//  * one u32 table entry per row holds BOTH slides as XOR deltas (slide toward nibble 0 ^ row in the low half, toward
//    nibble 3 in the high half), so 8 lookups (4 rows of the board, 4 rows of its transpose) give all four candidate
//    boards: no legality table (a direction is legal iff the OR of its four deltas is nonzero), no mirror(), no second
//    round of lookups for the chosen direction -- 8 LDS.32 per move instead of 8 LDS.U8 + 4 LDS.U16;
//  * the table covers rows < K3_ROWS (224 KB of shared memory, filled by cp.async.bulk); a larger row index needs a tile
//    sum >= 16384, and a game hands over to the exact global-table path before that is possible (the hand-over that
//    already guards the 32768+32768 annihilation, so it costs nothing per move);
//  * games are prepared 32 at a time by the whole warp (a pool of one stash entry per lane, handed out by shuffles) and
//    finished games are written out in batches, so the divergent start / finish code runs at full lane occupancy;
//  * K3_UNROLL moves per pass of the outer loop (the refill vote is paid once per pass), the spawn's draw counter folded
//    into the key (two IMADs per draw), transposes as 4-instruction delta swaps, a 3-instruction empty mask.
// Bit-exact with variants 0 and 1 (tests/test_engine_gpu.py).
// ---------------------------------------------------------------------------------------
#define K2V3_SMEM_BYTES (K3_ROWS * 4)
#ifndef K3_UNROLL
#define K3_UNROLL 8
#endif
#ifndef K3_REFILL_MIN
#define K3_REFILL_MIN 2
#endif

__device__ __forceinline__ void k3_bulk_load(void *smem_dst, const void *gsrc, u32 bytes, unsigned long long *bar) {
    const u32 b = (u32)__cvta_generic_to_shared(bar), d = (u32)__cvta_generic_to_shared(smem_dst);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        for (u32 off = 0; off < bytes; off += 32768u) {
            const u32 n = bytes - off < 32768u ? bytes - off : 32768u;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d + off),
                         "l"((const char *)gsrc + off), "r"(n), "r"(b)
                         : "memory");
        }
    }
    __syncthreads();  // the barrier is initialised before anyone polls it
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "K3_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0, 0x989680;\n\t"
        "@P1 bra K3_DONE;\n\t"
        "bra K3_WAIT;\n\t"
        "K3_DONE:\n\t"
        "}" ::"r"(b)
        : "memory");
}

// byte offsets of the two rows of a 32-bit half in the pair table
__device__ __forceinline__ u32 k3_off_lo(u32 h) { return (h << 2) & 0x3FFFCu; }
__device__ __forceinline__ u32 k3_off_hi(u32 h) { return (h >> 14) & 0x3FFFCu; }  // (IMAD.HI or PRMT+IMAD instead of SHF+LOP3: measured slower / same)
__device__ __forceinline__ u32 k3_ld(u32 smem_base, u32 off) {
    u32 v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_base + off));  // read-only after k3_bulk_load
    return v;
}

// One move of one game on the pair table.  Returns false when the game cannot go on here: the board is dead, or the
// move budget of the table path is used up (`count >= limit`: tile sum may reach K3_TILESUM_MAX).
template <bool LUDR>
__device__ __forceinline__ bool k3_step(u32 sbase, u64 &b, u64 keyc, u32 &count, u32 &n4, u32 limit, u32 o0, u32 o1, u32 o2,
                                        u32 o3) {
    if (count >= limit) return false;
    const u64 bt = transpose(b);
    const u32 lo = lo32(b), hi = hi32(b), tlo = lo32(bt), thi = hi32(bt);
    const u32 e0 = k3_ld(sbase, k3_off_lo(lo)), e1 = k3_ld(sbase, k3_off_hi(lo));
    const u32 e2 = k3_ld(sbase, k3_off_lo(hi)), e3 = k3_ld(sbase, k3_off_hi(hi));
    const u32 c0 = k3_ld(sbase, k3_off_lo(tlo)), c1 = k3_ld(sbase, k3_off_hi(tlo));
    const u32 c2 = k3_ld(sbase, k3_off_lo(thi)), c3 = k3_ld(sbase, k3_off_hi(thi));
    // entries hold (slide ^ row): a direction is legal iff some row's delta is nonzero
    const u32 he = e0 | e1 | e2 | e3, ve = c0 | c1 | c2 | c3;
    const bool lL = (he & 0xFFFFu) != 0, lR = (he & 0xFFFF0000u) != 0;
    const bool lU = (ve & 0xFFFFu) != 0, lD = (ve & 0xFFFF0000u) != 0;
    if ((he | ve) == 0) return false;
    bool hfwd, vfwd, vert;
    if (LUDR) {  // Left, else Up, else Down, else Right
        hfwd = lL, vfwd = lU;
        vert = !lL & (lU | lD);
    } else {
        const u32 mask = (u32)lL | ((u32)lU << 1) | ((u32)lR << 2) | ((u32)lD << 3);
        const u32 d = ((mask >> o0) & 1u) ? o0 : ((mask >> o1) & 1u) ? o1 : ((mask >> o2) & 1u) ? o2 : o3;
        hfwd = vfwd = (d & 2u) == 0;
        vert = (d & 1u) != 0;
    }
    const u32 hsel = hfwd ? 0x5410u : 0x7632u, vsel = vfwd ? 0x5410u : 0x7632u;
    const u32 hl = lo ^ __byte_perm(e0, e1, hsel), hh = hi ^ __byte_perm(e2, e3, hsel);
    const u32 vl = tlo ^ __byte_perm(c0, c1, vsel), vh = thi ^ __byte_perm(c2, c3, vsel);
    const u64 tv = transpose(pack64(vl, vh));
    const u64 y = vert ? tv : pack64(hl, hh);
    u32 four;
    b = spawn_tile(y, mix64(keyc + (u64)count * B2048_GOLDEN), four);  // = rng_draw(key, INIT_DRAWS + spawn0 + count)
    n4 += four;
    count++;
    return true;
}

template <bool LUDR>
__global__ void __launch_bounds__(1024, 1) playout_fixed_kernel3(const PlayParams p) {
    extern __shared__ __align__(128) u32 sP[];
    __shared__ unsigned long long s_bar;
    k3_bulk_load(sP, p.tb.pair, K2V3_SMEM_BYTES, &s_bar);
    const u32 sbase = (u32)__cvta_generic_to_shared(sP);

    const unsigned lane = threadIdx.x & 31;
    const long total_threads = (long)gridDim.x * blockDim.x;
    bool first = true;
    bool have = false, fin = false;  // fin: the lane holds a finished game whose results are not written yet
    u64 b = 0, keyc = 0;
    u32 count = 0, n4 = 0, pot0 = 0, limit = 0, rs = 0;
    long slot = 0;
    u64 sb = 0, skeyc = 0;  // this lane's entry of the warp's pool of prepared games
    u32 spot0 = 0, slimit = 0, srs = 0;
    long sslot = 0;
    int pool = 0;            // warp-uniform: entries 0 .. pool-1 of the pool are still to be handed out
    bool pool_done = false;  // warp-uniform: the launch has no further games
    const u32 o0 = p.order[0], o1 = p.order[1], o2 = p.order[2], o3 = p.order[3];
    const u64 key_off = (u64)(B2048_INIT_DRAWS + p.spawn0 + 1u) * B2048_GOLDEN;  // keyc = key + key_off

    // All 32 lanes stay in the loop until the warp has no game left; dead lanes wait until K3_REFILL_MIN of them can write
    // their results and take a new game together.  New games come from a per-warp pool of 32 prepared games (one stash
    // entry per lane, in registers): preparing a game (key, start board, potential, move budget: a few hundred
    // instructions and four L2 loads) is done by all 32 lanes at once whenever the pool runs dry, not by the few dead lanes.
    for (;;) {
        unsigned alive_mask = __ballot_sync(0xffffffffu, have);
        if (__popc(~alive_mask) >= K3_REFILL_MIN || alive_mask == 0u) {
            if (fin) {
                // ---- game over, or hand-over to the exact path (which also finds a dead board dead) ----------
                u32 score = board_stat(b, p.tb).x - pot0 - 4u * n4;
                if (count >= limit) {
                    ExactResult r = finish_exact(p, b, keyc - key_off, count, score);
                    b = r.b;
                    count = r.count;
                    score = r.score;
                    if (r.err) *p.err_flag = 1;
                }
                if (p.out_final) p.out_final[slot] = b;
                if (p.out_score) p.out_score[slot] = score;
                if (p.out_moves) p.out_moves[slot] = count;
                if (p.root_mode) {
                    long r = slot / p.number;
                    if (p.out_logsum) {
                        double t = log10((double)score + (double)rs + 1.0);  // mcts.py:31
                        long long q = __double2ll_rn(ldexp(t, 40));
                        atomicAdd(reinterpret_cast<unsigned long long *>(p.out_logsum + r), (unsigned long long)q);
                    }
                    if (p.out_count) atomicAdd(p.out_count + r, 1);
                    if (p.out_minmov) atomicMin(p.out_minmov + r, (int)count);
                }
                fin = false;
            }
            // Hand out prepared games; whenever the pool runs dry ALL 32 lanes prepare one game each (key, start board,
            // potential, move budget) at full lane occupancy, instead of the few dead lanes preparing their own.
            for (;;) {
                const unsigned need = __ballot_sync(0xffffffffu, !have);
                if (need == 0u) break;
                if (pool == 0) {
                    if (pool_done) break;
                    long item;
                    if (first) {
                        item = (long)blockIdx.x * blockDim.x + threadIdx.x;  // first games: static assignment
                        first = false;
                    } else {
                        unsigned long long base = 0;
                        if (lane == 0) base = atomicAdd(p.counter, 32ULL);
                        base = __shfl_sync(0xffffffffu, base, 0);
                        item = total_threads + (long)base + lane;
                    }
                    const bool valid = item < p.n;  // items grow with the lane: the valid lanes are 0 .. V-1
                    const unsigned vm = __ballot_sync(0xffffffffu, valid);
                    if (vm != 0xffffffffu) pool_done = true;  // the launch has no further games
                    if (vm == 0u) break;
                    sb = 0;  // 0 = no game in this stash entry (skipped root move)
                    if (valid) {
                        sslot = item;
                        srs = 0;
                        bool skip = false;
                        u64 key, nb = 0;
                        if (p.root_mode) {
                            long r = item / p.line_count;
                            int j = p.line_begin + (int)(item - r * p.line_count);
                            sslot = r * p.number + j;
                            key = rng_key(p.seed, (p.gid_base + (u64)r) * (u64)p.number + (u64)j);
                            if (!p.rlegal[r]) {
                                skip = true;
                            } else {
                                u32 rs_;
                                u64 rb = move_exact(p.origins[r >> 2], (int)(r & 3), p.tb, rs_);  // mcts.py:21
                                srs = p.rscore[r];
                                if (countzero(rb) == 0) {
                                    *p.err_flag = 1;
                                    skip = true;
                                } else {
                                    u32 four;
                                    nb = spawn_tile(rb, rng_draw(key, B2048_INIT_DRAWS), four);  // mcts.py:23
                                }
                            }
                        } else {
                            key = rng_key(p.seed, p.gid_base + (u64)item);
                            u64 s0 = p.start ? p.start[item] : 0ULL;
                            nb = s0 ? s0 : init_board(key);  // board.py:205 `if not board`
                        }
                        skeyc = key + key_off;
                        if (skip) {
                            if (p.out_final) p.out_final[sslot] = 0;
                            if (p.out_score) p.out_score[sslot] = 0;
                            if (p.out_moves) p.out_moves[sslot] = 0;
                        } else {
                            const uint2 st = board_stat(nb, p.tb);
                            spot0 = st.x;
                            // moves that can neither annihilate (tile sum < 65536) nor leave the pair table (tile sum < 16384)
                            slimit = st.y >= K3_TILESUM_MAX ? 0u : ((K3_TILESUM_MAX - st.y + 3u) >> 2);
                            sb = nb;
                        }
                    }
                    pool = __popc(vm);
                }
                // the j-th lane in need takes stash entry pool-1-j
                const int src = pool - 1 - __popc(need & ((1u << lane) - 1u));
                const int sl = src < 0 ? 0 : src;
                const u64 tb_ = __shfl_sync(0xffffffffu, sb, sl), tk = __shfl_sync(0xffffffffu, skeyc, sl);
                const u32 tp = __shfl_sync(0xffffffffu, spot0, sl), tl = __shfl_sync(0xffffffffu, slimit, sl);
                const u32 tr = __shfl_sync(0xffffffffu, srs, sl);
                const long tsl = __shfl_sync(0xffffffffu, sslot, sl);
                if (!have && src >= 0 && tb_ != 0) {
                    b = tb_, keyc = tk, pot0 = tp, limit = tl, rs = tr, slot = tsl;
                    count = 0, n4 = 0;
                    have = true;
                }
                const int taken = __popc(need);
                pool = pool > taken ? pool - taken : 0;
            }
            alive_mask = __ballot_sync(0xffffffffu, have);
            if (alive_mask == 0u && pool_done && pool == 0) break;

        }
        if (!have) continue;

        bool going = k3_step<LUDR>(sbase, b, keyc, count, n4, limit, o0, o1, o2, o3);
#pragma unroll
        for (int u = 1; u < K3_UNROLL; u++)
            if (going) going = k3_step<LUDR>(sbase, b, keyc, count, n4, limit, o0, o1, o2, o3);
        if (going) continue;
        // the game is over here (or must be handed over to the exact path): its results are written in the batched block
        // above, together with the other finished lanes' and the refill
        have = false;
        fin = true;
    }
}

}  // namespace b2048
