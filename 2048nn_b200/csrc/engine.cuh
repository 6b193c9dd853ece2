// engine.cuh -- device-side 2048 board engine for sm_100a.
//
// One game per thread; the board is one nibble-packed uint64 held in registers.  Every
// function here is a from-scratch CUDA formulation of the reference semantics cited beside it
// (file:line into fqjin/2048NN); parity is checked bit-for-bit by tests/test_engine_gpu.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2048 {

typedef unsigned long long u64;
typedef uint32_t u32;

// ---------------------------------------------------------------------------------------
// Row tables (board.py:104-148).  Global copies live in the context; the playout kernels
// stage the forward table in shared memory once per CTA.
//   fin_fwd / fin_rev : u16[65536] final row for a slide toward nibble 0 / nibble 3
//   score             : u32[65536] merge score of the row (same for both directions)
// `moved` is not stored: moved == (final != row) (board.py:131).
// ---------------------------------------------------------------------------------------
struct Tables {
    const uint16_t *fin_fwd;
    const uint16_t *fin_rev;
    const uint32_t *score;
    const uint8_t *legal;  // u8[65536]: bit 0 = row moves toward nibble 0, bit 1 = toward nibble 3
    const uint2 *rowstat;  // [65536] per row: .x = sum (t-1)*2^t ("potential"), .y = sum 2^t (tile sum)
    const uint32_t *pair;  // u32[K3_ROWS]: (fin_fwd ^ row) | (fin_rev ^ row) << 16 for the rows whose nibble 3 is below 14
};
// K2 variant 2 keeps `pair` in shared memory: 0xE000 rows x 4 B = 224 KB of the 227 KB a CTA may have.  A row index
// >= K3_ROWS needs a tile 2^14 or 2^15 in board column 3 / row 3, i.e. a tile sum >= 16384; games hand over to the
// exact global-table path before that can happen (the same hand-over that guards the 32768+32768 annihilation).
#define K3_ROWS 0xE000u
#define K3_TILESUM_MAX 16384u

// board.py:104-132 move_row restated for one thread (used only to BUILD the tables on device).
__device__ __forceinline__ void row_slide(u32 row, bool rev, u32 &fin, u32 &score) {
    u32 out = 0, sc = 0, pending = 0;
    int n_out = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        u32 t = (row >> (4 * (rev ? 3 - i : i))) & 0xF;
        if (t == 0) continue;
        if (pending == t) {
            u32 nt = t + 1;
            sc += 1u << nt;
            out |= (nt & 0xF) << (4 * (rev ? 3 - n_out : n_out));  // 15+15 -> 0 keeps its slot
            n_out++;
            pending = 0;
        } else {
            if (pending) {
                out |= pending << (4 * (rev ? 3 - n_out : n_out));
                n_out++;
            }
            pending = t;
        }
    }
    if (pending) out |= pending << (4 * (rev ? 3 - n_out : n_out));
    fin = out;
    score = sc;
}

// ---------------------------------------------------------------------------------------
// Bit tricks on the two 32-bit halves (rows 0,1 in lo; rows 2,3 in hi).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ u32 lo32(u64 x) { return (u32)x; }
__device__ __forceinline__ u32 hi32(u64 x) { return (u32)(x >> 32); }
__device__ __forceinline__ u64 pack64(u32 lo, u32 hi) { return ((u64)hi << 32) | lo; }

// board.py:46-54 transpose.  Stage 1 swaps nibbles inside each 2x2 block and never leaves a
// 32-bit half; stage 2 swaps the off-diagonal 2x2 blocks, which is a byte permutation of the
// two halves (PRMT): bytes [B0..B7] -> [B0,B4,B2,B6 | B1,B5,B3,B7].
// (written as a delta swap in PTX lop3 so that it stays 4 instructions -- SHF, LOP3, IMAD.SHL, LOP3 -- one of them off the
// ALU pipe; the C expression is re-associated into 5)
template <int LUT>
__device__ __forceinline__ u32 lop3(u32 a, u32 b, u32 c) {
    u32 r;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return r;
}
__device__ __forceinline__ u32 tr_stage1(u32 h) {
    const u32 t = lop3<0x28>(h, h >> 12, 0x0000F0F0u);  // (h ^ (h >> 12)) & mask
    return lop3<0x96>(h, t, t << 12);                   // h ^ t ^ (t << 12)
}
__device__ __forceinline__ u64 transpose(u64 x) {
    u32 a = tr_stage1(lo32(x)), b = tr_stage1(hi32(x));
    return pack64(__byte_perm(a, b, 0x6240), __byte_perm(a, b, 0x7351));
}

// mirror: reverse the 4 nibbles of every 16-bit row (col c <-> col 3-c).
__device__ __forceinline__ u32 mirror32(u32 h) {
    u32 s = ((h & 0x0F0F0F0Fu) << 4) | ((h >> 4) & 0x0F0F0F0Fu);
    return __byte_perm(s, 0, 0x2301);
}
__device__ __forceinline__ u64 mirror(u64 x) { return pack64(mirror32(lo32(x)), mirror32(hi32(x))); }

// bit 4k set when nibble k is empty
__device__ __forceinline__ u32 empty_mask32(u32 h) {
    u32 z = h | (h >> 2);
    z |= z >> 1;
    return ~z & 0x11111111u;
}
// board.py:57-65 countzero INCLUDING its wrap: 16 empties -> 0
__device__ __forceinline__ int countzero(u64 x) {
    return (__popc(empty_mask32(lo32(x))) + __popc(empty_mask32(hi32(x)))) & 15;
}

// ---------------------------------------------------------------------------------------
// Counter-based per-game RNG (boundary spec in include/b2048.h).
// ---------------------------------------------------------------------------------------
#define B2048_GOLDEN 0x9E3779B97F4A7C15ULL
#define B2048_INIT_DRAWS 8

__device__ __forceinline__ u64 mix64(u64 z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ u64 rng_key(u64 seed, u64 gid) {
    return mix64(mix64(seed + B2048_GOLDEN) ^ (gid * 0xD1342543DE82EF95ULL + 1ULL));
}
__device__ __forceinline__ u64 rng_draw(u64 key, u64 i) { return mix64(key + (i + 1ULL) * B2048_GOLDEN); }
__device__ __forceinline__ u32 rng_word(u64 key, u32 w) {
    u64 d = rng_draw(key, w >> 1);
    return (w & 1) ? lo32(d) : hi32(d);
}

// ---------------------------------------------------------------------------------------
// generate_tile (board.py:85-101): position = randrange(#empty) counted from nibble 0 upward,
// then randrange(10) != 0 -> exponent 1 (a 2) else exponent 2 (a 4).
// `draw` is the 64-bit draw of this spawn: hi word -> position, lo word -> ten.
// Precondition: 1 <= #empty <= 15 (callers route 0/16 to the error path).
// Selecting the pos-th empty cell: prefix counts of the empty mask by one multiply
// (mask * 0x11111111 puts #empties in nibbles 0..k into nibble k), then a SWAR compare.
// Written for the ALU pipe, which bounds the playout kernels: empty mask in 3 instructions (the add may issue on the FMA
// pipe), the 2-or-4 choice folded into the final shift, the tile ORed into its half by two predicated instructions.
// ---------------------------------------------------------------------------------------
// bit 4k+3 set when nibble k is empty: (h & 7..7) + 7..7 carries into bit 3 of every nonzero nibble (3 instructions,
// the add can go to the FMA pipe)
__device__ __forceinline__ u32 empty_mask32_hi(u32 h) {
    return lop3<0x02>((h & 0x77777777u) + 0x77777777u, h, 0x88888888u);  // ~(a | b) & c in one LOP3
}
__device__ __forceinline__ u64 spawn_tile(u64 board, u64 draw, u32 &is_four) {
    u32 lo = lo32(board), hi = hi32(board);
    const u32 zl = empty_mask32_hi(lo), zh = empty_mask32_hi(hi);
    const u32 cl = __popc(zl);
    const u32 n = cl + __popc(zh);
    u32 pos = __umulhi(hi32(draw), n);
    const bool four = __umulhi(lo32(draw), 10u) == 0u;
    is_four = four ? 1u : 0u;
    const bool in_hi = pos >= cl;
    const u32 z = (in_hi ? zh : zl) >> 3;
    pos -= in_hi ? cl : 0u;
    const u32 t = (z + 7u - pos) * 0x11111111u;  // nibble k >= 8  <=>  #empties in nibbles 0..k > pos
    const u32 m = t & 0x88888888u;
    const u32 tile = (m & (0u - m)) >> (four ? 2u : 3u);
    asm("{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.u32 p, %2, 0;\n\t"
        "@p or.b32 %0, %0, %3;\n\t"
        "@!p or.b32 %1, %1, %3;\n\t"
        "}"
        : "+r"(hi), "+r"(lo)
        : "r"((u32)in_hi), "r"(tile));  // two predicated ORs instead of two selects and two ORs
    return pack64(lo, hi);
}

// generate_init_tiles (board.py:68-82) on the game's word stream (words 0..15 reserved).
__device__ u64 init_board_slow(u64 key);
__device__ __forceinline__ u64 init_board(u64 key) {
    // common case (15/16): pos1 != pos2 on the first try -> words 0..3 = two draws
    const u64 d0 = rng_draw(key, 0), d1 = rng_draw(key, 1);
    const u32 pos1 = __umulhi(hi32(d0), 16u), pos2 = __umulhi(lo32(d0), 16u);
    if (pos1 == pos2) return init_board_slow(key);
    const u64 t1 = 1ULL << (4 * pos1 + (__umulhi(hi32(d1), 10u) ? 0 : 1));
    const u64 t2 = 1ULL << (4 * pos2 + (__umulhi(lo32(d1), 10u) ? 0 : 1));
    return t1 | t2;
}
__device__ __noinline__ u64 init_board_slow(u64 key) {
    u32 w = 0;
    u32 pos1 = __umulhi(rng_word(key, w++), 16u);
    u32 pos2 = __umulhi(rng_word(key, w++), 16u);
    while (pos1 == pos2) {
        if (w >= 2 * B2048_INIT_DRAWS - 2) {
            pos2 = (pos1 + 1) & 15;
            break;
        }
        pos2 = __umulhi(rng_word(key, w++), 16u);
    }
    u64 t1 = __umulhi(rng_word(key, w++), 10u) ? (1ULL << (4 * pos1)) : (1ULL << (4 * pos1 + 1));
    u64 t2 = __umulhi(rng_word(key, w++), 10u) ? (1ULL << (4 * pos2)) : (1ULL << (4 * pos2 + 1));
    return t1 | t2;
}

// ---------------------------------------------------------------------------------------
// Moves.
// ---------------------------------------------------------------------------------------
// Slide all four rows toward nibble 0 with a u16 table (shared or global memory).
template <typename TablePtr>
__device__ __forceinline__ u64 slide_left(u64 x, TablePtr T) {
    u32 lo = lo32(x), hi = hi32(x);
    u32 f0 = T[lo & 0xFFFFu], f1 = T[lo >> 16], f2 = T[hi & 0xFFFFu], f3 = T[hi >> 16];
    return pack64(f0 | (f1 << 16), f2 | (f3 << 16));
}
// A slide toward nibble 3 is a mirrored slide toward nibble 0:
// merge_table_rev[x] == mirror(merge_table[mirror(x)]) (checked exhaustively in the tests).
template <typename TablePtr>
__device__ __forceinline__ u64 slide_right(u64 x, TablePtr T) {
    return mirror(slide_left(mirror(x), T));
}

// Exact, fully general move (board.py:151-189) with score, from the global tables.
__device__ __forceinline__ u64 move_exact(u64 x, int dir, const Tables &tb, u32 &score) {
    const uint16_t *T = (dir & 2) ? tb.fin_rev : tb.fin_fwd;
    u64 src = (dir & 1) ? transpose(x) : x;
    u32 lo = lo32(src), hi = hi32(src);
    u32 r0 = lo & 0xFFFFu, r1 = lo >> 16, r2 = hi & 0xFFFFu, r3 = hi >> 16;
    u32 f0 = __ldg(T + r0), f1 = __ldg(T + r1), f2 = __ldg(T + r2), f3 = __ldg(T + r3);
    score = __ldg(tb.score + r0) + __ldg(tb.score + r1) + __ldg(tb.score + r2) + __ldg(tb.score + r3);
    u64 fin = pack64(f0 | (f1 << 16), f2 | (f3 << 16));
    return (dir & 1) ? transpose(fin) : fin;
}

// Sum over tiles of (t-1)*2^t (t = exponent >= 1).  For a game without a 32768+32768
// annihilation:  score gained = potential(final) - potential(start) - 4 * (#spawned 4-tiles),
// because a tile 2^t assembled from 2-tiles has paid (t-1)*2^t in merges and every spawned 4
// skips one merge worth 4.  The playout kernel uses this closed form while no annihilation is
// possible (tile sum < 65536) and falls back to table scores otherwise.
__device__ __forceinline__ u32 potential(u64 b) {
    u32 s = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        u32 t = (u32)(b >> (4 * k)) & 0xF;
        s += t ? ((t - 1u) << t) : 0u;
    }
    return s;
}
// potential and tile sum of a board from the per-row table (4 L2-resident loads instead of two
// 16-step loops: this runs for a whole warp whenever one lane starts or finishes a game)
__device__ __forceinline__ uint2 board_stat(u64 b, const Tables &tb) {
    u32 lo = lo32(b), hi = hi32(b);
    uint2 a = __ldg(tb.rowstat + (lo & 0xFFFFu)), c = __ldg(tb.rowstat + (lo >> 16));
    uint2 d = __ldg(tb.rowstat + (hi & 0xFFFFu)), e = __ldg(tb.rowstat + (hi >> 16));
    return make_uint2(a.x + c.x + d.x + e.x, a.y + c.y + d.y + e.y);
}
__device__ __forceinline__ u32 tile_sum(u64 b) {
    u32 s = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        u32 t = (u32)(b >> (4 * k)) & 0xF;
        s += t ? (1u << t) : 0u;
    }
    return s;
}

}  // namespace b2048
