// convnet_tc.cuh -- c64b3 forward on 5th-gen tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Replaces ConvNet.forward (network.py:78-87) + the inline one-hot encoder (mcts_nn.py:25-35)
// for a "tile-set" of 32 boards per CTA pass, all 8 layers on chip:
//
//   * Implicit GEMM without im2col.  M rows of one MMA = 4 board-rows y x 32 boards for ONE
//     board column x ("column slab"); activations live in shared memory as fp16, one 128-byte
//     row (64 channels, K-major, SWIZZLE_128B) per (x, y, board), slabs separated by 32 zero
//     rows.  A 3x3 tap (dy,dx) is then just a descriptor base: slab x+dx, start row
//     (1+dy)*32 -- the zero rows supply the y padding, out-of-range columns are skipped, so
//     no shifted copies and no masked loads.
//   * One A operand feeds up to three output columns: for input slab s and a given dy,
//     D[x=s-1 | x=s | x=s+1] += A(s,dy) * [W(dy,+1); W(dy,0); W(dy,-1)]^T is ONE
//     tcgen05.mma with N = 192 (128 at the borders) into adjacent TMEM column blocks.
//   * Accumulators: TMEM, 4 column blocks x 64 fp32 columns (one per board column x).
//   * Weights (fp16, BatchNorm folded) stream from L2 as 24 KB pre-swizzled chunks
//     [W(dy,+1); W(dy,0); W(dy,-1)] through a cp.async.bulk + mbarrier ring.
//   * Epilogue warps: tcgen05.ld -> +bias (+residual) -> ReLU -> fp16 -> swizzled st.shared
//     as the next layer's A operand; the head (1x1 conv, FC) runs in fp32 on CUDA cores.
//   * L0 (one-hot input, 16 channels) is made exact on the tensor cores by splitting its folded
//     weights into fp16 hi + lo and feeding the one-hot row twice along K (K = 32).
//
// Warp roles (320 threads): warp 0 = weight producer, warp 1 = MMA issuer (+TMEM owner),
// warps 2..9 = workers (encode boards, epilogues, head).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2048 {
namespace tc {

// ---- geometry --------------------------------------------------------------------------------
constexpr int TILE_BOARDS = 32;                    // boards per tile-set
constexpr int ROW_BYTES = 128;                     // 64 fp16 channels
constexpr int SLAB_ROWS = 32;                      // one (x, y) group: 32 boards
constexpr int ACT_ROWS = 21 * SLAB_ROWS;           // [Z][y0..y3] x 4 slabs + trailing Z
constexpr int ACT_BYTES = ACT_ROWS * ROW_BYTES;    // 86,016
constexpr int CHUNK_ROWS = 192;                    // [dx=+1 | dx=0 | dx=-1] x 64 co
constexpr int CHUNK_BYTES = CHUNK_ROWS * ROW_BYTES;  // 24,576
constexpr int N_LAYERS = 7;                        // L0 + 6 residual-block convs
constexpr int N_CHUNKS = N_LAYERS * 3;
constexpr int RING = 2;
constexpr int TMEM_COLS = 256;

// blob: [21 chunks fp16, pre-swizzled] [fp32 tail: bias 7x64 | w7 4x64 | b7 4 | fcw 4x64 | fcb 4]
constexpr int BLOB_W_BYTES = N_CHUNKS * CHUNK_BYTES;
constexpr int TAIL_BIAS = 0, TAIL_W7 = 7 * 64, TAIL_B7 = TAIL_W7 + 256, TAIL_FCW = TAIL_B7 + 4,
              TAIL_FCB = TAIL_FCW + 256, TAIL_FLOATS = TAIL_FCB + 4;
constexpr int BLOB_BYTES = BLOB_W_BYTES + TAIL_FLOATS * 4;

// shared memory map (offsets from a 1024-aligned base)
constexpr int SM_ACT_A = 0;
constexpr int SM_ACT_B = SM_ACT_A + ACT_BYTES;
constexpr int SM_RING = SM_ACT_B + ACT_BYTES;            // 172,032 (1024-aligned)
constexpr int SM_TAIL = SM_RING + RING * CHUNK_BYTES;    // fp32 tail copy
constexpr int SM_BOARDS = SM_TAIL + TAIL_FLOATS * 4;     // 32 x u64
constexpr int SM_LOGITS = SM_BOARDS + 32 * 8;            // 32 x 4 f32
constexpr int SM_BARS = SM_LOGITS + 32 * 16;             // mbarriers
constexpr int SM_TOTAL = SM_BARS + 16 * RING + 16 + 16;  // ring full/empty, acc, act, tmem base
constexpr int SMEM_BYTES = SM_TOTAL + 1024;              // + alignment slack

constexpr int HEAD_STRIDE = 65;  // floats per board in the head staging area (bank-conflict padding)

// ---- PTX helpers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16 (fp16 operands, fp32 accumulate), issued by one thread
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 columns of fp32 -> 32 registers per thread (thread t <-> TMEM lane base+t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 [0,14), LBO>>4 [16,30) = 0, SBO>>4 [32,46) = 1024 B (8 rows x 128 B), version 1 [46,48),
// layout SWIZZLE_128B = 2 [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | (1ULL << 46) | (2ULL << 61);
}
// kind::f16 instruction descriptor: D fp32 (bit 4), A/B fp16 K-major, N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// byte offset of 16-byte chunk c of activation row R (SWIZZLE_128B: chunk ^= row % 8)
__device__ __forceinline__ uint32_t act_off(int R, int c) { return (uint32_t)R * ROW_BYTES + (uint32_t)((c ^ (R & 7)) << 4); }
__device__ __forceinline__ int act_row(int x, int y, int b) { return (5 * x + 1 + y) * SLAB_ROWS + b; }

struct FwdParams {
    const uint8_t *blob;             // BLOB_BYTES image
    const unsigned long long *boards;  // N boards
    float *out_logp;                 // N x 4 log-probs (may be NULL)
    float *out_logits;               // N x 4 pre-softmax (may be NULL)
    long n;
    // debug: dump the fp32 post-activation output of layer dbg_layer as [board][co][pos]
    float *dbg;
    int dbg_layer;
};

// The CTA-wide forward of one tile-set.  Called by ALL 320 threads; `s_boards` holds the 32
// boards (0 for unused slots), logits land in `s_logits` [32][4].
// Pipeline state (ring slot / phases) persists across calls in `st`.
struct PipeState {
    uint32_t ring_prod;  // chunks produced
    uint32_t ring_cons;  // chunks consumed
    uint32_t layer_phase;  // parity of acc_full / act_ready
};

__device__ __forceinline__ void forward_tile(uint8_t *sm, uint32_t tmem_base, const uint8_t *blob, PipeState &st,
                                             float *dbg, int dbg_layer, long dbg_base, int dbg_n) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sm_base = smem_u32(sm);
    const uint32_t bar_full0 = sm_base + SM_BARS;          // ring full  [RING]
    const uint32_t bar_empty0 = bar_full0 + 8 * RING;      // ring empty [RING]
    const uint32_t bar_acc = bar_empty0 + 8 * RING;        // accumulators complete
    const uint32_t bar_act = bar_acc + 8;                  // layer input ready (8 worker warps)
    const float *tail = reinterpret_cast<const float *>(sm + SM_TAIL);
    const unsigned long long *s_boards = reinterpret_cast<const unsigned long long *>(sm + SM_BOARDS);
    float *s_logits = reinterpret_cast<float *>(sm + SM_LOGITS);

    if (warp == 0) {
        // ===== weight producer: 21 chunks per tile-set through the ring ========================
        if (lane == 0) {
            for (int c = 0; c < N_CHUNKS; c++) {
                uint32_t slot = st.ring_prod % RING, use = st.ring_prod / RING;
                if (use > 0) mbar_wait(bar_empty0 + 8 * slot, (use - 1) & 1);
                mbar_expect_tx(bar_full0 + 8 * slot, CHUNK_BYTES);
                bulk_g2s(sm_base + SM_RING + slot * CHUNK_BYTES, blob + (size_t)c * CHUNK_BYTES, CHUNK_BYTES,
                         bar_full0 + 8 * slot);
                st.ring_prod++;
            }
        }
        st.ring_prod = __shfl_sync(0xffffffffu, st.ring_prod, 0);
    } else if (warp == 1) {
        // ===== MMA issuer ===========================================================================
        uint32_t phase = st.layer_phase;
        for (int l = 0; l < N_LAYERS; l++) {
            // layer input: l even (incl. L0, whose input is the one-hot image) reads B, l odd reads A
            const uint32_t in_addr = sm_base + ((l & 1) ? SM_ACT_A : SM_ACT_B);
            mbar_wait(bar_act, phase & 1);  // workers finished writing this layer's input
            tc_fence_after();
            const int nk = (l == 0) ? 2 : 4;  // K steps of 16 channels (L0: one-hot x [hi | lo])
            for (int dyi = 0; dyi < 3; dyi++) {
                uint32_t slot = st.ring_cons % RING, use = st.ring_cons / RING;
                mbar_wait(bar_full0 + 8 * slot, use & 1);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t w_addr = sm_base + SM_RING + slot * CHUNK_BYTES;
#pragma unroll 1
                    for (int si = 0; si < 4; si++) {
                        // slab order 0,3,1,2: the k==0 MMAs of slabs 0 and 3 (dy=-1) are the first writers
                        // of column blocks {0,1} and {2,3}, so they overwrite; everything else accumulates.
                        const int s = (si == 0) ? 0 : (si == 1) ? 3 : si - 1;
                        const int x_lo = s > 0 ? s - 1 : 0, x_hi = s < 3 ? s + 1 : 3;
                        const int N = (x_hi - x_lo + 1) * 64;   // 128 at the borders, 192 inside
                        const int brow = (s == 0) ? 64 : 0;     // rows [dx=+1|dx=0|dx=-1]: drop dx=+1 for s==0
                        const uint32_t idesc = make_idesc(128, N);
                        const uint32_t a_addr = in_addr + (uint32_t)((5 * s + dyi) * SLAB_ROWS) * ROW_BYTES;
                        const uint32_t b_addr = w_addr + (uint32_t)brow * ROW_BYTES;
                        const uint32_t d_addr = tmem_base + (uint32_t)(x_lo * 64);
                        for (int k = 0; k < nk; k++) {
                            const uint32_t acc = (dyi == 0 && k == 0 && si < 2) ? 0u : 1u;
                            umma_f16(d_addr, make_desc(a_addr + k * 32), make_desc(b_addr + k * 32), idesc, acc);
                        }
                    }
                    umma_commit(bar_empty0 + 8 * slot);  // ring slot reusable once these MMAs retire
                }
                __syncwarp();
                st.ring_cons++;
            }
            if (lane == 0) umma_commit(bar_acc);
            __syncwarp();
            phase++;
        }
        st.layer_phase = phase;
    } else {
        // ===== workers: 8 warps, thread <-> (y = warp%4 TMEM quarter, board = lane), 2 x-tiles ======
        const int q = warp & 3;                 // TMEM lane quarter this warp may read == board row y
        const int xg = (warp - 2) >> 2;         // 0: x in {0,1}; 1: x in {2,3}
        uint32_t phase = st.layer_phase;
        // ---- encode: one-hot rows of buffer B (channels [0,16) and again [16,32) for the hi/lo split)
        {
            const unsigned long long brd = s_boards[lane];
            for (int xi = 0; xi < 2; xi++) {
                const int x = xg * 2 + xi, y = q;
                const int t = (int)((brd >> (4 * (y * 4 + x))) & 0xF);
                const int R = act_row(x, y, lane);
                uint4 lo8 = make_uint4(0, 0, 0, 0), hi8 = make_uint4(0, 0, 0, 0);
                const uint32_t one = 0x3C00u << ((t & 1) * 16);  // fp16 1.0 in the right half-word
                uint32_t *dst = (t < 8) ? &lo8.x : &hi8.x;
                dst[(t & 7) >> 1] = one;
                *reinterpret_cast<uint4 *>(sm + SM_ACT_B + act_off(R, 0)) = lo8;
                *reinterpret_cast<uint4 *>(sm + SM_ACT_B + act_off(R, 1)) = hi8;
                *reinterpret_cast<uint4 *>(sm + SM_ACT_B + act_off(R, 2)) = lo8;
                *reinterpret_cast<uint4 *>(sm + SM_ACT_B + act_off(R, 3)) = hi8;
            }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_act);

        float head_acc[2][4];
        for (int l = 0; l < N_LAYERS; l++) {
            uint8_t *out = sm + (((l & 1) == 0) ? SM_ACT_A : SM_ACT_B);  // l=0 ->A, l=1 ->B, l=2 ->A ...
            const bool residual = (l >= 2) && ((l & 1) == 0);           // second conv of a block writes A in place
            const bool last = (l == N_LAYERS - 1);
            mbar_wait(bar_acc, phase & 1);
            tc_fence_after();
            for (int xi = 0; xi < 2; xi++) {
                const int x = xg * 2 + xi, y = q;
                const int R = act_row(x, y, lane);
                float v[64];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(x * 64), reinterpret_cast<uint32_t *>(v));
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(x * 64 + 32),
                          reinterpret_cast<uint32_t *>(v + 32));
                tmem_ld_wait();
                const float *bias = tail + TAIL_BIAS + l * 64;
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    float r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    if (residual) {
                        uint4 xr = *reinterpret_cast<const uint4 *>(out + act_off(R, c));
                        const __half2 *h = reinterpret_cast<const __half2 *>(&xr);
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            float2 f = __half22float2(h[j]);
                            r[2 * j] = f.x, r[2 * j + 1] = f.y;
                        }
                    }
                    uint4 pk;
                    __half2 *hp = reinterpret_cast<__half2 *>(&pk);
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        float a = fmaxf(v[c * 8 + 2 * j] + bias[c * 8 + 2 * j] + r[2 * j], 0.f);
                        float b2 = fmaxf(v[c * 8 + 2 * j + 1] + bias[c * 8 + 2 * j + 1] + r[2 * j + 1], 0.f);
                        v[c * 8 + 2 * j] = a, v[c * 8 + 2 * j + 1] = b2;
                        hp[j] = __floats2half2_rn(a, b2);
                    }
                    if (!last) *reinterpret_cast<uint4 *>(out + act_off(R, c)) = pk;
                }
                if (dbg && dbg_layer == l && lane < dbg_n) {
                    float *o = dbg + ((dbg_base + lane) * 64) * 16 + (y * 4 + x);
#pragma unroll
                    for (int co = 0; co < 64; co++) o[co * 16] = v[co];
                }
                if (last) {  // head 1x1 conv 64->4 on the fp32 values of this (board, position)
#pragma unroll
                    for (int c4 = 0; c4 < 4; c4++) {
                        float s = tail[TAIL_B7 + c4];
                        const float *w = tail + TAIL_W7 + c4 * 64;
#pragma unroll
                        for (int ci = 0; ci < 64; ci++) s = fmaf(v[ci], w[ci], s);
                        head_acc[xi][c4] = fmaxf(s, 0.f);
                    }
                }
            }
            tc_fence_before();
            if (!last) {
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_act);
            }
            phase++;
        }
        st.layer_phase = phase;
        // ---- head: stage [board][c4*16 + pos] in buffer B's first data slab, then FC by warp 2 -------
        float *hb = reinterpret_cast<float *>(sm + SM_ACT_B + SLAB_ROWS * ROW_BYTES);
        for (int xi = 0; xi < 2; xi++) {
            const int pos = q * 4 + xg * 2 + xi;
#pragma unroll
            for (int c4 = 0; c4 < 4; c4++) hb[lane * HEAD_STRIDE + c4 * 16 + pos] = head_acc[xi][c4];
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 worker warps
        if (warp == 2) {
            float z[4];
#pragma unroll
            for (int o = 0; o < 4; o++) {
                float s = tail[TAIL_FCB + o];
                const float *w = tail + TAIL_FCW + o * 64;
                for (int j = 0; j < 64; j++) s = fmaf(hb[lane * HEAD_STRIDE + j], w[j], s);
                z[o] = s;
            }
            *reinterpret_cast<float4 *>(s_logits + lane * 4) = make_float4(z[0], z[1], z[2], z[3]);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
    }
}

// ---- CTA setup / teardown shared by the forward and playout kernels ---------------------------
__device__ __forceinline__ uint8_t *tc_setup(uint8_t *smem_raw, const uint8_t *blob, uint32_t &tmem_base) {
    uint8_t *sm = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sm_base = smem_u32(sm);
    const int tid = threadIdx.x;
    // zero both activation buffers (the zero slabs stay zero for the kernel's lifetime)
    for (int i = tid; i < 2 * ACT_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4 *>(sm)[i] = make_uint4(0, 0, 0, 0);
    const float *gtail = reinterpret_cast<const float *>(blob + BLOB_W_BYTES);
    for (int i = tid; i < TAIL_FLOATS; i += blockDim.x) reinterpret_cast<float *>(sm + SM_TAIL)[i] = gtail[i];
    if (tid == 0) {
        const uint32_t bars = sm_base + SM_BARS;
        for (int i = 0; i < RING; i++) {
            mbar_init(bars + 8 * i, 1);           // full
            mbar_init(bars + 8 * (RING + i), 1);  // empty
        }
        mbar_init(bars + 16 * RING, 1);      // acc complete (tcgen05.commit)
        mbar_init(bars + 16 * RING + 8, 8);  // layer input ready (8 worker warps)
        fence_mbar_init();
    }
    if ((tid >> 5) == 1) tmem_alloc(sm_base + SM_BARS + 16 * RING + 16, TMEM_COLS);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    tmem_base = *reinterpret_cast<volatile uint32_t *>(sm + SM_BARS + 16 * RING + 16);
    return sm;
}
__device__ __forceinline__ void tc_teardown(uint32_t tmem_base) {
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// Forward-only kernel (config 5): boards u64[N] -> log-probs f32[N][4].
__global__ void __launch_bounds__(320, 1) convnet_tc_forward_kernel(const FwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint32_t tmem_base;
    uint8_t *sm = tc_setup(smem_raw, p.blob, tmem_base);
    PipeState st = {0, 0, 0};
    unsigned long long *s_boards = reinterpret_cast<unsigned long long *>(sm + SM_BOARDS);
    const float *s_logits = reinterpret_cast<const float *>(sm + SM_LOGITS);
    const long tiles = (p.n + TILE_BOARDS - 1) / TILE_BOARDS;
    for (long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long base = tile * TILE_BOARDS;
        if (threadIdx.x < TILE_BOARDS) s_boards[threadIdx.x] = (base + threadIdx.x < p.n) ? p.boards[base + threadIdx.x] : 0ULL;
        __syncthreads();
        int dbg_n = (int)((p.n - base) < TILE_BOARDS ? (p.n - base) : TILE_BOARDS);
        forward_tile(sm, tmem_base, p.blob, st, p.dbg, p.dbg_layer, base, dbg_n);
        __syncthreads();
        if (threadIdx.x < TILE_BOARDS && base + threadIdx.x < p.n) {
            float4 z = reinterpret_cast<const float4 *>(s_logits)[threadIdx.x];
            if (p.out_logits) reinterpret_cast<float4 *>(p.out_logits)[base + threadIdx.x] = z;
            if (p.out_logp) {
                float m = fmaxf(fmaxf(z.x, z.y), fmaxf(z.z, z.w));
                float l = m + logf(expf(z.x - m) + expf(z.y - m) + expf(z.z - m) + expf(z.w - m));
                reinterpret_cast<float4 *>(p.out_logp)[base + threadIdx.x] = make_float4(z.x - l, z.y - l, z.z - l, z.w - l);
            }
        }
        __syncthreads();
    }
    tc_teardown(tmem_base);
}

}  // namespace tc
}  // namespace b2048
