// train_tc.cuh -- "mixed" mode of the training step: the 13 forward / data-gradient 3x3 convolutions of a step on
// the tcgen05 tensor cores (16-bit operands, fp32 accumulation in TMEM); BatchNorm, the head, the weight gradient
// and Adam stay fp32 (train.cuh).  Same implicit GEMM as the inference kernel (convnet_tc.cuh: column slabs, zero
// slabs for the y padding, [dx=+1 | 0 | -1] weight chunks, two MMA-issuing warps over accumulator halves), but one
// layer per launch -- BatchNorm needs the whole batch between two convolutions -- and one tile of 32 boards per CTA:
//   * the operand image of a tile is produced in global memory already in shared-memory order by the elementwise
//     kernel before it (bn_apply / bn_bwd_apply / one-hot encoder): 16 slabs x 32 boards x 128 B, SWIZZLE_128B
//     pre-applied, so the loader is 4 bulk copies (one per board column) + 1 bulk copy of the layer's 3 chunks;
//   * forward operands are fp16 (activations are O(1) after BatchNorm+ReLU), data-gradient operands bf16 (the
//     gradient's range, ~1e-8..1e-3, does not fit fp16 without scaling);
//   * the epilogue writes the raw fp32 accumulators ([board][16 positions][64 channels]) for the fp32 kernels.
#pragma once
#include <cuda_bf16.h>

#include "convnet_tc.cuh"

namespace b2048 {
namespace train {

constexpr int TCL_THREADS = 256;  // warps 0..3 epilogue (TMEM lane quarters), 4 loader, 5/6 MMA issuers, 7 idle
constexpr int TCL_IMG_BYTES = 16 * tc::SLAB_ROWS * tc::ROW_BYTES;  // 65,536: data slabs of one tile in global memory
constexpr int TCL_W_BYTES = 3 * tc::CHUNK_BYTES;                   // the layer's three (dy) weight chunks
constexpr int TCL_SM_W = tc::ACT_BYTES;
constexpr int TCL_SM_BARS = TCL_SM_W + TCL_W_BYTES;
constexpr int TCL_SMEM_BYTES = TCL_SM_BARS + 64 + 1024;

// element (board b of the tile, position p, channel ch) of a tile image: byte offset inside the tile
__device__ __forceinline__ size_t tile_image_offset(int b, int p, int ch) {
    const int x = p & 3, y = p >> 2;
    const int row = (x * 4 + y) * tc::SLAB_ROWS + b;             // data slabs only, board-column major
    const int chunk = (ch >> 3) ^ (b & 7);                       // SWIZZLE_128B: row index % 8 == b % 8 in shared memory
    return (size_t)row * tc::ROW_BYTES + (size_t)chunk * 16 + (size_t)(ch & 7) * 2;
}

// STATS: per-tile per-channel sum / sum of squares of the outputs (BatchNorm batch statistics) to part[tile][2][64]: the
// epilogue threads own rows, so the tile is staged in shared memory (the operand image and weights are dead once the
// accumulators are committed) and summed by columns in a fixed order.
constexpr int TCL_STAGE_STRIDE = 65;  // floats per staged row: column sums walk consecutive banks
static_assert((512 * TCL_STAGE_STRIDE + 4 * 2 * 64) * 4 <= TCL_SM_BARS, "the staged tile must fit in the operand region");

template <bool BF16, bool STATS>
__global__ void __launch_bounds__(TCL_THREADS, 1)
tc_conv_layer_kernel(const uint8_t *__restrict__ image, const uint8_t *__restrict__ wchunks, float *__restrict__ out,
                     float *__restrict__ part, long B, int nk) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *sm = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sm_base = tc::smem_u32(sm);
    const uint32_t bar_full = sm_base + TCL_SM_BARS, bar_acc = bar_full + 8, tmem_slot = bar_full + 16;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long tile = blockIdx.x;
    // zero slabs [Z] of the shared image: slabs 0, 5, 10, 15, 20
    for (int i = tid; i < 5 * tc::SLAB_ROWS * tc::ROW_BYTES / 16; i += TCL_THREADS) {
        const int z = i / (tc::SLAB_ROWS * tc::ROW_BYTES / 16), r = i % (tc::SLAB_ROWS * tc::ROW_BYTES / 16);
        reinterpret_cast<uint4 *>(sm + (size_t)z * 5 * tc::SLAB_ROWS * tc::ROW_BYTES)[r] = make_uint4(0, 0, 0, 0);
    }
    if (tid == 0) {
        tc::mbar_init(bar_full, 1);
        tc::mbar_init(bar_acc, 2);
        tc::fence_mbar_init();
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(sm + TCL_SM_BARS + 16);
    float *stage = reinterpret_cast<float *>(sm);

    if (warp == 4) {
        if (lane == 0) {
            tc::mbar_expect_tx(bar_full, TCL_IMG_BYTES + TCL_W_BYTES);
            const uint8_t *src = image + (size_t)tile * TCL_IMG_BYTES;
#pragma unroll
            for (int x = 0; x < 4; x++)  // board column x: 4 data slabs (y) behind that column's zero slab
                tc::bulk_g2s(sm_base + (uint32_t)((5 * x + 1) * tc::SLAB_ROWS * tc::ROW_BYTES), src + (size_t)x * 4 * tc::SLAB_ROWS * tc::ROW_BYTES,
                             4 * tc::SLAB_ROWS * tc::ROW_BYTES, bar_full);
            tc::bulk_g2s(sm_base + TCL_SM_W, wchunks, TCL_W_BYTES, bar_full);
        }
    } else if (warp == 5 || warp == 6) {
        const int half = warp - 5;
        const uint64_t desc_hi = ((uint64_t)(1024 >> 4) << 32) | (1ULL << 46) | (2ULL << 61);
        tc::mbar_wait(bar_full, 0);
        tc::tc_fence_after();
#pragma unroll 1
        for (int dyi = 0; dyi < 3; dyi++)
            tc::issue_half(sm_base >> 4, (sm_base + TCL_SM_W + dyi * tc::CHUNK_BYTES) >> 4, tmem_base, desc_hi, dyi, nk, half, BF16);
        tc::umma_commit_elect(bar_acc);
    } else if (warp < 4) {
        // thread = (board row y = warp, board = lane): TMEM lane y*32 + board, column block x holds position (y, x)
        tc::mbar_wait(bar_acc, 0);  // both accumulator halves are committed: every MMA has read its operands
        tc::tc_fence_after();
        const long board = tile * tc::TILE_BOARDS + lane;
        const bool live = board < B;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
        for (int x = 0; x < 4; x++) {
            float *o = out + (board * 16 + warp * 4 + x) * 64;
            float *srow = stage + (size_t)((x * 4 + warp) * 32 + lane) * TCL_STAGE_STRIDE;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                float v[8];
                tc::tmem_ld8(lane_addr + (uint32_t)(x * 64 + c * 8), reinterpret_cast<uint32_t *>(v));
                tc::tmem_ld_wait();
                if (live) {
                    *reinterpret_cast<float4 *>(o + c * 8) = make_float4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<float4 *>(o + c * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
                }
                if (STATS) {
#pragma unroll
                    for (int k = 0; k < 8; k++) srow[c * 8 + k] = live ? v[k] : 0.f;  // rows past the batch end count as 0
                }
            }
        }
        tc::tc_fence_before();
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
    if (STATS) {
        // 256 threads = 4 row quarters x 64 channels; quarter sums added in a fixed order
        float *red = stage + 512 * TCL_STAGE_STRIDE;  // [4][2][64] behind the staged tile
        const int c = tid & 63, q = tid >> 6;
        float s1 = 0.f, s2 = 0.f;
        const float *col = stage + (size_t)(q * 128) * TCL_STAGE_STRIDE + c;
#pragma unroll 8
        for (int r = 0; r < 128; r++) {
            const float v = col[(size_t)r * TCL_STAGE_STRIDE];
            s1 += v;
            s2 = fmaf(v, v, s2);
        }
        red[(q * 2 + 0) * 64 + c] = s1;
        red[(q * 2 + 1) * 64 + c] = s2;
        __syncthreads();
        if (tid < 128) {
            const int which = tid >> 6;
            part[(size_t)tile * 128 + tid] = (red[(0 * 2 + which) * 64 + c] + red[(1 * 2 + which) * 64 + c]) +
                                             (red[(2 * 2 + which) * 64 + c] + red[(3 * 2 + which) * 64 + c]);
        }
    }
}

// ---- weight gradient on the tensor cores ----------------------------------------------------------------------
// dW[tap][ci][co] = sum over rows (board, position) of a[row + tap][ci] * dz[row][co]: a GEMM whose K dimension is the
// ROW index of both operand images, i.e. both operands are MN-major for tcgen05 (instruction-descriptor bits 15/16):
// a 128-byte image row is 64 channels along M (resp. N), 8 rows form the SWIZZLE_128B atom along K (SBO = 1024 B).
// M = 128 stacks two taps: the second 64-channel atom along M sits LBO = 4096 B further = the same window one board row
// lower, so one MMA accumulates taps (dy, dx) and (dy+1, dx) at once; (dy=+1, dx) is paired with a dummy window whose
// accumulator rows are ignored.  Per tile of 32 boards: 10 valid (output column, dx) pairs x 8 K-steps x 2 = 160 MMAs
// (M128 N64 K16), split over two issuing warps by accumulator block; partial dW per tile -> wpart[tile][tap][ci][co].
constexpr int TCW_SM_B = tc::ACT_BYTES;                 // dz image behind the a image
constexpr int TCW_SM_BARS = 2 * tc::ACT_BYTES;
constexpr int TCW_SMEM_BYTES = TCW_SM_BARS + 64 + 1024;

__global__ void __launch_bounds__(TCL_THREADS, 1)
tc_wgrad_kernel(const uint8_t *__restrict__ img_a, const uint8_t *__restrict__ img_dz, float *__restrict__ wpart) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *sm = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sm_base = tc::smem_u32(sm);
    const uint32_t bar_full = sm_base + TCW_SM_BARS, bar_acc = bar_full + 8, tmem_slot = bar_full + 16;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long tile = blockIdx.x;
    for (int i = tid; i < 5 * tc::SLAB_ROWS * tc::ROW_BYTES / 16; i += TCL_THREADS) {  // zero slabs of the a image
        const int z = i / (tc::SLAB_ROWS * tc::ROW_BYTES / 16), r = i % (tc::SLAB_ROWS * tc::ROW_BYTES / 16);
        reinterpret_cast<uint4 *>(sm + (size_t)z * 5 * tc::SLAB_ROWS * tc::ROW_BYTES)[r] = make_uint4(0, 0, 0, 0);
    }
    if (tid == 0) {
        tc::mbar_init(bar_full, 1);
        tc::mbar_init(bar_acc, 2);
        tc::fence_mbar_init();
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(sm + TCW_SM_BARS + 16);
    constexpr uint32_t SLAB4 = 4 * tc::SLAB_ROWS * tc::ROW_BYTES;  // one board column: 4 data slabs

    if (warp == 4) {
        if (lane == 0) {
            tc::mbar_expect_tx(bar_full, 2 * TCL_IMG_BYTES);
#pragma unroll
            for (int x = 0; x < 4; x++) {
                const uint32_t dst = (uint32_t)((5 * x + 1) * tc::SLAB_ROWS * tc::ROW_BYTES);
                tc::bulk_g2s(sm_base + dst, img_a + (size_t)tile * TCL_IMG_BYTES + (size_t)x * SLAB4, SLAB4, bar_full);
                tc::bulk_g2s(sm_base + TCW_SM_B + dst, img_dz + (size_t)tile * TCL_IMG_BYTES + (size_t)x * SLAB4, SLAB4, bar_full);
            }
        }
    } else if (warp == 5 || warp == 6) {
        const int which = warp - 5;  // 0: taps (dy=-1, dy=0) stacked along M; 1: tap dy=+1 (+ ignored dummy rows)
        // MN-major SWIZZLE_128B descriptors: LBO = 4096 B (next 64-channel atom along M = the window one board row lower),
        // SBO = 1024 B (next 8 rows along K), version 1, layout SWIZZLE_128B
        const uint64_t desc_hi = ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ULL << 46) | (2ULL << 61);
        const uint32_t idesc = tc::make_idesc(128, 64, true) | (1u << 15) | (1u << 16);
        tc::mbar_wait(bar_full, 0);
        tc::tc_fence_after();
        const uint32_t a_lo = sm_base >> 4, b_lo = (sm_base + TCW_SM_B) >> 4;
#pragma unroll 1
        for (int dxi = 0; dxi < 3; dxi++) {
            const uint32_t d_addr = tmem_base + (uint32_t)((dxi * 2 + which) * 64);
            bool first = true;
#pragma unroll 1
            for (int xs = 0; xs < 4; xs++) {
                const int xin = xs + dxi - 1;
                if (xin < 0 || xin > 3) continue;
                // 16-byte units: a slab-row group is 256 units, a K step of 16 rows is 128 units
                const uint32_t a0 = a_lo + (uint32_t)((5 * xin + (which ? 2 : 0)) * 256);
                const uint32_t b0 = b_lo + (uint32_t)((5 * xs + 1) * 256);
                for (int ks = 0; ks < 8; ks++) {
                    tc::umma_f16_elect(d_addr, desc_hi | (uint64_t)(a0 + ks * 128), desc_hi | (uint64_t)(b0 + ks * 128), idesc,
                                       first ? 0u : 1u);
                    first = false;
                }
            }
        }
        tc::umma_commit_elect(bar_acc);
    } else if (warp < 4) {
        tc::mbar_wait(bar_acc, 0);
        tc::tc_fence_after();
        // TMEM lane = row of the stacked M: lanes 0..63 first tap of the block, 64..127 second tap
        const int m = warp * 32 + lane, ci = m & 63, second = m >> 6;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        float *base = wpart + (size_t)tile * 9 * 4096;
#pragma unroll 1
        for (int blk = 0; blk < 6; blk++) {
            const int dxi = blk >> 1, which = blk & 1;
            const int dyi = which ? 2 : second;  // block 0: dy = -1 / 0 ; block 1: dy = +1 / dummy
            const bool valid = !(which && second);
            float *o = base + (size_t)(dyi * 3 + dxi) * 4096 + ci * 64;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                float v[8];
                tc::tmem_ld8(lane_addr + (uint32_t)(blk * 64 + c * 8), reinterpret_cast<uint32_t *>(v));
                tc::tmem_ld_wait();
                if (valid) {
                    *reinterpret_cast<float4 *>(o + c * 8) = make_float4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<float4 *>(o + c * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
                }
            }
        }
        tc::tc_fence_before();
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// ---- operand images -----------------------------------------------------------------------------------------
// boards -> first-layer image: the one-hot row twice along K (channels [0,16) and [16,32)) against [W_hi | W_lo]
// image_bf16 (optional): the plain one-hot row (channels [0,16), bf16) = the first layer's weight-gradient operand
__global__ void onehot_image_kernel(const unsigned long long *__restrict__ boards, uint8_t *__restrict__ image,
                                    uint8_t *__restrict__ image_bf16, long B) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;  // (board, position)
    const long tiles = (B + 31) / 32;
    if (i >= tiles * 32 * 16) return;
    const long board = i >> 4;
    const int p = (int)(i & 15), b = (int)(board & 31);
    const int t = board < B ? (int)((boards[board] >> (4 * p)) & 0xF) : 0;
    uint8_t *row = image + (size_t)(board >> 5) * TCL_IMG_BYTES + tile_image_offset(b, p, 0) - (size_t)((0 ^ (b & 7)) * 16);
    const uint32_t one = 0x3C00u << ((t & 1) * 16);  // fp16 1.0 in the right half-word
    const int wsel = (t & 7) >> 1;
    const uint4 hot = make_uint4(wsel == 0 ? one : 0u, wsel == 1 ? one : 0u, wsel == 2 ? one : 0u, wsel == 3 ? one : 0u);
    const uint4 zero = make_uint4(0, 0, 0, 0);
    const uint4 lo8 = (t < 8) ? hot : zero, hi8 = (t < 8) ? zero : hot;
    const int sw = b & 7;
    *reinterpret_cast<uint4 *>(row + ((0 ^ sw) << 4)) = lo8;
    *reinterpret_cast<uint4 *>(row + ((1 ^ sw) << 4)) = hi8;
    *reinterpret_cast<uint4 *>(row + ((2 ^ sw) << 4)) = lo8;
    *reinterpret_cast<uint4 *>(row + ((3 ^ sw) << 4)) = hi8;
#pragma unroll
    for (int c = 4; c < 8; c++) *reinterpret_cast<uint4 *>(row + ((c ^ sw) << 4)) = zero;
    if (image_bf16) {
        uint8_t *row2 = image_bf16 + (row - image);
        const uint32_t one2 = 0x3F80u << ((t & 1) * 16);  // bf16 1.0
        const uint4 hot2 = make_uint4(wsel == 0 ? one2 : 0u, wsel == 1 ? one2 : 0u, wsel == 2 ? one2 : 0u, wsel == 3 ? one2 : 0u);
        *reinterpret_cast<uint4 *>(row2 + ((0 ^ sw) << 4)) = (t < 8) ? hot2 : zero;
        *reinterpret_cast<uint4 *>(row2 + ((1 ^ sw) << 4)) = (t < 8) ? zero : hot2;
#pragma unroll
        for (int c = 2; c < 8; c++) *reinterpret_cast<uint4 *>(row2 + ((c ^ sw) << 4)) = zero;
    }
}

// four consecutive channels of (board, position) into a tile image, fp16 or bf16
template <bool BF16>
__device__ __forceinline__ void store_image4(uint8_t *image, long board, int p, int ch, float4 v) {
    uint2 w;
    if (BF16) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b2 = __floats2bfloat162_rn(v.z, v.w);
        w.x = *reinterpret_cast<const uint32_t *>(&a), w.y = *reinterpret_cast<const uint32_t *>(&b2);
    } else {
        const __half2 a = __floats2half2_rn(v.x, v.y), b2 = __floats2half2_rn(v.z, v.w);
        w.x = *reinterpret_cast<const uint32_t *>(&a), w.y = *reinterpret_cast<const uint32_t *>(&b2);
    }
    *reinterpret_cast<uint2 *>(image + (size_t)(board >> 5) * TCL_IMG_BYTES + tile_image_offset((int)(board & 31), p, ch)) = w;
}

// images of the fused finalize+apply kernels (train.cuh): forward = fp16 operand of the next convolution (+ a bf16 copy
// kept for the weight gradient); backward = bf16 image of dz
__device__ __forceinline__ void store_images(const Image16 &im, long i, float4 v, bool first_bf16) {
    const long board = i >> 8;
    const int p = (int)((i >> 4) & 15), c = (int)(i & 15) * 4;
    if (first_bf16) {
        store_image4<true>(im.img, board, p, c, v);
    } else {
        store_image4<false>(im.img, board, p, c, v);
        if (im.img2) store_image4<true>(im.img2, board, p, c, v);
    }
}

// ---- weight chunks ------------------------------------------------------------------------------------------
// fwd[l][dy] fp16 (21 chunks): rows [dx=+1 | 0 | -1] x 64 co, K = ci; first layer rows are [W_hi(16) | W_lo(16) | 0].
// bwd[l-1][dy] bf16 (18 chunks, l = 1..6): the data gradient is the same convolution over dz with flipped taps and
// the channel roles exchanged: rows [dx=+1 | 0 | -1] x 64 ci, K = co, value W[co][ci][2-ky][2-kx].
__global__ void pack_weights_tc_kernel(const float *__restrict__ params, uint16_t *__restrict__ fwd, uint16_t *__restrict__ bwd) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (chunk, row, k)
    constexpr int PER = tc::CHUNK_ROWS * 64;
    if (idx >= (21 + 18) * PER) return;
    const int chunk = idx / PER, row = (idx % PER) >> 6, k = idx & 63;
    const int blk = row >> 6, n = row & 63, kx = 2 - blk;  // blk 0,1,2 <-> dx = +1, 0, -1 <-> kx = 2, 1, 0
    const size_t dst = (size_t)chunk * PER + (size_t)row * 64 + (size_t)((((k >> 3) ^ (row & 7)) << 3) | (k & 7));
    if (chunk < 21) {
        const int l = chunk / 3, ky = chunk % 3;
        float v = 0.f;
        if (l == 0) {
            if (k < 32) {
                const float w = params[P_W(0) + ((n * 16 + (k & 15)) * 3 + ky) * 3 + kx];
                const float hi = __half2float(__float2half_rn(w));
                v = k < 16 ? hi : w - hi;
            }
        } else {
            v = params[P_W(l) + ((n * 64 + k) * 3 + ky) * 3 + kx];
        }
        const __half h = __float2half_rn(v);
        fwd[dst] = *reinterpret_cast<const uint16_t *>(&h);
    } else {
        const int c2 = chunk - 21, l = c2 / 3 + 1, ky = c2 % 3;
        const float v = params[P_W(l) + ((k * 64 + n) * 3 + (2 - ky)) * 3 + (2 - kx)];  // n = ci, k = co
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        bwd[dst - (size_t)21 * PER] = *reinterpret_cast<const uint16_t *>(&h);
    }
}

}  // namespace train
}  // namespace b2048
