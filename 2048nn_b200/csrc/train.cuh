// train.cuh -- one c64b3 training step on the B200 in fp32 (trainer.py:55-62: forward in train mode,
// KLDivLoss(batchmean), backward, Adam), hand-written: no autograd, no cuDNN.
//
// Why CUDA cores: the reference trains in fp32 (trainer.py never casts); one step is 46 GFLOP of 64-channel
// 3x3 convolutions on 4x4 maps plus BatchNorm over the batch.  The convolutions run as register-tiled
// outer products on the packed FFMA2 pipe (2 fp32 FMA per lane per issue, 148 x 128 FMA/clk), everything else
// is a handful of elementwise / reduction kernels; the host enqueues the ~65 launches of a step on one stream
// (graph-capturable: no host sync, no allocation).
//
// Layout: every activation-shaped tensor is [board][16 positions][64 channels] fp32 (channels fastest), position
// p = 4*y + x as board.py:23-33 / mcts_nn.py:25-35 (nibble p).  Parameters, gradients and Adam moments are flat
// fp32 vectors in `parameters()` order of network.py:43-76 (offsets below), so the nn.Module's tensors can alias
// the flat buffer.  All reductions over the batch use per-CTA partials summed in a fixed order: a step is
// deterministic run to run.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2048 {
namespace train {

constexpr int NL = 7;  // 3x3 conv layers: in_block + 6 block convs

// ---- flat parameter vector (network.py:55-76 registration order) --------------------------------------------
__host__ __device__ constexpr int P_W(int l) { return l == 0 ? 0 : 9344 + (l - 1) * 36992; }  // [co][ci][3][3]
__host__ __device__ constexpr int P_G(int l) { return l == 0 ? 9216 : P_W(l) + 36864; }      // BN weight [64]
__host__ __device__ constexpr int P_B(int l) { return P_G(l) + 64; }                          // BN bias [64]
constexpr int P_W7 = 9344 + 6 * 36992;  // out_block conv [4][64]
constexpr int P_G7 = P_W7 + 256, P_B7 = P_G7 + 4;
constexpr int P_FCW = P_B7 + 4, P_FCB = P_FCW + 256;  // policy Linear [4][64], [4]
constexpr int N_PARAMS = P_FCB + 4;
static_assert(N_PARAMS == 231820, "c64b3 has 231,820 parameters (network.py:138)");
// BatchNorm buffers: running_mean [8][64] then running_var [8][64] (layer 7 = out_block BN uses 4 channels)
constexpr int N_BN_FLOATS = 2 * 8 * 64;

constexpr int ACT = 16 * 64;  // floats per board in an activation tensor

// ---- workspace -------------------------------------------------------------------------------------------
constexpr int MAX_PARTS = 512;     // per-CTA partial slots of the channel reductions (conv stats need ceil(B/8))
constexpr int WGRAD_SPLITS = 98;   // K-splits of the weight-gradient GEMM (x 3 tap rows = 294 CTAs = 2 per SM)

struct Workspace {
    float *x0;            // [B][16][16] one-hot input
    float *z[NL], *a[NL];  // pre-BN conv outputs, post-activation outputs
    float *dyA, *dyB, *dz, *dz2, *da;  // dz double-buffered: the weight gradient reads it on a side stream
    float *z7, *dy7;      // [B][16][4]
    float *h;             // [B][64] head activations, flattened c4*16+p (network.py:85)
    float *dlogits;       // [B][4]
    float *logp;          // [B][4]
    float *stats;         // [8][4][64]: mean, invstd, scale, shift
    float *coef;          // [8][2][64]: mean(dy), mean(dy*xhat)
    float *part;          // [max_parts(B)][2][64]
    float *wpart;         // [max(WGRAD_SPLITS, ceil(B/32))][9][64][64]
    float *hpart;         // [MAX_PARTS][1024] head / loss partials
    float *wf, *wd;       // packed weights [7][9][64 ci][64 co] forward / [7][9][64 co][64 ci] data-gradient
    float *grad;          // [N_PARAMS]
    // mixed mode (train_tc.cuh): 16-bit operand images of the current conv input / dz, tensor-core weight chunks
    uint8_t *img_a, *img_dz, *img_dz2;  // [ceil(B/32)][65536]; dz double-buffered (weight gradient runs on a side stream)
    uint8_t *img_abf[NL];     // bf16 copies of every conv layer's input image (weight gradient operands)
    uint16_t *wtc_fwd, *wtc_bwd;  // [21][192*64], [18][192*64]
    size_t bytes;
};

inline int max_parts(int64_t B) { return (int)((B + 7) / 8 > MAX_PARTS ? (B + 7) / 8 : MAX_PARTS); }

inline Workspace carve(void *base, int64_t B) {
    Workspace w;
    const uintptr_t p = reinterpret_cast<uintptr_t>(base);  // base may be NULL: only offsets are wanted then
    size_t off = 0;
    auto take = [&](size_t floats) {
        float *r = reinterpret_cast<float *>(p + off);
        off += (floats * sizeof(float) + 255) & ~(size_t)255;
        return r;
    };
    w.x0 = take((size_t)B * 16 * 16);
    for (int l = 0; l < NL; l++) w.z[l] = take((size_t)B * ACT), w.a[l] = take((size_t)B * ACT);
    w.dyA = take((size_t)B * ACT), w.dyB = take((size_t)B * ACT), w.dz = take((size_t)B * ACT), w.dz2 = take((size_t)B * ACT);
    w.da = take((size_t)B * ACT);
    w.z7 = take((size_t)B * 64), w.dy7 = take((size_t)B * 64);
    w.h = take((size_t)B * 64), w.dlogits = take((size_t)B * 4), w.logp = take((size_t)B * 4);
    w.stats = take(8 * 4 * 64), w.coef = take(8 * 2 * 64);
    w.part = take((size_t)max_parts(B) * 128);
    w.wpart = take((size_t)((B + 31) / 32 > WGRAD_SPLITS ? (B + 31) / 32 : WGRAD_SPLITS) * 9 * 64 * 64);
    w.hpart = take((size_t)MAX_PARTS * 1024);
    w.wf = take(7 * 9 * 64 * 64), w.wd = take(7 * 9 * 64 * 64);
    w.grad = take(N_PARAMS);
    const size_t tiles = (size_t)((B + 31) / 32);
    w.img_a = reinterpret_cast<uint8_t *>(take(tiles * 16384)), w.img_dz = reinterpret_cast<uint8_t *>(take(tiles * 16384));
    w.img_dz2 = reinterpret_cast<uint8_t *>(take(tiles * 16384));
    for (int l = 0; l < NL; l++) w.img_abf[l] = reinterpret_cast<uint8_t *>(take(tiles * 16384));
    w.wtc_fwd = reinterpret_cast<uint16_t *>(take(21 * 192 * 32)), w.wtc_bwd = reinterpret_cast<uint16_t *>(take(18 * 192 * 32));
    w.bytes = off;
    return w;
}

__device__ __forceinline__ float2 ffma2(float a, float2 b, float2 c) { return __ffma2_rn(make_float2(a, a), b, c); }

// ---- weight packing ----------------------------------------------------------------------------------------
// wf[l][tap][ci][co] = W_l[co][ci][tap]                (forward: out[p] += in[p+tap] * W[tap])
// wd[l][tap][co][ci] = W_l[co][ci][8-tap]              (data gradient = the same conv over dz with flipped taps)
__global__ void pack_weights_kernel(const float *__restrict__ params, float *__restrict__ wf, float *__restrict__ wd) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 7 * 9 * 64 * 64) return;
    const int n = idx & 63, k = (idx >> 6) & 63, tap = (idx >> 12) % 9, l = idx / (9 * 4096);
    const int cin = l == 0 ? 16 : 64;
    const float *W = params + P_W(l);
    wf[idx] = (k < cin) ? W[(n * cin + k) * 9 + tap] : 0.f;            // k = ci, n = co
    wd[idx] = (l > 0) ? W[(k * cin + n) * 9 + (8 - tap)] : 0.f;        // k = co, n = ci
}

// boards -> one-hot [B][16][16] (mcts_nn.py:25-35: channel = exponent, position p = nibble p)
__global__ void onehot_kernel(const unsigned long long *__restrict__ boards, float *__restrict__ x0, long B) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;  // (board, position)
    if (i >= B * 16) return;
    const int t = (int)((boards[i >> 4] >> (4 * (i & 15))) & 0xF);
    float4 *o = reinterpret_cast<float4 *>(x0 + i * 16);
#pragma unroll
    for (int q = 0; q < 4; q++)
        o[q] = make_float4(t == 4 * q ? 1.f : 0.f, t == 4 * q + 1 ? 1.f : 0.f, t == 4 * q + 2 ? 1.f : 0.f, t == 4 * q + 3 ? 1.f : 0.f);
}

// ---- 3x3 convolution, pad 1, 4x4 maps, CIN -> 64 channels ---------------------------------------------------
// CTA = 128 threads = 8 boards, inputs [8][16][CIN] in shared memory.  Warp w owns board rows {2h, 2h+1} (h = w & 1)
// of four boards; thread (board, cg): 8 output positions x 8 output channels {4cg..4cg+3, 32+4cg..}; per 4 input
// channels 8 + 8 LDS.128 feed 128 FFMA2.  The tap loop is unrolled, so "input position outside the board" is known
// per (tap, row) at compile time in x and per warp in y: those rows are skipped instead of multiplying zeros
// (100 of 144 tap-positions are in bounds).  The tap's [CIN][64] weight tile streams through a 2-deep cp.async
// ring.  STATS: per-CTA per-channel sum / sum of squares of the outputs (BatchNorm batch statistics) to
// part[cta][2][64].
#ifndef B2048_CONV_NQ
#define B2048_CONV_NQ 2
#endif
constexpr int CONV_NQ = B2048_CONV_NQ;            // float4 channel groups per thread: 2 -> 8x8 register tile, 4 -> 8x16
constexpr int CONV_THREADS = 256 / CONV_NQ, CONV_BOARDS = 8;
constexpr int CONV_CGS = 16 / CONV_NQ;             // threads along the 64 output channels
constexpr int CONV_CSTRIDE = 64 / CONV_NQ;         // channel distance between a thread's groups
template <int CIN>
struct ConvSmem {
    static constexpr int PIXC = CIN + 4;          // floats per position (+4 keeps 16-byte alignment, spreads banks)
    static constexpr int BSTR = 16 * PIXC + 8;    // floats per board: the 4 boards of a warp hit distinct banks
    static constexpr int A_FLOATS = CONV_BOARDS * BSTR;
    static constexpr int W_FLOATS = CIN * 64;
    static constexpr int RED_FLOATS = 2 * 16 * 64;  // stats staging, reuses the input tile
    static constexpr int BYTES = ((A_FLOATS > RED_FLOATS ? A_FLOATS : RED_FLOATS) + 2 * W_FLOATS) * 4;
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// one tap for the rows in MASK (bit i: output row i = (y0 + i/4, i%4) has its input position on the board)
template <int CIN, int MASK>
__device__ __forceinline__ void conv_tap(const float *ap, const float *wt, float2 (&acc)[8][2 * CONV_NQ]) {
    constexpr int PIXC = ConvSmem<CIN>::PIXC;
#pragma unroll 2
    for (int c4 = 0; c4 < CIN / 4; c4++) {
        float4 av[8];
#pragma unroll
        for (int i = 0; i < 8; i++)
            if ((MASK >> i) & 1) av[i] = *reinterpret_cast<const float4 *>(ap + ((i >> 2) * 4 + (i & 3)) * PIXC + c4 * 4);
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
            float2 w[2 * CONV_NQ];
#pragma unroll
            for (int j = 0; j < CONV_NQ; j++) {
                const float4 t = *reinterpret_cast<const float4 *>(wt + (c4 * 4 + kk) * 64 + j * CONV_CSTRIDE);
                w[2 * j] = make_float2(t.x, t.y), w[2 * j + 1] = make_float2(t.z, t.w);
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (!((MASK >> i) & 1)) continue;
                const float a = kk == 0 ? av[i].x : kk == 1 ? av[i].y : kk == 2 ? av[i].z : av[i].w;
#pragma unroll
                for (int j = 0; j < 2 * CONV_NQ; j++) acc[i][j] = ffma2(a, w[j], acc[i][j]);
            }
        }
    }
}

template <int CIN, bool STATS>
__global__ void __launch_bounds__(CONV_THREADS) conv3x3_kernel(const float *__restrict__ in, const float *__restrict__ wp,
                                                               float *__restrict__ out, float *__restrict__ part, long B) {
    typedef ConvSmem<CIN> S;
    extern __shared__ __align__(16) float smem[];
    constexpr int A_REGION = S::A_FLOATS > S::RED_FLOATS ? S::A_FLOATS : S::RED_FLOATS;
    float *a_s = smem, *w_s = smem + A_REGION, *red = smem;
    const int tid = threadIdx.x;
    const long base = (long)blockIdx.x * CONV_BOARDS;
    const int nb = (int)((B - base) < CONV_BOARDS ? (B - base) : CONV_BOARDS);

    auto load_w = [&](int tap) {
        const float *src = wp + (size_t)tap * 64 * 64;
        float *dst = w_s + (tap & 1) * S::W_FLOATS;
        for (int i = tid; i < S::W_FLOATS / 4; i += CONV_THREADS) cp_async16(dst + i * 4, src + i * 4);
        cp_async_commit();
    };
    load_w(0);
    {  // input tile: board b, position p, channels [0, CIN); boards past the batch end read as zeros
        constexpr int V = CIN / 4;  // float4 per position
        for (int i = tid; i < CONV_BOARDS * 16 * V; i += CONV_THREADS) {
            const int c4 = i % V, p = (i / V) & 15, b = i / (V * 16);
            const float4 v = b < nb ? reinterpret_cast<const float4 *>(in + (base + b) * 16 * CIN)[p * V + c4] : make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4 *>(a_s + b * S::BSTR + p * S::PIXC + c4 * 4) = v;
        }
    }
    const int warp = tid >> 5, lane = tid & 31;
    const int half = warp & 1, b = (warp >> 1) * (32 / CONV_CGS) + lane / CONV_CGS, cg = lane % CONV_CGS;
    const int y0 = half * 2;
    float2 acc[8][2 * CONV_NQ];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 2 * CONV_NQ; j++) acc[i][j] = make_float2(0.f, 0.f);

#pragma unroll
    for (int tap = 0; tap < 9; tap++) {
        cp_async_wait<0>();
        __syncthreads();  // tap's weights (and, first time, the input tile) are visible; the other buffer is free
        if (tap + 1 < 9) load_w(tap + 1);
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        // row masks: x + dx on the board -> 0xEE (dx=-1) / 0xFF / 0x77 (dx=+1); y + dy on the board -> & 0xF0 / & 0x0F
        const float *ap = a_s + b * S::BSTR + ((y0 + dy) * 4 + dx) * S::PIXC;
        const float *wt = w_s + (tap & 1) * S::W_FLOATS + cg * 4;
        if (dy == -1) {
            if (half == 0) {  // y = 0 has no row above
                if (dx == -1) conv_tap<CIN, 0xEE & 0xF0>(ap, wt, acc);
                else if (dx == 0) conv_tap<CIN, 0xFF & 0xF0>(ap, wt, acc);
                else conv_tap<CIN, 0x77 & 0xF0>(ap, wt, acc);
            } else {
                if (dx == -1) conv_tap<CIN, 0xEE>(ap, wt, acc);
                else if (dx == 0) conv_tap<CIN, 0xFF>(ap, wt, acc);
                else conv_tap<CIN, 0x77>(ap, wt, acc);
            }
        } else if (dy == 1) {
            if (half == 1) {  // y = 3 has no row below
                if (dx == -1) conv_tap<CIN, 0xEE & 0x0F>(ap, wt, acc);
                else if (dx == 0) conv_tap<CIN, 0xFF & 0x0F>(ap, wt, acc);
                else conv_tap<CIN, 0x77 & 0x0F>(ap, wt, acc);
            } else {
                if (dx == -1) conv_tap<CIN, 0xEE>(ap, wt, acc);
                else if (dx == 0) conv_tap<CIN, 0xFF>(ap, wt, acc);
                else conv_tap<CIN, 0x77>(ap, wt, acc);
            }
        } else {
            if (dx == -1) conv_tap<CIN, 0xEE>(ap, wt, acc);
            else if (dx == 0) conv_tap<CIN, 0xFF>(ap, wt, acc);
            else conv_tap<CIN, 0x77>(ap, wt, acc);
        }
    }
    if (b < nb) {
        float *o = out + ((base + b) * 16 + y0 * 4) * 64 + cg * 4;
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < CONV_NQ; j++)
                *reinterpret_cast<float4 *>(o + i * 64 + j * CONV_CSTRIDE) =
                    make_float4(acc[i][2 * j].x, acc[i][2 * j].y, acc[i][2 * j + 1].x, acc[i][2 * j + 1].y);
    }
    if (STATS) {
        // red[0][rg][64], red[1][rg][64] with rg = (board, half); pair j of a thread = channels 4cg + (j/2)*CSTRIDE + 2*(j%2) + {0,1}
        const int rg = b * 2 + half;
        __syncthreads();  // every thread is done reading the input tile
#pragma unroll
        for (int j = 0; j < 2 * CONV_NQ; j++) {
            float sx = 0.f, sy = 0.f, qx = 0.f, qy = 0.f;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                sx += acc[i][j].x, sy += acc[i][j].y;
                qx = fmaf(acc[i][j].x, acc[i][j].x, qx), qy = fmaf(acc[i][j].y, acc[i][j].y, qy);
            }
            const int c = cg * 4 + (j >> 1) * CONV_CSTRIDE + 2 * (j & 1);
            red[rg * 64 + c] = sx, red[rg * 64 + c + 1] = sy;
            red[16 * 64 + rg * 64 + c] = qx, red[16 * 64 + rg * 64 + c + 1] = qy;
        }
        __syncthreads();
        for (int t = tid; t < 128; t += CONV_THREADS) {
            const int which = t >> 6, c = t & 63;
            float v = 0.f;
#pragma unroll
            for (int r = 0; r < 16; r++) v += red[which * 16 * 64 + r * 64 + c];
            part[(size_t)blockIdx.x * 128 + which * 64 + c] = v;
        }
    }
}

// ---- BatchNorm (training mode, network.py:56,63,66,72: nn.BatchNorm2d defaults eps 1e-5, momentum 0.1) -------
// partials [nparts][2][64] -> mean, invstd, scale = gamma*invstd, shift = beta - mean*scale; running stats update
// with the unbiased variance.  One CTA; sums in double, fixed order.
// Sum partials [nparts][stride] (sums at [c], squares at [stride/2 + c]) over nparts in double: 1024 threads =
// 16 lanes x 64 channels, lane k takes partials k, k+16, ...; the 16 lane sums are added in a fixed order.
__device__ __forceinline__ void sum_partials(const float *__restrict__ part, int nparts, int stride, int C, double &s, double &q) {
    __shared__ double sh[2][16][64];
    const int c = threadIdx.x & 63, lane = threadIdx.x >> 6;
    double ls = 0.0, lq = 0.0;
    if (c < C)
        for (int i0 = lane; i0 < nparts; i0 += 16 * 8) {  // 8 partials' loads in flight, then added in index order
            float vs[8], vq[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int i = i0 + 16 * k;
                vs[k] = i < nparts ? part[(size_t)i * stride + c] : 0.f;
                vq[k] = i < nparts ? part[(size_t)i * stride + stride / 2 + c] : 0.f;
            }
#pragma unroll
            for (int k = 0; k < 8; k++) ls += (double)vs[k], lq += (double)vq[k];
        }
    sh[0][lane][c] = ls, sh[1][lane][c] = lq;
    __syncthreads();
    s = 0.0, q = 0.0;
    if (lane == 0)
        for (int k = 0; k < 16; k++) s += sh[0][k][c], q += sh[1][k][c];
}

__global__ void __launch_bounds__(1024) bn_finalize_kernel(const float *__restrict__ part, int nparts, int part_stride, double count,
                                                           const float *__restrict__ gamma, const float *__restrict__ beta,
                                                           float *__restrict__ run_mean, float *__restrict__ run_var, float momentum,
                                                           float eps, float *__restrict__ stats, int C) {
    double s, q;
    sum_partials(part, nparts, part_stride, C, s, q);
    const int c = threadIdx.x;
    if (c >= C) return;  // (threads 64.. belong to lanes > 0)
    const double mean = s / count;
    double var = q / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float scale = gamma[c] * invstd;
    stats[c] = (float)mean;
    stats[64 + c] = invstd;
    stats[128 + c] = scale;
    stats[192 + c] = beta[c] - (float)mean * scale;
    run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * (float)mean;
    run_var[c] = (1.f - momentum) * run_var[c] + momentum * (float)(var * count / (count - 1.0));
}

// ---- finalize + apply in one launch ----------------------------------------------------------------------------
// Every CTA (1024 threads) re-reduces the per-CTA partials itself (<= 151 KB from L2, the same fixed order everywhere)
// instead of waiting for a one-CTA finalize kernel: one launch and one dependency less per layer and direction.
// CTA 0 publishes the statistics / parameter gradients and updates the running statistics.
struct Image16 {  // optional 16-bit operand images written next to the fp32 result (mixed mode, train_tc.cuh)
    uint8_t *img, *img2;
};
__device__ __forceinline__ void store_images(const Image16 &im, long i, float4 v, bool first_bf16);

template <bool IMAGE>
__global__ void __launch_bounds__(1024) bn_apply_fused_kernel(const float *__restrict__ part, int nparts, double count,
                                                              const float *__restrict__ gamma, const float *__restrict__ beta,
                                                              float *__restrict__ run_mean, float *__restrict__ run_var, float momentum,
                                                              float eps, float *__restrict__ stats, const float4 *__restrict__ z,
                                                              const float4 *__restrict__ res, float4 *__restrict__ a, Image16 im, long n4) {
    __shared__ __align__(16) float s_sc[64], s_sh[64];
    double s, q;
    sum_partials(part, nparts, 128, 64, s, q);
    if (threadIdx.x < 64) {
        const int c = threadIdx.x;
        const double mean = s / count;
        double var = q / count - mean * mean;
        if (var < 0.0) var = 0.0;
        const float invstd = (float)(1.0 / sqrt(var + (double)eps));
        const float scale = gamma[c] * invstd, shift = beta[c] - (float)mean * scale;
        s_sc[c] = scale, s_sh[c] = shift;
        if (blockIdx.x == 0) {
            stats[c] = (float)mean, stats[64 + c] = invstd, stats[128 + c] = scale, stats[192 + c] = shift;
            run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * (float)mean;
            run_var[c] = (1.f - momentum) * run_var[c] + momentum * (float)(var * count / (count - 1.0));
        }
    }
    __syncthreads();
    for (long i = (long)blockIdx.x * 1024 + threadIdx.x; i < n4; i += (long)gridDim.x * 1024) {
        const int c = (int)(i & 15) * 4;
        const float4 sc = *reinterpret_cast<const float4 *>(s_sc + c), sh = *reinterpret_cast<const float4 *>(s_sh + c);
        float4 v = z[i];
        v.x = fmaf(v.x, sc.x, sh.x), v.y = fmaf(v.y, sc.y, sh.y), v.z = fmaf(v.z, sc.z, sh.z), v.w = fmaf(v.w, sc.w, sh.w);
        if (res) {
            const float4 r = res[i];
            v.x += r.x, v.y += r.y, v.z += r.z, v.w += r.w;
        }
        v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
        a[i] = v;
        if (IMAGE) store_images(im, i, v, false);
    }
}

template <bool IMAGE>
__global__ void __launch_bounds__(1024) bn_bwd_apply_fused_kernel(const float *__restrict__ part, int nparts, double count,
                                                                  float *__restrict__ dgamma, float *__restrict__ dbeta,
                                                                  const float *__restrict__ stats, const float4 *__restrict__ dy,
                                                                  const float4 *__restrict__ z, float4 *__restrict__ dz, Image16 im, long n4) {
    __shared__ __align__(16) float s_c1[64], s_c2[64];
    double s, q;
    sum_partials(part, nparts, 128, 64, s, q);
    if (threadIdx.x < 64) {
        const int c = threadIdx.x;
        s_c1[c] = (float)(s / count), s_c2[c] = (float)(q / count);
        if (blockIdx.x == 0) dbeta[c] = (float)s, dgamma[c] = (float)q;
    }
    __syncthreads();
    for (long i = (long)blockIdx.x * 1024 + threadIdx.x; i < n4; i += (long)gridDim.x * 1024) {
        const int c = (int)(i & 15) * 4;
        const float4 mean = *reinterpret_cast<const float4 *>(stats + c), istd = *reinterpret_cast<const float4 *>(stats + 64 + c),
                     sc = *reinterpret_cast<const float4 *>(stats + 128 + c), c1 = *reinterpret_cast<const float4 *>(s_c1 + c),
                     c2 = *reinterpret_cast<const float4 *>(s_c2 + c);
        const float4 g = dy[i], zv = z[i];
        float4 o;
        o.x = sc.x * (g.x - c1.x - (zv.x - mean.x) * istd.x * c2.x);
        o.y = sc.y * (g.y - c1.y - (zv.y - mean.y) * istd.y * c2.y);
        o.z = sc.z * (g.z - c1.z - (zv.z - mean.z) * istd.z * c2.z);
        o.w = sc.w * (g.w - c1.w - (zv.w - mean.w) * istd.w * c2.w);
        dz[i] = o;
        if (IMAGE) store_images(im, i, o, true);
    }
}

// Backward through ReLU and the BatchNorm reductions: dy = (g1 (+ g2)) * (a > 0); partial sums of dy and dy*xhat
// per channel.  256 threads = 16 channel quads x 16 row lanes; rows strided over the grid.
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float4 *__restrict__ g1, const float4 *g2,
                                                            const float4 *__restrict__ a, const float4 *__restrict__ z,
                                                            const float *__restrict__ stats, float4 *dy,  // dy may alias g2
                                                            float *__restrict__ part, long rows) {
    __shared__ float red[2][16][64];
    const int cq = threadIdx.x & 15, rl = threadIdx.x >> 4;
    const float4 mean = *reinterpret_cast<const float4 *>(stats + cq * 4), istd = *reinterpret_cast<const float4 *>(stats + 64 + cq * 4);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    for (long r = (long)blockIdx.x * 16 + rl; r < rows; r += (long)gridDim.x * 16) {
        const long i = r * 16 + cq;
        float4 g = g1[i];
        if (g2) {
            const float4 t = g2[i];
            g.x += t.x, g.y += t.y, g.z += t.z, g.w += t.w;
        }
        const float4 av = a[i], zv = z[i];
        g.x = av.x > 0.f ? g.x : 0.f, g.y = av.y > 0.f ? g.y : 0.f, g.z = av.z > 0.f ? g.z : 0.f, g.w = av.w > 0.f ? g.w : 0.f;
        dy[i] = g;
        s.x += g.x, s.y += g.y, s.z += g.z, s.w += g.w;
        q.x = fmaf(g.x, (zv.x - mean.x) * istd.x, q.x), q.y = fmaf(g.y, (zv.y - mean.y) * istd.y, q.y);
        q.z = fmaf(g.z, (zv.z - mean.z) * istd.z, q.z), q.w = fmaf(g.w, (zv.w - mean.w) * istd.w, q.w);
    }
    *reinterpret_cast<float4 *>(&red[0][rl][cq * 4]) = s;
    *reinterpret_cast<float4 *>(&red[1][rl][cq * 4]) = q;
    __syncthreads();
    if (threadIdx.x < 128) {
        const int which = threadIdx.x >> 6, c = threadIdx.x & 63;
        float t = 0.f;
#pragma unroll
        for (int r = 0; r < 16; r++) t += red[which][r][c];
        part[(size_t)blockIdx.x * 128 + which * 64 + c] = t;
    }
}

// partials -> dgamma = sum dy*xhat, dbeta = sum dy (into the flat gradient), coef = {mean(dy), mean(dy*xhat)}
__global__ void __launch_bounds__(1024) bn_bwd_finalize_kernel(const float *__restrict__ part, int nparts, int part_stride, double count,
                                                               float *__restrict__ dgamma, float *__restrict__ dbeta,
                                                               float *__restrict__ coef, int C) {
    double s, q;
    sum_partials(part, nparts, part_stride, C, s, q);
    const int c = threadIdx.x;
    if (c >= C) return;
    dbeta[c] = (float)s;
    dgamma[c] = (float)q;
    coef[c] = (float)(s / count);
    coef[64 + c] = (float)(q / count);
}

// ---- weight gradient: dW[tap][ci][co] = sum over (board, position) of a[p+tap][ci] * dz[p][co] ---------------
// grid (WGRAD_SPLITS, 3): blockIdx.y = tap row dy; CTA 256 threads = RS row-splits x (CIN/4 ci quads) x (16 co quads),
// thread accumulates its 4 ci x 4 co for the 3 dx taps (48 fp32, FFMA2).  Boards stream through a 2-deep cp.async
// ring, 4 boards per stage.  Partial sums per CTA to wpart[split][tap][64][64].
constexpr int WG_THREADS = 256, WG_STAGE = 4;
template <int CIN>
__global__ void __launch_bounds__(WG_THREADS) wgrad_kernel(const float *__restrict__ a, const float *__restrict__ dz,
                                                           float *__restrict__ wpart, long B, int boards_per_cta) {
    constexpr int RS = 64 / CIN;                 // row splits (1 for 64 input channels, 4 for the one-hot layer)
    constexpr int A_ST = WG_STAGE * 16 * CIN;    // floats of `a` per stage
    constexpr int D_ST = WG_STAGE * 16 * 64;
    extern __shared__ __align__(16) float smem[];
    float *a_s = smem, *d_s = smem + 2 * A_ST;
    const int tid = threadIdx.x;
    const int coq = tid & 15, ciq = (tid >> 4) % (CIN / 4), rs = (tid >> 4) / (CIN / 4);
    const int dy = (int)blockIdx.y - 1;
    const long b0 = (long)blockIdx.x * boards_per_cta;
    long b1 = b0 + boards_per_cta;
    if (b1 > B) b1 = B;
    const int nstage = b1 > b0 ? (int)((b1 - b0 + WG_STAGE - 1) / WG_STAGE) : 0;
    auto load = [&](int st) {
        const long sb = b0 + (long)st * WG_STAGE;
        const int nb = (int)((b1 - sb) < WG_STAGE ? (b1 - sb) : WG_STAGE);
        float *as = a_s + (st & 1) * A_ST, *ds = d_s + (st & 1) * D_ST;
        for (int i = tid; i < nb * 16 * CIN / 4; i += WG_THREADS) cp_async16(as + i * 4, a + sb * 16 * CIN + i * 4);
        for (int i = tid; i < nb * 16 * 16; i += WG_THREADS) cp_async16(ds + i * 4, dz + sb * ACT + i * 4);
        cp_async_commit();
    };
    float2 acc[3][4][2];
#pragma unroll
    for (int d = 0; d < 3; d++)
#pragma unroll
        for (int i = 0; i < 4; i++) acc[d][i][0] = acc[d][i][1] = make_float2(0.f, 0.f);
    if (nstage > 0) load(0);
    for (int st = 0; st < nstage; st++) {
        cp_async_wait<0>();
        __syncthreads();
        if (st + 1 < nstage) load(st + 1);
        const long sb = b0 + (long)st * WG_STAGE;
        const int nb = (int)((b1 - sb) < WG_STAGE ? (b1 - sb) : WG_STAGE);
        const float *as = a_s + (st & 1) * A_ST + ciq * 4, *ds = d_s + (st & 1) * D_ST + coq * 4;
        for (int by = rs; by < nb * 4; by += RS) {  // (board, output row y); positions x unrolled
            const int bb = by >> 2, y = by & 3;
            const int yy = y + dy;
            if (yy < 0 || yy > 3) continue;
            const float *drow = ds + (bb * 16 + y * 4) * 64;
            const float *arow = as + (bb * 16 + yy * 4) * CIN;
            float4 av[4];
#pragma unroll
            for (int xx = 0; xx < 4; xx++) av[xx] = *reinterpret_cast<const float4 *>(arow + xx * CIN);
#pragma unroll
            for (int x = 0; x < 4; x++) {
                const float4 d = *reinterpret_cast<const float4 *>(drow + x * 64);
                const float2 d0 = make_float2(d.x, d.y), d1 = make_float2(d.z, d.w);
#pragma unroll
                for (int dxi = 0; dxi < 3; dxi++) {
                    const int xx = x + dxi - 1;
                    if (xx < 0 || xx > 3) continue;
                    acc[dxi][0][0] = ffma2(av[xx].x, d0, acc[dxi][0][0]), acc[dxi][0][1] = ffma2(av[xx].x, d1, acc[dxi][0][1]);
                    acc[dxi][1][0] = ffma2(av[xx].y, d0, acc[dxi][1][0]), acc[dxi][1][1] = ffma2(av[xx].y, d1, acc[dxi][1][1]);
                    acc[dxi][2][0] = ffma2(av[xx].z, d0, acc[dxi][2][0]), acc[dxi][2][1] = ffma2(av[xx].z, d1, acc[dxi][2][1]);
                    acc[dxi][3][0] = ffma2(av[xx].w, d0, acc[dxi][3][0]), acc[dxi][3][1] = ffma2(av[xx].w, d1, acc[dxi][3][1]);
                }
            }
        }
    }
    __syncthreads();
    float *outp = wpart + ((size_t)blockIdx.x * 9 + (size_t)(dy + 1) * 3) * 4096;
    if (RS == 1) {
#pragma unroll
        for (int dxi = 0; dxi < 3; dxi++)
#pragma unroll
            for (int i = 0; i < 4; i++)
                *reinterpret_cast<float4 *>(outp + dxi * 4096 + (ciq * 4 + i) * 64 + coq * 4) =
                    make_float4(acc[dxi][i][0].x, acc[dxi][i][0].y, acc[dxi][i][1].x, acc[dxi][i][1].y);
    } else {
        // reduce the row splits through shared memory (fixed order), [rs][3][CIN][64]
        float *red = smem;
#pragma unroll
        for (int dxi = 0; dxi < 3; dxi++)
#pragma unroll
            for (int i = 0; i < 4; i++)
                *reinterpret_cast<float4 *>(red + ((rs * 3 + dxi) * CIN + ciq * 4 + i) * 64 + coq * 4) =
                    make_float4(acc[dxi][i][0].x, acc[dxi][i][0].y, acc[dxi][i][1].x, acc[dxi][i][1].y);
        __syncthreads();
        for (int i = tid; i < 3 * CIN * 64; i += WG_THREADS) {
            float t = 0.f;
#pragma unroll
            for (int k = 0; k < RS; k++) t += red[k * 3 * CIN * 64 + i];
            const int dxi = i / (CIN * 64), rem = i % (CIN * 64);
            outp[dxi * 4096 + rem] = t;
        }
    }
}
template <int CIN>
constexpr int wgrad_smem_bytes() {
    constexpr int ring = 2 * WG_STAGE * 16 * CIN + 2 * WG_STAGE * 16 * 64;
    constexpr int red = (64 / CIN > 1) ? (64 / CIN) * 3 * CIN * 64 : 0;  // row-split reduction staging (one-hot layer only)
    return (ring > red ? ring : red) * 4;
}

// wpart[splits][tap][ci][co] -> grad[co][ci][tap] (PyTorch conv weight layout), fixed summation order
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float *__restrict__ wpart, int splits, float *__restrict__ gw, int cin) {
    // 4 lanes per (tap, ci, co): consecutive threads -> consecutive co x 4 split phases; fixed shuffle tree
    const int idx = blockIdx.x * 64 + (threadIdx.x & 63), lane = threadIdx.x >> 6;
    __shared__ float sh[4][64];
    float t = 0.f;
    for (int s = lane; s < splits; s += 4) t += wpart[(size_t)s * 36864 + idx];
    sh[lane][threadIdx.x & 63] = t;
    __syncthreads();
    if (lane == 0) {
        const int co = idx & 63, ci = (idx >> 6) & 63, tap = idx >> 12;
        if (ci < cin) gw[(co * cin + ci) * 9 + tap] = (sh[0][co] + sh[1][co]) + (sh[2][co] + sh[3][co]);
    }
}

// ---- head (network.py:72-76,84-87): 1x1 conv 64->4, BN, ReLU, flatten, Linear 64->4, log-softmax; KLDiv loss ----
// z7[row][c4] = sum_ci a6[row][ci] * W7[c4][ci]; per-CTA partial sum / sum of squares per c4 (part stride 8).
__global__ void __launch_bounds__(256) head_conv_kernel(const float *__restrict__ a6, const float *__restrict__ w7,
                                                        float *__restrict__ z7, float *__restrict__ part, long rows) {
    __shared__ float w_s[256];
    __shared__ float red[8][8];
    w_s[threadIdx.x] = w7[threadIdx.x];
    __syncthreads();
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
    for (long r = (long)blockIdx.x * 256 + threadIdx.x; r < rows; r += (long)gridDim.x * 256) {
        const float4 *x = reinterpret_cast<const float4 *>(a6 + r * 64);
        float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const float4 v = x[k];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const float4 w = *reinterpret_cast<const float4 *>(w_s + c * 64 + k * 4);
                o[c] = fmaf(v.x, w.x, o[c]), o[c] = fmaf(v.y, w.y, o[c]), o[c] = fmaf(v.z, w.z, o[c]), o[c] = fmaf(v.w, w.w, o[c]);
            }
        }
        *reinterpret_cast<float4 *>(z7 + r * 4) = make_float4(o[0], o[1], o[2], o[3]);
#pragma unroll
        for (int c = 0; c < 4; c++) s[c] += o[c], q[c] = fmaf(o[c], o[c], q[c]);
    }
#pragma unroll
    for (int c = 0; c < 4; c++)
        for (int o = 16; o > 0; o >>= 1) s[c] += __shfl_xor_sync(0xffffffffu, s[c], o), q[c] += __shfl_xor_sync(0xffffffffu, q[c], o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
#pragma unroll
        for (int c = 0; c < 4; c++) red[warp][c] = s[c], red[warp][4 + c] = q[c];
    __syncthreads();
    if (threadIdx.x < 8) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) t += red[w][threadIdx.x];
        part[(size_t)blockIdx.x * 8 + threadIdx.x] = t;
    }
}

// one thread per board: h = relu(bn7(z7)) flattened c4*16+p, logits = FC(h), log-softmax, KL term, dlogits.
// KLDivLoss(reduction='batchmean') on log-probs: loss = sum_b sum_k xlogy(y,y) - y*logp over B; its gradient
// through log-softmax is dlogits = (softmax * sum_k y - y) / B.  Loss partial per CTA -> hpart[cta*1024].
__global__ void __launch_bounds__(128) head_fc_kernel(const float *__restrict__ z7, const float *__restrict__ stats7,
                                                      const float *__restrict__ fcw, const float *__restrict__ fcb,
                                                      const float *__restrict__ target, float *__restrict__ h,
                                                      float *__restrict__ logp, float *__restrict__ dlogits,
                                                      float *__restrict__ hpart, long B) {
    __shared__ float w_s[256];
    __shared__ float red[4];
    for (int i = threadIdx.x; i < 256; i += 128) w_s[i] = fcw[i];
    __syncthreads();
    const long b = (long)blockIdx.x * 128 + threadIdx.x;
    float term = 0.f;
    if (b < B) {
        float zl[4] = {fcb[0], fcb[1], fcb[2], fcb[3]};
        const float4 sc = *reinterpret_cast<const float4 *>(stats7 + 128), sh = *reinterpret_cast<const float4 *>(stats7 + 192);
        const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
        for (int p = 0; p < 16; p++) {
            const float4 v = *reinterpret_cast<const float4 *>(z7 + (b * 16 + p) * 4);
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const float hv = fmaxf(fmaf(vv[c], scv[c], shv[c]), 0.f);
                h[b * 64 + c * 16 + p] = hv;
#pragma unroll
                for (int o = 0; o < 4; o++) zl[o] = fmaf(hv, w_s[o * 64 + c * 16 + p], zl[o]);
            }
        }
        const float m = fmaxf(fmaxf(zl[0], zl[1]), fmaxf(zl[2], zl[3]));
        const float e[4] = {expf(zl[0] - m), expf(zl[1] - m), expf(zl[2] - m), expf(zl[3] - m)};
        const float se = e[0] + e[1] + e[2] + e[3], lse = m + logf(se);
        const float4 y = *reinterpret_cast<const float4 *>(target + b * 4);
        const float yv[4] = {y.x, y.y, y.z, y.w};
        const float sy = yv[0] + yv[1] + yv[2] + yv[3], invB = 1.f / (float)B;
        float lp[4], dl[4];
#pragma unroll
        for (int o = 0; o < 4; o++) {
            lp[o] = zl[o] - lse;
            dl[o] = (e[o] / se * sy - yv[o]) * invB;
            term += (yv[o] > 0.f ? yv[o] * logf(yv[o]) : 0.f) - yv[o] * lp[o];
        }
        *reinterpret_cast<float4 *>(logp + b * 4) = make_float4(lp[0], lp[1], lp[2], lp[3]);
        *reinterpret_cast<float4 *>(dlogits + b * 4) = make_float4(dl[0], dl[1], dl[2], dl[3]);
    }
    for (int o = 16; o > 0; o >>= 1) term += __shfl_xor_sync(0xffffffffu, term, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = term;
    __syncthreads();
    if (threadIdx.x == 0) hpart[(size_t)blockIdx.x * 1024] = red[0] + red[1] + red[2] + red[3];
}

// loss = sum of CTA partials / B
__global__ void loss_finalize_kernel(const float *__restrict__ hpart, int nparts, double B, float *__restrict__ loss_out) {
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < nparts; i++) s += (double)hpart[(size_t)i * 1024];
        *loss_out = (float)(s / B);
    }
}

// Linear backward: thread (o, j) of 256 accumulates dW[o][j] = sum_b dlogits[b][o]*h[b][j] over the CTA's boards,
// threads 0..3 also db[o]; partials hpart[cta*1024 + 0..259].
__global__ void __launch_bounds__(256) fc_wgrad_kernel(const float *__restrict__ h, const float *__restrict__ dlogits,
                                                       float *__restrict__ hpart, long B, int boards_per_cta) {
    const int o = threadIdx.x >> 6, j = threadIdx.x & 63;
    const long b0 = (long)blockIdx.x * boards_per_cta;
    long b1 = b0 + boards_per_cta;
    if (b1 > B) b1 = B;
    float w = 0.f, bias = 0.f;
    for (long b = b0; b < b1; b++) {
        const float d = dlogits[b * 4 + o];
        w = fmaf(d, h[b * 64 + j], w);
        bias += d;
    }
    hpart[(size_t)blockIdx.x * 1024 + threadIdx.x] = w;
    if (j == 0) hpart[(size_t)blockIdx.x * 1024 + 256 + o] = bias;
}

// generic: out[i] = sum_k hpart[k*1024 + i], i < n.  8 lanes per output (fixed shuffle tree -> deterministic).
__global__ void __launch_bounds__(256) hpart_reduce_kernel(const float *__restrict__ hpart, int nparts, int n, float *__restrict__ out) {
    const int i = blockIdx.x * 32 + (threadIdx.x >> 3), lane = threadIdx.x & 7;
    float t = 0.f;
    if (i < n)
        for (int k = lane; k < nparts; k += 8) t += hpart[(size_t)k * 1024 + i];
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    t += __shfl_xor_sync(0xffffffffu, t, 4);
    if (i < n && lane == 0) out[i] = t;
}

// dh = dlogits * FCW, through ReLU: dy7[row][c4] = dh[c4*16+p] * (h > 0); BN7 partial sums (stride 8)
__global__ void __launch_bounds__(256) head_bwd1_kernel(const float *__restrict__ dlogits, const float *__restrict__ fcw,
                                                        const float *__restrict__ h, const float *__restrict__ z7,
                                                        const float *__restrict__ stats7, float *__restrict__ dy7,
                                                        float *__restrict__ part, long rows) {
    __shared__ float w_s[256];
    __shared__ float red[8][8];
    w_s[threadIdx.x] = fcw[threadIdx.x];
    __syncthreads();
    const float4 mean = *reinterpret_cast<const float4 *>(stats7), istd = *reinterpret_cast<const float4 *>(stats7 + 64);
    const float mv[4] = {mean.x, mean.y, mean.z, mean.w}, iv[4] = {istd.x, istd.y, istd.z, istd.w};
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
    for (long r = (long)blockIdx.x * 256 + threadIdx.x; r < rows; r += (long)gridDim.x * 256) {
        const long b = r >> 4;
        const int p = (int)(r & 15);
        const float4 d = *reinterpret_cast<const float4 *>(dlogits + b * 4);
        const float4 zv = *reinterpret_cast<const float4 *>(z7 + r * 4);
        const float zz[4] = {zv.x, zv.y, zv.z, zv.w};
        float g[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int j = c * 16 + p;
            const float dh = d.x * w_s[j] + d.y * w_s[64 + j] + d.z * w_s[128 + j] + d.w * w_s[192 + j];
            g[c] = h[b * 64 + j] > 0.f ? dh : 0.f;
            s[c] += g[c];
            q[c] = fmaf(g[c], (zz[c] - mv[c]) * iv[c], q[c]);
        }
        *reinterpret_cast<float4 *>(dy7 + r * 4) = make_float4(g[0], g[1], g[2], g[3]);
    }
#pragma unroll
    for (int c = 0; c < 4; c++)
        for (int o = 16; o > 0; o >>= 1) s[c] += __shfl_xor_sync(0xffffffffu, s[c], o), q[c] += __shfl_xor_sync(0xffffffffu, q[c], o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
#pragma unroll
        for (int c = 0; c < 4; c++) red[warp][c] = s[c], red[warp][4 + c] = q[c];
    __syncthreads();
    if (threadIdx.x < 8) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) t += red[w][threadIdx.x];
        part[(size_t)blockIdx.x * 8 + threadIdx.x] = t;
    }
}

// dz7 = BN7 backward of dy7; da6[row][ci] = sum_c4 dz7[c4]*W7[c4][ci]; dW7[c4][ci] partial = sum_rows dz7[c4]*a6[row][ci].
// 256 threads = 4 row lanes x 64 ci; CTA partial -> hpart[cta*1024 + c4*64 + ci].
__global__ void __launch_bounds__(256) head_bwd2_kernel(const float *__restrict__ dy7, const float *__restrict__ z7,
                                                        const float *__restrict__ stats7, const float *__restrict__ coef7,
                                                        const float *__restrict__ w7, const float *__restrict__ a6,
                                                        float *__restrict__ da6, float *__restrict__ hpart, long rows,
                                                        int rows_per_cta) {
    __shared__ float red[4][4][64];
    const int ci = threadIdx.x & 63, rl = threadIdx.x >> 6;
    const float w[4] = {w7[ci], w7[64 + ci], w7[128 + ci], w7[192 + ci]};
    const float4 mean = *reinterpret_cast<const float4 *>(stats7), istd = *reinterpret_cast<const float4 *>(stats7 + 64),
                 sc = *reinterpret_cast<const float4 *>(stats7 + 128), c1 = *reinterpret_cast<const float4 *>(coef7),
                 c2 = *reinterpret_cast<const float4 *>(coef7 + 64);
    const long r0 = (long)blockIdx.x * rows_per_cta;
    long r1 = r0 + rows_per_cta;
    if (r1 > rows) r1 = rows;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (long r = r0 + rl; r < r1; r += 4) {
        const float4 g = *reinterpret_cast<const float4 *>(dy7 + r * 4), zv = *reinterpret_cast<const float4 *>(z7 + r * 4);
        float d[4];
        d[0] = sc.x * (g.x - c1.x - (zv.x - mean.x) * istd.x * c2.x);
        d[1] = sc.y * (g.y - c1.y - (zv.y - mean.y) * istd.y * c2.y);
        d[2] = sc.z * (g.z - c1.z - (zv.z - mean.z) * istd.z * c2.z);
        d[3] = sc.w * (g.w - c1.w - (zv.w - mean.w) * istd.w * c2.w);
        const float av = a6[r * 64 + ci];
        da6[r * 64 + ci] = d[0] * w[0] + d[1] * w[1] + d[2] * w[2] + d[3] * w[3];
#pragma unroll
        for (int c = 0; c < 4; c++) acc[c] = fmaf(d[c], av, acc[c]);
    }
#pragma unroll
    for (int c = 0; c < 4; c++) red[rl][c][ci] = acc[c];
    __syncthreads();
    {
        const int c = threadIdx.x >> 6;
        hpart[(size_t)blockIdx.x * 1024 + threadIdx.x] = red[0][c][ci] + red[1][c][ci] + red[2][c][ci] + red[3][c][ci];
    }
}

// ---- Adam (trainer.py:36-38 torch.optim.Adam(lr, weight_decay): L2 added to the gradient, bias-corrected) ------
__global__ void adam_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
                            int n, float lr_over_bc1, float one_minus_beta1, float beta2, float one_minus_beta2, float sqrt_bc2,
                            float eps, float wd) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float w = p[i];
    const float gr = fmaf(wd, w, g[i]);
    const float mm = m[i] + (gr - m[i]) * one_minus_beta1;          // exp_avg.lerp_(grad, 1 - beta1)
    const float vv = fmaf(one_minus_beta2, gr * gr, beta2 * v[i]);  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    m[i] = mm, v[i] = vv;
    p[i] = w - lr_over_bc1 * (mm / (sqrtf(vv) / sqrt_bc2 + eps));
}

}  // namespace train
}  // namespace b2048
