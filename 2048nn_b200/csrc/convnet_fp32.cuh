// convnet_fp32.cuh -- exact-mode c64b3 forward on CUDA cores (fp32 FMA), sm_100a.
//
// ConvNet(channels=64, blocks=3) of network.py:43-87 with BatchNorm folded on the host in
// float64 (eval-mode running stats, eps 1e-5) and the inline one-hot encoder of
// mcts_nn.py:25-35 fused in.  This is the bit-close mode (|dlogp| ~1e-5 vs PyTorch fp32) used
// for parity anchoring and for stream-exact NN playouts; the throughput mode is the tcgen05
// kernel in convnet_tc.cuh.
//
// One CTA = 512 threads = 8 boards x 64 output channels; thread (b, co) owns the 16 spatial
// outputs of its channel in registers.  Activations ping-pong through shared memory as
// [board][channel][16 positions] fp32; folded weights are read through the read-only path as
// [ci][tap][co] (co fastest -> coalesced, L1-resident across the CTA's warps).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2048 {

// fp32 blob layout (floats); see 2048nn_b200/convnet.py:pack_fp32
#define CN_C 64
#define CN_L0_W 0                              // [16 ci][9 tap][64 co]
#define CN_L0_B (CN_L0_W + 16 * 9 * 64)        // [64]
#define CN_LW(l) (CN_L0_B + 64 + (l) * (64 * 9 * 64 + 64))  // l = 0..5: [64 ci][9 tap][64 co] then bias [64]
#define CN_HEAD (CN_LW(6))                     // w7 [4][64], b7 [4], fcw [4 out][64 in], fcb [4]
#define CN_FP32_FLOATS (CN_HEAD + 256 + 4 + 256 + 4)

#define CN_BOARDS_PER_CTA 8

struct ConvAcc {
    float v[16];
};

// 3x3 conv, pad 1, over a 4x4 map for one (board, co): acc[p] += sum_tap in[p+tap] * w[tap].
// Only in-bounds taps are enumerated (100 of 144), resolved at compile time.
__device__ __forceinline__ void conv_ci(ConvAcc &acc, const float *__restrict__ in16, const float *__restrict__ w,
                                        int wstride) {
    float x[16];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        float4 t = reinterpret_cast<const float4 *>(in16)[q];
        x[4 * q] = t.x, x[4 * q + 1] = t.y, x[4 * q + 2] = t.z, x[4 * q + 3] = t.w;
    }
    float wt[9];
#pragma unroll
    for (int t = 0; t < 9; t++) wt[t] = __ldg(w + t * wstride);
#pragma unroll
    for (int py = 0; py < 4; py++)
#pragma unroll
        for (int px = 0; px < 4; px++)
#pragma unroll
            for (int ky = 0; ky < 3; ky++)
#pragma unroll
                for (int kx = 0; kx < 3; kx++) {
                    int iy = py + ky - 1, ix = px + kx - 1;
                    if (iy >= 0 && iy < 4 && ix >= 0 && ix < 4)
                        acc.v[py * 4 + px] = fmaf(x[iy * 4 + ix], wt[ky * 3 + kx], acc.v[py * 4 + px]);
                }
}

// CTA-cooperative forward of up to 8 boards.  `boards` are the CTA's boards (shared or
// registers of threads 0..7 broadcast through smem by the caller); results: logits[b][4]
// (pre-softmax FC outputs) written to smem `s_logits`.
// smem: act0, act1: [8][64][16] floats each; s_head: [8][64] floats.
__device__ __forceinline__ void convnet_fp32_cta(const float *__restrict__ blob, const unsigned long long *s_boards,
                                                 int nb, float *act0, float *act1, float *s_head, float *s_logits) {
    const int tid = threadIdx.x;
    const int co = tid & 63, b = tid >> 6;
    ConvAcc acc;
    const bool live = b < nb;

    // ---- L0: one-hot gather (exactly the non-zero products of the dense conv) ----------------
    {
        const float *w = blob + CN_L0_W;
        float bias = __ldg(blob + CN_L0_B + co);
#pragma unroll
        for (int p = 0; p < 16; p++) acc.v[p] = bias;
        if (live) {
            unsigned long long brd = s_boards[b];
#pragma unroll
            for (int q = 0; q < 16; q++) {  // input cell q holds exponent t -> channel t is 1 there
                int t = (int)((brd >> (4 * q)) & 0xF);
                const float *wq = w + (t * 9) * 64 + co;
                int qy = q >> 2, qx = q & 3;
#pragma unroll
                for (int ky = 0; ky < 3; ky++)
#pragma unroll
                    for (int kx = 0; kx < 3; kx++) {
                        int py = qy - (ky - 1), px = qx - (kx - 1);  // output cell fed by (q, tap)
                        if (py >= 0 && py < 4 && px >= 0 && px < 4) acc.v[py * 4 + px] += __ldg(wq + (ky * 3 + kx) * 64);
                    }
            }
        }
        float *o = act0 + (b * 64 + co) * 16;
#pragma unroll
        for (int p = 0; p < 16; p++) o[p] = fmaxf(acc.v[p], 0.f);
    }
    __syncthreads();

    // ---- 3 residual blocks: x (act0) -> h (act1) -> y (act0) ---------------------------------
#pragma unroll 1
    for (int l = 0; l < 6; l++) {
        const float *w = blob + CN_LW(l);
        const float *in = (l & 1) ? act1 : act0;
        float *out = (l & 1) ? act0 : act1;
        float bias = __ldg(w + 64 * 9 * 64 + co);
#pragma unroll
        for (int p = 0; p < 16; p++) acc.v[p] = bias;
        if (live) {
#pragma unroll 2
            for (int ci = 0; ci < 64; ci++) conv_ci(acc, in + (b * 64 + ci) * 16, w + ci * 9 * 64 + co, 64);
        }
        float *o = out + (b * 64 + co) * 16;
        if (l & 1) {  // second conv of a block: residual add with the block input (in act0 == out)
#pragma unroll
            for (int p = 0; p < 16; p++) o[p] = fmaxf(acc.v[p] + o[p], 0.f);
        } else {
#pragma unroll
            for (int p = 0; p < 16; p++) o[p] = fmaxf(acc.v[p], 0.f);
        }
        __syncthreads();
    }

    // ---- head: 1x1 conv 64->4 (+BN+ReLU), flatten c*16+p, Linear 64->4 -----------------------
    {
        const float *hw = blob + CN_HEAD;
        // thread (b, j): j = c4*16 + p, 64 per board
        int j = co, c4 = j >> 4, p = j & 15;
        float s = __ldg(hw + 256 + c4);
        if (live) {
            const float *x = act0 + (b * 64) * 16 + p;
#pragma unroll 8
            for (int ci = 0; ci < 64; ci++) s = fmaf(x[ci * 16], __ldg(hw + c4 * 64 + ci), s);
        }
        s_head[b * 64 + j] = fmaxf(s, 0.f);
    }
    __syncthreads();
    if (co < 4 && live) {
        const float *fw = blob + CN_HEAD + 260 + co * 64;
        float s = __ldg(blob + CN_HEAD + 260 + 256 + co);
#pragma unroll 8
        for (int j = 0; j < 64; j++) s = fmaf(s_head[b * 64 + j], __ldg(fw + j), s);
        s_logits[b * 4 + co] = s;
    }
    __syncthreads();
}

__device__ __forceinline__ void log_softmax4(const float *z, float *out) {
    float m = fmaxf(fmaxf(z[0], z[1]), fmaxf(z[2], z[3]));
    float s = expf(z[0] - m) + expf(z[1] - m) + expf(z[2] - m) + expf(z[3] - m);
    float l = m + logf(s);
#pragma unroll
    for (int k = 0; k < 4; k++) out[k] = z[k] - l;
}

// Forward-only kernel: boards u64[N] -> log-probs f32[N][4] (network.py:78-87 output).
__global__ void __launch_bounds__(512) convnet_fp32_forward_kernel(const float *__restrict__ blob,
                                                                   const unsigned long long *__restrict__ boards,
                                                                   float *__restrict__ out_logp, long n) {
    extern __shared__ __align__(16) float smem[];
    float *act0 = smem, *act1 = smem + 8 * 64 * 16, *s_head = act1 + 8 * 64 * 16, *s_logits = s_head + 8 * 64;
    unsigned long long *s_boards = reinterpret_cast<unsigned long long *>(s_logits + 32);
    for (long base = (long)blockIdx.x * CN_BOARDS_PER_CTA; base < n; base += (long)gridDim.x * CN_BOARDS_PER_CTA) {
        int nb = (int)((n - base) < CN_BOARDS_PER_CTA ? (n - base) : CN_BOARDS_PER_CTA);
        if (threadIdx.x < nb) s_boards[threadIdx.x] = boards[base + threadIdx.x];
        __syncthreads();
        convnet_fp32_cta(blob, s_boards, nb, act0, act1, s_head, s_logits);
        if (threadIdx.x < nb) {
            float lp[4];
            log_softmax4(s_logits + threadIdx.x * 4, lp);
            reinterpret_cast<float4 *>(out_logp)[base + threadIdx.x] = make_float4(lp[0], lp[1], lp[2], lp[3]);
        }
        __syncthreads();
    }
}
#define CN_FP32_SMEM_BYTES ((2 * 8 * 64 * 16 + 8 * 64 + 32) * 4 + 8 * 8)

}  // namespace b2048
