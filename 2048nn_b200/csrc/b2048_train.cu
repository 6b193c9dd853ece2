// b2048_train.cu -- extern "C" entry points of the c64b3 training step (trainer.py:14-97), kernels in train.cuh.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/b2048.h"
#include "capi_internal.h"
#include "train.cuh"
#include "train_tc.cuh"

using namespace b2048::train;

static bool g_attr_done[64] = {};
// mixed mode: the tensor-core weight gradient of layer l overlaps the data gradient / BatchNorm backward of the layers
// below it on a side stream (fork/join with events; legal under stream capture, so the CUDA graph keeps the branches)
static cudaStream_t g_side[64] = {};
static cudaEvent_t g_fork[64][NL] = {}, g_done[64][NL] = {};

static int ensure_attrs(const b2048_ctx *c) {
    if (c->device >= 0 && c->device < 64 && g_attr_done[c->device]) return 0;
    CK(cudaFuncSetAttribute(conv3x3_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvSmem<64>::BYTES));
    CK(cudaFuncSetAttribute(conv3x3_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvSmem<64>::BYTES));
    CK(cudaFuncSetAttribute(conv3x3_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvSmem<16>::BYTES));
    CK(cudaFuncSetAttribute(wgrad_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, wgrad_smem_bytes<64>()));
    CK(cudaFuncSetAttribute(wgrad_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, wgrad_smem_bytes<16>()));
    CK(cudaFuncSetAttribute(tc_conv_layer_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCL_SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc_conv_layer_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCL_SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TCW_SMEM_BYTES));
    if (c->device >= 64) return b2048_fail_msg("b2048 training: device index above 63");
    CK(cudaStreamCreateWithFlags(&g_side[c->device], cudaStreamNonBlocking));
    for (int l = 0; l < NL; l++) {
        CK(cudaEventCreateWithFlags(&g_fork[c->device][l], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&g_done[c->device][l], cudaEventDisableTiming));
    }
    g_attr_done[c->device] = true;
    return 0;
}

extern "C" int64_t b2048_train_workspace_bytes(int64_t max_batch) {
    if (max_batch <= 0) return 0;
    return (int64_t)carve(nullptr, max_batch).bytes;
}
extern "C" int64_t b2048_train_grad_offset(int64_t max_batch) {
    if (max_batch <= 0) return -1;
    return (int64_t)reinterpret_cast<uintptr_t>(carve(nullptr, max_batch).grad);
}
extern "C" int64_t b2048_train_logp_offset(int64_t max_batch) {
    if (max_batch <= 0) return -1;
    return (int64_t)reinterpret_cast<uintptr_t>(carve(nullptr, max_batch).logp);
}

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

extern "C" int b2048_train_forward_backward(b2048_ctx *c, const float *params, float *bn_buffers, void *workspace,
                                            int64_t max_batch, const uint64_t *boards, const float *targets, int64_t B,
                                            float bn_momentum, int mixed, float *loss_out, void *stream) {
    DeviceGuard dg(c->device);
    if (!params || !bn_buffers || !workspace || !boards || !targets || !loss_out)
        return b2048_fail_msg("b2048_train_forward_backward: NULL argument");
    if (B < 2 || B > max_batch) return b2048_fail_msg("b2048_train_forward_backward: need 2 <= B <= max_batch");
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) || (reinterpret_cast<uintptr_t>(targets) & 15))
        return b2048_fail_msg("b2048_train_forward_backward: workspace must be 256-byte aligned, targets 16-byte aligned");
    if (int rc = ensure_attrs(c)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const Workspace w = carve(workspace, max_batch);
    const long rows = (long)B * 16, n4 = (long)B * 256;
    const double count = (double)rows;
    const float eps = 1e-5f;
    float *run_mean = bn_buffers, *run_var = bn_buffers + 8 * 64;

    const int tiles = cdiv(B, 32);
    const int g4 = 2 * c->num_sms < MAX_PARTS ? 2 * c->num_sms : MAX_PARTS;
    const size_t chunk3 = (size_t)3 * 192 * 64;  // 16-bit elements of one layer's three weight chunks
    if (mixed == 2 && (B & 31)) {
        // the tensor-core weight gradient sums over all 32 rows of a tile: rows of boards past the batch end must be 0
        const size_t last = (size_t)(tiles - 1) * TCL_IMG_BYTES;
        CK(cudaMemsetAsync(w.img_dz + last, 0, TCL_IMG_BYTES, st));
        CK(cudaMemsetAsync(w.img_dz2 + last, 0, TCL_IMG_BYTES, st));
        for (int l = 0; l < NL; l++) CK(cudaMemsetAsync(w.img_abf[l] + last, 0, TCL_IMG_BYTES, st));
    }
    if (mixed) {
        pack_weights_tc_kernel<<<LC(c, cdiv((21 + 18) * 192 * 64, 256)), 256, 0, st>>>(params, w.wtc_fwd, w.wtc_bwd);
        onehot_image_kernel<<<LC(c, cdiv((long)tiles * 512, 256)), 256, 0, st>>>((const unsigned long long *)boards, w.img_a, w.img_abf[0],
                                                                          (long)B);
    } else {
        pack_weights_kernel<<<LC(c, cdiv(7 * 9 * 4096, 256)), 256, 0, st>>>(params, w.wf, w.wd);
    }
    onehot_kernel<<<LC(c, cdiv(rows, 256)), 256, 0, st>>>((const unsigned long long *)boards, w.x0, (long)B);
    // ---- forward (network.py:78-87 in training mode: BatchNorm uses batch statistics) -----------------------
    const int conv_grid = cdiv(B, CONV_BOARDS);
    const int fat_grid = cdiv(n4, 1024) < c->num_sms ? cdiv(n4, 1024) : c->num_sms;  // finalize + apply kernels: 1024 threads per CTA
    for (int l = 0; l < NL; l++) {
        const float *res = (l >= 2 && (l & 1) == 0) ? w.a[l - 2] : nullptr;
        if (mixed) {
            tc_conv_layer_kernel<false, true><<<LC(c, tiles), TCL_THREADS, TCL_SMEM_BYTES, st>>>(
                w.img_a, reinterpret_cast<const uint8_t *>(w.wtc_fwd + (size_t)l * chunk3), w.z[l], w.part, (long)B, l == 0 ? 2 : 4);
            const Image16 im = {w.img_a, l + 1 < NL ? w.img_abf[l + 1] : nullptr};
            bn_apply_fused_kernel<true><<<LC(c, fat_grid), 1024, 0, st>>>(w.part, tiles, count, params + P_G(l), params + P_B(l),
                                                                   run_mean + l * 64, run_var + l * 64, bn_momentum, eps,
                                                                   w.stats + l * 256, (const float4 *)w.z[l], (const float4 *)res,
                                                                   (float4 *)w.a[l], im, n4);
            continue;
        }
        if (l == 0)
            conv3x3_kernel<16, true><<<LC(c, conv_grid), CONV_THREADS, ConvSmem<16>::BYTES, st>>>(w.x0, w.wf, w.z[0], w.part, (long)B);
        else
            conv3x3_kernel<64, true><<<LC(c, conv_grid), CONV_THREADS, ConvSmem<64>::BYTES, st>>>(w.a[l - 1], w.wf + (size_t)l * 36864,
                                                                                           w.z[l], w.part, (long)B);
        bn_apply_fused_kernel<false><<<LC(c, fat_grid), 1024, 0, st>>>(w.part, conv_grid, count, params + P_G(l), params + P_B(l),
                                                                run_mean + l * 64, run_var + l * 64, bn_momentum, eps, w.stats + l * 256,
                                                                (const float4 *)w.z[l], (const float4 *)res, (float4 *)w.a[l], Image16{},
                                                                n4);
    }
    float *stats7 = w.stats + 7 * 256, *coef7 = w.coef + 7 * 128;
    const int g1 = cdiv(rows, 256) < MAX_PARTS ? cdiv(rows, 256) : MAX_PARTS;
    head_conv_kernel<<<LC(c, g1), 256, 0, st>>>(w.a[6], params + P_W7, w.z7, w.part, rows);
    bn_finalize_kernel<<<LC(c, 1), 1024, 0, st>>>(w.part, g1, 8, count, params + P_G7, params + P_B7, run_mean + 7 * 64, run_var + 7 * 64,
                                         bn_momentum, eps, stats7, 4);
    const int gfc = cdiv(B, 128);
    if (gfc > MAX_PARTS) return b2048_fail_msg("b2048_train_forward_backward: batch too large for the loss partials");
    head_fc_kernel<<<LC(c, gfc), 128, 0, st>>>(w.z7, stats7, params + P_FCW, params + P_FCB, targets, w.h, w.logp, w.dlogits, w.hpart, (long)B);
    loss_finalize_kernel<<<LC(c, 1), 32, 0, st>>>(w.hpart, gfc, (double)B, loss_out);
    // ---- backward: head -----------------------------------------------------------------------------------
    const int g2 = c->num_sms < MAX_PARTS ? c->num_sms : MAX_PARTS;
    fc_wgrad_kernel<<<LC(c, g2), 256, 0, st>>>(w.h, w.dlogits, w.hpart, (long)B, cdiv(B, g2));
    hpart_reduce_kernel<<<LC(c, cdiv(260, 32)), 256, 0, st>>>(w.hpart, g2, 260, w.grad + P_FCW);
    head_bwd1_kernel<<<LC(c, g1), 256, 0, st>>>(w.dlogits, params + P_FCW, w.h, w.z7, stats7, w.dy7, w.part, rows);
    bn_bwd_finalize_kernel<<<LC(c, 1), 1024, 0, st>>>(w.part, g1, 8, count, w.grad + P_G7, w.grad + P_B7, coef7, 4);
    const int g3 = 2 * c->num_sms < MAX_PARTS ? 2 * c->num_sms : MAX_PARTS;
    head_bwd2_kernel<<<LC(c, g3), 256, 0, st>>>(w.dy7, w.z7, stats7, coef7, params + P_W7, w.a[6], w.da, w.hpart, rows, cdiv(rows, g3));
    hpart_reduce_kernel<<<LC(c, cdiv(256, 32)), 256, 0, st>>>(w.hpart, g3, 256, w.grad + P_W7);
    // ---- backward: trunk ----------------------------------------------------------------------------------
    // a[l2] = relu(bn(z[l2]) + a[l2-2]) for l2 = 2,4,6: dy of a block output (dyA) is also the residual gradient of
    // the block input, added when that layer's own dy is formed.
    const int bpc = cdiv(B, WGRAD_SPLITS), splits = cdiv(B, bpc);
    for (int l = NL - 1; l >= 0; l--) {
        const bool even = (l & 1) == 0;
        float *dybuf = even ? w.dyA : w.dyB;
        const float *gres = (even && l != NL - 1) ? w.dyA : nullptr;
        bn_bwd_reduce_kernel<<<LC(c, g4), 256, 0, st>>>((const float4 *)w.da, (const float4 *)gres, (const float4 *)w.a[l],
                                                 (const float4 *)w.z[l], w.stats + l * 256, (float4 *)dybuf, w.part, rows);
        // The weight gradient of layer l runs on a side stream, overlapping the data gradient and the BatchNorm backward
        // of the layers below; dz (fp32 and image) is double-buffered by layer parity, and the buffer's previous
        // reader -- the weight gradient of layer l+2 -- must be done before it is rewritten.
        uint8_t *dz_img = (l & 1) ? w.img_dz2 : w.img_dz;
        float *dz = (l & 1) ? w.dz2 : w.dz;
        // (fp32 mode keeps it on the main stream: its FFMA2 weight-gradient and convolution kernels each fill the SMs, and
        // running them concurrently measured 8 % slower.)
        const bool forked = mixed != 0;
        cudaStream_t side = forked ? g_side[c->device] : st;
        if (forked && l + 2 < NL) CK(cudaStreamWaitEvent(st, g_done[c->device][l + 2], 0));
        if (mixed)
            bn_bwd_apply_fused_kernel<true><<<LC(c, fat_grid), 1024, 0, st>>>(w.part, g4, count, w.grad + P_G(l), w.grad + P_B(l),
                                                                       w.stats + l * 256, (const float4 *)dybuf, (const float4 *)w.z[l],
                                                                       (float4 *)dz, Image16{dz_img, nullptr}, n4);
        else
            bn_bwd_apply_fused_kernel<false><<<LC(c, fat_grid), 1024, 0, st>>>(w.part, g4, count, w.grad + P_G(l), w.grad + P_B(l),
                                                                        w.stats + l * 256, (const float4 *)dybuf, (const float4 *)w.z[l],
                                                                        (float4 *)dz, Image16{}, n4);
        if (forked) {
            CK(cudaEventRecord(g_fork[c->device][l], st));
            CK(cudaStreamWaitEvent(side, g_fork[c->device][l], 0));
        }
        if (mixed == 2) {  // tensor cores, bf16 operand images
            tc_wgrad_kernel<<<LC(c, tiles), TCL_THREADS, TCW_SMEM_BYTES, side>>>(w.img_abf[l], dz_img, w.wpart);
            wgrad_reduce_kernel<<<LC(c, 36864 / 64), 256, 0, side>>>(w.wpart, tiles, w.grad + P_W(l), l == 0 ? 16 : 64);
        } else if (l == 0) {
            wgrad_kernel<16><<<LC(c, dim3(splits, 3)), WG_THREADS, wgrad_smem_bytes<16>(), side>>>(w.x0, dz, w.wpart, (long)B, bpc);
            wgrad_reduce_kernel<<<LC(c, 36864 / 64), 256, 0, side>>>(w.wpart, splits, w.grad + P_W(0), 16);
        } else {
            wgrad_kernel<64><<<LC(c, dim3(splits, 3)), WG_THREADS, wgrad_smem_bytes<64>(), side>>>(w.a[l - 1], dz, w.wpart, (long)B, bpc);
            wgrad_reduce_kernel<<<LC(c, 36864 / 64), 256, 0, side>>>(w.wpart, splits, w.grad + P_W(l), 64);
        }
        if (forked) {
            CK(cudaEventRecord(g_done[c->device][l], side));
            if (l == 0) CK(cudaStreamWaitEvent(st, g_done[c->device][0], 0));  // join: the side stream is in order
        }
        if (l > 0) {
            if (mixed)
                tc_conv_layer_kernel<true, false><<<LC(c, tiles), TCL_THREADS, TCL_SMEM_BYTES, st>>>(
                    dz_img, reinterpret_cast<const uint8_t *>(w.wtc_bwd + (size_t)(l - 1) * chunk3), w.da, nullptr, (long)B, 4);
            else
                conv3x3_kernel<64, false><<<LC(c, conv_grid), CONV_THREADS, ConvSmem<64>::BYTES, st>>>(dz, w.wd + (size_t)l * 36864, w.da,
                                                                                                nullptr, (long)B);
        }
    }
    CK(cudaGetLastError());
    return 0;
}

extern "C" int b2048_train_adam(b2048_ctx *c, float *params, const float *grad, float *adam_m, float *adam_v, double lr,
                                double beta1, double beta2, double eps, double weight_decay, int64_t step, void *stream) {
    DeviceGuard dg(c->device);
    if (!params || !grad || !adam_m || !adam_v) return b2048_fail_msg("b2048_train_adam: NULL argument");
    if (step < 1) return b2048_fail_msg("b2048_train_adam: step counts from 1");
    // scalars are formed in double and rounded once, as torch.optim.Adam does with its Python floats
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    adam_kernel<<<LC(c, cdiv(N_PARAMS, 256)), 256, 0, (cudaStream_t)stream>>>(params, grad, adam_m, adam_v, N_PARAMS, (float)(lr / bc1),
                                                                       (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2),
                                                                       (float)sqrt(bc2), (float)eps, (float)weight_decay);
    CK(cudaGetLastError());
    return 0;
}
