// b2048_capi.cu -- the extern "C" boundary of lib2048nn_b200.so (declared in include/b2048.h).
// Plain pointers and sizes only; no torch types.  Every call enqueues on the caller's stream.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/b2048.h"
#include "capi_internal.h"
#include "engine_kernels.cuh"
#include "playout_nn.cuh"
#include "selfplay_kernels.cuh"
#include "probe.cuh"

using namespace b2048;

static thread_local char g_err[512] = "";

int b2048_fail(const char *what, cudaError_t e) {
    snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
    return 1;
}
int b2048_fail_msg(const char *what) {
    snprintf(g_err, sizeof g_err, "%s", what);
    return 2;
}
#define fail b2048_fail
#define fail_msg b2048_fail_msg

static Tables tables_of(const b2048_ctx *c) {
    Tables t;
    t.fin_fwd = c->fin_fwd;
    t.fin_rev = c->fin_rev;
    t.score = c->score;
    t.legal = c->legal;
    t.rowstat = c->rowstat;
    t.pair = c->pair;
    return t;
}

extern "C" int b2048_abi_version(void) { return B2048_ABI_VERSION; }
extern "C" const char *b2048_last_error(void) { return g_err; }

extern "C" int b2048_create(int device, b2048_ctx **out) {
    if (!out) return fail_msg("b2048_create: out is NULL");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail_msg("b2048_create: no such CUDA device");
    DeviceGuard dg(device);  // the caller's current device is restored on return
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail_msg("b2048_create: this library is built for sm_100a (Blackwell) only");
    b2048_ctx *c = new b2048_ctx();
    c->device = device;
    c->next_counter = 0;
    c->launches = 0;
    c->num_sms = prop.multiProcessorCount;
    CK(cudaMalloc(&c->fin_fwd, 65536 * sizeof(uint16_t)));
    CK(cudaMalloc(&c->fin_rev, 65536 * sizeof(uint16_t)));
    CK(cudaMalloc(&c->score, 65536 * sizeof(uint32_t)));
    CK(cudaMalloc(&c->moved_fwd, 65536));
    CK(cudaMalloc(&c->moved_rev, 65536));
    CK(cudaMalloc(&c->legal, 65536));
    CK(cudaMalloc(&c->rowstat, 65536 * sizeof(uint2)));
    CK(cudaMalloc(&c->pair, K3_ROWS * sizeof(uint32_t)));
    c->k2_variant = 2;
    c->tc_wide = 3;
    if (const char *e = getenv("B2048_LATENCY_MODE")) c->tc_wide = atoi(e) < 0 ? 0 : (atoi(e) > 3 ? 3 : atoi(e));  // A/B runs only
    // Measured on B200 (tools/debug_maingame.py): two tile-sets per CTA reach 1.9e8 policy moves/s when all 9,472 slots
    // are busy but take ~50 us per step; one tile-set (16-warp latency variant) takes ~31 us with 4,736 slots.  With a
    // line queue behind the slots the one-set kernel is the faster one up to ~10,000 lines in flight (profiles/r02_main_single_pct.txt).
    c->main_single_pct = 250;
    if (const char *e = getenv("B2048_MAIN_SINGLE_PCT")) c->main_single_pct = atoi(e);  // bring-up / A-B runs only
    CK(cudaMalloc(&c->counters, N_COUNTERS * sizeof(unsigned long long)));
    CK(cudaMemset(c->counters, 0, N_COUNTERS * sizeof(unsigned long long)));
    CK(cudaMalloc(&c->err_flag, 2 * sizeof(int)));
    CK(cudaMemset(c->err_flag, 0, 2 * sizeof(int)));
    build_tables_kernel<<<LC(c, 65536 / 256), 256>>>(c->fin_fwd, c->fin_rev, c->score, c->moved_fwd, c->moved_rev,
                                              c->legal, c->rowstat, c->pair, c->err_flag + 1);
    CK(cudaGetLastError());
    int mism = 0;
    CK(cudaMemcpy(&mism, c->err_flag + 1, sizeof(int), cudaMemcpyDeviceToHost));
    if (mism) return fail_msg("b2048_create: row-table invariants violated");
    CK(cudaFuncSetAttribute(playout_fixed_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    CK(cudaFuncSetAttribute(playout_fixed_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    CK(cudaFuncSetAttribute(playout_fixed_kernel2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, K2V2_SMEM_BYTES));
    CK(cudaFuncSetAttribute(playout_fixed_kernel2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, K2V2_SMEM_BYTES));
    CK(cudaFuncSetAttribute(playout_fixed_kernel3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, K2V3_SMEM_BYTES));
    CK(cudaFuncSetAttribute(playout_fixed_kernel3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, K2V3_SMEM_BYTES));
    CK(cudaFuncSetAttribute(convnet_fp32_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            CN_FP32_SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc::tc_persistent_kernel<tc::FwdWork, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            tc::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc::tc_persistent_kernel<tc::FwdWork, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            tc::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc::tc_persistent_kernel<PlayWork, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            tc::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc::tc_persistent_kernel<MainWork, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            tc::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc::tc_persistent_kernel<MainWork, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            tc::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc::tc_persistent_kernel<PlayWork, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            tc::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc::tc_persistent_kernel<tc::FwdWork, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            tc::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc::tc_persistent_kernel<PlaySpecWork, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            tc::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc::tc_persistent_kernel<MainSpecWork, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            tc::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc::tc_persistent_kernel<PlaySpec3Work, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            tc::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc::tc_persistent_kernel<PlayAdaptWork, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            tc::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc::tc_persistent_kernel<PlayAdaptWork, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            tc::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc::tc_persistent_kernel<PlaySpecWork, false, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc::tc_persistent_kernel<MainSpecWork, false, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc::tc_persistent_kernel<PlaySpec3Work, false, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc::tc_persistent_kernel<MainSpec3Work, false, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc::tc_persistent_kernel<PlayAdaptWork, false, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc::tc_persistent_kernel<MainSpec3Work, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            tc::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc::tc_persistent_kernel<tc::FwdWork, false, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            tc::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc::tc_persistent_kernel<PlayWork, false, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            tc::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc::tc_persistent_kernel<MainWork, false, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            tc::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc::tc_persistent_kernel<tc::FwdWork, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            tc::SMEM_BYTES));
    CK(cudaFuncSetAttribute(tc::tc_persistent_kernel<PlayWork, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            tc::SMEM_BYTES));
    CK(cudaFuncSetAttribute(playout_nn_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CN_FP32_SMEM_BYTES));
    *out = c;
    return 0;
}

extern "C" int b2048_destroy(b2048_ctx *c) {
    if (!c) return 0;
    DeviceGuard dg(c->device);
    cudaFree(c->fin_fwd);
    cudaFree(c->fin_rev);
    cudaFree(c->score);
    cudaFree(c->moved_fwd);
    cudaFree(c->moved_rev);
    cudaFree(c->legal);
    cudaFree(c->rowstat);
    cudaFree(c->pair);
    cudaFree(c->counters);
    cudaFree(c->err_flag);
    delete c;
    return 0;
}

extern "C" int b2048_num_sms(const b2048_ctx *c) { return c ? c->num_sms : 0; }
extern "C" int64_t b2048_launch_count(const b2048_ctx *c) { return c ? (int64_t)c->launches.load() : 0; }

/* Measurement aid (bench.py roofline.executed.measured_pipe_rates): one launch of the pipe-rate probe, #SMs CTAs x 1024
 * threads x iters x 64 probe instructions (kind 0: LOP3 = ALU pipe; kind 1: LOP3 + IMAD alternating = issue limit). */
extern "C" int b2048_probe_pipe_rate(b2048_ctx *c, int kind, uint32_t iters, uint32_t *sink, uint64_t *out_thread_instr,
                                     void *stream) {
    if (!c || !sink || !out_thread_instr) return fail_msg("b2048_probe_pipe_rate: NULL argument");
    if (kind < 0 || kind > 1) return fail_msg("b2048_probe_pipe_rate: kind must be 0 (LOP3) or 1 (LOP3 + IMAD)");
    DeviceGuard dg(c->device);
    *out_thread_instr = (uint64_t)c->num_sms * 1024ULL * (uint64_t)iters * (uint64_t)PROBE_UNROLL;
    if (iters == 0) return 0;
    if (kind == 0)
        pipe_probe_kernel<0><<<LC(c, c->num_sms), 1024, 0, (cudaStream_t)stream>>>(iters, 0x9E3779B9u, 0x7F4A7C15u, sink);
    else
        pipe_probe_kernel<1><<<LC(c, c->num_sms), 1024, 0, (cudaStream_t)stream>>>(iters, 0x9E3779B9u, 0x7F4A7C15u, sink);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int b2048_set_latency_mode(b2048_ctx *c, int mode) {
    if (mode < 0 || mode > 3) return fail_msg("b2048_set_latency_mode: mode must be 0, 1, 2 or 3");
    c->tc_wide.store(mode);
    return 0;
}

extern "C" int b2048_set_playout_variant(b2048_ctx *c, int variant) {
    if (variant < 0 || variant > 2) return fail_msg("b2048_set_playout_variant: variant must be 0, 1 or 2");
    c->k2_variant.store(variant);
    return 0;
}

/* device error flag: set when a spawn hit countzero == 0 (ValueError in the reference).
 * Reads and clears it; synchronises `stream`. */
extern "C" int b2048_take_error_flag(b2048_ctx *c, void *stream, int *out) {
    DeviceGuard dg(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemcpyAsync(out, c->err_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (*out) CK(cudaMemsetAsync(c->err_flag, 0, sizeof(int), st));
    return 0;
}

static inline int grid_for(long n, int threads, int cap) {
    long g = (n + threads - 1) / threads;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

extern "C" int b2048_tables(b2048_ctx *c, uint16_t *fin_fwd, uint16_t *fin_rev, uint32_t *score,
                            uint8_t *moved_fwd, uint8_t *moved_rev, void *stream) {
    DeviceGuard dg(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (fin_fwd) CK(cudaMemcpyAsync(fin_fwd, c->fin_fwd, 65536 * 2, cudaMemcpyDeviceToDevice, st));
    if (fin_rev) CK(cudaMemcpyAsync(fin_rev, c->fin_rev, 65536 * 2, cudaMemcpyDeviceToDevice, st));
    if (score) CK(cudaMemcpyAsync(score, c->score, 65536 * 4, cudaMemcpyDeviceToDevice, st));
    if (moved_fwd) CK(cudaMemcpyAsync(moved_fwd, c->moved_fwd, 65536, cudaMemcpyDeviceToDevice, st));
    if (moved_rev) CK(cudaMemcpyAsync(moved_rev, c->moved_rev, 65536, cudaMemcpyDeviceToDevice, st));
    return 0;
}

extern "C" int b2048_move_batch(b2048_ctx *c, const uint64_t *boards, const uint8_t *dirs, int dir_scalar,
                                uint64_t *out_boards, uint32_t *out_score, uint8_t *out_moved, int64_t n,
                                void *stream) {
    DeviceGuard dg(c->device);
    if (n <= 0) return 0;
    move_batch_kernel<<<LC(c, grid_for(n, 256, c->num_sms * 16)), 256, 0, (cudaStream_t)stream>>>(
        tables_of(c), (const u64 *)boards, dirs, dir_scalar, (u64 *)out_boards, out_score, out_moved, (long)n);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int b2048_board_ops(b2048_ctx *c, const uint64_t *boards, uint64_t *out_transpose,
                               uint8_t *out_countzero, uint8_t *out_legal_mask, int64_t n, void *stream) {
    DeviceGuard dg(c->device);
    if (n <= 0) return 0;
    board_ops_kernel<<<LC(c, grid_for(n, 256, c->num_sms * 16)), 256, 0, (cudaStream_t)stream>>>(
        tables_of(c), (const u64 *)boards, (u64 *)out_transpose, out_countzero, out_legal_mask, (long)n);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int b2048_init_boards(b2048_ctx *c, uint64_t seed, uint64_t gid_base, uint64_t *out_boards,
                                 int64_t n, void *stream) {
    DeviceGuard dg(c->device);
    if (n <= 0) return 0;
    init_boards_kernel<<<LC(c, grid_for(n, 256, c->num_sms * 16)), 256, 0, (cudaStream_t)stream>>>(
        seed, gid_base, (u64 *)out_boards, (long)n);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int b2048_spawn_batch(b2048_ctx *c, const uint64_t *boards, uint64_t seed, const uint64_t *gids,
                                 uint64_t gid_base, const uint32_t *k, uint32_t k_scalar, uint64_t *out_boards,
                                 int32_t *err_flag, int64_t n, void *stream) {
    DeviceGuard dg(c->device);
    if (n <= 0) return 0;
    spawn_batch_kernel<<<LC(c, grid_for(n, 256, c->num_sms * 16)), 256, 0, (cudaStream_t)stream>>>(
        (const u64 *)boards, seed, (const u64 *)gids, gid_base, k, k_scalar, (u64 *)out_boards, err_flag,
        (long)n);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int b2048_rng_draws(b2048_ctx *c, uint64_t seed, uint64_t gid_base, uint64_t first, int count,
                               uint64_t *out, int64_t n, void *stream) {
    DeviceGuard dg(c->device);
    if (n <= 0 || count <= 0) return 0;
    rng_draws_kernel<<<LC(c, grid_for(n, 256, c->num_sms * 16)), 256, 0, (cudaStream_t)stream>>>(
        seed, gid_base, first, count, (u64 *)out, (long)n);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int b2048_policy_step(b2048_ctx *c, const uint64_t *boards, const uint8_t *prio, uint64_t seed,
                                 const uint64_t *gids, uint64_t gid_base, const uint32_t *k, uint64_t *out_boards,
                                 uint32_t *out_gain, uint8_t *out_moved, int64_t n, void *stream) {
    DeviceGuard dg(c->device);
    if (n <= 0) return 0;
    if (!boards || !prio || !k || !out_boards || !out_gain || !out_moved)
        return fail_msg("b2048_policy_step: NULL argument");
    policy_step_kernel<<<LC(c, grid_for(n, 256, c->num_sms * 16)), 256, 0, (cudaStream_t)stream>>>(
        tables_of(c), (const u64 *)boards, prio, seed, (const u64 *)gids, gid_base, k, (u64 *)out_boards, out_gain,
        out_moved, c->err_flag, (long)n);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int b2048_encode_boards(b2048_ctx *c, const uint64_t *boards, int mode, float *out, int64_t n, void *stream) {
    DeviceGuard dg(c->device);
    if (n <= 0) return 0;
    if (mode != 0 && mode != 1) return fail_msg("b2048_encode_boards: mode must be 0 (exponents) or 1 (one-hot)");
    const long total = (long)n * (mode ? 64 : 4);
    encode_boards_kernel<<<LC(c, grid_for(total, 256, c->num_sms * 8)), 256, 0, (cudaStream_t)stream>>>(
        (const u64 *)boards, mode, reinterpret_cast<float4 *>(out), (long)n);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int b2048_soft_targets(b2048_ctx *c, const float *results, float soft, float *out, int64_t n, void *stream) {
    DeviceGuard dg(c->device);
    if (n <= 0) return 0;
    soft_targets_kernel<<<LC(c, grid_for(n, 256, c->num_sms * 8)), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(results), soft, reinterpret_cast<float4 *>(out), (long)n);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int b2048_game_summary(b2048_ctx *c, const uint64_t *finals, const uint32_t *score, const uint32_t *moves,
                                  uint8_t *out_maxtile, uint64_t *acc20, int64_t n, void *stream) {
    DeviceGuard dg(c->device);
    if (n <= 0) return 0;
    if (!finals || !score || !acc20) return fail_msg("b2048_game_summary: NULL argument");
    game_summary_kernel<<<LC(c, grid_for(n, 256, c->num_sms * 8)), 256, 0, (cudaStream_t)stream>>>(
        (const u64 *)finals, score, moves, out_maxtile, (unsigned long long *)acc20, (long)n);
    CK(cudaGetLastError());
    return 0;
}

static int launch_playout(b2048_ctx *c, PlayParams &p, cudaStream_t st) {
    p.tb = tables_of(c);
    p.err_flag = c->err_flag;
    p.counter = take_counters(c, 1);
    CK(cudaMemsetAsync(p.counter, 0, sizeof(unsigned long long), st));
    // one CTA per SM; fewer games than 1024 x #SMs are spread over all SMs with narrower CTAs instead of filling a few
    // SMs (the kernels take their stride from blockDim): 64 K games = 148 x 448 threads rather than 64 x 1024
    int grid = c->num_sms, block = 1024;
    if (p.n < (long)c->num_sms * 1024) {
        long per_sm = (p.n + c->num_sms - 1) / c->num_sms;
        block = (int)(((per_sm + 31) / 32) * 32);
        if (block < 64) block = 64;
        grid = (int)((p.n + block - 1) / block);
    }
    bool ludr = p.order[0] == 0 && p.order[1] == 1 && p.order[2] == 3 && p.order[3] == 2;
    const int variant = c->k2_variant.load();
    if (variant == 2) {
        if (ludr)
            playout_fixed_kernel3<true><<<LC(c, grid), block, K2V3_SMEM_BYTES, st>>>(p);
        else
            playout_fixed_kernel3<false><<<LC(c, grid), block, K2V3_SMEM_BYTES, st>>>(p);
    } else if (variant == 1) {
        if (ludr)
            playout_fixed_kernel2<true><<<LC(c, grid), block, K2V2_SMEM_BYTES, st>>>(p);
        else
            playout_fixed_kernel2<false><<<LC(c, grid), block, K2V2_SMEM_BYTES, st>>>(p);
    } else if (ludr) {
        playout_fixed_kernel<true><<<LC(c, grid), block, 131072, st>>>(p);
    } else {
        playout_fixed_kernel<false><<<LC(c, grid), block, 131072, st>>>(p);
    }
    CK(cudaGetLastError());
    return 0;
}

extern "C" int b2048_playout_fixed(b2048_ctx *c, const uint64_t *start, const uint8_t *order4_host,
                                   uint64_t seed, uint64_t gid_base, uint32_t spawn0, uint64_t *out_final,
                                   uint32_t *out_score, uint32_t *out_moves, int64_t n, void *stream) {
    DeviceGuard dg(c->device);
    if (n <= 0) return 0;
    PlayParams p;
    memset(&p, 0, sizeof p);
    p.n = (long)n;
    p.start = (const u64 *)start;
    p.seed = seed;
    p.gid_base = gid_base;
    p.spawn0 = spawn0;
    static const uint8_t ludr[4] = {0, 1, 3, 2};
    for (int a = 0; a < 4; a++) p.order[a] = (order4_host ? order4_host[a] : ludr[a]) & 3;
    p.out_final = (u64 *)out_final;
    p.out_score = out_score;
    p.out_moves = out_moves;
    return launch_playout(c, p, (cudaStream_t)stream);
}

extern "C" int b2048_mcts_fixed(b2048_ctx *c, const uint64_t *origins, int64_t G, int number, int line_begin,
                                int line_end, uint64_t seed, uint64_t eval_id_base, uint8_t *out_legal,
                                uint32_t *out_rscore, uint32_t *out_lscore, uint32_t *out_lmoves,
                                int64_t *out_logsum, int32_t *out_count, int32_t *out_minmov, void *stream) {
    DeviceGuard dg(c->device);
    if (G <= 0) return 0;
    if (number <= 0 || line_begin < 0 || line_end > number || line_begin > line_end)
        return fail_msg("b2048_mcts_fixed: bad line range");
    if (!out_legal || !out_rscore) return fail_msg("b2048_mcts_fixed: out_legal and out_rscore are required");
    cudaStream_t st = (cudaStream_t)stream;
    root_moves_kernel<<<LC(c, (unsigned)((G * 4 + 255) / 256)), 256, 0, st>>>(tables_of(c), (const u64 *)origins, (long)G,
                                                                       out_legal, out_rscore);
    CK(cudaGetLastError());
    if (line_begin == line_end) return 0;
    PlayParams p;
    memset(&p, 0, sizeof p);
    p.root_mode = 1;
    p.origins = (const u64 *)origins;
    p.rlegal = out_legal;
    p.rscore = out_rscore;
    p.number = number;
    p.line_begin = line_begin;
    p.line_count = line_end - line_begin;
    p.n = (long)G * 4 * p.line_count;
    p.seed = seed;
    p.gid_base = eval_id_base * 4ULL;
    p.spawn0 = 1;
    p.order[0] = 0, p.order[1] = 1, p.order[2] = 3, p.order[3] = 2;  // mcts.py:26
    p.out_score = out_lscore;
    p.out_moves = out_lmoves;
    p.out_logsum = (long long *)out_logsum;
    p.out_count = out_count;
    p.out_minmov = out_minmov;
    return launch_playout(c, p, st);
}

extern "C" int b2048_selfplay_fixed(b2048_ctx *c, const uint64_t *starts, int64_t G, uint64_t game_id_base,
                                    uint64_t games_total, int number, int times, int mode, uint64_t seed, int max_steps,
                                    uint64_t *out_boards, float *out_results, uint64_t *out_final, int64_t *out_score,
                                    int32_t *out_steps, int32_t *out_status, uint64_t *move_counter, void *stream) {
    DeviceGuard dg(c->device);
    if (G <= 0) return 0;
    if (number <= 0 || times <= 0 || times > SELFPLAY_MAX_TIMES || max_steps <= 0)
        return fail_msg("b2048_selfplay_fixed: need number > 0, 1 <= times <= 64, max_steps > 0");
    if (mode != B2048_MODE_MEANLOG && mode != B2048_MODE_MEDIAN && mode != B2048_MODE_MIN)
        return fail_msg("b2048_selfplay_fixed: unknown mode");
    if (mode != B2048_MODE_MIN && times != 1) return fail_msg("b2048_selfplay_fixed: times > 1 is defined for the min mode only");
    if (mode == B2048_MODE_MEDIAN && number > 2048) return fail_msg("b2048_selfplay_fixed: median mode supports number <= 2048");
    if (!out_final || !out_score || !out_steps || !out_status) return fail_msg("b2048_selfplay_fixed: NULL output");
    if (games_total < game_id_base + (uint64_t)G) return fail_msg("b2048_selfplay_fixed: games_total < game_id_base + G");
    SelfplayParams p;
    memset(&p, 0, sizeof p);
    p.tb = tables_of(c);
    p.starts = (const u64 *)starts;
    p.G = (long)G;
    p.g0 = game_id_base;
    p.g_total = games_total;
    p.number = number, p.times = times, p.mode = mode, p.max_steps = max_steps;
    p.seed = seed;
    p.out_boards = (u64 *)out_boards;
    p.out_results = out_results;
    p.out_final = (u64 *)out_final;
    p.out_score = (long long *)out_score;
    p.out_steps = out_steps;
    p.out_status = out_status;
    p.move_counter = (unsigned long long *)move_counter;
    p.err_flag = c->err_flag;
    long lines = (long)times * 4 * number;
    int threads = (int)(lines < 1024 ? ((lines + 31) / 32) * 32 : 1024);
    size_t smem = mode == B2048_MODE_MEDIAN ? (size_t)16 * number : 0;
    selfplay_fixed_kernel<<<LC(c, (unsigned)G), threads, smem, (cudaStream_t)stream>>>(p);
    CK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// Policy network (network.py:43-87) and NN-policy playouts (mcts_nn.py, eval_nn.py)
// ------------------------------------------------------------------------------------------
extern "C" int64_t b2048_convnet_blob_bytes(int precision) {
    if (precision == B2048_PREC_FP32) return (int64_t)CN_FP32_FLOATS * 4;
    if (precision == B2048_PREC_FP16_TC || precision == B2048_PREC_BF16_TC) return (int64_t)tc::BLOB_BYTES;
    if (precision == B2048_PREC_FP16X2_TC) return (int64_t)tc::BLOB_BYTES_PRECISE;
    return -1;
}

static int launch_tc_forward(b2048_ctx *c, const void *blob, const uint64_t *boards, float *out_logp, float *out_logits,
                             float *dbg, int dbg_layer, int64_t n, cudaStream_t st, bool bf16 = false, bool precise = false) {
    tc::FwdWork::Params p;
    p.boards = (const unsigned long long *)boards;
    p.out_logp = out_logp;
    p.out_logits = out_logits;
    p.n = (long)n;
    p.counter = take_counters(c, 1);
    CK(cudaMemsetAsync(p.counter, 0, sizeof(unsigned long long), st));
    int grid = grid_for(n, tc::TILE_BOARDS * tc::NSETS, c->num_sms);
    if (precise) {  // one tile-set per CTA whatever the batch size
        grid = grid_for(n, tc::TILE_BOARDS, c->num_sms);
        tc::tc_persistent_kernel<tc::FwdWork, false, false, true, true><<<LC(c, grid), tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(
            p, (const uint8_t *)blob, nullptr, -1, (long)n);
        CK(cudaGetLastError());
        return 0;
    }
    if (!dbg && !bf16 && c->tc_wide.load() && n <= (int64_t)c->num_sms * tc::TILE_BOARDS) {
        // latency mode: at most one 32-board tile per SM -> one tile-set per CTA, all 16 worker warps on its epilogue
        grid = grid_for(n, tc::TILE_BOARDS, c->num_sms);
        tc::tc_persistent_kernel<tc::FwdWork, false, false, true><<<LC(c, grid), tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(
            p, (const uint8_t *)blob, nullptr, -1, (long)n);
        CK(cudaGetLastError());
        return 0;
    }
    if (dbg)
        tc::tc_persistent_kernel<tc::FwdWork, true><<<LC(c, grid), tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(
            p, (const uint8_t *)blob, dbg, dbg_layer, (long)n);
    else if (bf16)
        tc::tc_persistent_kernel<tc::FwdWork, false, true><<<LC(c, grid), tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(
            p, (const uint8_t *)blob, nullptr, -1, (long)n);
    else
        tc::tc_persistent_kernel<tc::FwdWork, false><<<LC(c, grid), tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(
            p, (const uint8_t *)blob, nullptr, -1, (long)n);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int b2048_convnet_forward(b2048_ctx *c, int precision, const void *blob, const uint64_t *boards,
                                     float *out_logp, int64_t n, void *stream) {
    DeviceGuard dg(c->device);
    if (n <= 0) return 0;
    if (!blob || !boards || !out_logp) return fail_msg("b2048_convnet_forward: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == B2048_PREC_FP32) {
        int grid = grid_for(n, CN_BOARDS_PER_CTA, c->num_sms * 2);
        convnet_fp32_forward_kernel<<<LC(c, grid), 512, CN_FP32_SMEM_BYTES, st>>>((const float *)blob, (const u64 *)boards,
                                                                           out_logp, (long)n);
        CK(cudaGetLastError());
        return 0;
    }
    if (precision == B2048_PREC_FP16_TC || precision == B2048_PREC_BF16_TC || precision == B2048_PREC_FP16X2_TC)
        return launch_tc_forward(c, blob, boards, out_logp, nullptr, nullptr, -1, n, st, precision == B2048_PREC_BF16_TC,
                                 precision == B2048_PREC_FP16X2_TC);
    return fail_msg("b2048_convnet_forward: unknown precision");
}

/* Debug/bring-up entry of the tensor-core forward: pre-softmax logits and, when dbg != NULL, the
 * fp32 post-activation output of conv layer `dbg_layer` (0..6) as [board][64 co][16 pos]. */
extern "C" int b2048_convnet_tc_debug(b2048_ctx *c, const void *blob, const uint64_t *boards, float *out_logits,
                                      float *dbg, int dbg_layer, int64_t n, void *stream) {
    DeviceGuard dg(c->device);
    if (n <= 0) return 0;
    return launch_tc_forward(c, blob, boards, nullptr, out_logits, dbg, dbg ? dbg_layer : -1, n, (cudaStream_t)stream);
}

static int launch_nn(b2048_ctx *c, int precision, NNPlayParams &p, cudaStream_t st) {
    p.tb = tables_of(c);
    p.err_flag = c->err_flag;
    const int models = p.n_models > 1 ? p.n_models : 1;
    if (models > N_COUNTERS || models > c->num_sms) return fail_msg("b2048 NN playout: too many models in one launch");
    // `models` consecutive work cursors out of the rotating pool
    p.counter = take_counters(c, (unsigned)models);
    CK(cudaMemsetAsync(p.counter, 0, sizeof(unsigned long long) * models, st));
    const long total = p.n * models;
    if (precision == B2048_PREC_FP32) {
        int grid = models > 1 ? c->num_sms * 2 : grid_for(p.n, CN_BOARDS_PER_CTA, c->num_sms * 2);
        playout_nn_fp32_kernel<<<LC(c, grid), 512, CN_FP32_SMEM_BYTES, st>>>(p);
        CK(cudaGetLastError());
        return 0;
    }
    if (precision == B2048_PREC_FP16X2_TC) {
        if (models > 1) return fail_msg("b2048 NN playout: the precise tensor-core mode plays one model per launch");
        p.single_set = 1;
        int grid = grid_for(total, tc::TILE_BOARDS, c->num_sms);
        const int lat = c->tc_wide.load();
        // the latency regimes of the throughput mode (below): three- / two-ply speculation for small launches, adaptive
        // tail speculation for whole-game launches
        if (lat >= 3 && total <= (long)c->num_sms * SPEC3_LINES) {
            grid = grid_for(total, SPEC3_LINES, c->num_sms);
            if (p.main_mode)
                tc::tc_persistent_kernel<MainSpec3Work, false, false, true, true><<<LC(c, grid), tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(
                    p, (const uint8_t *)p.blob, nullptr, -1, 0);
            else
                tc::tc_persistent_kernel<PlaySpec3Work, false, false, true, true><<<LC(c, grid), tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(
                    p, (const uint8_t *)p.blob, nullptr, -1, 0);
        } else if (lat >= 2 && total <= (long)c->num_sms * SPEC_LINES) {
            grid = grid_for(total, SPEC_LINES, c->num_sms);
            if (p.main_mode)
                tc::tc_persistent_kernel<MainSpecWork, false, false, true, true><<<LC(c, grid), tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(
                    p, (const uint8_t *)p.blob, nullptr, -1, 0);
            else
                tc::tc_persistent_kernel<PlaySpecWork, false, false, true, true><<<LC(c, grid), tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(
                    p, (const uint8_t *)p.blob, nullptr, -1, 0);
        } else if (lat >= 2 && !p.main_mode && !p.root_mode)
            tc::tc_persistent_kernel<PlayAdaptWork, false, false, true, true><<<LC(c, grid), tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(
                p, (const uint8_t *)p.blob, nullptr, -1, 0);
        else if (p.main_mode)
            tc::tc_persistent_kernel<MainWork, false, false, true, true><<<LC(c, grid), tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(
                p, (const uint8_t *)p.blob, nullptr, -1, 0);
        else
            tc::tc_persistent_kernel<PlayWork, false, false, true, true><<<LC(c, grid), tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(
                p, (const uint8_t *)p.blob, nullptr, -1, 0);
        CK(cudaGetLastError());
        return 0;
    }
    if (precision == B2048_PREC_FP16_TC || precision == B2048_PREC_BF16_TC) {
        int grid;
        if (models > 1) {
            // every CTA is bound to one model; a model's games spread over its num_sms / models CTAs
            grid = c->num_sms;
            p.single_set = p.n <= (long)(grid / models) * tc::TILE_BOARDS;
        } else {
            // few items: spread them over the SMs, one tile-set per CTA (lower per-step latency)
            long limit = (long)c->num_sms * tc::TILE_BOARDS;
            if (p.main_mode) limit = limit * c->main_single_pct / 100;
            p.single_set = total <= limit;
            grid = p.single_set ? grid_for(total, tc::TILE_BOARDS, c->num_sms)
                                : grid_for(total, tc::TILE_BOARDS * tc::NSETS, c->num_sms);
        }
        const int lat = c->tc_wide.load();
        const bool wide = p.single_set && models == 1 && precision == B2048_PREC_FP16_TC && lat >= 1;
        if (wide && lat >= 3 && total <= (long)c->num_sms * SPEC3_LINES) {
            // deepest latency regime (config 3: 200 lines on 148 SMs): three-ply speculation, 2 lines x 16 slots per tile
            grid = grid_for(total, SPEC3_LINES, c->num_sms);
            if (p.main_mode)
                tc::tc_persistent_kernel<MainSpec3Work, false, false, true><<<LC(c, grid), tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(
                    p, (const uint8_t *)p.blob, nullptr, -1, 0);
            else
                tc::tc_persistent_kernel<PlaySpec3Work, false, false, true><<<LC(c, grid), tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(
                    p, (const uint8_t *)p.blob, nullptr, -1, 0);
            CK(cudaGetLastError());
            return 0;
        }
        if (wide && lat >= 2 && total <= (long)c->num_sms * SPEC_LINES) {
            // latency regime (config 3: 200 lines): two-ply speculation, 6 lines x 5 slots per tile, one tile per CTA
            grid = grid_for(total, SPEC_LINES, c->num_sms);
            if (p.main_mode)
                tc::tc_persistent_kernel<MainSpecWork, false, false, true><<<LC(c, grid), tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(
                    p, (const uint8_t *)p.blob, nullptr, -1, 0);
            else
                tc::tc_persistent_kernel<PlaySpecWork, false, false, true><<<LC(c, grid), tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(
                    p, (const uint8_t *)p.blob, nullptr, -1, 0);
            CK(cudaGetLastError());
            return 0;
        }
        if (p.main_mode) {
            if (precision != B2048_PREC_FP16_TC) return fail_msg("b2048_selfplay_nn: precision must be FP32, FP16_TC or FP16X2_TC");
            if (wide)
                tc::tc_persistent_kernel<MainWork, false, false, true><<<LC(c, grid), tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(
                    p, (const uint8_t *)p.blob, nullptr, -1, 0);
            else
                tc::tc_persistent_kernel<MainWork, false><<<LC(c, grid), tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(
                    p, (const uint8_t *)p.blob, nullptr, -1, 0);
        } else if (wide) {
            // latency modes >= 2, whole-game launches (eval_nn / population evaluation: long games of very different
            // lengths): the tail of the launch (queue empty, few lines left per tile) speculates (PlayAdaptWork).  Root
            // evaluations keep the plain kernel: their early stop makes the tail short, and the adaptive policy's longer
            // engine phase measured 1.5 % slower there.
            if (lat >= 2 && !p.root_mode)
                tc::tc_persistent_kernel<PlayAdaptWork, false, false, true><<<LC(c, grid), tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(
                    p, (const uint8_t *)p.blob, nullptr, -1, 0);
            else
                tc::tc_persistent_kernel<PlayWork, false, false, true><<<LC(c, grid), tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(
                    p, (const uint8_t *)p.blob, nullptr, -1, 0);
        } else if (precision == B2048_PREC_BF16_TC)
            tc::tc_persistent_kernel<PlayWork, false, true><<<LC(c, grid), tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(
                p, (const uint8_t *)p.blob, nullptr, -1, 0);
        else if (lat >= 2 && !p.root_mode)
            tc::tc_persistent_kernel<PlayAdaptWork, false><<<LC(c, grid), tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(
                p, (const uint8_t *)p.blob, nullptr, -1, 0);
        else
            tc::tc_persistent_kernel<PlayWork, false><<<LC(c, grid), tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(
                p, (const uint8_t *)p.blob, nullptr, -1, 0);
        CK(cudaGetLastError());
        return 0;
    }
    return fail_msg("b2048 NN playout: unknown precision");
}

extern "C" int b2048_playout_nn(b2048_ctx *c, int precision, const void *blob, const uint64_t *start, uint64_t seed,
                                uint64_t gid_base, uint32_t spawn0, int group_size, int32_t *group_min,
                                uint64_t *out_final, uint32_t *out_score, uint32_t *out_moves,
                                uint64_t *move_counter, int64_t n, void *stream) {
    DeviceGuard dg(c->device);
    if (n <= 0) return 0;
    if (!blob) return fail_msg("b2048_playout_nn: blob is NULL");
    NNPlayParams p;
    memset(&p, 0, sizeof p);
    p.blob = blob;
    p.n = (long)n;
    p.start = (const u64 *)start;
    p.seed = seed;
    p.gid_base = gid_base;
    p.spawn0 = spawn0;
    p.group_size = group_size;
    p.group_min = group_min;
    p.out_final = (u64 *)out_final;
    p.out_score = out_score;
    p.out_moves = out_moves;
    p.move_counter = (unsigned long long *)move_counter;
    return launch_nn(c, precision, p, (cudaStream_t)stream);
}

extern "C" int b2048_playout_nn_population(b2048_ctx *c, int precision, const void *blobs, int64_t blob_stride,
                                           int n_models, const uint64_t *start, uint64_t seed, uint64_t gid_base,
                                           uint64_t gid_model_stride, int group_size, int32_t *group_min,
                                           uint64_t *out_final, uint32_t *out_score, uint32_t *out_moves,
                                           uint64_t *move_counter, int64_t n_per_model, void *stream) {
    DeviceGuard dg(c->device);
    if (n_per_model <= 0 || n_models <= 0) return 0;
    if (!blobs) return fail_msg("b2048_playout_nn_population: blobs is NULL");
    if (blob_stride <= 0 || (blob_stride & 15)) return fail_msg("b2048_playout_nn_population: blob_stride must be a positive multiple of 16");
    if (group_min && group_size > 0 && n_per_model % group_size)
        return fail_msg("b2048_playout_nn_population: n_per_model must be a multiple of group_size");
    const int per_launch = c->num_sms < N_COUNTERS ? c->num_sms : N_COUNTERS;
    for (int m0 = 0; m0 < n_models; m0 += per_launch) {
        const int mc = (n_models - m0) < per_launch ? (n_models - m0) : per_launch;
        NNPlayParams p;
        memset(&p, 0, sizeof p);
        p.blob = (const uint8_t *)blobs + (size_t)m0 * (size_t)blob_stride;
        p.n = (long)n_per_model;
        p.n_models = mc;
        p.blob_stride = (long)blob_stride;
        p.gid_model_stride = gid_model_stride;
        p.start = (const u64 *)start;
        p.seed = seed;
        p.gid_base = gid_base + (uint64_t)m0 * gid_model_stride;
        p.group_size = group_size;
        const size_t o = (size_t)m0 * (size_t)n_per_model;
        p.group_min = (group_min && group_size > 0) ? group_min + o / group_size : nullptr;
        p.out_final = out_final ? (u64 *)out_final + o : nullptr;
        p.out_score = out_score ? out_score + o : nullptr;
        p.out_moves = out_moves ? out_moves + o : nullptr;
        p.move_counter = (unsigned long long *)move_counter;
        if (mc == 1) p.n_models = 0;
        if (int rc = launch_nn(c, precision, p, (cudaStream_t)stream)) return rc;
    }
    return 0;
}

extern "C" int b2048_mcts_nn(b2048_ctx *c, int precision, const void *blob, const uint64_t *origins, int64_t G,
                             int number, int line_begin, int line_end, uint64_t seed, uint64_t eval_id_base,
                             uint8_t *out_legal, uint32_t *out_rscore, uint32_t *out_lmoves, int32_t *out_minmov,
                             uint64_t *move_counter, void *stream) {
    DeviceGuard dg(c->device);
    if (G <= 0) return 0;
    if (number <= 0 || line_begin < 0 || line_end > number || line_begin > line_end)
        return fail_msg("b2048_mcts_nn: bad line range");
    if (!blob || !out_legal || !out_rscore || !out_minmov) return fail_msg("b2048_mcts_nn: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    root_moves_kernel<<<LC(c, (unsigned)((G * 4 + 255) / 256)), 256, 0, st>>>(tables_of(c), (const u64 *)origins, (long)G,
                                                                       out_legal, out_rscore);
    CK(cudaGetLastError());
    if (line_begin == line_end) return 0;
    NNPlayParams p;
    memset(&p, 0, sizeof p);
    p.blob = blob;
    p.root_mode = 1;
    p.origins = (const u64 *)origins;
    p.rlegal = out_legal;
    p.number = number;
    p.line_begin = line_begin;
    p.line_count = line_end - line_begin;
    p.n = (long)G * 4 * p.line_count;
    p.seed = seed;
    p.gid_base = eval_id_base * 4ULL;
    p.spawn0 = 1;
    p.group_min = out_minmov;
    p.out_moves = out_lmoves;
    p.move_counter = (unsigned long long *)move_counter;
    return launch_nn(c, precision, p, st);
}

// ------------------------------------------------------------------------------------------
// Device-resident NN main games (selfplay.py:62-115 driven by mcts_nn.py:4-43)
// ------------------------------------------------------------------------------------------
static inline uint32_t mg_ring_cap(int64_t G, int number) {
    uint64_t need = 2ULL * (uint64_t)G * 4ULL * (uint64_t)number;
    uint32_t cap = 1024;
    while ((uint64_t)cap < need) cap <<= 1;
    return cap;
}
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" int64_t b2048_selfplay_nn_workspace_bytes(int64_t G, int number) {
    if (G <= 0 || number <= 0 || (uint64_t)G * 4ULL * (uint64_t)number >= (1ULL << 30)) return -1;
    return (int64_t)(align256(sizeof(MainQueue)) + align256((size_t)G * sizeof(MainGame)) + align256((size_t)G * 4 * sizeof(int)) +
                     (size_t)mg_ring_cap(G, number) * sizeof(uint32_t));
}

extern "C" int b2048_selfplay_nn(b2048_ctx *c, int precision, const void *blob, const uint64_t *starts, int64_t G,
                                 uint64_t game_id_base, uint64_t games_total, int number, uint64_t seed, int max_steps,
                                 uint64_t *out_boards, float *out_results, uint64_t *out_final, int64_t *out_score,
                                 int32_t *out_steps, int32_t *out_status, uint64_t *move_counter, void *workspace,
                                 void *stream) {
    DeviceGuard dg(c->device);
    if (G <= 0) return 0;
    if (!blob || !workspace) return fail_msg("b2048_selfplay_nn: blob and workspace are required");
    if (number <= 0 || max_steps < 0) return fail_msg("b2048_selfplay_nn: need number > 0, max_steps >= 0");
    if (!out_final || !out_score || !out_steps || !out_status) return fail_msg("b2048_selfplay_nn: NULL output");
    if (games_total < game_id_base + (uint64_t)G) return fail_msg("b2048_selfplay_nn: games_total < game_id_base + G");
    const int64_t wbytes = b2048_selfplay_nn_workspace_bytes(G, number);
    if (wbytes < 0) return fail_msg("b2048_selfplay_nn: G * 4 * number too large");
    if ((uintptr_t)workspace & 255) return fail_msg("b2048_selfplay_nn: workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    NNPlayParams p;
    memset(&p, 0, sizeof p);
    uint8_t *w = (uint8_t *)workspace;
    p.mgq = (MainQueue *)w;
    w += align256(sizeof(MainQueue));
    p.mg = (MainGame *)w;
    w += align256((size_t)G * sizeof(MainGame));
    p.group_min = (int *)w;
    w += align256((size_t)G * 4 * sizeof(int));
    p.mg_items = (uint32_t *)w;
    p.mg_cap = mg_ring_cap(G, number);
    CK(cudaMemsetAsync(workspace, 0, (size_t)wbytes, st));
    p.main_mode = 1;
    p.blob = blob;
    p.n = (long)G * 4 * number;  // lines in flight at most (launch sizing only)
    p.number = number;
    p.seed = seed;
    p.spawn0 = 1;
    p.mg_g0 = game_id_base;
    p.mg_total = games_total;
    p.mg_max_steps = max_steps;
    p.mg_out_boards = (u64 *)out_boards;
    p.mg_out_results = out_results;
    p.mg_out_final = (u64 *)out_final;
    p.mg_out_score = (long long *)out_score;
    p.mg_out_steps = out_steps;
    p.mg_out_status = out_status;
    p.move_counter = (unsigned long long *)move_counter;
    p.tb = tables_of(c);
    p.err_flag = c->err_flag;
    mg_init_kernel<<<LC(c, (unsigned)((G + 127) / 128)), 128, 0, st>>>(p, (const u64 *)starts, (long)G);
    CK(cudaGetLastError());
    return launch_nn(c, precision, p, st);
}

#ifdef B2048_TRACE
extern "C" int b2048_debug_trace(long long *out_host, int n) {
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyFromSymbol(out_host, tc::g_trace, sizeof(long long) * (n < 512 ? n : 512)));
    return 0;
}
#endif
