// capi_internal.h -- shared by the translation units of lib2048nn_b200.so: the context and error reporting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define N_COUNTERS 256

struct b2048_ctx {
    int device;
    int num_sms;
    uint16_t *fin_fwd, *fin_rev;
    uint32_t *score;
    uint8_t *moved_fwd, *moved_rev, *legal;
    uint2 *rowstat;
    int k2_variant;  // 0 = attempt cascade, 1 = legality-first uniform slide
    unsigned long long *counters;  // N_COUNTERS work cursors, one per in-flight launch
    int *err_flag;
    unsigned next_counter;
    // root-search scratch (grown on demand outside of stream capture; see b2048_mcts_fixed)
    uint64_t *rboard;
    int64_t rboard_cap;
};

// record the message returned by b2048_last_error(); return the status code of the failing entry point
int b2048_fail(const char *what, cudaError_t e);
int b2048_fail_msg(const char *what);

#define CK(call)                                       \
    do {                                               \
        cudaError_t e_ = (call);                       \
        if (e_ != cudaSuccess) return b2048_fail(#call, e_); \
    } while (0)
