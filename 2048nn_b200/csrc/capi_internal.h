// capi_internal.h -- shared by the translation units of lib2048nn_b200.so: the context and error reporting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

// Work cursors: every launch of a persistent kernel takes the next slot(s) of a rotating pool (atomic cursor, so
// host threads and streams may share a context).  A slot is reused after N_COUNTERS further launches; the CUDA
// launch queue is far shallower than that, so a slot is never shared by two launches in flight.
#define N_COUNTERS 4096

struct b2048_ctx {
    int device;
    int num_sms;
    uint16_t *fin_fwd, *fin_rev;
    uint32_t *score;
    uint8_t *moved_fwd, *moved_rev, *legal;
    uint2 *rowstat;
    uint32_t *pair;               // u32[K3_ROWS] XOR deltas of both slides of a row (engine.cuh)
    std::atomic<int> k2_variant;  // 0 = attempt cascade, 1 = legality-first uniform slide, 2 = pair table (default)
    int main_single_pct;          // main-game launches up to this % of (32 x #SMs) lines in flight run one tile-set per CTA
    std::atomic<int> tc_wide;     // 1 = small tensor-core launches use the 16-warp one-set (latency) kernel variant
    unsigned long long *counters;  // N_COUNTERS work cursors, one per in-flight launch
    int *err_flag;                 // [0] spawn on a full board (ValueError in the reference), [1] table self-check
    std::atomic<unsigned> next_counter;
    std::atomic<unsigned long long> launches;  // kernels launched through this context (b2048_launch_count)
};

// record the message returned by b2048_last_error(); return the status code of the failing entry point
int b2048_fail(const char *what, cudaError_t e);
int b2048_fail_msg(const char *what);

#define CK(call)                                       \
    do {                                               \
        cudaError_t e_ = (call);                       \
        if (e_ != cudaSuccess) return b2048_fail(#call, e_); \
    } while (0)

// Every entry point runs with the context's device current and restores the caller's device on return, so a
// context of GPU 1 can be used from a thread whose current device is GPU 0 (b2048_create used to leave
// the device switched and no other call set it).
struct DeviceGuard {
    int prev;
    explicit DeviceGuard(int device) : prev(-1) {
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != device) {
            prev = cur;
            cudaSetDevice(device);
        }
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};

// `count` consecutive work cursors out of the rotating pool (count <= N_COUNTERS)
static inline unsigned long long *take_counters(b2048_ctx *c, unsigned count) {
    for (;;) {
        unsigned cur = c->next_counter.load(std::memory_order_relaxed);
        unsigned slot = cur % N_COUNTERS;
        unsigned skip = (slot + count > N_COUNTERS) ? (N_COUNTERS - slot) : 0;  // do not wrap inside one launch
        if (c->next_counter.compare_exchange_weak(cur, cur + skip + count, std::memory_order_relaxed))
            return c->counters + ((slot + skip) % N_COUNTERS);
    }
}

// launch accounting: every kernel launch of the library is written kernel<<<LC(c, grid), ...>>>
template <typename G>
static inline G LC(b2048_ctx *c, G grid) {
    c->launches.fetch_add(1, std::memory_order_relaxed);
    return grid;
}
