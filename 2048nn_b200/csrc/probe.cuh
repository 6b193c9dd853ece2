// Pipe-rate probes: the MEASURED denominators of the K2 roofline (bench.py `roofline.executed`).
// SURVEY 8d bounds the fixed-order playout kernel by SM integer issue; ncu shows the ALU pipe (LOP3 / SHF / PRMT / SEL /
// ISETP: 16 lanes per clock and SM sub-partition) as the pipe that fills first.  These kernels measure what one B200
// sustains on exactly those instruction classes, under its own power management, so the fractions bench.py reports are
// against a measured rate next to the formula `SMs x 4 x 32 x max clock`.
//   kind 0: LOP3 only                  -> the ALU pipe alone
//   kind 1: LOP3 and IMAD alternating  -> ALU pipe + FMA pipe together: one warp-instruction per clock and sub-partition,
//                                         the issue limit the survey's 250-instruction model is written against
#pragma once
#include <stdint.h>

namespace b2048 {

constexpr int PROBE_UNROLL = 64;  // probe instructions per loop trip (the loop's own IADD / ISETP / BRA are not counted)

template <int KIND>
__global__ void __launch_bounds__(1024, 1) pipe_probe_kernel(uint32_t iters, uint32_t k1, uint32_t k2, uint32_t *sink) {
    uint32_t a[8];  // eight independent chains per thread, 32 warps per SM: latency is covered many times over
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0x9E3779B9u + i * 0x85EBCA6Bu + blockIdx.x;
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < PROBE_UNROLL / 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (KIND == 1 && (i & 1))
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(k1), "r"(k2));
                else
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(k1), "r"(k2));
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= a[i];
    if (r == 0x12345678u) sink[0] = r;  // keeps the chains alive; practically never taken
}

}  // namespace b2048
