// Device helpers shared by the polarizability kernels (sm_100a).
#pragma once

#include "rn_common.cuh"

namespace rn {

struct Alpha0 {
    double v[9];
};

// wrap(pos - ref) into (-0.5, 0.5]: apply_pbc_displacement(calc_displacement(...)) of the
// reference (structure/utils.py:46,132-135); ties resolve to +0.5 exactly as `d % 1 > 0.5`.
__device__ __forceinline__ double wrap_disp(double pos, double ref) {
    const double d = pos - ref;
    return d - ceil(d - 0.5);
}

// The same value with ONE instruction on the FP64 pipe in the common case: |d| < 0.5 is decided on the
// exponent bits (integer ALU) and d is then already the minimum image (d - ceil(d - 0.5) = d - (-0) = d;
// only the sign of a zero may differ).  Kept as an A/B variant of the dense producers
// (rn_debug_set_dense_config): measured 3-4 % SLOWER there than the branch-free formula above — the
// branch serialises the eight rows a producer thread wraps per chunk, and what the MMA warps wait for is
// the producers' latency per chunk, not their FP64 issue slots.
__device__ __forceinline__ double wrap_disp_fast(double pos, double ref) {
    const double d = pos - ref;
    const unsigned mag = (unsigned)__double2hiint(d) & 0x7fffffffu;
    if (mag < 0x3fe00000u) return d;
    return d - ceil(d - 0.5);
}

// FP64 tensor-pipe MMA (SASS: DMMA.8x8x4).  A 8x4 row-major: lane holds A[lane>>2][lane&3];
// B 4x8 col-major: lane holds B[lane&3][lane>>2]; C 8x8: lane holds C[lane>>2][2*(lane&3)+{0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_8(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// mbarrier + TMA bulk-copy helpers (cp.async.bulk global -> shared with complete_tx)
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

}  // namespace rn
