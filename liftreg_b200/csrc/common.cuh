// Shared device helpers for the LiftReg resampling kernels (sm_100a).
//
// Numerical contract: the reference evaluates every coordinate and weight as a chain of separately
// rounded fp32 torch ops (SURVEY.md §7 "coordinate parity").  To keep ray and voxel indexing bit-exact
// the kernels replay that chain with explicitly rounded intrinsics (__fadd_rn/__fmul_rn are never
// contracted into FMAs by nvcc) and use a fused multiply-add only where torch's CPU kernels do.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lr {

__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// Correctly rounded x / c for a loop-invariant divisor c, given rc = RN(1/c) (Markstein: q0 = x*rc,
// r = x - c*q0 exactly, q = q0 + r*rc).  Bit-identical to __fdiv_rn for normal operands (validated over
// 6M values per divisor, tests/test_host_logic.py::test_markstein_division); 3 instructions instead of ~10.
struct ConstDiv {
    float c, rc;
};
__host__ __device__ __forceinline__ ConstDiv make_const_div(float c) {
    ConstDiv k;
    k.c = c;
    k.rc = 1.0f / c;
    return k;
}
__device__ __forceinline__ float div_const(float x, ConstDiv k) {
    float q0 = __fmul_rn(x, k.rc);
    float r = __fmaf_rn(-k.c, q0, x);
    return __fmaf_rn(r, k.rc, q0);
}

// floor(x) as float and int for |x| < 2^22, on the FP32/ALU pipes only.  FRND/F2I/I2F run on the XU pipe at 16
// lanes/clk/SM and saturated it in the first version of these kernels (ncu: xu 95 %).  Adding 1.5*2^23 rounds x to
// an integer that can be read straight out of the mantissa; one compare turns round-to-nearest into floor.
// Callers clamp x to a few voxels around the volume first (samples further out have no in-bounds tap anyway).
__device__ __forceinline__ void floor_fi(float x, float &f, int &i) {
    const float M = 12582912.0f;                  // 1.5 * 2^23 = 0x4B400000
    const float t = __fadd_rn(x, M);
    const float r = __fsub_rn(t, M);              // exact: nearest integer to x
    const int ri = __float_as_int(t) - 0x4B400000;
    const bool up = r > x;                        // rounded up -> step back
    f = up ? __fsub_rn(r, 1.0f) : r;
    i = up ? ri - 1 : ri;
}
// nearbyint(x) (half to even) for |x| < 2^22, same trick without the compare
__device__ __forceinline__ int rint_i(float x) {
    return __float_as_int(__fadd_rn(x, 12582912.0f)) - 0x4B400000;
}
__device__ __forceinline__ float clamp_index(float x, float hi) {   // keep |x| < 2^22; NaN -> -2
    return fminf(fmaxf(x, -2.0f), hi);
}

// Streaming (read-once / write-once) accesses: keep them out of L1 so the gather working set stays resident.
__device__ __forceinline__ float ld_stream(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_stream4(const float4 *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(float *p, float v) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void st_stream4(float4 *p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

// fire-and-forget fp32 reduction to global memory (RED.E.ADD.F32)
__device__ __forceinline__ void red_add(float *p, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

}  // namespace lr

// ---- host-side plumbing shared by the .cu files -------------------------------------------------
#include "../../include/liftreg_b200.h"
namespace lr {
void set_error(const char *fmt, ...);
int check_launch(const char *what);  // cudaGetLastError -> lr_status; bumps the per-thread launch counter
inline cudaStream_t as_stream(lr_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
}  // namespace lr

#define LR_REQUIRE(cond, ...)                \
    do {                                     \
        if (!(cond)) {                       \
            lr::set_error(__VA_ARGS__);      \
            return LR_ERR_BAD_ARGUMENT;      \
        }                                    \
    } while (0)
