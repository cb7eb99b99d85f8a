// Measurement probes (diagnostics, no reference counterpart): micro-kernels that measure the on-chip peaks the DRR
// kernel is bound by, so that bench.py can report it against the resource that actually limits it (the
// DRR's compulsory HBM bytes are tiny; its binding resources are the L1 gather path and instruction issue).
#include "common.cuh"

namespace lr {

// Every block owns `floats_per_block` consecutive floats (a few KB: L1-resident after the first pass) and reads them
// `iters` times with fully coalesced warp-wide 32-bit loads (one 128 B wavefront per LDG, 8 independent loads in
// flight per thread): the ceiling of a kernel that gathers with scalar loads.  bytes = gridDim.x * 256 * iters * 32.
__global__ void __launch_bounds__(256) probe_l1_gather_kernel(const float *__restrict__ buf, int floats_per_block, int iters,
                                                              int zero, float *__restrict__ sink) {
    // thread t reads floats t, t+256, ..., t+7*256 of the block's slice (2048 floats = 8 KB) over and over.  `zero` is a
    // run-time 0 folded into the address so that the loads cannot be hoisted out of the loop; the eight addresses differ
    // by immediates, so the loop body is 8 LDG + 8 FADD + 2 integer instructions.
    const float *p = buf + (size_t)blockIdx.x * floats_per_block + threadIdx.x;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
    for (int it = 0; it < iters; ++it) {
        const float *q = p + (it & zero);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] += __ldca(q + u * 256);
    }
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) s += acc[u];
    if (s == 123.456f) sink[0] = s;      // keeps the loads alive
}

// Issue-rate probe: `iters` x 8 dependent-free FFMA chains per thread (8 independent accumulators): measures the
// sustained warp-instruction issue rate of the chip with 32 resident warps per SM.
__global__ void __launch_bounds__(256) probe_issue_kernel(int iters, float a, float *__restrict__ sink) {
    float acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = (float)(threadIdx.x + u);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = __fmaf_rn(acc[u], a, 1.0f);
    }
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) s += acc[u];
    if (s == 123.456f) sink[0] = s;
}

// The same with packed fp32x2 FMAs (FFMA2): tells whether a packed instruction costs one or two FP32-pipe cycles.
__global__ void __launch_bounds__(256) probe_issue_packed_kernel(int iters, float a, float *__restrict__ sink) {
    f32x2 acc[8];
    const f32x2 a2 = splat2(a), one = splat2(1.0f);
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = pack2((float)(threadIdx.x + u), (float)(threadIdx.x - u));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = fma2(acc[u], a2, one);
    }
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        float lo, hi;
        unpack2(acc[u], lo, hi);
        s += lo + hi;
    }
    if (s == 123.456f) sink[0] = s;
}

}  // namespace lr

using namespace lr;

extern "C" int lr_probe_l1_gather(const float *buf, int64_t buf_floats, int blocks, int floats_per_block, int iters,
                                  float *sink, lr_stream_t stream) {
    LR_REQUIRE(buf && sink && blocks > 0 && iters > 0, "probe_l1_gather: bad argument");
    LR_REQUIRE(floats_per_block >= 2048 && (floats_per_block & (floats_per_block - 1)) == 0,
               "probe_l1_gather: floats_per_block must be a power of two >= 2048");
    LR_REQUIRE((int64_t)blocks * floats_per_block <= buf_floats, "probe_l1_gather: buffer too small");
    probe_l1_gather_kernel<<<blocks, 256, 0, as_stream(stream)>>>(buf, floats_per_block, iters, 0, sink);
    return check_launch("probe_l1_gather_kernel");
}

extern "C" int lr_probe_issue(int blocks, int iters, float *sink, lr_stream_t stream) {
    LR_REQUIRE(sink && blocks > 0 && iters > 0, "probe_issue: bad argument");
    probe_issue_kernel<<<blocks, 256, 0, as_stream(stream)>>>(iters, 0.999f, sink);
    return check_launch("probe_issue_kernel");
}

extern "C" int lr_probe_issue_packed(int blocks, int iters, float *sink, lr_stream_t stream) {
    LR_REQUIRE(sink && blocks > 0 && iters > 0, "probe_issue_packed: bad argument");
    probe_issue_packed_kernel<<<blocks, 256, 0, as_stream(stream)>>>(iters, 0.999f, sink);
    return check_launch("probe_issue_packed_kernel");
}
