// Similarity loss on the warp output: the reference's NCCLoss, sm_100a.
//
// Replaces reference src/liftreg/layers/losses.py:14-29 (NCCLoss.forward), used as the training similarity
// (src/liftreg/losses/SubspaceLoss.py:12,27) and as the validation score (src/liftreg/networks/RegistrationNet.py:210-212):
//     a = x - mean(x) + 1e-10,  b = y - mean(y) + 1e-10                       (per batch item, over all voxels)
//     ncc = mean(a*b) / sqrt(mean(a^2) * mean(b^2)),   loss = 1 - mean_over_batch(ncc)
// The reference runs ~10 elementwise / reduction kernels over the two volumes (each re-reading 16 MB per item at 160^3);
// here the forward is two passes over x and y (sums, then centred second moments: the centring needs the means first)
// and the backward is one pass.  Per-thread partial sums are fp32 over a handful of elements, everything above that
// is accumulated in fp64 (warp shuffles, shared memory, one fp64 atomic per block), so the result does not depend on the launch shape
// beyond fp64 round-off.
#include "common.cuh"

namespace lr {

constexpr int NCC_THREADS = 256;
constexpr int NCC_SUMS = 7;          // per item: sum x, sum y, sum a*b, sum a*a, sum b*b, sum a, sum b

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block reduction of NV per-thread values, then ONE fp64 atomic per value and block (an atomic per warp made the few
// accumulator addresses the bottleneck: 96 us instead of 17 us per 160^3 pair).
template <int NV>
__device__ __forceinline__ void block_accumulate(const float (&v)[NV], double *__restrict__ dst) {
    __shared__ double part[NCC_THREADS / 32][NV];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const double w = warp_sum((double)v[k]);
        if (lane == 0) part[warp][k] = w;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < NCC_THREADS / 32; ++w) t += part[w][threadIdx.x];
        atomicAdd(dst + threadIdx.x, t);
    }
}

// pass 1: sums[b][0..1] += sum x, sum y
__global__ void __launch_bounds__(NCC_THREADS) ncc_sum_kernel(const float *__restrict__ x, const float *__restrict__ y, int64_t N,
                                                              double *__restrict__ sums) {
    const int b = blockIdx.y;
    const float *xb = x + (int64_t)b * N, *yb = y + (int64_t)b * N;
    float sx = 0.0f, sy = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * NCC_THREADS + threadIdx.x; i < N; i += (int64_t)gridDim.x * NCC_THREADS) {
        sx += ld_stream(xb + i);
        sy += ld_stream(yb + i);
    }
    const float v[2] = {sx, sy};
    block_accumulate<2>(v, sums + b * NCC_SUMS);
}

// the centred values exactly as the reference forms them in fp32 (losses.py:21-22): (x - mean) + 1e-10
__device__ __forceinline__ float centred(float v, float mean) { return add_rn(sub_rn(v, mean), 1e-10f); }

// pass 2: sums[b][2..6] += sum a*b, a*a, b*b, a, b
__global__ void __launch_bounds__(NCC_THREADS) ncc_moment_kernel(const float *__restrict__ x, const float *__restrict__ y, int64_t N,
                                                                 double *__restrict__ sums) {
    const int b = blockIdx.y;
    const float *xb = x + (int64_t)b * N, *yb = y + (int64_t)b * N;
    const float mx = (float)(sums[b * NCC_SUMS + 0] / (double)N), my = (float)(sums[b * NCC_SUMS + 1] / (double)N);
    float sab = 0.0f, saa = 0.0f, sbb = 0.0f, sa = 0.0f, sb = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * NCC_THREADS + threadIdx.x; i < N; i += (int64_t)gridDim.x * NCC_THREADS) {
        const float a = centred(__ldg(xb + i), mx), c = centred(__ldg(yb + i), my);
        sab = fmaf(a, c, sab); saa = fmaf(a, a, saa); sbb = fmaf(c, c, sbb);
        sa += a; sb += c;
    }
    const float v[5] = {sab, saa, sbb, sa, sb};
    block_accumulate<5>(v, sums + b * NCC_SUMS + 2);
}

// backward wrt x:  ncc = Sab / sqrt(Saa Sbb);  d ncc / d a_i = c1 b_i - c2 a_i  with c1 = 1/sqrt(Saa Sbb),
// c2 = Sab / (Saa sqrt(Saa Sbb));  a_i = x_i - mean(x) + eps  =>  d ncc / d x_j = g_j - mean(g).
// grad_x = scale * (c1 b - c2 a - c3),  c3 = c1 Sb/N - c2 Sa/N,  scale = -grad_loss / B  (loss = 1 - mean_b ncc).
__global__ void __launch_bounds__(NCC_THREADS) ncc_backward_kernel(const float *__restrict__ x, const float *__restrict__ y, int64_t N,
                                                                   const double *__restrict__ sums, const float *__restrict__ grad_loss,
                                                                   int B, float *__restrict__ gx) {
    const int b = blockIdx.y;
    const double *s = sums + b * NCC_SUMS;
    const float mx = (float)(s[0] / (double)N), my = (float)(s[1] / (double)N);
    const double root = sqrt(s[3] * s[4]);
    const double c1 = 1.0 / root, c2 = s[2] / (s[3] * root);
    const double c3 = c1 * s[6] / (double)N - c2 * s[5] / (double)N;
    const double scale = -(double)grad_loss[0] / (double)B;
    const float f1 = (float)(scale * c1), f2 = (float)(scale * c2), f3 = (float)(scale * c3);
    const float *xb = x + (int64_t)b * N, *yb = y + (int64_t)b * N;
    float *gb = gx + (int64_t)b * N;
    for (int64_t i = (int64_t)blockIdx.x * NCC_THREADS + threadIdx.x; i < N; i += (int64_t)gridDim.x * NCC_THREADS) {
        const float a = centred(ld_stream(xb + i), mx), c = centred(ld_stream(yb + i), my);
        st_stream(gb + i, fmaf(f1, c, fmaf(-f2, a, -f3)));
    }
}

static unsigned ncc_blocks(int64_t N, int B) {
    int64_t want = (N + NCC_THREADS * 8 - 1) / (NCC_THREADS * 8);          // >= 8 elements per thread
    const int64_t cap = (int64_t)sm_count() * 8 / (B < 8 ? B : 8) + 1;      // ~8 resident blocks per SM over the whole grid
    if (want > cap) want = cap;
    return (unsigned)(want < 1 ? 1 : want);
}

}  // namespace lr

using namespace lr;

extern "C" int lr_ncc_sums(const float *x, const float *y, int B, int64_t N, double *sums, lr_stream_t stream) {
    LR_REQUIRE(x && y && sums, "ncc_sums: null pointer");
    LR_REQUIRE(B > 0 && B <= 65535 && N > 0, "ncc_sums: bad size (B=%d N=%lld)", B, (long long)N);
    cudaStream_t st = as_stream(stream);
    cudaError_t ce = cudaMemsetAsync(sums, 0, sizeof(double) * NCC_SUMS * (size_t)B, st);
    if (ce != cudaSuccess) { set_error("ncc_sums: memset failed: %s", cudaGetErrorString(ce)); return LR_ERR_CUDA; }
    const dim3 grid(ncc_blocks(N, B), (unsigned)B);
    ncc_sum_kernel<<<grid, NCC_THREADS, 0, st>>>(x, y, N, sums);
    if (int e = check_launch("ncc_sum_kernel")) return e;
    ncc_moment_kernel<<<grid, NCC_THREADS, 0, st>>>(x, y, N, sums);
    return check_launch("ncc_moment_kernel");
}

extern "C" int lr_ncc_backward(const float *x, const float *y, int B, int64_t N, const double *sums, const float *grad_loss,
                               float *grad_x, lr_stream_t stream) {
    LR_REQUIRE(x && y && sums && grad_loss && grad_x, "ncc_backward: null pointer");
    LR_REQUIRE(B > 0 && B <= 65535 && N > 0, "ncc_backward: bad size (B=%d N=%lld)", B, (long long)N);
    const dim3 grid(ncc_blocks(N, B), (unsigned)B);
    ncc_backward_kernel<<<grid, NCC_THREADS, 0, as_stream(stream)>>>(x, y, N, sums, grad_loss, B, grad_x);
    return check_launch("ncc_backward_kernel");
}
