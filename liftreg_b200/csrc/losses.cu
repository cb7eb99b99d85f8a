// Similarity loss on the warp output: the reference's NCCLoss, sm_100a.
//
// Replaces reference src/liftreg/layers/losses.py:14-29 (NCCLoss.forward), used as the training similarity
// (src/liftreg/losses/SubspaceLoss.py:12,27) and as the validation score (src/liftreg/networks/RegistrationNet.py:210-212):
//     a = x - mean(x) + 1e-10,  b = y - mean(y) + 1e-10                       (per batch item, over all voxels)
//     ncc = mean(a*b) / sqrt(mean(a^2) * mean(b^2)),   loss = 1 - mean_over_batch(ncc)
// The reference runs ~10 elementwise / reduction kernels over the two volumes (each re-reading 16 MB per item at 160^3);
// here the forward is ONE pass over x and y (raw moments about the item's first voxel, rewritten into the centred ones by a
// B-thread kernel; the two-pass op-for-op form -- sums, then fp32-centred second moments -- stays behind
// LIFTREG_B200_NCC_TWO_PASS=1) and the backward is one pass.  Per-thread partial sums are fp32 over a handful of elements, everything above that
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

// Single-pass forward.  The centring needs the means, which is why the restatement above reads x and y twice; but with
// u = x - k, v = y - l for ANY constants k, l the centred moments follow from the raw moments of u and v:
//     a = x - mean + eps = u - dx  (dx = mean - k - eps)   =>   sum a*a = sum u*u - 2 dx sum u + N dx^2, etc.
// and the cancellation in these differences is mild when k, l lie within a few standard deviations of the means.  k and l are
// the item's first voxels (every thread reads them); per-thread fp32 partials over a few elements, everything above in fp64
// as before; a B-thread kernel then rewrites the five raw sums into the seven-entry layout the backward and the host use.
// What is not reproduced is the reference's fp32 rounding of each (x - mean) + 1e-10 (the 1e-10 is below half an ulp of all
// but the elements within 1e-3 of the mean): the loss differs from the two-pass kernels by ~1e-8.
__global__ void __launch_bounds__(NCC_THREADS) ncc_shifted_kernel(const float *__restrict__ x, const float *__restrict__ y, int64_t N,
                                                                  double *__restrict__ sums) {
    const int b = blockIdx.y;
    const float *xb = x + (int64_t)b * N, *yb = y + (int64_t)b * N;
    const float k = __ldg(xb), l = __ldg(yb);
    float su = 0.0f, sv = 0.0f, suv = 0.0f, suu = 0.0f, svv = 0.0f;
    const int64_t t = (int64_t)blockIdx.x * NCC_THREADS + threadIdx.x, stride = (int64_t)gridDim.x * NCC_THREADS;
    // 16-byte loads where the item's rows allow it (four times the bytes in flight per thread: the scalar loop was bound by
    // memory latency at full occupancy, 2.6 TB/s), scalar tail
    const int64_t n4 = (((uintptr_t)xb | (uintptr_t)yb) % 16 == 0) ? N / 4 : 0;
    for (int64_t i = t; i < n4; i += stride) {
        const float4 a = ld_stream4(reinterpret_cast<const float4 *>(xb) + i), c = ld_stream4(reinterpret_cast<const float4 *>(yb) + i);
        const float u0 = a.x - k, u1 = a.y - k, u2 = a.z - k, u3 = a.w - k, v0 = c.x - l, v1 = c.y - l, v2 = c.z - l, v3 = c.w - l;
        su += (u0 + u1) + (u2 + u3); sv += (v0 + v1) + (v2 + v3);
        suv = fmaf(u0, v0, fmaf(u1, v1, fmaf(u2, v2, fmaf(u3, v3, suv))));
        suu = fmaf(u0, u0, fmaf(u1, u1, fmaf(u2, u2, fmaf(u3, u3, suu))));
        svv = fmaf(v0, v0, fmaf(v1, v1, fmaf(v2, v2, fmaf(v3, v3, svv))));
    }
    for (int64_t i = 4 * n4 + t; i < N; i += stride) {
        const float u = ld_stream(xb + i) - k, v = ld_stream(yb + i) - l;
        su += u; sv += v;
        suv = fmaf(u, v, suv); suu = fmaf(u, u, suu); svv = fmaf(v, v, svv);
    }
    const float vals[5] = {su, sv, suv, suu, svv};
    block_accumulate<5>(vals, sums + b * NCC_SUMS);
}

__global__ void ncc_finalize_kernel(const float *__restrict__ x, const float *__restrict__ y, int64_t N, int B, double *__restrict__ sums) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double *s = sums + b * NCC_SUMS;
    const double n = (double)N, k = (double)x[(int64_t)b * N], l = (double)y[(int64_t)b * N];
    const double Su = s[0], Sv = s[1], Suv = s[2], Suu = s[3], Svv = s[4];
    const double Sx = Su + n * k, Sy = Sv + n * l;
    const double dx = (double)(float)(Sx / n) - k, dy = (double)(float)(Sy / n) - l;      // the means as the fp32 values the backward uses
    s[0] = Sx; s[1] = Sy;
    s[2] = Suv - dy * Su - dx * Sv + n * dx * dy;
    s[3] = Suu - 2.0 * dx * Su + n * dx * dx;
    s[4] = Svv - 2.0 * dy * Sv + n * dy * dy;
    s[5] = Su - n * dx; s[6] = Sv - n * dy;
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
    const int64_t t = (int64_t)blockIdx.x * NCC_THREADS + threadIdx.x, stride = (int64_t)gridDim.x * NCC_THREADS;
    const int64_t n4 = (((uintptr_t)xb | (uintptr_t)yb | (uintptr_t)gb) % 16 == 0) ? N / 4 : 0;
    for (int64_t i = t; i < n4; i += stride) {
        const float4 xv = ld_stream4(reinterpret_cast<const float4 *>(xb) + i), yv = ld_stream4(reinterpret_cast<const float4 *>(yb) + i);
        float4 o;
        o.x = fmaf(f1, centred(yv.x, my), fmaf(-f2, centred(xv.x, mx), -f3));
        o.y = fmaf(f1, centred(yv.y, my), fmaf(-f2, centred(xv.y, mx), -f3));
        o.z = fmaf(f1, centred(yv.z, my), fmaf(-f2, centred(xv.z, mx), -f3));
        o.w = fmaf(f1, centred(yv.w, my), fmaf(-f2, centred(xv.w, mx), -f3));
        st_stream4(reinterpret_cast<float4 *>(gb) + i, o);
    }
    for (int64_t i = 4 * n4 + t; i < N; i += stride) {
        const float a = centred(ld_stream(xb + i), mx), c = centred(ld_stream(yb + i), my);
        st_stream(gb + i, fmaf(f1, c, fmaf(-f2, a, -f3)));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Displacement regulariser of the subspace loss, reference src/liftreg/losses/SubspaceLoss.py:51-67:
//     fd = mermaid.finite_differences.FD_torch(spacing*2),  spacing = 1/(shape-1)
//     reg = mean_{b,voxel} sum_{c<3} dXc(disp_c)^2 + dYc(disp_c)^2 + dZc(disp_c)^2
// dXc = (xp - xm) * (0.5/spacing[0]) is mermaid's central difference along the FIRST spatial axis (Y: second, Z: last).
// mermaid is a third-party dependency that is absent from /root/reference (requirements.txt:61 pins `mermaid==0.3.2`);
// its central difference is restated here.  What the call site fixes: interior voxels use (I[i+1] - I[i-1]) * scale.
// What only mermaid's source fixes is the boundary rule, so both of its published modes are offered:
//     LR_FD_LINEAR (0, FD_torch's default mode='linear'): the missing neighbour is extrapolated linearly,
//                  xp[n-1] = 2 I[n-1] - I[n-2], xm[0] = 2 I[0] - I[1]  (a one-sided difference at the faces)
//     LR_FD_NEUMANN_ZERO (1): the central difference is zero on the faces.
// The reference needs 9 difference kernels, 9 squares, 8 sums and a mean (each a pass over 16 MB * B at 160^3); here the
// forward is ONE pass over the displacement (12 B per voxel from HBM, neighbours from L1/L2) and the backward one more.
constexpr int REG_TX = 32, REG_TY = 8;                 // block tile: 32 voxels along the fastest axis, 8 rows
#ifndef LR_REG_PF
#define LR_REG_PF 4
#endif
#ifndef LR_REG_WAVES
#define LR_REG_WAVES 1
#endif
// A thread marches planes serially and has only three first-touch loads in flight per plane (one per channel), so the
// march is bound by DRAM latency, not bandwidth; the planes LR_REG_PF ahead are pulled into L2 while it works.
__device__ __forceinline__ void prefetch_l2(const float *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

struct RegDims {
    int d, h, w;              // first, second, last spatial axis (the reference calls them D, W, H)
    int zchunk, n_zchunks;    // planes of the first axis marched by one block
    float sd, sh, sw;         // 0.5 / (2/(n-1)) per axis, rounded to fp32 like the Python scalar entering the tensor op
};

// unscaled (xp - xm) from the upper / lower neighbour u, d (loaded with offsets clamped at the faces) and the centre c
template <int MODE>
__device__ __forceinline__ float fd_delta(float u, float d, float c, bool lo, bool hi) {
    if (MODE != 0) return (lo || hi) ? 0.0f : sub_rn(u, d);
    const float xp = hi ? fmaf(2.0f, c, -d) : u;          // 2 I[n-1] - I[n-2]   (2c is exact, so this is sub_rn(mul_rn(2,c), d))
    const float xm = lo ? fmaf(2.0f, c, -u) : d;          // 2 I[0] - I[1]
    return sub_rn(xp, xm);
}

__device__ __forceinline__ void block_accumulate_f64(double v, double *__restrict__ dst) {
    __shared__ double part64[NCC_THREADS / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double w = warp_sum(v);
    if (lane == 0) part64[warp] = w;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < NCC_THREADS / 32; ++k) t += part64[k];
        atomicAdd(dst, t);
    }
}

// Forward: a block owns a 32 x 8 tile of (last, second) axis positions and marches a chunk of planes of the first axis
// with a three-plane register window per channel (each plane is read from HBM once; the in-plane neighbours hit L1).
template <int MODE>
__global__ void __launch_bounds__(NCC_THREADS) diffusion_reg_kernel(const float *__restrict__ disp, RegDims g, double *__restrict__ sum) {
    const int x = blockIdx.x * REG_TX + (threadIdx.x & 31), y = blockIdx.y * REG_TY + (threadIdx.x >> 5);
    const int zc = blockIdx.z % g.n_zchunks, b = blockIdx.z / g.n_zchunks;
    const int z0 = zc * g.zchunk, z1 = min(g.d, z0 + g.zchunk);
    const int64_t plane = (int64_t)g.h * g.w, cstride = plane * g.d;
    double acc = 0.0;
    if (x < g.w && y < g.h) {
        const bool xlo = x == 0, xhi = x == g.w - 1, ylo = y == 0, yhi = y == g.h - 1;
        const int oxu = xhi ? 0 : 1, oxd = xlo ? 0 : -1;
        const int oyu = yhi ? 0 : g.w, oyd = ylo ? 0 : -g.w;
        // 32-bit element indices (the host checks B*3*D*H*W < 2^31): one IMAD.WIDE per address instead of 64-bit chains
        const unsigned pl = (unsigned)plane, cs = (unsigned)cstride;
        unsigned i0 = ((unsigned)b * 3u * (unsigned)g.d + (unsigned)z0) * pl + (unsigned)y * (unsigned)g.w + (unsigned)x;
        float vm[3], vc[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            vc[c] = __ldg(disp + (i0 + c * cs));
            vm[c] = z0 > 0 ? __ldg(disp + (i0 + c * cs - pl)) : vc[c];
        }
        for (int z = z0; z < z1; ++z, i0 += pl) {
            const bool zlo = z == 0, zhi = z == g.d - 1;
            float l2 = 0.0f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const unsigned ic = i0 + c * cs;
                const float cc = vc[c];
                const float vp = zhi ? cc : __ldg(disp + (ic + pl));      // allocating load: the next iteration's in-plane neighbours hit L1
                if (LR_REG_PF > 0 && z + LR_REG_PF <= z1) prefetch_l2(disp + (ic + LR_REG_PF * pl));
                const float yu = __ldg(disp + (ic + oyu)), yd = __ldg(disp + (ic + oyd));
                const float xu = __ldg(disp + (ic + oxu)), xd = __ldg(disp + (ic + oxd));
                const float dd = mul_rn(fd_delta<MODE>(vp, vm[c], cc, zlo, zhi), g.sd);
                const float dh = mul_rn(fd_delta<MODE>(yu, yd, cc, ylo, yhi), g.sh);
                const float dw = mul_rn(fd_delta<MODE>(xu, xd, cc, xlo, xhi), g.sw);
                l2 = add_rn(l2, mul_rn(dd, dd));          // SubspaceLoss.py:55-63, left to right
                l2 = add_rn(l2, mul_rn(dh, dh));
                l2 = add_rn(l2, mul_rn(dw, dw));
                vm[c] = cc; vc[c] = vp;
            }
            acc += (double)l2;
        }
    }
    block_accumulate_f64(acc, sum);
}

// d reg / d disp.  With D the difference operator of one axis (rows q, columns p) and s its scale,
//     reg = (1/M) sum_q (s (D I)_q)^2   =>   d reg / d I_p = (2 s^2 / M) sum_q D_qp (D I)_q ;
// row q of D has +1 at q+1 and -1 at q-1 in the interior, and (linear mode) -2, +2 at (0, 1) resp. (n-2, n-1) on the faces.
// Interior: sum_q D_qp (D I)_q = (I_p - I_{p-2}) - (I_{p+2} - I_p).  Within two entries of a face the rows of D that
// follow the face rule enter with their own coefficients (the rare predicated block below).
template <int MODE>
__device__ __forceinline__ float fd_adjoint(const float *__restrict__ base, unsigned p, float c, int i, int n, unsigned stride) {
    if (i >= 2 && i <= n - 3) return sub_rn(sub_rn(c, __ldg(base + (p - 2u * stride))), sub_rn(__ldg(base + (p + 2u * stride)), c));
    float tm = 0.0f, tp = 0.0f, t0 = 0.0f;
    if (i >= 2) tm = sub_rn(c, __ldg(base + (p - 2u * stride)));          // row p-1 is an interior row
    if (i <= n - 3) tp = sub_rn(__ldg(base + (p + 2u * stride)), c);      // row p+1 is an interior row
    if (MODE == 0) {
        const float bb = i >= 1 ? __ldg(base + (p - stride)) : c, dd = i <= n - 2 ? __ldg(base + (p + stride)) : c;
        if (i == 1) tm = 2.0f * sub_rn(c, fmaf(2.0f, bb, -c));            // +2 * row 0:   I[1] - (2 I[0] - I[1])
        if (i == n - 2) tp = 2.0f * sub_rn(fmaf(2.0f, dd, -c), c);        // -2 * row n-1: (2 I[n-1] - I[n-2]) - I[n-2]  (sign below)
        if (i == 0) t0 = -2.0f * sub_rn(dd, fmaf(2.0f, c, -dd));          // -2 * row 0
        if (i == n - 1) t0 = add_rn(t0, 2.0f * sub_rn(fmaf(2.0f, c, -bb), bb));   // +2 * row n-1
    }
    return add_rn(sub_rn(tm, tp), t0);
}

template <int MODE>
__global__ void __launch_bounds__(NCC_THREADS) diffusion_reg_backward_kernel(const float *__restrict__ disp, RegDims g, int B,
                                                                             const float *__restrict__ grad_loss,
                                                                             float *__restrict__ grad_disp) {
    const int x = blockIdx.x * REG_TX + (threadIdx.x & 31), y = blockIdx.y * REG_TY + (threadIdx.x >> 5);
    const int zc = blockIdx.z % g.n_zchunks, b = blockIdx.z / g.n_zchunks;
    const int z0 = zc * g.zchunk, z1 = min(g.d, z0 + g.zchunk);
    const int64_t plane = (int64_t)g.h * g.w, cstride = plane * g.d;
    if (x >= g.w || y >= g.h) return;
    const float k = (float)(2.0 * (double)grad_loss[0] / ((double)B * (double)g.d * (double)plane));
    const float kd = k * g.sd * g.sd, kh = k * g.sh * g.sh, kw = k * g.sw * g.sw;
    const unsigned pl = (unsigned)plane, cs = (unsigned)cstride;      // 32-bit element indices, as in the forward
    unsigned off = ((unsigned)b * 3u * (unsigned)g.d + (unsigned)z0) * pl + (unsigned)y * (unsigned)g.w + (unsigned)x;
    for (int z = z0; z < z1; ++z, off += pl) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const unsigned p = off + c * cs;
            const float cc = __ldg(disp + p);
            if (LR_REG_PF > 0 && z + LR_REG_PF + 2 < min(g.d, z1 + 2)) prefetch_l2(disp + (p + (LR_REG_PF + 2) * pl));
            float gr = kd * fd_adjoint<MODE>(disp, p, cc, z, g.d, pl);
            gr = fmaf(kh, fd_adjoint<MODE>(disp, p, cc, y, g.h, (unsigned)g.w), gr);
            gr = fmaf(kw, fd_adjoint<MODE>(disp, p, cc, x, g.w, 1u), gr);
            st_stream(grad_disp + p, gr);
        }
    }
}

// ---- two voxels per thread (W even, 8-byte-aligned field): 64-bit loads and packed fp32x2 arithmetic -------------------
// The scalar kernels above issue ~140 instructions per voxel, most of them addresses and face selects, and are bound by
// the issue rate (ncu: 55 % issue-active at 0.2 of the HBM roofline).  Here a thread owns the voxels (x, x+1), x even:
// every load of the z / y neighbours is one aligned LDG.64, the differences of both voxels run as one FADD2 / FMUL2 /
// FFMA2 stream, and the face rules are a few predicated instructions (x: the first and last pair of a row).
constexpr int REGP_TX = 16, REGP_TY = 16;        // block tile: 16 pairs (32 voxels, one 128-byte line) x 16 rows
#ifndef LR_REGP_MINB
#define LR_REGP_MINB 1
#endif
// measured at 160^3 (profiles/README.md): the adjoint is fastest with 6 resident blocks, two waves of shorter chunks
// and a deeper L2 prefetch (37 us vs 42); the forward with the compiler's own register budget and one wave (26.6 us)
#ifndef LR_REGP_BWD_MINB
#define LR_REGP_BWD_MINB 6
#endif
#ifndef LR_REG_BWD_PF
#define LR_REG_BWD_PF 8
#endif
#ifndef LR_REG_BWD_WAVES
#define LR_REG_BWD_WAVES 2
#endif

__device__ __forceinline__ f32x2 ldg2(const float *p) { return __ldg(reinterpret_cast<const unsigned long long *>(p)); }
__device__ __forceinline__ f32x2 twice_minus(f32x2 c, f32x2 d) { return sub2(add2(c, c), d); }      // 2c - d, 2c exact

template <int MODE>
__device__ __forceinline__ f32x2 fd_delta2(f32x2 u, f32x2 d, f32x2 c, bool lo, bool hi) {
    if (MODE != 0) return (lo || hi) ? 0ull : sub2(u, d);
    if (hi) u = twice_minus(c, d);
    else if (lo) d = twice_minus(c, u);
    return sub2(u, d);
}

template <int MODE>
__global__ void __launch_bounds__(NCC_THREADS, LR_REGP_MINB) diffusion_reg_pair_kernel(const float *__restrict__ disp, RegDims g, double *__restrict__ sum) {
    const int x = (blockIdx.x * REGP_TX + (threadIdx.x & (REGP_TX - 1))) * 2, y = blockIdx.y * REGP_TY + (threadIdx.x / REGP_TX);
    const int zc = blockIdx.z % g.n_zchunks, b = blockIdx.z / g.n_zchunks;
    const int z0 = zc * g.zchunk, z1 = min(g.d, z0 + g.zchunk);
    double acc = 0.0;
    if (x < g.w && y < g.h) {
        const bool xlo = x == 0, xhi = x + 2 == g.w, ylo = y == 0, yhi = y == g.h - 1;
        const int oxl = xlo ? 0 : -1, oxr = xhi ? 1 : 2;                 // I[x-1], I[x+2], clamped into the row
        const int oyu = yhi ? 0 : g.w, oyd = ylo ? 0 : -g.w;
        const unsigned pl = (unsigned)(g.h * g.w), cs = pl * (unsigned)g.d;
        unsigned i0 = ((unsigned)b * 3u * (unsigned)g.d + (unsigned)z0) * pl + (unsigned)y * (unsigned)g.w + (unsigned)x;
        const f32x2 sd2 = splat2(g.sd), sh2 = splat2(g.sh), sw2 = splat2(g.sw);
#ifndef LR_REG_WINDOW
#define LR_REG_WINDOW 1
#endif
        f32x2 vm[3], vc[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            vc[c] = LR_REG_WINDOW ? ldg2(disp + (i0 + c * cs)) : 0ull;
            vm[c] = LR_REG_WINDOW ? (z0 > 0 ? ldg2(disp + (i0 + c * cs - pl)) : vc[c]) : 0ull;
        }
        f32x2 part = 0ull;                                               // fp32 partial sums of at most 16 planes
        for (int z = z0; z < z1; ++z, i0 += pl) {
            const bool zlo = z == 0, zhi = z == g.d - 1;
            f32x2 l2 = 0ull;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const unsigned ic = i0 + c * cs;
                const f32x2 cc = LR_REG_WINDOW ? vc[c] : ldg2(disp + ic);
                if (!LR_REG_WINDOW) vm[c] = zlo ? cc : ldg2(disp + (ic - pl));
                const f32x2 vp = zhi ? cc : ldg2(disp + (ic + pl));      // allocating load: next iteration's row neighbours hit L1
                if (LR_REG_PF > 0 && z + LR_REG_PF <= z1) prefetch_l2(disp + (ic + LR_REG_PF * pl));
                const f32x2 yu = ldg2(disp + (ic + oyu)), yd = ldg2(disp + (ic + oyd));
                float xl = __ldg(disp + (ic + oxl)), xr = __ldg(disp + (ic + oxr));
                float c0, c1;
                unpack2(cc, c0, c1);
                f32x2 dx;
                if (MODE == 0) {
                    if (xlo) xl = fmaf(2.0f, c0, -c1);                   // xm of voxel 0:     2 I[0] - I[1]
                    if (xhi) xr = fmaf(2.0f, c1, -c0);                   // xp of voxel W-1:   2 I[W-1] - I[W-2]
                    dx = sub2(pack2(c1, xr), pack2(xl, c0));
                } else {
                    dx = pack2(xlo ? 0.0f : sub_rn(c1, xl), xhi ? 0.0f : sub_rn(xr, c0));
                }
                const f32x2 dd = mul2(fd_delta2<MODE>(vp, vm[c], cc, zlo, zhi), sd2);
                const f32x2 dh = mul2(fd_delta2<MODE>(yu, yd, cc, ylo, yhi), sh2);
                const f32x2 dw = mul2(dx, sw2);
                l2 = fma2(dd, dd, l2);                                   // SubspaceLoss.py:55-63 in order, squares fused into the sum
                l2 = fma2(dh, dh, l2);
                l2 = fma2(dw, dw, l2);
                vm[c] = cc; vc[c] = vp;
            }
            part = add2(part, l2);
            if (((z - z0) & 15) == 15 || z == z1 - 1) {
                float p0, p1;
                unpack2(part, p0, p1);
                acc += (double)p0 + (double)p1;
                part = 0ull;
            }
        }
    }
    block_accumulate_f64(acc, sum);      // one fp64 atomic per block (spreading them over 64 addresses changed nothing)
}

// x term of the adjoint for the pair (x, x+1): interior (I_p - I_{p-2}) - (I_{p+2} - I_p); the first pair of a row sees
// row 0 of the difference operator, the last pair row W-1 (see fd_adjoint)
template <int MODE>
__device__ __forceinline__ f32x2 fd_adjoint_x2(const float *__restrict__ base, unsigned p, f32x2 cc, int x, int w) {
    if (x >= 2 && x + 4 <= w) return sub2(sub2(cc, ldg2(base + (p - 2u))), sub2(ldg2(base + (p + 2u)), cc));
    float c0, c1, tm0 = 0.0f, tm1 = 0.0f, tp0 = 0.0f, tp1 = 0.0f, t00 = 0.0f, t01 = 0.0f;
    unpack2(cc, c0, c1);
    if (x >= 2) {
        float a0, a1;
        unpack2(ldg2(base + (p - 2u)), a0, a1);
        tm0 = sub_rn(c0, a0); tm1 = sub_rn(c1, a1);
    } else if (MODE == 0) {
        const float q = sub_rn(c1, fmaf(2.0f, c0, -c1));                 // row 0:   I[1] - (2 I[0] - I[1])
        tm1 = 2.0f * q; t00 = -2.0f * q;
    }
    if (x + 4 <= w) {
        float e0, e1;
        unpack2(ldg2(base + (p + 2u)), e0, e1);
        tp0 = sub_rn(e0, c0); tp1 = sub_rn(e1, c1);
    } else if (MODE == 0) {
        const float r = sub_rn(fmaf(2.0f, c1, -c0), c0);                 // row W-1: (2 I[W-1] - I[W-2]) - I[W-2]
        tp0 = 2.0f * r; t01 = 2.0f * r;
    }
    return pack2(add_rn(sub_rn(tm0, tp0), t00), add_rn(sub_rn(tm1, tp1), t01));
}

// y / z term for the pair: both voxels share the index i along the axis
template <int MODE>
__device__ __forceinline__ f32x2 fd_adjoint_s2(const float *__restrict__ base, unsigned p, f32x2 cc, int i, int n, unsigned stride) {
    if (i >= 2 && i <= n - 3) return sub2(sub2(cc, ldg2(base + (p - 2u * stride))), sub2(ldg2(base + (p + 2u * stride)), cc));
    float c0, c1;
    unpack2(cc, c0, c1);
    return pack2(fd_adjoint<MODE>(base, p, c0, i, n, stride), fd_adjoint<MODE>(base, p + 1u, c1, i, n, stride));
}

template <int MODE>
__global__ void __launch_bounds__(NCC_THREADS, LR_REGP_BWD_MINB) diffusion_reg_backward_pair_kernel(const float *__restrict__ disp, RegDims g, int B,
                                                                                  const float *__restrict__ grad_loss,
                                                                                  float *__restrict__ grad_disp) {
    const int x = (blockIdx.x * REGP_TX + (threadIdx.x & (REGP_TX - 1))) * 2, y = blockIdx.y * REGP_TY + (threadIdx.x / REGP_TX);
    const int zc = blockIdx.z % g.n_zchunks, b = blockIdx.z / g.n_zchunks;
    const int z0 = zc * g.zchunk, z1 = min(g.d, z0 + g.zchunk);
    if (x >= g.w || y >= g.h) return;
    const unsigned pl = (unsigned)(g.h * g.w), cs = pl * (unsigned)g.d;
    const float k = (float)(2.0 * (double)grad_loss[0] / ((double)B * (double)g.d * (double)pl));
    const f32x2 kd = splat2(k * g.sd * g.sd), kh = splat2(k * g.sh * g.sh), kw = splat2(k * g.sw * g.sw);
    unsigned off = ((unsigned)b * 3u * (unsigned)g.d + (unsigned)z0) * pl + (unsigned)y * (unsigned)g.w + (unsigned)x;
    for (int z = z0; z < z1; ++z, off += pl) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const unsigned p = off + c * cs;
            const f32x2 cc = ldg2(disp + p);
            if (LR_REG_BWD_PF > 0 && z + LR_REG_BWD_PF + 2 < min(g.d, z1 + 2)) prefetch_l2(disp + (p + (LR_REG_BWD_PF + 2) * pl));
            f32x2 gr = mul2(kd, fd_adjoint_s2<MODE>(disp, p, cc, z, g.d, pl));
            gr = fma2(kh, fd_adjoint_s2<MODE>(disp, p, cc, y, g.h, (unsigned)g.w), gr);
            gr = fma2(kw, fd_adjoint_x2<MODE>(disp, p, cc, x, g.w), gr);
            float g0, g1;
            unpack2(gr, g0, g1);
            __stcs(reinterpret_cast<float2 *>(grad_disp + p), make_float2(g0, g1));
        }
    }
}

template <typename K>
static int reg_dims(K kernel, bool pairs, int waves, int B, int D, int H, int W, RegDims &g, dim3 &grid) {
    const int tile_w = pairs ? 2 * REGP_TX : REG_TX, tile_h = pairs ? REGP_TY : REG_TY;
    g.d = D; g.h = H; g.w = W;
    const double sp[3] = {1.0 / (double)(D - 1) * 2.0, 1.0 / (double)(H - 1) * 2.0, 1.0 / (double)(W - 1) * 2.0};   // spacing*2 (:53-54)
    g.sd = (float)(0.5 / sp[0]); g.sh = (float)(0.5 / sp[1]); g.sw = (float)(0.5 / sp[2]);
    const int64_t tiles = (int64_t)((W + tile_w - 1) / tile_w) * ((H + tile_h - 1) / tile_h) * B;
    // ONE wave of blocks when the tiles allow it (a second, nearly empty wave would double the time of this short
    // kernel): as many chunks of planes as fit the resident-block capacity, chunks of at least 4 planes
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, NCC_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
    const int64_t capacity = (int64_t)sm_count() * per_sm * waves;
    int64_t chunks = capacity / tiles;
    if (chunks < 1) chunks = 1;
    if (chunks > (D + 3) / 4) chunks = (D + 3) / 4;
    g.zchunk = (int)((D + chunks - 1) / chunks);
    g.n_zchunks = (D + g.zchunk - 1) / g.zchunk;
    const int64_t gz = (int64_t)g.n_zchunks * B;
    if (gz > 65535 || (H + tile_h - 1) / tile_h > 65535) return -1;
    if ((int64_t)B * 3 * D * H * W >= ((int64_t)1 << 31)) return -1;          // the kernels index with 32 bits
    grid = dim3((unsigned)((W + tile_w - 1) / tile_w), (unsigned)((H + tile_h - 1) / tile_h), (unsigned)gz);
    return 0;
}

static unsigned ncc_blocks(int64_t N, int B) {
    int64_t want = (N + NCC_THREADS * 8 - 1) / (NCC_THREADS * 8);          // >= 8 elements per thread
    const int64_t cap = (int64_t)sm_count() * 8 / (B < 8 ? B : 8) + 1;      // ~8 resident blocks per SM over the whole grid
    if (want > cap) want = cap;
    return (unsigned)(want < 1 ? 1 : want);
}

// the pair kernels need an even row length and 8-byte-aligned rows (LIFTREG_B200_REG_PAIRS=0 forces the scalar kernels)
static bool reg_use_pairs(const void *a, const void *b, int W) {
    static const bool enabled = [] { const char *e = getenv("LIFTREG_B200_REG_PAIRS"); return !(e && e[0] == '0'); }();
    return enabled && (W % 2 == 0) && ((uintptr_t)a % 8 == 0) && ((uintptr_t)b % 8 == 0);
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
    // LIFTREG_B200_NCC_TWO_PASS=1: the op-for-op restatement (means first, then the fp32-centred moments)
    static const bool two_pass = [] { const char *e = getenv("LIFTREG_B200_NCC_TWO_PASS"); return e && e[0] == '1'; }();
    if (!two_pass) {
        ncc_shifted_kernel<<<grid, NCC_THREADS, 0, st>>>(x, y, N, sums);
        if (int e = check_launch("ncc_shifted_kernel")) return e;
        ncc_finalize_kernel<<<(B + 127) / 128, 128, 0, st>>>(x, y, N, B, sums);
        return check_launch("ncc_finalize_kernel");
    }
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

extern "C" int lr_diffusion_reg_sum(const float *disp, int B, int D, int H, int W, int boundary, double *sum, lr_stream_t stream) {
    LR_REQUIRE(disp && sum, "diffusion_reg_sum: null pointer");
    LR_REQUIRE(B > 0 && D >= 2 && H >= 2 && W >= 2, "diffusion_reg_sum: every axis needs >= 2 entries (B=%d, %dx%dx%d)", B, D, H, W);
    LR_REQUIRE(boundary == LR_FD_LINEAR || boundary == LR_FD_NEUMANN_ZERO, "diffusion_reg_sum: unknown boundary mode %d", boundary);
    RegDims g; dim3 grid;
    const bool pairs = reg_use_pairs(disp, nullptr, W);
    const int rd = pairs ? (boundary == LR_FD_LINEAR ? reg_dims(diffusion_reg_pair_kernel<0>, true, LR_REG_WAVES, B, D, H, W, g, grid)
                                                     : reg_dims(diffusion_reg_pair_kernel<1>, true, LR_REG_WAVES, B, D, H, W, g, grid))
                         : (boundary == LR_FD_LINEAR ? reg_dims(diffusion_reg_kernel<0>, false, LR_REG_WAVES, B, D, H, W, g, grid)
                                                     : reg_dims(diffusion_reg_kernel<1>, false, LR_REG_WAVES, B, D, H, W, g, grid));
    LR_REQUIRE(rd == 0, "diffusion_reg_sum: field too large for one launch (B=%d, %dx%dx%d; B*3*D*H*W must stay below 2^31)", B, D, H, W);
    cudaStream_t st = as_stream(stream);
    cudaError_t ce = cudaMemsetAsync(sum, 0, sizeof(double), st);
    if (ce != cudaSuccess) { set_error("diffusion_reg_sum: memset failed: %s", cudaGetErrorString(ce)); return LR_ERR_CUDA; }
    if (pairs) {
        if (boundary == LR_FD_LINEAR) diffusion_reg_pair_kernel<0><<<grid, NCC_THREADS, 0, st>>>(disp, g, sum);
        else diffusion_reg_pair_kernel<1><<<grid, NCC_THREADS, 0, st>>>(disp, g, sum);
        return check_launch("diffusion_reg_pair_kernel");
    }
    if (boundary == LR_FD_LINEAR) diffusion_reg_kernel<0><<<grid, NCC_THREADS, 0, st>>>(disp, g, sum);
    else diffusion_reg_kernel<1><<<grid, NCC_THREADS, 0, st>>>(disp, g, sum);
    return check_launch("diffusion_reg_kernel");
}

extern "C" int lr_diffusion_reg_backward(const float *disp, int B, int D, int H, int W, int boundary, const float *grad_loss,
                                         float *grad_disp, lr_stream_t stream) {
    LR_REQUIRE(disp && grad_loss && grad_disp, "diffusion_reg_backward: null pointer");
    LR_REQUIRE(B > 0 && D >= 2 && H >= 2 && W >= 2, "diffusion_reg_backward: every axis needs >= 2 entries (B=%d, %dx%dx%d)", B, D, H, W);
    LR_REQUIRE(boundary == LR_FD_LINEAR || boundary == LR_FD_NEUMANN_ZERO, "diffusion_reg_backward: unknown boundary mode %d", boundary);
    RegDims g; dim3 grid;
    const bool pairs = reg_use_pairs(disp, grad_disp, W);
    const int rd = pairs ? (boundary == LR_FD_LINEAR ? reg_dims(diffusion_reg_backward_pair_kernel<0>, true, LR_REG_BWD_WAVES, B, D, H, W, g, grid)
                                                     : reg_dims(diffusion_reg_backward_pair_kernel<1>, true, LR_REG_BWD_WAVES, B, D, H, W, g, grid))
                         : (boundary == LR_FD_LINEAR ? reg_dims(diffusion_reg_backward_kernel<0>, false, LR_REG_BWD_WAVES, B, D, H, W, g, grid)
                                                     : reg_dims(diffusion_reg_backward_kernel<1>, false, LR_REG_BWD_WAVES, B, D, H, W, g, grid));
    LR_REQUIRE(rd == 0, "diffusion_reg_backward: field too large for one launch (B=%d, %dx%dx%d; B*3*D*H*W must stay below 2^31)", B, D, H, W);
    cudaStream_t st = as_stream(stream);
    if (pairs) {
        if (boundary == LR_FD_LINEAR) diffusion_reg_backward_pair_kernel<0><<<grid, NCC_THREADS, 0, st>>>(disp, g, B, grad_loss, grad_disp);
        else diffusion_reg_backward_pair_kernel<1><<<grid, NCC_THREADS, 0, st>>>(disp, g, B, grad_loss, grad_disp);
        return check_launch("diffusion_reg_backward_pair_kernel");
    }
    if (boundary == LR_FD_LINEAR) diffusion_reg_backward_kernel<0><<<grid, NCC_THREADS, 0, st>>>(disp, g, B, grad_loss, grad_disp);
    else diffusion_reg_backward_kernel<1><<<grid, NCC_THREADS, 0, st>>>(disp, g, B, grad_loss, grad_disp);
    return check_launch("diffusion_reg_backward_kernel");
}
