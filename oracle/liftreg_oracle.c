/*
 * liftreg_oracle.c -- CPU restatement of LiftReg's geometric resampling path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under liftreg_b200/ may import, link or
 * call this file; it is the checker used by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.
 *
 * What it restates (reference = uncbiag/LiftReg, paths relative to its root;
 * "sdct" = src/liftreg/utils/sdct_projection_utils.py):
 *
 *   lro_project_grid        sdct:15-57   project_grid_multi (ray sample grid + dx)
 *                           layers.py:194-236 proj_layer._project_grid_multi (y_mode=1)
 *   lro_drr_forward         sdct:59-86   calculate_projection (flip, grid_sample, sum, *dx, *0.1)
 *                           layers.py:182-187 proj_layer.forward (before F.interpolate)
 *   lro_drr_backward        autograd adjoint of the above wrt the volume
 *                           (models/previous/RegNet2D3D.py:161-171 needs it)
 *   lro_backproj_grid       sdct:227-250 backproj_grids_with_poses
 *   lro_backproject_forward models/LiftRegDeformSubspaceBackproj.py:85-93
 *   lro_backproject_backward adjoint wrt the projections
 *   lro_warp_forward        utils/net_utils.py:26-56 Bilinear.forward / forward_stn
 *   lro_warp_backward       autograd adjoint wrt image and phi
 *   lro_identity_map        utils/net_utils.py:59-87
 *
 * The sampling arithmetic itself lives in a THIRD-PARTY dependency that is not
 * under /root/reference: PyTorch ATen grid_sampler_{2d,3d} (reference pins
 * torch==1.9.0+cu111, requirements.txt:133; this image has 2.11.0).  The
 * restatement follows the ATen CPU kernels as observed in this image:
 *   3-D: scalar kernel, aten/src/ATen/native/GridSampler.cpp
 *        unnormalise ((c+1)/2)*(size-1); taps tnw,tne,tsw,tse,bnw,bne,bsw,bse;
 *        weight (x1-x)*(y1-y)*(z1-z) left to right; out += val*w with NO fma.
 *   2-D: vectorised kernel, aten/src/ATen/native/cpu/GridSamplerKernel.cpp
 *        unnormalise (c+1)*((size-1)/2); w=x-floor(x), e=1-w, n=y-floor(y), s=1-n;
 *        out = fma(se_v,n*w, fma(sw_v,n*e, fma(ne_v,s*w, nw_v*(s*e)))).
 * Both orders were pinned bit-exactly against torch 2.11 CPU F.grid_sample
 * (tests/test_oracle_golden.py re-checks that on every run) and against
 * golden vectors produced by importing the reference itself
 * (oracle/make_golden.py -> tests/golden/).
 *
 * Everything is fp32 with the reference's op order; compile with
 * -ffp-contract=off so that the only fused multiply-adds are the explicit
 * fmaf() calls below (torch's CPU norm and 2-D sampler do contract).
 * The one order this file does NOT reproduce is torch.sum's vectorised
 * cascade over the ray (sdct:81); rays are summed sequentially in fp32
 * (acc64=0) or in double (acc64=1), optionally in runs of seg_len planes
 * whose sums are then added in order (the CUDA kernel's ray-segment
 * order).  All of these differ by <=2e-7 relative.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LRO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* Blend order                                                        */
/* ------------------------------------------------------------------ */
/* 0 (default): ATen's operation order -- what the reference computes, pinned against torch and the goldens.
 * 1: the operation order of the CUDA library's LR_NUMERICS_FAST mode (include/liftreg_b200.h): identical
 *    coordinates, floor indices and weights; the 4-tap / 8-tap blend evaluated as fused linear interpolations
 *    (backprojection: detector rows interpolated along axis 1, then blended along axis 0; warp: x, then y, then z,
 *    with `using_scale` folded away and skipped taps entering as -1).  Exists so that the tests can demand BIT-EXACT
 *    agreement from the fast kernels too (which proves their indexing), and bound fast-vs-ATen on the CPU. */
static int g_blend = 0;
LRO_API void lro_set_blend(int mode) { g_blend = mode != 0; }
LRO_API int lro_get_blend(void) { return g_blend; }

/* ------------------------------------------------------------------ */
/* ATen grid_sampler semantics                                        */
/* ------------------------------------------------------------------ */

/* GridSampler.h:27-36 grid_sampler_unnormalize, align_corners=True (3-D scalar path) */
static inline float unnorm3(float c, int size) {
    return ((c + 1.0f) / 2.0f) * (float)(size - 1);
}
/* GridSamplerKernel.cpp ComputeLocationBase<align_corners=true>::unnormalize (2-D vector path) */
static inline float unnorm2(float c, int size) {
    return (c + 1.0f) * ((float)(size - 1) / 2.0f);
}
/* GridSampler.h:58-60 clip_coordinates (border padding) */
static inline float clipc(float c, int size) {
    return fminf((float)(size - 1), fmaxf(c, 0.0f));
}
static inline int inb3(int64_t z, int64_t y, int64_t x, int D, int H, int W) {
    return z >= 0 && z < D && y >= 0 && y < H && x >= 0 && x < W;
}

/* One trilinear / nearest sample of vol[D][H][W] at normalised (gx->W, gy->H, gz->D).
 * padding: 0 zeros, 1 border.  mode: 0 linear, 1 nearest. */
static inline float sample3(const float *vol, int D, int H, int W,
                            float gx, float gy, float gz, int padding, int mode) {
    float ix = unnorm3(gx, W), iy = unnorm3(gy, H), iz = unnorm3(gz, D);
    if (padding == 1) { ix = clipc(ix, W); iy = clipc(iy, H); iz = clipc(iz, D); }
    if (mode == 1) {
        int64_t xn = (int64_t)nearbyintf(ix), yn = (int64_t)nearbyintf(iy), zn = (int64_t)nearbyintf(iz);
        return inb3(zn, yn, xn, D, H, W) ? vol[((size_t)zn * H + yn) * W + xn] : 0.0f;
    }
    int64_t x0 = (int64_t)floorf(ix), y0 = (int64_t)floorf(iy), z0 = (int64_t)floorf(iz);
    int64_t x1 = x0 + 1, y1 = y0 + 1, z1 = z0 + 1;
    float wx0 = (float)x1 - ix, wx1 = ix - (float)x0;
    float wy0 = (float)y1 - iy, wy1 = iy - (float)y0;
    float wz0 = (float)z1 - iz, wz1 = iz - (float)z0;
    float tnw = wx0 * wy0 * wz0, tne = wx1 * wy0 * wz0, tsw = wx0 * wy1 * wz0, tse = wx1 * wy1 * wz0;
    float bnw = wx0 * wy0 * wz1, bne = wx1 * wy0 * wz1, bsw = wx0 * wy1 * wz1, bse = wx1 * wy1 * wz1;
    float out = 0.0f;
#define TAP(zz, yy, xx, ww) if (inb3(zz, yy, xx, D, H, W)) out += vol[((size_t)(zz) * H + (yy)) * W + (xx)] * (ww);
    TAP(z0, y0, x0, tnw) TAP(z0, y0, x1, tne) TAP(z0, y1, x0, tsw) TAP(z0, y1, x1, tse)
    TAP(z1, y0, x0, bnw) TAP(z1, y0, x1, bne) TAP(z1, y1, x0, bsw) TAP(z1, y1, x1, bse)
#undef TAP
    return out;
}

/* Linear sample in the fast kernels' order (g_blend == 1): taps that padding skips enter as `masked`. */
static inline float sample3_fast(const float *vol, int D, int H, int W,
                                 float gx, float gy, float gz, int padding, float masked) {
    float ix = unnorm3(gx, W), iy = unnorm3(gy, H), iz = unnorm3(gz, D);
    if (padding == 1) { ix = clipc(ix, W); iy = clipc(iy, H); iz = clipc(iz, D); }
    float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
    /* far outside: every tap is skipped whatever the weights are (keeps the int casts defined) */
    if (!(fx > -4.0f && fx < (float)W + 4.0f && fy > -4.0f && fy < (float)H + 4.0f && fz > -4.0f && fz < (float)D + 4.0f))
        return masked;
    int64_t x0 = (int64_t)fx, y0 = (int64_t)fy, z0 = (int64_t)fz;
    float wx1 = ix - fx, wy1 = iy - fy, wz1 = iz - fz;
    float v[8];
    for (int t = 0; t < 8; ++t) {
        int64_t xx = x0 + (t & 1), yy = y0 + ((t >> 1) & 1), zz = z0 + (t >> 2);
        v[t] = inb3(zz, yy, xx, D, H, W) ? vol[((size_t)zz * H + yy) * W + xx] : masked;
    }
    float c00 = fmaf(wx1, v[1] - v[0], v[0]), c10 = fmaf(wx1, v[3] - v[2], v[2]);
    float c01 = fmaf(wx1, v[5] - v[4], v[4]), c11 = fmaf(wx1, v[7] - v[6], v[6]);
    float d0 = fmaf(wy1, c10 - c00, c00), d1 = fmaf(wy1, c11 - c01, c01);
    return fmaf(wz1, d1 - d0, d0);
}

/* One bilinear sample of img[H][W] at normalised (gx->W, gy->H), zeros padding. */
static inline float sample2(const float *img, int H, int W, float gx, float gy) {
    float ix = unnorm2(gx, W), iy = unnorm2(gy, H);
    float xw = floorf(ix), yn = floorf(iy);
    float w = ix - xw, e = 1.0f - w, n = iy - yn, s = 1.0f - n;
    float nw = s * e, ne = s * w, sw = n * e, se = n * w;
    /* integer conversion as the vector kernel does it: convert_to_int_of_same_size */
    int64_t x0 = (int64_t)xw, y0 = (int64_t)yn, x1 = x0 + 1, y1 = y0 + 1;
    /* guard the casts for wildly out-of-range coordinates */
    if (!(xw > -4.0f && xw < (float)W + 4.0f && yn > -4.0f && yn < (float)H + 4.0f)) return 0.0f;
    float a = (y0 >= 0 && y0 < H && x0 >= 0 && x0 < W) ? img[(size_t)y0 * W + x0] : 0.0f;
    float b = (y0 >= 0 && y0 < H && x1 >= 0 && x1 < W) ? img[(size_t)y0 * W + x1] : 0.0f;
    float c = (y1 >= 0 && y1 < H && x0 >= 0 && x0 < W) ? img[(size_t)y1 * W + x0] : 0.0f;
    float d = (y1 >= 0 && y1 < H && x1 >= 0 && x1 < W) ? img[(size_t)y1 * W + x1] : 0.0f;
    if (g_blend) {   /* fast kernels: rows interpolated along x first, then blended along y */
        float t_lo = fmaf(b, w, a * e), t_up = fmaf(d, w, c * e);
        return fmaf(t_up, n, t_lo * s);
    }
    return fmaf(d, se, fmaf(c, sw, fmaf(b, ne, a * nw)));
}

/* Generic entry points so the tests can pin sample3/sample2 against torch's
 * CPU F.grid_sample directly.  grid is (N,3) resp. (N,2) in grid_sample's
 * own (x,y,z) order. */
LRO_API void lro_grid_sample_3d(const float *vol, int D, int H, int W, const float *grid, int64_t N,
                                int padding, int mode, float *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i)
        out[i] = sample3(vol, D, H, W, grid[3 * i], grid[3 * i + 1], grid[3 * i + 2], padding, mode);
}
LRO_API void lro_grid_sample_2d(const float *img, int H, int W, const float *grid, int64_t N, float *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i)
        out[i] = sample2(img, H, W, grid[2 * i], grid[2 * i + 1]);
}

/* ------------------------------------------------------------------ */
/* Cone-beam ray geometry: sdct:15-57 / layers.py:194-236             */
/* ------------------------------------------------------------------ */

typedef struct {
    float sx, sy, sz;   /* emitter, voxel units, fp32 cast of the f64 pose (sdct:28) */
    float Dx, Dy, Dz;   /* normalised direction (sdct:40) */
    float r2;           /* 1/dot(dir,N) (sdct:50) */
    float dx;           /* mm per coronal step (sdct:39,41) */
} lro_ray;

static inline lro_ray ray_setup(const double *pose, int u, int v, int rd, int rh, const float *sp) {
    lro_ray r;
    r.sx = (float)pose[0]; r.sy = (float)pose[1]; r.sz = (float)pose[2];
    /* lin_x = linspace(-rd/2, rd/2-1, rd): unit step, exact (sdct:32-33) */
    float lx = (float)((double)u - (double)rd / 2.0), ly = (float)((double)v - (double)rh / 2.0);
    float Ix = lx + (-r.sx), Iy = 0.0f + (-r.sy), Iz = ly + (-r.sz);       /* sdct:35-38 */
    float rc = 1.0f / Iy;                                                   /* sdct:39 */
    float ax = (Ix * rc) * sp[0], ay = (Iy * rc) * sp[1], az = (Iz * rc) * sp[2];
    r.dx = sqrtf(fmaf(az, az, fmaf(ay, ay, ax * ax)));                      /* sdct:41, torch CPU norm order */
    float n = sqrtf(fmaf(Iz, Iz, fmaf(Iy, Iy, Ix * Ix)));                   /* sdct:40 */
    r.Dx = Ix / n; r.Dy = Iy / n; r.Dz = Iz / n;
    r.r2 = 1.0f / r.Dy;                                                     /* sdct:50, dot with N=(0,1,0) is exact */
    return r;
}

/* Normalised grid point of ray r at coronal plane j, in the reference's
 * (axis0, axis1, axis2) order (before the flip of sdct:76).
 * y_mode 0: /(w-1) (sdct:55); 1: /w (layers.py:234). */
static inline void ray_point(const lro_ray *r, int j, int d, int w, int h, int y_mode, float *g) {
    float T = r->r2 * ((float)j - r->sy);                                   /* sdct:50 (K=1 matmul = one product) */
    float X = r->Dx * T + r->sx, Y = r->Dy * T + r->sy, Z = r->Dz * T + r->sz; /* sdct:51 */
    g[0] = X / (float)d * 2.0f;                                             /* sdct:54 */
    g[1] = (y_mode == 0) ? (Y - 0.0f) / ((float)w - 1.0f) * 2.0f + -1.0f    /* sdct:55 */
                         : (Y - 0.0f) / (float)w * 2.0f + -1.0f;            /* layers.py:234 */
    g[2] = Z / (float)h * 2.0f;                                             /* sdct:56 */
}

/* grid: (P,rd,rh,w,3) or NULL; dx: (P,rd,rh) or NULL */
LRO_API void lro_project_grid(const double *poses, int P, int rd, int rh, int d, int w, int h,
                              const float *spacing, int y_mode, float *grid, float *dx) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int p = 0; p < P; ++p)
        for (int u = 0; u < rd; ++u)
            for (int v = 0; v < rh; ++v) {
                lro_ray r = ray_setup(poses + 3 * p, u, v, rd, rh, spacing);
                size_t ray = ((size_t)p * rd + u) * rh + v;
                if (dx) dx[ray] = r.dx;
                if (grid)
                    for (int j = 0; j < w; ++j) ray_point(&r, j, d, w, h, y_mode, grid + (ray * w + j) * 3);
            }
}

/* vol (B,d,w,h); proj (B,P,rd,rh); samples (optional, B==1 only): (P,rd,rh,w) pre-sum values.
 * proj = ((sum_j sample) * dx) * out_scale   -- sdct:81,85
 * seg_len > 0: the ray is summed in runs of seg_len planes, run sums added in run order (fp32). */
/* The CUDA kernel's conservative clip of a ray to the coronal planes [j0, j1] in which it can touch the volume
 * (liftreg_b200/csrc/drr.cu ray_setup; samples outside contribute exactly +0 in the reference).  Restated only because
 * the fast-numerics kernel cuts each ray pair's clipped range into equal runs (seg_len < 0 below): the run boundaries,
 * and with them the fp32 summation order, depend on it.  Same fp32 expressions in the same order. */
static inline void kernel_ray_clip(const lro_ray *r, int u, int v, int rd, int rh, int d, int w, int h, int *j0, int *j1) {
    float half_rd = (float)((double)rd / 2.0), half_rh = (float)((double)rh / 2.0);
    float lx = (float)u - half_rd, lz = (float)v - half_rh;
    float lim_x = d > 1 ? (float)d / 2.0f + 2.0f + (float)d / (float)(d - 1) : 3.0e38f;
    float lim_z = h > 1 ? (float)h / 2.0f + 2.0f + (float)h / (float)(h - 1) : 3.0e38f;
    float t0 = 0.0f, t1 = (float)(w - 1);
    float inv_sy = 1.0f / r->sy;
    float bx = (r->sx - lx) * inv_sy, bz = (r->sz - lz) * inv_sy;
    if (fabsf(bx) > 1e-12f) {
        float a = (-lim_x - lx) / bx, b = (lim_x - lx) / bx;
        t0 = fmaxf(t0, fminf(a, b)); t1 = fminf(t1, fmaxf(a, b));
    } else if (fabsf(lx) > lim_x) t1 = -1.0f;
    if (fabsf(bz) > 1e-12f) {
        float a = (-lim_z - lz) / bz, b = (lim_z - lz) / bz;
        t0 = fmaxf(t0, fminf(a, b)); t1 = fminf(t1, fmaxf(a, b));
    } else if (fabsf(lz) > lim_z) t1 = -1.0f;
    *j0 = (int)floorf(t0) - 1; if (*j0 < 0) *j0 = 0;
    *j1 = (int)ceilf(t1) + 1; if (*j1 > w - 1) *j1 = w - 1;
    if (!(t1 >= t0)) { *j0 = 0; *j1 = -1; }
    if (!(r->sy > (float)(w - 1))) { *j0 = 0; *j1 = w - 1; }     /* emitter inside the slab: no clipping */
}

LRO_API void lro_drr_forward(const float *vol, int B, int d, int w, int h, const double *poses, int P,
                             int rd, int rh, const float *spacing, int y_mode, float out_scale,
                             int acc64, int seg_len, float *proj, float *samples) {
    for (int b = 0; b < B; ++b) {
        const float *V = vol + (size_t)b * d * w * h;
#pragma omp parallel for collapse(2) schedule(dynamic, 8)
        for (int p = 0; p < P; ++p)
            for (int u = 0; u < rd; ++u)
                for (int v = 0; v < rh; ++v) {
                    lro_ray r = ray_setup(poses + 3 * p, u, v, rd, rh, spacing);
                    size_t ray = ((size_t)p * rd + u) * rh + v;
                    float acc = 0.0f, run = 0.0f; double acc_d = 0.0;
                    /* seg_len < 0: -seg_len balanced runs over the clipped range of the ray PAIR (u & ~1, u | 1) -- the
                     * fast-numerics kernel's order; rays of an unclipped view keep fixed runs of ceil(w / -seg_len) */
                    int bal = 0, bj0 = 0, blen = 1, fixed_len = seg_len;
                    if (seg_len < 0) {
                        int nseg = -seg_len;
                        fixed_len = (w + nseg - 1) / nseg;
                        if (r.sy > (float)(w - 1)) {
                            int ua = u & ~1, ub = (ua + 1 < rd) ? ua + 1 : ua, a0, a1, b0, b1;
                            lro_ray qa = ray_setup(poses + 3 * p, ua, v, rd, rh, spacing), qb = ray_setup(poses + 3 * p, ub, v, rd, rh, spacing);
                            kernel_ray_clip(&qa, ua, v, rd, rh, d, w, h, &a0, &a1);
                            kernel_ray_clip(&qb, ub, v, rd, rh, d, w, h, &b0, &b1);
                            int ea = a1 < a0, eb = b1 < b0;
                            int j0 = ea ? b0 : (eb ? a0 : (a0 < b0 ? a0 : b0));
                            int j1 = ea ? b1 : (eb ? a1 : (a1 > b1 ? a1 : b1));
                            int n = j1 - j0 + 1;
                            bal = 1; bj0 = j0; blen = n > 0 ? (n + nseg - 1) / nseg : 1;
                        }
                    }
                    for (int j = 0; j < w; ++j) {
                        float g[3];
                        ray_point(&r, j, d, w, h, y_mode, g);
                        /* flip (sdct:76): grid_sample x<-axis2 (W=h), y<-axis1 (H=w), z<-axis0 (D=d) */
                        float s = g_blend ? sample3_fast(V, d, w, h, g[2], g[1], g[0], 0, 0.0f)
                                          : sample3(V, d, w, h, g[2], g[1], g[0], 0, 0);
                        if (samples && b == 0) samples[ray * w + j] = s;
                        acc_d += (double)s;
                        if (bal) {                /* runs [bj0 + k*blen, bj0 + (k+1)*blen) */
                            run += s;
                            if ((j >= bj0 && (j - bj0 + 1) % blen == 0) || j == w - 1) { acc += run; run = 0.0f; }
                        } else if (fixed_len > 0) {
                            run += s;
                            if ((j + 1) % fixed_len == 0 || j == w - 1) { acc += run; run = 0.0f; }
                        } else {
                            acc += s;
                        }
                    }
                    float sum = acc64 ? (float)acc_d : acc;
                    float o = sum * r.dx;
                    if (out_scale != 1.0f) o = o * out_scale;
                    proj[(size_t)b * P * rd * rh + ray] = o;
                }
    }
}

/* grad_vol (B,d,w,h) += adjoint; accumulates in double internally (deterministic), caller zeroes. */
LRO_API void lro_drr_backward(const float *grad_proj, int B, int d, int w, int h, const double *poses, int P,
                              int rd, int rh, const float *spacing, int y_mode, float out_scale,
                              float *grad_vol) {
    size_t nv = (size_t)d * w * h;
    double *acc = (double *)calloc(nv, sizeof(double));
    for (int b = 0; b < B; ++b) {
        memset(acc, 0, nv * sizeof(double));
        for (int p = 0; p < P; ++p)
            for (int u = 0; u < rd; ++u)
                for (int v = 0; v < rh; ++v) {
                    lro_ray r = ray_setup(poses + 3 * p, u, v, rd, rh, spacing);
                    size_t ray = ((size_t)p * rd + u) * rh + v;
                    float go = grad_proj[(size_t)b * P * rd * rh + ray];
                    float gs = (out_scale != 1.0f) ? (go * out_scale) * r.dx : go * r.dx;
                    for (int j = 0; j < w; ++j) {
                        float g[3];
                        ray_point(&r, j, d, w, h, y_mode, g);
                        float ix = unnorm3(g[2], h), iy = unnorm3(g[1], w), iz = unnorm3(g[0], d);
                        int64_t x0 = (int64_t)floorf(ix), y0 = (int64_t)floorf(iy), z0 = (int64_t)floorf(iz);
                        float wx[2] = {(float)(x0 + 1) - ix, ix - (float)x0};
                        float wy[2] = {(float)(y0 + 1) - iy, iy - (float)y0};
                        float wz[2] = {(float)(z0 + 1) - iz, iz - (float)z0};
                        for (int c = 0; c < 8; ++c) {
                            int64_t xx = x0 + (c & 1), yy = y0 + ((c >> 1) & 1), zz = z0 + (c >> 2);
                            if (inb3(zz, yy, xx, d, w, h))
                                acc[((size_t)zz * w + yy) * h + xx] += (double)(wx[c & 1] * wy[(c >> 1) & 1] * wz[c >> 2] * gs);
                        }
                    }
                }
        for (size_t i = 0; i < nv; ++i) grad_vol[(size_t)b * nv + i] += (float)acc[i];
    }
    free(acc);
}

/* ------------------------------------------------------------------ */
/* Backprojection: sdct:227-250 + LiftRegDeformSubspaceBackproj.py:85-93 */
/* ------------------------------------------------------------------ */

/* normalised detector coordinates of voxel (i,j,k) for pose s (fp32): gu along detector axis 0 (pw), gv along axis 1 (ph) */
static inline void backproj_point(const float *s, int i, int j, int k, int d, int w, int h, int pw, int ph,
                                  float *gu, float *gv) {
    float x = (float)((double)i - (double)d / 2.0);      /* sdct:231 linspace(-d/2, d/2-1, d) */
    float y = (float)(w - 1 - j);                        /* sdct:232 linspace(w-1, 0, w) -- reversed */
    float z = (float)((double)k - (double)h / 2.0);      /* sdct:233 */
    float scale = s[1] / (s[1] - y);                     /* sdct:239 */
    float a = (x - s[0]) * scale + s[0];                 /* sdct:241-242 */
    float c = (z - s[2]) * scale + s[2];
    *gu = a / (float)pw * 2.0f;                          /* sdct:247 */
    *gv = c / (float)ph * 2.0f;                          /* sdct:248 */
}

/* grid (P,2,d,w,h) in the reference's returned order (after flip(2)): channel 0 = gv, channel 1 = gu */
LRO_API void lro_backproj_grid(const float *poses, int P, int d, int w, int h, int pw, int ph, float *grid) {
    size_t nv = (size_t)d * w * h;
#pragma omp parallel for collapse(2) schedule(static)
    for (int p = 0; p < P; ++p)
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < w; ++j)
                for (int k = 0; k < h; ++k) {
                    float gu, gv;
                    backproj_point(poses + 3 * p, i, j, k, d, w, h, pw, ph, &gu, &gv);
                    size_t vox = ((size_t)i * w + j) * h + k;
                    grid[((size_t)p * 2 + 0) * nv + vox] = gv;
                    grid[((size_t)p * 2 + 1) * nv + vox] = gu;
                }
}

/* proj (B,P,pw,ph); poses (P,3) fp32 (geometry frozen from batch item 0, model :85-87); out (B,P,d,w,h) */
LRO_API void lro_backproject_forward(const float *proj, const float *poses, int B, int P, int pw, int ph,
                                     int d, int w, int h, float *out) {
    size_t nv = (size_t)d * w * h;
#pragma omp parallel for collapse(3) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int p = 0; p < P; ++p)
            for (int i = 0; i < d; ++i) {
                const float *img = proj + ((size_t)b * P + p) * pw * ph;
                float *o = out + ((size_t)b * P + p) * nv;
                for (int j = 0; j < w; ++j)
                    for (int k = 0; k < h; ++k) {
                        float gu, gv;
                        backproj_point(poses + 3 * p, i, j, k, d, w, h, pw, ph, &gu, &gv);
                        /* grid_sample 2-D: x<-gv (W=ph), y<-gu (H=pw) */
                        o[((size_t)i * w + j) * h + k] = sample2(img, pw, ph, gv, gu);
                    }
            }
}

/* grad_proj (B,P,pw,ph) += adjoint (double accumulation); caller zeroes */
LRO_API void lro_backproject_backward(const float *grad_out, const float *poses, int B, int P, int pw, int ph,
                                      int d, int w, int h, float *grad_proj) {
    size_t nv = (size_t)d * w * h, np_ = (size_t)pw * ph;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int p = 0; p < P; ++p) {
            double *acc = (double *)calloc(np_, sizeof(double));
            const float *go = grad_out + ((size_t)b * P + p) * nv;
            for (int i = 0; i < d; ++i)
                for (int j = 0; j < w; ++j)
                    for (int k = 0; k < h; ++k) {
                        float gu, gv;
                        backproj_point(poses + 3 * p, i, j, k, d, w, h, pw, ph, &gu, &gv);
                        float ix = unnorm2(gv, ph), iy = unnorm2(gu, pw);
                        float xw = floorf(ix), yn = floorf(iy);
                        if (!(xw > -4.0f && xw < (float)ph + 4.0f && yn > -4.0f && yn < (float)pw + 4.0f)) continue;
                        float wq = ix - xw, e = 1.0f - wq, n = iy - yn, s = 1.0f - n;
                        int64_t x0 = (int64_t)xw, y0 = (int64_t)yn;
                        float g = go[((size_t)i * w + j) * h + k];
                        float ws[4] = {s * e, s * wq, n * e, n * wq};
                        for (int c = 0; c < 4; ++c) {
                            int64_t xx = x0 + (c & 1), yy = y0 + (c >> 1);
                            if (yy >= 0 && yy < pw && xx >= 0 && xx < ph) acc[(size_t)yy * ph + xx] += (double)(ws[c] * g);
                        }
                    }
            float *gp = grad_proj + ((size_t)b * P + p) * np_;
            for (size_t q = 0; q < np_; ++q) gp[q] += (float)acc[q];
            free(acc);
        }
}

/* ------------------------------------------------------------------ */
/* Warp: net_utils.py:9-56 Bilinear                                   */
/* ------------------------------------------------------------------ */

/* img (B,C,D,H,W); phi (B,3,D,H,W) channel c <-> volume axis c, in [-1,1]; out (B,C,D,H,W)
 * padding 0 zeros / 1 border (net_utils.py:21); mode 0 bilinear / 1 nearest (:23);
 * using_scale: sample (img+1)/2 and return out*2-1 (:48-52). */
LRO_API void lro_warp_forward(const float *img, const float *phi, int B, int C, int D, int H, int W,
                              int padding, int mode, int using_scale, float *out) {
    size_t nv = (size_t)D * H * W;
    float *pre = NULL;
    const int fast = g_blend && mode == 0;
    if (using_scale && !fast) {
        pre = (float *)malloc((size_t)B * C * nv * sizeof(float));
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < (int64_t)((size_t)B * C * nv); ++i) pre[i] = (img[i] + 1.0f) / 2.0f;   /* :50 */
    }
    const float *src = pre ? pre : img;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int z = 0; z < D; ++z)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    size_t vox = ((size_t)z * H + y) * W + x;
                    const float *ph = phi + (size_t)b * 3 * nv;
                    /* forward_stn :27-30 reverses channels: grid x<-phi[2], y<-phi[1], z<-phi[0] */
                    float gx = ph[2 * nv + vox], gy = ph[nv + vox], gz = ph[vox];
                    if (fast) {   /* :50 and :52 cancel because the weights sum to 1; a skipped tap is intensity 0 = -1 */
                        for (int c = 0; c < C; ++c)
                            out[((size_t)b * C + c) * nv + vox] = sample3_fast(img + ((size_t)b * C + c) * nv, D, H, W, gx, gy, gz,
                                                                               padding, using_scale ? -1.0f : 0.0f);
                        continue;
                    }
                    for (int c = 0; c < C; ++c) {
                        float s = sample3(src + ((size_t)b * C + c) * nv, D, H, W, gx, gy, gz, padding, mode);
                        out[((size_t)b * C + c) * nv + vox] = using_scale ? s * 2.0f - 1.0f : s;   /* :52 */
                    }
                }
    free(pre);
}

/* Adjoint of lro_warp_forward (mode 0 only for grad_phi; nearest has zero grid gradient).
 * grad_img (B,C,D,H,W) nullable, += (double accumulation, caller zeroes);
 * grad_phi (B,3,D,H,W) nullable, written.  Follows ATen grid_sampler_3d_backward (CPU):
 * gix -= tnw_val*(y1-y)*(z1-z)*gOut ... ; grad_grid = (size-1)/2 * gi (0 where border-clipped). */
LRO_API void lro_warp_backward(const float *grad_out, const float *img, const float *phi, int B, int C,
                               int D, int H, int W, int padding, int mode, int using_scale,
                               float *grad_img, float *grad_phi) {
    size_t nv = (size_t)D * H * W;
    double *acc = grad_img ? (double *)calloc((size_t)B * C * nv, sizeof(double)) : NULL;
    for (int b = 0; b < B; ++b)
        for (size_t vox = 0; vox < nv; ++vox) {
            const float *ph = phi + (size_t)b * 3 * nv;
            float gx = ph[2 * nv + vox], gy = ph[nv + vox], gz = ph[vox];
            float ix = unnorm3(gx, W), iy = unnorm3(gy, H), iz = unnorm3(gz, D);
            float mx = (float)(W - 1) / 2.0f, my = (float)(H - 1) / 2.0f, mz = (float)(D - 1) / 2.0f;
            if (padding == 1) {
                /* clip_coordinates_set_grad */
                if (ix <= 0.0f) { ix = 0.0f; mx = 0.0f; } else if (ix >= (float)(W - 1)) { ix = (float)(W - 1); mx = 0.0f; }
                if (iy <= 0.0f) { iy = 0.0f; my = 0.0f; } else if (iy >= (float)(H - 1)) { iy = (float)(H - 1); my = 0.0f; }
                if (iz <= 0.0f) { iz = 0.0f; mz = 0.0f; } else if (iz >= (float)(D - 1)) { iz = (float)(D - 1); mz = 0.0f; }
            }
            double gix = 0, giy = 0, giz = 0;
            if (mode == 1) {
                int64_t xn = (int64_t)nearbyintf(ix), yn = (int64_t)nearbyintf(iy), zn = (int64_t)nearbyintf(iz);
                if (acc && inb3(zn, yn, xn, D, H, W))
                    for (int c = 0; c < C; ++c) {
                        float g = grad_out[((size_t)b * C + c) * nv + vox];
                        if (using_scale) g = g * 2.0f;
                        acc[((size_t)b * C + c) * nv + ((size_t)zn * H + yn) * W + xn] += (double)g;
                    }
            } else {
                int64_t x0 = (int64_t)floorf(ix), y0 = (int64_t)floorf(iy), z0 = (int64_t)floorf(iz);
                float wx[2] = {(float)(x0 + 1) - ix, ix - (float)x0};
                float wy[2] = {(float)(y0 + 1) - iy, iy - (float)y0};
                float wz[2] = {(float)(z0 + 1) - iz, iz - (float)z0};
                for (int c = 0; c < C; ++c) {
                    float g = grad_out[((size_t)b * C + c) * nv + vox];
                    if (using_scale) g = g * 2.0f;              /* d(out*2-1)/d(out) */
                    const float *src = img + ((size_t)b * C + c) * nv;
                    for (int t = 0; t < 8; ++t) {
                        int tx = t & 1, ty = (t >> 1) & 1, tz = t >> 2;
                        int64_t xx = x0 + tx, yy = y0 + ty, zz = z0 + tz;
                        if (!inb3(zz, yy, xx, D, H, W)) continue;
                        size_t o = ((size_t)zz * H + yy) * W + xx;
                        if (acc) acc[((size_t)b * C + c) * nv + o] += (double)(wx[tx] * wy[ty] * wz[tz] * g);
                        float val = src[o];
                        if (using_scale) val = (val + 1.0f) / 2.0f;
                        gix += (double)((tx ? 1.0f : -1.0f) * val * wy[ty] * wz[tz] * g);
                        giy += (double)((ty ? 1.0f : -1.0f) * val * wx[tx] * wz[tz] * g);
                        giz += (double)((tz ? 1.0f : -1.0f) * val * wx[tx] * wy[ty] * g);
                    }
                }
            }
            if (grad_phi) {
                float *gp = grad_phi + (size_t)b * 3 * nv;
                gp[2 * nv + vox] = mx * (float)gix;
                gp[nv + vox] = my * (float)giy;
                gp[vox] = mz * (float)giz;
            }
        }
    if (acc) {
        size_t n = (size_t)B * C * nv;
        for (size_t i = 0; i < n; ++i) {
            float g = (float)acc[i];
            grad_img[i] += using_scale ? g / 2.0f : g;          /* d((img+1)/2)/d(img) */
        }
        free(acc);
    }
}

/* net_utils.py:59-87: id[c] = idx_c * (1/(sz_c-1)) * 2 - 1.  `id[d] *= spacing[d]` multiplies a float32 array by a
 * numpy float64 scalar: under numpy >= 2 (NEP 50; this image) the product is formed in float64 and rounded to fp32;
 * under the reference's pinned numpy 1.21 (value-based casting) it is an fp32 product.  The two differ by <= 1 ulp;
 * this restatement follows the behaviour observed when the reference runs in this image (f64 product). */
LRO_API void lro_identity_map(int D, int H, int W, float *out) {
    int sz[3] = {D, H, W};
    size_t nv = (size_t)D * H * W;
    for (int c = 0; c < 3; ++c) {
        double sp = 1.0 / (double)(sz[c] - 1);
        for (int z = 0; z < D; ++z)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    int idx = c == 0 ? z : (c == 1 ? y : x);
                    float v = (float)((double)idx * sp);
                    out[c * nv + ((size_t)z * H + y) * W + x] = v * 2.0f - 1.0f;
                }
    }
}

/* sdct:6-9 calc_relative_atten_coef: mu = (max(HU,-1000)+1000)/1000*0.2, fp32 */
LRO_API void lro_atten_coef(const float *hu, int64_t n, float *mu) {
    for (int64_t i = 0; i < n; ++i) {
        float v = hu[i] < -1000.0f ? -1000.0f : hu[i];
        mu[i] = (v + 1000.0f) / 1000.0f * 0.2f;
    }
}

/* models/LiftRegDeformSubspaceBackproj.py:102 (+ :68 if add_identity):
 *   out[b,n] = (sum_k coefs[b,k]*basis[n,k]) + mean[n] (+ identity_map(n)),  fp32, one fmaf chain over k ascending.
 * F.linear's own accumulation order is cuBLAS / MKL internal; any order agrees with this one to ~1e-7 relative. */
LRO_API void lro_pca_decode(const float *coefs, const float *basis, const float *mean, int B, int K, int64_t N,
                            int add_identity, int D, int H, int W, float *out) {
    int sz[3] = {D, H, W};
    int64_t nvox = (int64_t)D * H * W;
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N; ++n) {
        float idv = 0.0f;
        if (add_identity) {
            int c = (int)(n / nvox);
            int64_t v = n - (int64_t)c * nvox;
            int z = (int)(v / ((int64_t)H * W)), y = (int)((v / W) % H), x = (int)(v % W);
            int idx = c == 0 ? z : (c == 1 ? y : x);
            float t = (float)((double)idx * (1.0 / (double)(sz[c] - 1)));
            idv = t * 2.0f - 1.0f;
        }
        for (int b = 0; b < B; ++b) {
            float acc = 0.0f;
            for (int k = 0; k < K; ++k) acc = fmaf(coefs[(size_t)b * K + k], basis[(size_t)n * K + k], acc);
            float o = acc + (mean ? mean[n] : 0.0f);
            if (add_identity) o = o + idv;
            out[(size_t)b * N + n] = o;
        }
    }
}

LRO_API int lro_version(void) { return 1; }
