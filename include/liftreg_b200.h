/*
 * liftreg_b200.h -- C-ABI of the B200-native LiftReg resampling library
 * (libliftreg_b200.so, built from liftreg_b200/csrc/ for sm_100a).
 *
 * The reference (uncbiag/LiftReg) is pure Python: its "operator API" for this
 * path is the Python surface of
 *     src/liftreg/utils/sdct_projection_utils.py   ("sdct" below)
 *     src/liftreg/utils/net_utils.py               (Bilinear, identity_map)
 *     src/liftreg/layers/layers.py                 (proj_layer)
 *     src/liftreg/models/LiftRegDeformSubspaceBackproj.py:85-93,68-69
 * and every one of those bottoms out in torch.nn.functional.grid_sample.
 * Each entry point below replaces the cited reference lines; the Python mirror
 * in liftreg_b200/ binds them with ctypes (INTEGRATION.md shows the stub a
 * reference maintainer would add).
 *
 * Conventions
 *   - plain C types only; no torch / C++ types cross this boundary.
 *   - every buffer is owned by the caller; the library never allocates or frees
 *     device memory and keeps no state (geometry is passed by value per call).
 *   - all tensors are dense, C-contiguous fp32 unless a stride is given.
 *   - functions without the _host suffix take DEVICE pointers, are asynchronous
 *     on `stream` (a cudaStream_t) and never synchronise.  As with any CUDA
 *     launch, the calling thread's current device must be the device that owns
 *     `stream` and the buffers (the current device is per host thread).
 *   - _host functions take HOST pointers (pinned for full speed) plus a device
 *     workspace of lr_*_workspace_bytes(); they enqueue H2D, kernels and D2H on
 *     `stream` and synchronise it before returning, like the reference calls
 *     they replace (which end in .cpu().numpy()).  When the host buffers are
 *     pinned+mapped, lr_warp_forward_host lets the kernel stream the map and the
 *     result over PCIe itself (set LIFTREG_B200_ZERO_COPY=0 to force staging).
 *   - return 0 on success, a negative lr_status on failure; never throws or
 *     exits.  lr_last_error() returns a thread-local message.
 *   - volume axes (d,w,h) = (axial, coronal, sagittal); detector (rd,rh);
 *     emitter poses in voxel units, detector plane y = 0 (sdct:59-68).
 */
#ifndef LIFTREG_B200_H
#define LIFTREG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define LR_API __attribute__((visibility("default")))
#else
#define LR_API
#endif

typedef void *lr_stream_t; /* cudaStream_t */

typedef enum {
    LR_OK = 0,
    LR_ERR_BAD_ARGUMENT = -1,
    LR_ERR_CUDA = -2,
    LR_ERR_WORKSPACE = -3,
    LR_ERR_NO_DEVICE = -4
} lr_status;

enum { LR_PAD_ZEROS = 0, LR_PAD_BORDER = 1 };     /* net_utils.py:21  zero_boundary ? zeros : border */
enum { LR_MODE_LINEAR = 0, LR_MODE_NEAREST = 1 }; /* net_utils.py:23  mode = "bilinear" | "nearest"  */
enum { LR_YNORM_WM1 = 0, LR_YNORM_W = 1 };        /* sdct:55 (/(w-1))  vs  layers.py:234 (/w)        */
/* How the interpolation BLEND of lr_backproject_forward and lr_warp_forward (linear mode) is rounded.  Sample
 * coordinates, floor indices and interpolation weights replay the reference's fp32 op chain bit for bit in both modes.
 *   LR_NUMERICS_FAST  (default) the blend is a chain of fused linear interpolations (separable rows-then-columns for
 *                     the backprojection, x-y-z lerps for the warp; `using_scale` folds away where all taps are inside):
 *                     <= 1e-6 relative L2 from ATen's result (BASELINE.json's tolerance is 1e-5), half the instructions.
 *   LR_NUMERICS_EXACT ATen's own operation order (grid_sampler_2d vector kernel / grid_sampler_3d scalar kernel):
 *                     results bit-identical to the reference's torch CPU path.                                    */
enum { LR_NUMERICS_FAST = 0, LR_NUMERICS_EXACT = 1 };

/* ---- library ---------------------------------------------------------- */
LR_API int lr_abi_version(void);
LR_API const char *lr_last_error(void);
/* process-wide numerics mode (also: environment LIFTREG_B200_NUMERICS=exact|fast read at first use) */
LR_API int lr_set_numerics(int mode);
LR_API int lr_get_numerics(void);
/* number of CUDA devices visible, or a negative lr_status */
LR_API int lr_device_count(void);
/* kernels launched by this library (all threads) since the last reset (bench.py's gpu_launches) */
LR_API long long lr_launch_count(void);
LR_API void lr_launch_count_reset(void);

/* ---- cone-beam DRR ---------------------------------------------------- */
/* replaces sdct:59-86 calculate_projection (grid build sdct:15-57, flip :76, grid_sample+sum+*dx :81, *0.1 :85)
 * and layers.py:182-187 proj_layer.forward (y_norm_mode = LR_YNORM_W, out_scale = 1).
 *   vol   (B,d,w,h)       attenuation volume(s)
 *   poses (n_pose_sets,P,3) HOST pointer, float64, voxel units; n_pose_sets is 1 (shared by all B, as in
 *         proj_layer) or B (per item, models/previous/RegNet2D3D.py:161-171)
 *   proj  (B,P,rd,rh) = out_scale * dx[p,u,v] * sum_j trilinear_zeros(vol[b], ray point j)   */
LR_API int lr_drr_forward(const float *vol, int B, int d, int w, int h,
                          const double *poses, int n_pose_sets, int P, int rd, int rh,
                          const float spacing[3], int y_norm_mode, float out_scale,
                          float *proj, lr_stream_t stream);

/* Multi-GPU form of lr_drr_forward for view-sharded sweeps (BASELINE.json north star: "forward projection by view angle
 * ..., all-gather of the detector images over NVLink"; the reference itself is single-GPU, main.py:108-110).  The kernel
 * stores every detector pixel into n_outs buffers -- this rank's gather buffer and the peers' (device pointers obtained
 * with lr_peer_open, written as P2P stores over NVLink) -- so the images need no collective afterwards, only a barrier.
 * View k of the call (k = b*P + p) lands at outs[i] + k * view_stride * rd * rh: with view_stride = world size and
 * outs[i] pointing at this rank's first view, a rank that owns the views r, r+world, ... fills the full (n_views,rd,rh)
 * array in view order on every rank. */
#define LR_MAX_PEERS 8
LR_API int lr_drr_forward_peers(const float *vol, int B, int d, int w, int h,
                                const double *poses, int n_pose_sets, int P, int rd, int rh,
                                const float spacing[3], int y_norm_mode, float out_scale,
                                float *const *outs, int n_outs, int view_stride, lr_stream_t stream);
/* Peer-visible device buffers for the call above (CUDA IPC, one process per GPU on one box).
 * lr_peer_alloc: cudaMalloc on the current device + its 64-byte IPC handle (to be sent to the other ranks by any means).
 * lr_peer_open: maps another rank's buffer into this process (peer access enabled lazily); lr_peer_close unmaps it.
 * lr_peer_free releases a buffer obtained from lr_peer_alloc (after the peers closed it). */
#define LR_IPC_HANDLE_BYTES 64
LR_API int lr_peer_alloc(size_t bytes, void **dev_ptr, unsigned char handle[LR_IPC_HANDLE_BYTES]);
LR_API int lr_peer_open(const unsigned char handle[LR_IPC_HANDLE_BYTES], void **dev_ptr);
LR_API int lr_peer_close(void *dev_ptr);
LR_API int lr_peer_free(void *dev_ptr);

/* adjoint of lr_drr_forward wrt vol (autograd of sdct:81 / layers.py:187; RegNet2D3D.py:161-185 needs it).
 * grad_vol (B,d,w,h) is ACCUMULATED into (caller zero-initialises). */
LR_API int lr_drr_backward(const float *grad_proj, int B, int d, int w, int h,
                           const double *poses, int n_pose_sets, int P, int rd, int rh,
                           const float spacing[3], int y_norm_mode, float out_scale,
                           float *grad_vol, lr_stream_t stream);

/* replaces sdct:15-57 project_grid_multi / layers.py:194-236 (API parity and bit-exactness checks; the DRR
 * kernels never materialise this).  grid (P,rd,rh,w,3) nullable, dx (P,rd,rh) nullable.
 * flip != 0 writes the last axis reversed, as sdct:76 / forward_grids (sdct:223,263) return it. */
LR_API int lr_project_grid(const double *poses, int P, int rd, int rh, int d, int w, int h,
                           const float spacing[3], int y_norm_mode, int flip,
                           float *grid, float *dx, lr_stream_t stream);

/* host-buffer form of lr_drr_forward == calculate_projection's numpy-in / numpy-out contract (sdct:59-100). */
LR_API size_t lr_drr_forward_host_workspace_bytes(int B, int d, int w, int h, int P, int rd, int rh);
LR_API int lr_drr_forward_host(const float *vol_host, int B, int d, int w, int h,
                               const double *poses, int n_pose_sets, int P, int rd, int rh,
                               const float spacing[3], int y_norm_mode, float out_scale,
                               float *proj_host, void *workspace, size_t workspace_bytes, lr_stream_t stream);

/* ---- backprojection ---------------------------------------------------- */
/* replaces sdct:227-250 backproj_grids_with_poses + the grid_sample block at
 * models/LiftRegDeformSubspaceBackproj.py:85-93 (same block: models/previous/RegNet2D3D.py:105-112).
 *   proj  (B,P,pw,ph)
 *   poses (P,3) HOST pointer, float32 (Registration2D3DDataset.py:121 casts to fp32; geometry frozen from
 *         batch item 0, model :85-87)
 *   out[b*out_batch_stride + p*out_chan_stride + (i*w+j)*h+k] = bilinear_zeros(proj[b,p], u(p,i,j), v(p,j,k))
 *   strides are in elements; a dense (B,P,d,w,h) output uses P*d*w*h and d*w*h.  They let the caller aim at
 *   channels 1..P of the (B,1+P,d,w,h) encoder input and skip torch.cat (model :95-98). */
LR_API int lr_backproject_forward(const float *proj, const float *poses, int B, int P, int pw, int ph,
                                  int d, int w, int h, float *out,
                                  int64_t out_batch_stride, int64_t out_chan_stride, lr_stream_t stream);

/* z-slab form for multi-GPU sharding along axis 0: `out` holds planes [i_begin, i_begin+i_count)
 * only, i.e. a (B,P,i_count,w,h) tensor with the given strides; projections (<= 1 MB per item) are replicated on
 * every rank, so the op needs no halo and no collective. */
LR_API int lr_backproject_forward_slab(const float *proj, const float *poses, int B, int P, int pw, int ph,
                                       int d_total, int w, int h, int i_begin, int i_count, float *out,
                                       int64_t out_batch_stride, int64_t out_chan_stride, lr_stream_t stream);

/* Geometry plan.  Everything lr_backproject_forward derives from (poses, shapes) -- floor detector rows / columns and
 * interpolation weights of every plane and voxel column, per view and coronal row -- can be evaluated ONCE into a
 * caller-owned device buffer, exactly as the reference caches its (1,P,2,d,w,h) sample grid on the first batch
 * (LiftRegDeformSubspaceBackproj.py:85-87; 131 MB there, 3 MB here).  lr_backproject_forward_planned then only gathers,
 * blends and stores (fast numerics: the separable blend of LR_NUMERICS_FAST, whatever lr_set_numerics says; indices and
 * weights are the same bit-exact chain).  The plan depends on (poses, P, pw, ph, d_total, w, h) only: any batch size,
 * any z-slab [i_begin, i_begin+i_count) and any output strides may use it.  plan must be 16-byte aligned. */
LR_API size_t lr_backproject_plan_bytes(int P, int pw, int ph, int d_total, int w, int h);
LR_API int lr_backproject_plan_build(const float *poses, int P, int pw, int ph, int d_total, int w, int h,
                                     void *plan, size_t plan_bytes, lr_stream_t stream);
LR_API int lr_backproject_forward_planned(const float *proj, const void *plan, int B, int P, int pw, int ph,
                                          int d_total, int w, int h, int i_begin, int i_count, float *out,
                                          int64_t out_batch_stride, int64_t out_chan_stride, lr_stream_t stream);

/* adjoint wrt proj; grad_proj (B,P,pw,ph) is ACCUMULATED into (caller zero-initialises). */
LR_API int lr_backproject_backward(const float *grad_out, int64_t go_batch_stride, int64_t go_chan_stride,
                                   const float *poses, int B, int P, int pw, int ph, int d, int w, int h,
                                   float *grad_proj, lr_stream_t stream);

/* replaces sdct:227-250 itself: grid (P,2,d,w,h), channel 0 = detector axis 1 (ph), channel 1 = axis 0 (pw). */
LR_API int lr_backproj_grid(const float *poses, int P, int d, int w, int h, int pw, int ph,
                            float *grid, lr_stream_t stream);

LR_API size_t lr_backproject_forward_host_workspace_bytes(int B, int P, int pw, int ph, int d, int w, int h);
LR_API int lr_backproject_forward_host(const float *proj_host, const float *poses, int B, int P, int pw, int ph,
                                       int d, int w, int h, float *out_host,
                                       void *workspace, size_t workspace_bytes, lr_stream_t stream);
/* Overlapped form of the host-buffer calls: the *_host_async entry points enqueue H2D, kernels and D2H on `stream` and
 * return WITHOUT synchronising, so one host thread can put the backprojection on one stream and the warp on another and
 * the two transfers share the full-duplex PCIe link (the reference's calls, sdct:70-72,97-99, are serial and blocking).
 * Host buffers must be pinned for the copies to be asynchronous and must stay alive until the stream is synchronised:
 * lr_stream_synchronize(stream) (for callers without a CUDA runtime binding) or cudaStreamSynchronize. */
LR_API int lr_backproject_forward_host_async(const float *proj_host, const float *poses, int B, int P, int pw, int ph,
                                             int d, int w, int h, float *out_host,
                                             void *workspace, size_t workspace_bytes, lr_stream_t stream);
LR_API int lr_stream_synchronize(lr_stream_t stream);

/* ---- displacement-field warp ------------------------------------------- */
/* replaces net_utils.py:26-56 Bilinear.forward / forward_stn.
 *   img (B,C,D,H,W); phi (B,3,D,H,W), channel c addresses volume axis c in [-1,1] (the channel reversal of
 *   :27-30 is folded into the addressing); out (B,C,D,H,W).
 *   using_scale: sample (img+1)/2 and return 2*out-1 (:48-52), fused.
 *   disp_plus_identity != 0: phi holds the DISPLACEMENT and the normalised identity map
 *   (net_utils.py:59-87) is added in-kernel, i.e. the `disp_field + self.id_transform` of model :68 is fused;
 *   the sum is rounded to fp32 exactly as the torch add would. */
LR_API int lr_warp_forward(const float *img, const float *phi, int B, int C, int D, int H, int W,
                           int padding, int mode, int using_scale, int disp_plus_identity,
                           float *out, lr_stream_t stream);

/* adjoint (autograd of grid_sample at net_utils.py:32-35 plus the (x+1)/2 and *2-1 factors).
 * grad_img (B,C,D,H,W) nullable, ACCUMULATED into (caller zero-initialises);
 * grad_phi (B,3,D,H,W) nullable, overwritten (zero for LR_MODE_NEAREST). */
LR_API int lr_warp_backward(const float *grad_out, const float *img, const float *phi,
                            int B, int C, int D, int H, int W,
                            int padding, int mode, int using_scale, int disp_plus_identity,
                            float *grad_img, float *grad_phi, lr_stream_t stream);

/* z-slab forms for multi-GPU sharding along axis 0: phi / out / grad_out / grad_phi are
 * (B,*,z_count,H,W) tensors holding output planes [z_begin, z_begin+z_count); img (and grad_img) stay the full
 * (B,C,D,H,W) volume, which every rank holds (16.4 MB at 160^3), so no halo exchange and no reduction is needed. */
LR_API int lr_warp_forward_slab(const float *img, const float *phi, int B, int C, int D, int H, int W,
                                int z_begin, int z_count, int padding, int mode, int using_scale,
                                int disp_plus_identity, float *out, lr_stream_t stream);
LR_API int lr_warp_backward_slab(const float *grad_out, const float *img, const float *phi,
                                 int B, int C, int D, int H, int W, int z_begin, int z_count,
                                 int padding, int mode, int using_scale, int disp_plus_identity,
                                 float *grad_img, float *grad_phi, lr_stream_t stream);

/* replaces net_utils.py:59-87 identity_map: out (3,D,H,W) */
LR_API int lr_identity_map(int D, int H, int W, float *out, lr_stream_t stream);

LR_API size_t lr_warp_forward_host_workspace_bytes(int B, int C, int D, int H, int W);
LR_API int lr_warp_forward_host(const float *img_host, const float *phi_host, int B, int C, int D, int H, int W,
                                int padding, int mode, int using_scale, int disp_plus_identity, float *out_host,
                                void *workspace, size_t workspace_bytes, lr_stream_t stream);
LR_API int lr_warp_forward_host_async(const float *img_host, const float *phi_host, int B, int C, int D, int H, int W,
                                      int padding, int mode, int using_scale, int disp_plus_identity, float *out_host,
                                      void *workspace, size_t workspace_bytes, lr_stream_t stream);

/* ---- PCA-subspace displacement decode ------------- */
/* replaces models/LiftRegDeformSubspaceBackproj.py:102
 *     disp_field = F.linear(x, self.pca_vectors, self.pca_mean)          (x: (B,K) coefficients)
 * and, if add_identity != 0, the `disp_field + self.id_transform` of :68 (N must then be 3*D*H*W; the (B,N) result
 * reshaped to (B,3,D,H,W) is the map phi that lr_warp_forward consumes).
 *   coefs (B,K); basis (N,K) row-major = pca_vectors as the model stores it (:42); K <= 160; the fast path needs
 *   K % 4 == 0 and a 16-byte aligned basis (anything else takes a scalar staging path);
 *   mean (N) nullable; out (B,N).
 * out[b,n] = (sum_k coefs[b,k]*basis[n,k], fp32 FMA chain, k ascending) + mean[n] (+ identity(n)).
 * The 4*N*K-byte basis (2.75 GB at 160^3, K = 56) is streamed from HBM exactly once for up to 32 batch items. */
LR_API int lr_pca_decode(const float *coefs, const float *basis, const float *mean, int B, int K, int64_t N,
                         int add_identity, int D, int H, int W, float *out, lr_stream_t stream);

/* adjoint wrt the coefficients (autograd of F.linear at model :102): grad_coefs (B,K) is ACCUMULATED into (caller
 * zero-initialises): grad_coefs[b,k] += sum_n grad_out[b,n] * basis[n,k].  K <= 160; K % 4 == 0 with 16-byte aligned
 * basis / grad_out takes the TMA-pipelined path, anything else a scalar staging path. */
LR_API int lr_pca_decode_backward(const float *grad_out, const float *basis, int B, int K, int64_t N,
                                  float *grad_coefs, lr_stream_t stream);

/* ---- HU -> attenuation -------------------------------------------------- */
/* replaces sdct:6-13 calc_relative_atten_coef(_cuda): mu = (max(HU,-1000)+1000)/1000*0.2; in place allowed */
LR_API int lr_atten_coef(const float *hu, int64_t n, float *mu, lr_stream_t stream);

/* ---- launch plans (diagnostics for host-side tests; no device work) ------- */
/* how lr_warp_forward cuts the z_count planes of an item into z-blocks: plan = {size0, n0, size1, n1, size2, n2} */
LR_API int lr_warp_forward_plan(int B, int D, int H, int W, int z_count, int plan[6]);
/* how lr_backproject_forward tiles planes, columns and rows: plan = {ichunk, isub, by, bx, n_chunks, run0, n0, run1, n1,
 * run2, n2, grid} */
LR_API int lr_backproject_forward_plan(int B, int P, int pw, int ph, int d, int w, int h, int plan[12]);

/* ---- measurement probes (diagnostics; bench.py's DRR fractions) -------------- */
/* lr_probe_l1_gather: every block re-reads its own floats_per_block-float (power of two >= 2048) slice of buf `iters`
 * times with coalesced 32-bit loads that hit L1: bytes moved = blocks*256*iters*32.  Timed by the caller, it gives the
 * aggregate L1 gather bandwidth that SURVEY.md 8d names as the DRR kernel's binding resource.
 * lr_probe_issue: blocks*256 threads each issue iters*8 independent FFMAs: the chip's sustained issue rate. */
LR_API int lr_probe_l1_gather(const float *buf, int64_t buf_floats, int blocks, int floats_per_block, int iters,
                              float *sink, lr_stream_t stream);
LR_API int lr_probe_issue(int blocks, int iters, float *sink, lr_stream_t stream);
/* lr_probe_issue_packed: the same with packed fp32x2 FMAs (blocks*256 threads, iters*8 FFMA2 each). */
LR_API int lr_probe_issue_packed(int blocks, int iters, float *sink, lr_stream_t stream);

/* ---- similarity loss on the warp output -------------- */
/* replaces src/liftreg/layers/losses.py:14-29 NCCLoss.forward (training similarity, SubspaceLoss.py:27; validation score,
 * RegistrationNet.py:210-212):  a = x - mean(x) + 1e-10, b = y - mean(y) + 1e-10 per batch item over N voxels,
 * ncc = mean(a*b) / sqrt(mean(a*a) * mean(b*b)).
 * lr_ncc_sums (one pass over x and y) fills sums (B,7) float64, device memory owned by the caller (zeroed by the call):
 *   [sum x, sum y, sum a*b, sum a*a, sum b*b, sum a, sum b]  ->  ncc_b = s2 / sqrt(s3*s4); loss = 1 - mean_b ncc_b.
 * lr_ncc_backward: grad_x (B,N) = d(loss)/dx * grad_loss[0] (grad_loss: one float in device memory), from the same sums. */
LR_API int lr_ncc_sums(const float *x, const float *y, int B, int64_t N, double *sums, lr_stream_t stream);
LR_API int lr_ncc_backward(const float *x, const float *y, int B, int64_t N, const double *sums, const float *grad_loss,
                           float *grad_x, lr_stream_t stream);

/* ---- displacement regulariser of the subspace loss ---- */
/* replaces src/liftreg/losses/SubspaceLoss.py:51-67 compute_reg_loss (same code: losses/RegNet2D3DLoss.py:53-69):
 *   fd = mermaid.finite_differences.FD_torch(spacing*2), spacing = 1/(shape-1);
 *   reg = mean over (B, D, H, W) of sum_{c<3} dXc(disp_c)^2 + dYc(disp_c)^2 + dZc(disp_c)^2
 * disp (B,3,D,H,W) float32, dense.  mermaid (third party, requirements.txt:61) is not part of the reference tree; its
 * central difference is (I[i+1]-I[i-1]) * 0.5/spacing in the interior, and on the faces either a linear extrapolation
 * of the missing neighbour (LR_FD_LINEAR, FD_torch's default mode) or zero (LR_FD_NEUMANN_ZERO).
 * lr_diffusion_reg_sum: *sum (one float64 in device memory, zeroed by the call) = sum over every voxel of the nine
 *   squares; reg = *sum / (B*D*H*W).
 * lr_diffusion_reg_backward: grad_disp (B,3,D,H,W) = d(reg)/d(disp) * grad_loss[0] (one float in device memory). */
#define LR_FD_LINEAR 0
#define LR_FD_NEUMANN_ZERO 1
LR_API int lr_diffusion_reg_sum(const float *disp, int B, int D, int H, int W, int boundary, double *sum, lr_stream_t stream);
LR_API int lr_diffusion_reg_backward(const float *disp, int B, int D, int H, int W, int boundary, const float *grad_loss,
                                     float *grad_disp, lr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LIFTREG_B200_H */
