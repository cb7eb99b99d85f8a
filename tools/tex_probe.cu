// Experiment (not part of the library): can the texture units fetch the DRR's 2x2 footprints faster than scalar loads?
// tex2Dgather on a cudaArray (cudaArrayTextureGather) returns the four texels of a bilinear footprint, unfiltered, in
// one instruction.  Measures (a) that the footprint of (x0+1, row+1) is exactly texels (x0..x0+1, row..row+1),
// (b) gathers/s for a DRR-like access pattern, (c) the same pattern with 16 scalar LDGs from linear memory.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/tex_probe tools/tex_probe.cu && tools/tex_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int W = 160, H = 160 * 160;

__global__ void verify_kernel(cudaTextureObject_t tex, const float *lin, int *bad) {
    const int x0 = threadIdx.x % (W - 1), row = blockIdx.x * 37 % (H - 1);
    const float4 g = tex2Dgather<float4>(tex, (float)x0 + 1.0f, (float)row + 1.0f, 0);
    const float a = lin[row * W + x0], b = lin[row * W + x0 + 1], c = lin[(row + 1) * W + x0], d = lin[(row + 1) * W + x0 + 1];
    if (g.w != a || g.z != b || g.x != c || g.y != d) atomicAdd(bad, 1);     // (i,j+1) (i+1,j+1) (i+1,j) (i,j)
    // border: footprint partly outside returns 0 there
    const float4 e = tex2Dgather<float4>(tex, 0.0f, (float)row + 1.0f, 0);  // texels x = -1, 0
    if (e.w != 0.0f || e.x != 0.0f || e.z != lin[row * W] || e.y != lin[(row + 1) * W]) atomicAdd(bad, 1);
}

__global__ void __launch_bounds__(256, 4) tex_kernel(cudaTextureObject_t tex, int iters, float *sink) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * 8 + (threadIdx.x >> 5));
    float xa = 1.0f + 0.9f * lane + 0.01f * warp, ya = (float)((warp * 157) % (H - 800)) + 1.0f;
    float acc = 0.0f;
    for (int it = 0; it < iters; ++it) {
        const float x0 = floorf(xa) + 1.0f, x1 = floorf(xa + 0.4f) + 1.0f;
        const float4 g0 = tex2Dgather<float4>(tex, x0, ya, 0), g1 = tex2Dgather<float4>(tex, x0, ya + 160.0f, 0);
        const float4 g2 = tex2Dgather<float4>(tex, x1, ya + 1.0f, 0), g3 = tex2Dgather<float4>(tex, x1, ya + 161.0f, 0);
        acc += (g0.x + g0.y + g0.z + g0.w) + (g1.x + g1.y + g1.z + g1.w) + (g2.x + g2.y + g2.z + g2.w) + (g3.x + g3.y + g3.z + g3.w);
        xa += 0.27f; ya += 1.0f;
        if (xa > 120.0f) xa -= 100.0f;
    }
    if (acc == 123.456f) sink[0] = acc;
}

__global__ void __launch_bounds__(256, 4) ldg_kernel(const float *__restrict__ lin, int iters, float *sink) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * 8 + (threadIdx.x >> 5));
    float xa = 1.0f + 0.9f * lane + 0.01f * warp;
    int ya = (warp * 157) % (H - 800);
    float acc = 0.0f;
    for (int it = 0; it < iters; ++it) {
        const int x0 = (int)floorf(xa), x1 = (int)floorf(xa + 0.4f);
        const float *p0 = lin + ya * W + x0, *p1 = p0 + 160 * W, *p2 = lin + (ya + 1) * W + x1, *p3 = p2 + 160 * W;
        acc += (__ldg(p0) + __ldg(p0 + 1) + __ldg(p0 + W) + __ldg(p0 + W + 1)) + (__ldg(p1) + __ldg(p1 + 1) + __ldg(p1 + W) + __ldg(p1 + W + 1))
             + (__ldg(p2) + __ldg(p2 + 1) + __ldg(p2 + W) + __ldg(p2 + W + 1)) + (__ldg(p3) + __ldg(p3 + 1) + __ldg(p3 + W) + __ldg(p3 + W + 1));
        xa += 0.27f; ya += 1;
        if (xa > 120.0f) xa -= 100.0f;
    }
    if (acc == 123.456f) sink[0] = acc;
}

int main() {
    float *lin, *sink; int *bad;
    CK(cudaMalloc(&lin, sizeof(float) * W * H)); CK(cudaMalloc(&sink, 4)); CK(cudaMalloc(&bad, 4)); CK(cudaMemset(bad, 0, 4));
    float *h = (float *)malloc(sizeof(float) * W * H);
    for (int i = 0; i < W * H; ++i) h[i] = (float)(i % 9973) * 0.25f + 1.0f;
    CK(cudaMemcpy(lin, h, sizeof(float) * W * H, cudaMemcpyHostToDevice));
    cudaChannelFormatDesc cd = cudaCreateChannelDesc<float>();
    cudaArray_t arr;
    CK(cudaMallocArray(&arr, &cd, W, H, cudaArrayTextureGather));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    CK(cudaMemcpy2DToArrayAsync(arr, 0, 0, lin, sizeof(float) * W, sizeof(float) * W, H, cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("linear -> array copy of %.1f MB: %.1f us\n", W * H * 4e-6, ms * 1e3);
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
    cudaTextureDesc td = {}; td.filterMode = cudaFilterModePoint; td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder;
    td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
    cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    verify_kernel<<<512, 256>>>(tex, lin, bad); CK(cudaDeviceSynchronize());
    int hb; CK(cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost));
    printf("footprint / border check: %d mismatches\n", hb);
    const int blocks = 148 * 4, iters = 400;
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0)); tex_kernel<<<blocks, 256>>>(tex, iters, sink); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double gathers = (double)blocks * 256 * iters * 4;
        printf("tex2Dgather: %.1f us, %.1f G lane-gathers/s = %.2f per clk per SM at 1.9 GHz (%.1f G texels/s)\n", ms * 1e3, gathers / ms * 1e-6,
               gathers / (ms * 1e-3) / 148 / 1.9e9, 4 * gathers / ms * 1e-6);
        CK(cudaEventRecord(e0)); ldg_kernel<<<blocks, 256>>>(lin, iters, sink); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("16 x LDG:    %.1f us, %.1f G footprints/s (%.1f G loads/s)\n", ms * 1e3, gathers / ms * 1e-6, 4 * gathers / ms * 1e-6);
    }
    return 0;
}
