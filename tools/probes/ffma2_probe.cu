// Probe for the next round: is the blend's per-pair arithmetic bit-identical and cheaper to
// issue when two pixels are evaluated with packed FP32 (FFMA2 / FMUL2 / FADD2, sm_100a)?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --use_fast_math -o ffma2_probe tools/probes/ffma2_probe.cu && ./ffma2_probe
//
// Kernel A evaluates sigma / alpha for ONE pixel per thread with the scalar expression of csrc/blend.cu; kernel B
// evaluates TWO vertically adjacent pixels per thread with the packed intrinsics, spelling out the contraction the
// compiler chose for the scalar form (visible in its SASS: t = c*dy; u = a*dx; v = b*dx; t = dy*t; u = fma(dx,u,t);
// u *= 0.5; sigma = fma(dy,v,u)).  The host compares every alpha bit for bit and times both kernels.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

constexpr int kGauss = 4096;  // Gaussians evaluated per pixel (staged in shared memory in chunks)

struct G { float x, y, opac, pad0, a, b, c, pad1; };

__global__ void scalar_kernel(const G *g, int W, int H, float *out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= W || i >= H) return;
    const float px = j + 0.5f, py = i + 0.5f;
    float acc = 0.f;
    for (int t = 0; t < kGauss; ++t) {
        const G q = g[t];
        const float dx = q.x - px, dy = q.y - py;
        const float sigma = 0.5f * (q.a * dx * dx + q.c * dy * dy) + q.b * dx * dy;
        const float alpha = fminf(0.999f, q.opac * __expf(-sigma));
        if (!(sigma < 0.f || alpha < 1.f / 255.f)) acc += alpha;
    }
    out[i * W + j] = acc;
}

__global__ void packed_kernel(const G *g, int W, int H, float *out) {  // rows i and i+1 per thread
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = 2 * blockIdx.y;
    if (j >= W || i >= H) return;
    const float2 px = make_float2(j + 0.5f, j + 0.5f), py = make_float2(i + 0.5f, i + 1.5f);
    const float2 half = make_float2(0.5f, 0.5f), nlog2e = make_float2(-1.4426950216293334961f, -1.4426950216293334961f);
    float2 acc = make_float2(0.f, 0.f);
    for (int t = 0; t < kGauss; ++t) {
        const G q = g[t];
        const float2 dx = __fadd2_rn(make_float2(q.x, q.x), make_float2(-px.x, -px.y));
        const float2 dy = __fadd2_rn(make_float2(q.y, q.y), make_float2(-py.x, -py.y));
        float2 tt = __fmul2_rn(make_float2(q.c, q.c), dy);
        float2 u = __fmul2_rn(make_float2(q.a, q.a), dx);
        const float2 v = __fmul2_rn(make_float2(q.b, q.b), dx);
        tt = __fmul2_rn(dy, tt);
        u = __ffma2_rn(dx, u, tt);
        u = __fmul2_rn(u, half);
        const float2 sigma = __ffma2_rn(dy, v, u);
        const float2 e = __fmul2_rn(sigma, nlog2e);
        const float2 al = __fmul2_rn(make_float2(q.opac, q.opac), make_float2(exp2f(e.x), exp2f(e.y)));
        const float a0 = fminf(0.999f, al.x), a1 = fminf(0.999f, al.y);
        if (!(sigma.x < 0.f || a0 < 1.f / 255.f)) acc.x += a0;
        if (!(sigma.y < 0.f || a1 < 1.f / 255.f)) acc.y += a1;
    }
    out[i * W + j] = acc.x;
    if (i + 1 < H) out[(i + 1) * W + j] = acc.y;
}

int main() {
    const int W = 1920, H = 1080;
    G *h = (G *)malloc(sizeof(G) * kGauss), *d;
    srand(1);
    for (int t = 0; t < kGauss; ++t) {
        const float s = 4.f + 60.f * rand() / RAND_MAX;
        h[t] = {W * (float)rand() / RAND_MAX, H * (float)rand() / RAND_MAX, (float)rand() / RAND_MAX, 0.f,
                1.f / (s * s), 0.3f / (s * s) * ((float)rand() / RAND_MAX - 0.5f), 1.3f / (s * s), 0.f};
    }
    float *o1, *o2;
    cudaMalloc(&d, sizeof(G) * kGauss); cudaMalloc(&o1, sizeof(float) * W * H); cudaMalloc(&o2, sizeof(float) * W * H);
    cudaMemcpy(d, h, sizeof(G) * kGauss, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms1 = 0, ms2 = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); scalar_kernel<<<dim3((W + 127) / 128, H), 128>>>(d, W, H, o1); cudaEventRecord(e1);
        cudaEventSynchronize(e1); cudaEventElapsedTime(&ms1, e0, e1);
        cudaEventRecord(e0); packed_kernel<<<dim3((W + 127) / 128, (H + 1) / 2), 128>>>(d, W, H, o2); cudaEventRecord(e1);
        cudaEventSynchronize(e1); cudaEventElapsedTime(&ms2, e0, e1);
    }
    float *a = (float *)malloc(sizeof(float) * W * H), *b = (float *)malloc(sizeof(float) * W * H);
    cudaMemcpy(a, o1, sizeof(float) * W * H, cudaMemcpyDeviceToHost); cudaMemcpy(b, o2, sizeof(float) * W * H, cudaMemcpyDeviceToHost);
    long diff = 0;
    for (long k = 0; k < (long)W * H; ++k) diff += memcmp(a + k, b + k, 4) != 0;
    printf("scalar %.3f ms, packed (2 px/thread) %.3f ms, %.2fx; pixels whose accumulated alpha differs bitwise: %ld of %d\n", ms1, ms2,
           ms1 / ms2, diff, W * H);
    return cudaGetLastError() != cudaSuccess;
}
