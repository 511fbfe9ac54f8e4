// Does the fast-math reciprocal return exactly -1 for -1 (and exact results for other powers of two)?  The pair walk of the
// blend backward relies on it: an invalid pixel continues with alpha = 0, so T * (-1 / (0 - 1)) must be T bit for bit.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a --use_fast_math -o rcp_probe rcp_probe.cu
#include <cstdio>
#include <cstring>
__global__ void k(const float *x, float *y, int n) {
    const int i = threadIdx.x;
    if (i < n) y[i] = 1.f / (x[i] - 1.f);
}
int main() {
    const float h[8] = {0.f, -1.f, 3.f, -3.f, 0.5f, 1.5f, 0.999f, 1.f / 255.f};
    float *dx, *dy, r[8];
    cudaMalloc(&dx, sizeof h);
    cudaMalloc(&dy, sizeof h);
    cudaMemcpy(dx, h, sizeof h, cudaMemcpyHostToDevice);
    k<<<1, 32>>>(dx, dy, 8);
    cudaMemcpy(r, dy, sizeof r, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < 8; ++i) {
        unsigned b;
        memcpy(&b, &r[i], 4);
        printf("1/(%g - 1) = %.9g (0x%08x)\n", h[i], r[i], b);
    }
    if (r[0] != -1.f) bad = 1;
    printf(bad ? "RCP PROBE: NOT EXACT\n" : "RCP PROBE: exact for alpha = 0\n");
    return bad;
}
