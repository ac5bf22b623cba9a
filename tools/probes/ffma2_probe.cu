// Probe: issue rate of scalar FFMA vs packed FFMA2 / FMUL2+FADD2 on sm_100a (build: nvcc -arch=sm_100a -O3).
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b, int iters) {
    float2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { acc[i].x = fmaf(acc[i].x, a, b); acc[i].y = fmaf(acc[i].y, a, b); }
            if (MODE == 1) acc[i] = __ffma2_rn(acc[i], a2, b2);
            if (MODE == 2) { acc[i].x = __fadd_rn(__fmul_rn(acc[i].x, a), b); acc[i].y = __fadd_rn(__fmul_rn(acc[i].y, a), b); }
            if (MODE == 3) acc[i] = __fadd2_rn(__fmul2_rn(acc[i], a2), b2);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, float* d) {
    const int iters = 4096, grid = 148 * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(d, 0.999f, 0.001f, iters);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<grid, 256>>>(d, 0.999f, 0.001f, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double fl = (double)grid * 256 * iters * 16;  // scalar mul-add pairs
    printf("%-22s %8.3f ms  %7.2f T mul-add/s\n", name, ms, fl / ms * 1e-9);
}

int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("scalar FFMA", d);
    run<1>("FFMA2", d);
    run<2>("scalar FMUL+FADD", d);
    run<3>("FMUL2+FADD2", d);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
