#include <cuda_runtime.h>
__constant__ float c_t[64];
__constant__ float2 c_t2[32];
__global__ void k(const float2* __restrict__ in, float2* out) {
    float2 acc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 12; ++t) {
        const float2 v = in[threadIdx.x + 32 * t];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int j = t - i;
            if (j >= 0 && j < 9) {
                acc[i] = __ffma2_rn(v, make_float2(c_t[j], c_t[j]), acc[i]);
                acc[i] = __ffma2_rn(v, c_t2[j], acc[i]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) out[threadIdx.x + 32 * i] = acc[i];
}
