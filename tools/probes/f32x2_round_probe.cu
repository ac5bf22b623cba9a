// Does ptxas keep mul.rn.f32x2 + add.rn.f32x2 separately rounded?  (fused and unfused results differ for these inputs)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(const float* in, float* out) {
    const float x = in[0], a = in[1], b = in[2];
    out[0] = __fadd_rn(__fmul_rn(x, a), b);
    out[1] = fmaf(x, a, b);
    const float2 r = __fadd2_rn(__fmul2_rn(make_float2(x, x), make_float2(a, a)), make_float2(b, b));
    out[2] = r.x; out[3] = r.y;
    float2 m;
    unsigned long long mm, xx, aa, bb, rr;
    const float2 x2 = make_float2(x, x), a2 = make_float2(a, a), b2 = make_float2(b, b);
    xx = *reinterpret_cast<const unsigned long long*>(&x2); aa = *reinterpret_cast<const unsigned long long*>(&a2); bb = *reinterpret_cast<const unsigned long long*>(&b2);
    asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(mm) : "l"(xx), "l"(aa));
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(rr) : "l"(mm), "l"(bb));
    m = *reinterpret_cast<float2*>(&rr);
    out[4] = m.x; out[5] = m.y;
    const float2 t = __ffma2_rn(__ffma2_rn(x2, a2, make_float2(-0.f, -0.f)), make_float2(1.f, 1.f), b2);
    out[6] = t.x; out[7] = t.y;
}
int main() {
    float h[3] = {1.f + 1.f / 4096, 1.f + 1.f / 4096, -(1.f + 1.f / 2048)}, *d, *o, r[8];
    cudaMalloc(&d, 12); cudaMalloc(&o, 32);
    cudaMemcpy(d, h, 12, cudaMemcpyHostToDevice);
    k<<<1, 1>>>(d, o);
    cudaMemcpy(r, o, 32, cudaMemcpyDeviceToHost);
    printf("scalar unfused %g | fmaf %g | packed intrinsics %g %g | packed asm volatile %g %g | fma(fma(x,a,-0),1,b) %g %g\n", r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7]);
    return 0;
}
