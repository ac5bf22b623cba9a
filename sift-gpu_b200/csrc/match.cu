// match.cu -- brute-force kNN(k=2) over 128-d descriptors (reference driver src/main.cpp:25-40).
//
// BFMatcher(norm).knnMatch(query, train, 2): per query the two nearest train rows, ascending, exact ties to
// the lowest train index.  One warp per query row; each lane owns 4 of the 128 components; element
// differences and the 128-term sum are carried in fp64 (differences of floats are exact in double) so the
// ranking agrees with an exhaustive fp64 evaluation -- the parity bar for this stage is identical indices.
// Train rows stream through L2 (nt*512 B, shared by all warps).
#include "sift_internal.cuh"

namespace siftb200 {
namespace {

constexpr int MATCH_WARPS = 8;

template <int NORM>
__global__ void __launch_bounds__(MATCH_WARPS * 32)
    match_kernel(const float* __restrict__ q, int nq, const float* __restrict__ t, int nt, float* __restrict__ dist, int32_t* __restrict__ idx) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * MATCH_WARPS + warp;
    if (i >= nq) return;
    const float4 a = __ldg(reinterpret_cast<const float4*>(q + (size_t)i * 128) + lane);
    double b0 = INFINITY, b1 = INFINITY;
    int i0 = -1, i1 = -1;
    for (int j = 0; j < nt; ++j) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(t + (size_t)j * 128) + lane);
        const double e0 = (double)a.x - (double)b.x, e1 = (double)a.y - (double)b.y, e2 = (double)a.z - (double)b.z, e3 = (double)a.w - (double)b.w;
        double d = NORM == SIFT_B200_NORM_L1 ? fabs(e0) + fabs(e1) + fabs(e2) + fabs(e3) : e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) d += __shfl_xor_sync(0xffffffffu, d, s);
        if (NORM == SIFT_B200_NORM_L2) d = sqrt(d);
        if (d < b0) { b1 = b0; i1 = i0; b0 = d; i0 = j; }
        else if (d < b1) { b1 = d; i1 = j; }
    }
    if (lane == 0) {
        dist[2 * i] = (float)b0; dist[2 * i + 1] = (float)b1;
        idx[2 * i] = i0; idx[2 * i + 1] = i1;
    }
}

}  // namespace

int launch_match(const float* d_q, int nq, const float* d_t, int nt, int norm, float* d_dist, int32_t* d_idx, cudaStream_t st) {
    if (nq <= 0) return 0;
    const int blocks = (nq + MATCH_WARPS - 1) / MATCH_WARPS;
    if (norm == SIFT_B200_NORM_L1) match_kernel<SIFT_B200_NORM_L1><<<blocks, MATCH_WARPS * 32, 0, st>>>(d_q, nq, d_t, nt, d_dist, d_idx);
    else match_kernel<SIFT_B200_NORM_L2><<<blocks, MATCH_WARPS * 32, 0, st>>>(d_q, nq, d_t, nt, d_dist, d_idx);
    return 1;
}

}  // namespace siftb200
