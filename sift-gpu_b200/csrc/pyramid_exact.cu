// pyramid_exact.cu -- the reference's Gaussian pyramid in the reference's own summation order (opt-in "exact pyramid" mode).
//
// Gaussian_Blur (reference src/sift.cpp:110-153) is a NON-separable loop: every output pixel sums its (2w+1)^2 window in
// row-major order in float, product and sum rounded separately (x86-64 build without FMA), with the 2-D taps
// K[i][j] = float(8192 * exp(-(i^2+j^2)/den) / (2 PI s^2)) and a final /8192.  The default pipeline (pyramid.cu) uses the
// mathematically identical separable form, whose results differ from this loop in the last bits; a few of those differences
// survive to the descriptors' uchar quantisation step (src/sift.cpp:703-709).  The kernels here replay the loop literally --
// same taps, same order, same rounding -- so the pyramid is bit-identical to the reference's and the rest of the path can be
// checked end to end without that noise.  Cost: 2364 multiply-add pairs per octave pixel instead of 176 FFMAs, ~4x lower
// whole-path throughput.  Used only when sift_b200_set_exact_pyramid(h, 1) was called.
//
// A CTA = 128 threads owns a 32x16 output tile; a thread owns 4 adjacent pixels and keeps the current window row in registers,
// so one shared load feeds up to 4 taps; the taps of a kernel row are warp-uniform constant-bank operands.
#include "sift_internal.cuh"

namespace siftb200 {

// 2-D taps: base (w=4) and the four octave scales (w = 4, 8, 12, 18), row-major (2w+1)^2 each
constexpr int kK2dOff[5] = {0, 81, 162, 162 + 289, 162 + 289 + 625};
constexpr int kK2dTotal = 162 + 289 + 625 + 1369;
__constant__ float c_k2d[kK2dTotal];

namespace {

constexpr int XT = 32, YT = 16;  // output tile
constexpr int XP = 4;            // pixels per thread
constexpr int XNT = 128;         // threads per CTA
constexpr int XHALO = kMaxRadius;
constexpr int XPITCH = XT + 2 * XHALO + 1;  // 69: 4*tx + 5*ty (mod 32) is a bijection over the 8x4 lanes of a warp -> conflict-free
constexpr int XROWS = YT + 2 * XHALO;

__host__ __device__ constexpr int xrad(int s) { return s <= 1 ? 4 : s == 2 ? 8 : s == 3 ? 12 : 18; }
__host__ __device__ constexpr int xoff(int s) { return s == 0 ? 0 : s == 1 ? 81 : s == 2 ? 162 : s == 3 ? 162 + 289 : 162 + 289 + 625; }

// masked tile: zero padding AND the reference's ">= rows-1 / cols-1 reads as zero" window fetch (src/sift.cpp:116)
template <int HALO>
__device__ __forceinline__ void load_masked(float* __restrict__ tile, const float* __restrict__ src, const uint8_t* __restrict__ src8, int rows, int cols,
                                            int pitch, int ty0, int tx0, int tid) {
    constexpr int W = XT + 2 * HALO, H = YT + 2 * HALO;
    for (int idx = tid; idx < W * H; idx += XNT) {
        const int y = idx / W, x = idx - y * W;
        const int gy = ty0 - HALO + y, gx = tx0 - HALO + x;
        float v = 0.f;
        if (gy >= 0 && gy < rows - 1 && gx >= 0 && gx < cols - 1) v = src8 ? (float)src8[(size_t)gy * pitch + gx] : __ldg(src + (size_t)gy * pitch + gx);
        tile[y * XPITCH + x] = v;
    }
}

// dot[p] = sum over the window of pixel (y, x0+p) in the reference's row-major order (:142-145), then /8192 (:146).
// `win` points at tile element (y - R, x0 - R) of the thread's first pixel.
template <int S>
__device__ __forceinline__ void blur2d_exact(const float* __restrict__ win, float (&dot)[XP]) {
    constexpr int R = xrad(S), KS = 2 * R + 1;
#pragma unroll
    for (int p = 0; p < XP; ++p) dot[p] = 0.f;
#pragma unroll 1
    for (int i = 0; i < KS; ++i) {
        float seg[XP + 2 * R];
#pragma unroll
        for (int t = 0; t < XP + 2 * R; ++t) seg[t] = win[i * XPITCH + t];
        const float* k = c_k2d + xoff(S) + i * KS;
#pragma unroll
        for (int j = 0; j < KS; ++j) {
            const float kv = k[j];
#pragma unroll
            for (int p = 0; p < XP; ++p) dot[p] = __fadd_rn(dot[p], __fmul_rn(seg[p + j], kv));
        }
    }
#pragma unroll
    for (int p = 0; p < XP; ++p) dot[p] = __fmul_rn(dot[p], 1.f / 8192.f);  // == dot / 8192 (power of two; no subnormals here)
}

__global__ void __launch_bounds__(XNT) exact_base_kernel(const float* __restrict__ src, const uint8_t* __restrict__ src8, size_t src_frame_stride, int src_pitch,
                                                         float* __restrict__ dst, size_t dst_frame_stride, int dst_pitch, int rows, int cols) {
    __shared__ float tile[(YT + 8) * XPITCH];
    const int tid = threadIdx.x, tx0 = blockIdx.x * XT, ty0 = blockIdx.y * YT;
    load_masked<4>(tile, src ? src + (size_t)blockIdx.z * src_frame_stride : nullptr, src8 ? src8 + (size_t)blockIdx.z * src_frame_stride : nullptr, rows, cols,
                   src_pitch, ty0, tx0, tid);
    __syncthreads();
    const int x0 = (tid & 7) * XP, y = tid >> 3;
    float g[XP];
    blur2d_exact<0>(tile + y * XPITCH + x0, g);
    const int gy = ty0 + y;
    if (gy >= rows) return;
#pragma unroll
    for (int p = 0; p < XP; ++p)
        if (tx0 + x0 + p < cols) dst[(size_t)blockIdx.z * dst_frame_stride + (size_t)gy * dst_pitch + tx0 + x0 + p] = g[p];
}

struct XOctArgs {
    const float* G0;
    float *G1, *G2, *G3, *G4, *D0, *D1, *D2, *D3;
    float* nextG0;
    int rows, cols, pitch;
    size_t frame_stride;
    int nrows, ncols, npitch;
    size_t nframe_stride;
};

// scales 1..4 of one octave from the octave base (src/sift.cpp:257-258), DoG (:265-283) and the NEAREST-decimated next base (:253-254)
__global__ void __launch_bounds__(XNT) exact_octave_kernel(const XOctArgs a) {
    __shared__ float tile[XROWS * XPITCH];
    const int tid = threadIdx.x, tx0 = blockIdx.x * XT, ty0 = blockIdx.y * YT, f = blockIdx.z;
    const size_t foff = (size_t)f * a.frame_stride;
    load_masked<XHALO>(tile, a.G0 + foff, nullptr, a.rows, a.cols, a.pitch, ty0, tx0, tid);
    __syncthreads();
    const int x0 = (tid & 7) * XP, y = tid >> 3;
    float g1[XP], g2[XP], g3[XP], g4[XP];
    blur2d_exact<1>(tile + (y + XHALO - 4) * XPITCH + x0 + XHALO - 4, g1);
    blur2d_exact<2>(tile + (y + XHALO - 8) * XPITCH + x0 + XHALO - 8, g2);
    blur2d_exact<3>(tile + (y + XHALO - 12) * XPITCH + x0 + XHALO - 12, g3);
    blur2d_exact<4>(tile + y * XPITCH + x0, g4);
    const int gy = ty0 + y;
    if (gy >= a.rows) return;
#pragma unroll
    for (int p = 0; p < XP; ++p) {
        const int gx = tx0 + x0 + p;
        if (gx >= a.cols) break;
        const size_t q = foff + (size_t)gy * a.pitch + gx;
        const float g0 = __ldg(a.G0 + q);  // the real base value (the tile copy is masked on the last row/column)
        a.G1[q] = g1[p];
        a.G2[q] = g2[p];
        if (a.G3) { a.G3[q] = g3[p]; a.G4[q] = g4[p]; }
        a.D0[q] = g1[p] - g0;
        a.D1[q] = g2[p] - g1[p];
        a.D2[q] = g3[p] - g2[p];
        a.D3[q] = g4[p] - g3[p];
        if (a.nextG0 && !((gy | gx) & 1)) {
            const int ny = gy >> 1, nx = gx >> 1;
            if (ny < a.nrows && nx < a.ncols) a.nextG0[(size_t)f * a.nframe_stride + (size_t)ny * a.npitch + nx] = g2[p];
        }
    }
}

}  // namespace

// host_k2d: the five kernels back to back (base, scales 1..4), see kK2dOff
void upload_taps_2d(const float* host_k2d) { cudaMemcpyToSymbol(c_k2d, host_k2d, sizeof(float) * kK2dTotal); }
int k2d_total() { return kK2dTotal; }
int k2d_offset(int s) { return kK2dOff[s]; }

int launch_exact_base(const float* src, size_t src_frame_stride, int src_pitch, const uint8_t* src_u8, const OctaveView& o0, int n_frames, cudaStream_t st) {
    dim3 grid((o0.cols + XT - 1) / XT, (o0.rows + YT - 1) / YT, n_frames);
    exact_base_kernel<<<grid, XNT, 0, st>>>(src_u8 ? nullptr : src, src_u8, src_frame_stride, src_pitch, o0.G[0], o0.frame_stride, o0.pitch, o0.rows, o0.cols);
    return 1;
}

int launch_exact_octave(const PyrView& pv, int o, int n_frames, bool write_all_levels, cudaStream_t st) {
    const OctaveView& v = pv.oct[o];
    XOctArgs a;
    a.G0 = v.G[0]; a.G1 = v.G[1]; a.G2 = v.G[2];
    a.G3 = write_all_levels ? v.G[3] : nullptr;
    a.G4 = write_all_levels ? v.G[4] : nullptr;
    a.D0 = v.D[0]; a.D1 = v.D[1]; a.D2 = v.D[2]; a.D3 = v.D[3];
    a.rows = v.rows; a.cols = v.cols; a.pitch = v.pitch; a.frame_stride = v.frame_stride;
    if (o + 1 < pv.n_oct) {
        const OctaveView& n = pv.oct[o + 1];
        a.nextG0 = n.G[0]; a.nrows = n.rows; a.ncols = n.cols; a.npitch = n.pitch; a.nframe_stride = n.frame_stride;
    } else {
        a.nextG0 = nullptr; a.nrows = a.ncols = a.npitch = 0; a.nframe_stride = 0;
    }
    dim3 grid((v.cols + XT - 1) / XT, (v.rows + YT - 1) / YT, n_frames);
    exact_octave_kernel<<<grid, XNT, 0, st>>>(a);
    return 1;
}

}  // namespace siftb200
