// pyramid.cu -- Gaussian scale-space + DoG for sm_100a.
//
// Replaces Gaussian_Blur / buildGaussianPyramid / buildDoGPyramid (reference src/sift.cpp:95-153, 219-283).
// The reference blurs every scale of an octave from the OCTAVE BASE with an unnormalised, truncated
// (radius floor(3 sigma)) sampled 2-D Gaussian, zero padded, with source row rows-1 / col cols-1 read as
// zero (:116).  That kernel and that mask are both separable, so each scale is two 1-D passes here.
//
// octave_kernel: one CTA owns a 32x64 output tile of one frame.
//   phase 1  masked octave-base tile + 18-px halo  -> shared (coalesced 4-byte loads, odd pitch)
//   phase 2  horizontal pass, 4 scales (radii 4/8/12/18), register-blocked: a thread produces 8 adjacent
//            outputs of one row from 8+2r shared loads; lanes run down rows (odd pitch => conflict-free)
//   phase 3  vertical pass, lane = column, 8 output rows per thread, all 4 scales kept in registers so the
//            DoG subtraction, the G1/G2 stores and the NEAREST-decimated next-octave base (src[2y][2x],
//            :253-254) are fused into the epilogue; every store is a full 128-byte row segment.
// HBM traffic per pixel: 4 B read (+halo re-reads served by L2) and 28 B written (G1,G2,D0..D3, 1/4 G0').
#include "sift_internal.cuh"

namespace siftb200 {

__constant__ float c_taps[5][kTapStride];

void upload_taps(const float host_taps[5][kTapStride]) { cudaMemcpyToSymbol(c_taps, host_taps, sizeof(float) * 5 * kTapStride); }

namespace {

constexpr int TW = 32;   // tile width  (= warp width: one lane per column in the vertical pass)
constexpr int TH = 64;   // tile height
constexpr int NT = 256;  // threads per CTA
constexpr int GRP = 8;   // outputs per thread along the filter direction
constexpr int HP = TW + 1;  // pitch of the horizontal-pass results (odd)

__host__ __device__ constexpr int rad_of(int s) { return s <= 1 ? 4 : s == 2 ? 8 : s == 3 ? 12 : 18; }

// 8 outputs of a (2R+1)-tap FIR from 8+2R inputs at `in[t*stride]`; taps are compile-time constant-bank operands.
template <int S, int STRIDE>
__device__ __forceinline__ void fir8(const float* __restrict__ in, float (&acc)[GRP]) {
    constexpr int R = rad_of(S);
#pragma unroll
    for (int k = 0; k < GRP; ++k) acc[k] = 0.f;
#pragma unroll
    for (int t = 0; t < GRP + 2 * R; ++t) {
        const float v = in[t * STRIDE];
#pragma unroll
        for (int k = 0; k < GRP; ++k) {
            const int j = t - k;
            if (j >= 0 && j <= 2 * R) acc[k] = fmaf(v, c_taps[S][j], acc[k]);
        }
    }
}

// Horizontal pass of scale S over the rows the vertical pass will need: [HALO-R, HALO+TH+R) of the input tile.
template <int S, int HALO>
__device__ __forceinline__ void hpass(const float* __restrict__ sIn, float* __restrict__ sH, int tid) {
    constexpr int R = rad_of(S);
    constexpr int IP = TW + 2 * HALO + 1;
    constexpr int NROWS = TH + 2 * R;
    constexpr int NITEMS = NROWS * (TW / GRP);
    for (int id = tid; id < NITEMS; id += NT) {
        const int row = id % NROWS, g = id / NROWS;
        float acc[GRP];
        fir8<S, 1>(sIn + (HALO - R + row) * IP + (HALO - R + GRP * g), acc);
        float* out = sH + row * HP + GRP * g;
#pragma unroll
        for (int k = 0; k < GRP; ++k) out[k] = acc[k];
    }
}

// Masked tile load, scalar: thread = (column, row phase); 4-byte loads coalesced along the row.  Used for the base blur
// (arbitrary source pitch / u8 source).  Zero padding AND the reference's ">= rows-1 / cols-1 reads as zero" (src/sift.cpp:116).
template <int HALO>
__device__ __forceinline__ void load_tile(float* __restrict__ sIn, const float* __restrict__ src, const uint8_t* __restrict__ src8, int rows,
                                          int cols, int pitch, int ty0, int tx0, int tid) {
    constexpr int IW = TW + 2 * HALO, IH = TH + 2 * HALO, IP = IW + 1;
    constexpr int RPP = NT / IW;  // rows per pass
    const int x = tid % IW, y0 = tid / IW;
    if (y0 >= RPP) return;
    const int gx = tx0 - HALO + x;
    const bool col_ok = gx >= 0 && gx < cols - 1;
    for (int y = y0; y < IH; y += RPP) {
        const int gy = ty0 - HALO + y;
        float v = 0.f;
        if (col_ok && gy >= 0 && gy < rows - 1) v = src8 ? (float)src8[(size_t)gy * pitch + gx] : __ldg(src + (size_t)gy * pitch + gx);
        sIn[y * IP + x] = v;
    }
}

// Masked tile load in two halves so that the global loads can be issued one tile ahead: tile_fetch (16-byte loads into
// registers) and tile_stash (registers -> shared, applying the mask).  Needs a float4-aligned source: pitch % 4 == 0 and
// tx0 % 32 == 0, so the window starts at tx0 - HALO - SLACK with SLACK = (-HALO) mod 4 extra columns dropped on the way in.
// Mask = zero padding AND the reference's ">= rows-1 / cols-1 reads as zero" window fetch (src/sift.cpp:116).
template <int HALO>
struct TileIO {
    static constexpr int SLACK = (4 - HALO % 4) % 4;
    static constexpr int IW = TW + 2 * HALO, IH = TH + 2 * HALO, IP = IW + 1;
    static constexpr int Q = (IW + 2 * SLACK) / 4;        // float4 per row
    static constexpr int NSLOT = (IH * Q + NT - 1) / NT;  // float4 per thread

    static __device__ __forceinline__ void fetch(float4 (&pre)[NSLOT], const float* __restrict__ src, int rows, int pitch, int ty0, int tx0, int tid) {
#pragma unroll
        for (int k = 0; k < NSLOT; ++k) {
            const int idx = tid + k * NT;
            const int y = idx / Q, q = idx - y * Q;
            const int gy = ty0 - HALO + y, gx4 = tx0 - HALO - SLACK + 4 * q;
            pre[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < IH * Q && gy >= 0 && gy < rows - 1 && gx4 >= 0 && gx4 < pitch) pre[k] = __ldg(reinterpret_cast<const float4*>(src + (size_t)gy * pitch + gx4));
        }
    }
    static __device__ __forceinline__ void stash(const float4 (&pre)[NSLOT], float* __restrict__ sIn, int cols, int tx0, int tid) {
#pragma unroll
        for (int k = 0; k < NSLOT; ++k) {
            const int idx = tid + k * NT;
            if (idx >= IH * Q) break;
            const int y = idx / Q, q = idx - y * Q;
            const int gx4 = tx0 - HALO - SLACK + 4 * q;
            const float e[4] = {pre[k].x, pre[k].y, pre[k].z, pre[k].w};
            float* out = sIn + y * IP + 4 * q - SLACK;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int sx = 4 * q - SLACK + c, gx = gx4 + c;
                if (sx >= 0 && sx < IW) out[c] = (gx >= 0 && gx < cols - 1) ? e[c] : 0.f;
            }
        }
    }
};

constexpr int H_OFF1 = 0;
constexpr int H_OFF2 = H_OFF1 + (TH + 2 * 4) * HP;
constexpr int H_OFF3 = H_OFF2 + (TH + 2 * 8) * HP;
constexpr int H_OFF4 = H_OFF3 + (TH + 2 * 12) * HP;
constexpr int H_END = H_OFF4 + (TH + 2 * 18) * HP;
constexpr int OCT_HALO = kMaxRadius;
constexpr int OCT_IN = (TW + 2 * OCT_HALO + 1) * (TH + 2 * OCT_HALO);
constexpr int OCT_SMEM_BYTES = (OCT_IN + H_END) * 4;

struct OctArgs {
    const float* G0;
    float *G1, *G2, *G3, *G4, *D0, *D1, *D2, *D3;
    float* nextG0;
    int rows, cols, pitch;
    size_t frame_stride;
    int nrows, ncols, npitch;
    size_t nframe_stride;
};

// One CTA per tile, 3 CTAs per SM.  Two variants were measured slower and dropped (profiles/README.md): a persistent CTA that
// kept the next tile's loads in flight in registers (128 registers -> 2 CTAs/SM, 46.7 us), and a CTA marching down four blocks
// re-using the last 2R horizontal-pass rows (13 % fewer FFMAs but spills, an extra barrier and a row shift per block: 46 us).
__global__ void __launch_bounds__(NT, 3) octave_kernel(const OctArgs a) {
    extern __shared__ float smem[];
    float* sIn = smem;
    float* sH = smem + OCT_IN;
    using IO = TileIO<OCT_HALO>;
    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    const int f = blockIdx.z;
    const size_t foff = (size_t)f * a.frame_stride;
    {
        float4 pre[IO::NSLOT];
        IO::fetch(pre, a.G0 + foff, a.rows, a.pitch, ty0, tx0, tid);
        IO::stash(pre, sIn, a.cols, tx0, tid);
    }
    __syncthreads();
    hpass<4, OCT_HALO>(sIn, sH + H_OFF4, tid);
    hpass<3, OCT_HALO>(sIn, sH + H_OFF3, tid);
    hpass<2, OCT_HALO>(sIn, sH + H_OFF2, tid);
    hpass<1, OCT_HALO>(sIn, sH + H_OFF1, tid);
    __syncthreads();

    const int x = tid & 31, rg = tid >> 5;
    float g1[GRP], g2[GRP], g3[GRP], g4[GRP];
    fir8<1, HP>(sH + H_OFF1 + (rg * GRP) * HP + x, g1);
    fir8<2, HP>(sH + H_OFF2 + (rg * GRP) * HP + x, g2);
    fir8<3, HP>(sH + H_OFF3 + (rg * GRP) * HP + x, g3);
    fir8<4, HP>(sH + H_OFF4 + (rg * GRP) * HP + x, g4);

    const int gx = tx0 + x;
    constexpr int IP = TW + 2 * OCT_HALO + 1;
    if (gx >= a.cols) return;
#pragma unroll
    for (int k = 0; k < GRP; ++k) {
        const int gy = ty0 + rg * GRP + k;
        if (gy >= a.rows) break;
        const size_t p = foff + (size_t)gy * a.pitch + gx;
        // DoG level 0 uses the real base value; the masked copy in shared memory is zero on the last row/col.
        float g0 = sIn[(OCT_HALO + rg * GRP + k) * IP + OCT_HALO + x];
        if (gy == a.rows - 1 || gx == a.cols - 1) g0 = __ldg(a.G0 + p);
        a.G1[p] = g1[k];
        a.G2[p] = g2[k];
        if (a.G3) { a.G3[p] = g3[k]; a.G4[p] = g4[k]; }
        a.D0[p] = g1[k] - g0;
        a.D1[p] = g2[k] - g1[k];
        a.D2[p] = g3[k] - g2[k];
        a.D3[p] = g4[k] - g3[k];
        if (a.nextG0 && !((gy | gx) & 1)) {
            const int ny = gy >> 1, nx = gx >> 1;
            if (ny < a.nrows && nx < a.ncols) a.nextG0[(size_t)f * a.nframe_stride + (size_t)ny * a.npitch + nx] = g2[k];
        }
    }
}

// ---- base blur: image -> octave-0 base, sigma = sqrt(1.6^2 + 0.2^2), radius 4 (src/sift.cpp:237) ----------
constexpr int BASE_HALO = 4;
constexpr int BASE_IN = (TW + 2 * BASE_HALO + 1) * (TH + 2 * BASE_HALO);
constexpr int BASE_SMEM_BYTES = (BASE_IN + (TH + 2 * 4) * HP) * 4;

__global__ void __launch_bounds__(NT, 4)
    base_blur_kernel(const float* __restrict__ src, const uint8_t* __restrict__ src8, size_t src_frame_stride, int src_pitch, float* __restrict__ dst,
                     size_t dst_frame_stride, int dst_pitch, int rows, int cols, int vec) {
    extern __shared__ float smem[];
    float* sIn = smem;
    float* sH = smem + BASE_IN;
    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    if (vec) {  // float source, 16-byte aligned rows: float4 loads
        using IO = TileIO<BASE_HALO>;
        float4 pre[IO::NSLOT];
        IO::fetch(pre, src + (size_t)blockIdx.z * src_frame_stride, rows, src_pitch, ty0, tx0, tid);
        IO::stash(pre, sIn, cols, tx0, tid);
    } else {
        load_tile<BASE_HALO>(sIn, src ? src + (size_t)blockIdx.z * src_frame_stride : nullptr, src8 ? src8 + (size_t)blockIdx.z * src_frame_stride : nullptr,
                             rows, cols, src_pitch, ty0, tx0, tid);
    }
    __syncthreads();
    hpass<0, BASE_HALO>(sIn, sH, tid);
    __syncthreads();
    const int x = tid & 31, rg = tid >> 5;
    float g[GRP];
    fir8<0, HP>(sH + (rg * GRP) * HP + x, g);
    const int gx = tx0 + x;
    if (gx >= cols) return;
#pragma unroll
    for (int k = 0; k < GRP; ++k) {
        const int gy = ty0 + rg * GRP + k;
        if (gy >= rows) break;
        dst[(size_t)blockIdx.z * dst_frame_stride + (size_t)gy * dst_pitch + gx] = g[k];
    }
}

// ---- generic 1-D pass with run-time taps (stage-level Gaussian_Blur / Gaussian_Blur_1D for any sigma) ------
// out(y,x) = sum_{k=lo..hi} taps[k-lo] * src(y+k or x+k) over source positions p with 0 <= p < limit-1
// (the reference's mask).  Sequential accumulation in k order; `exact` rounds mul and add separately
// (the reference's non-FMA arithmetic) so Gaussian_Blur_1D is reproduced bit for bit.
__global__ void blur_pass_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols, const float* __restrict__ taps, int lo,
                                 int hi, int vertical, int exact, int zero_last_row) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= cols || y >= rows) return;
    float acc = 0.f;
    for (int k = lo; k <= hi; ++k) {
        float v = 0.f;
        if (vertical) {
            const int p = y + k;
            if (p >= 0 && p < rows - 1) v = src[(size_t)p * cols + x];
        } else {
            const int p = x + k;
            if (p >= 0 && p < cols - 1 && !(zero_last_row && y >= rows - 1)) v = src[(size_t)y * cols + p];
        }
        const float t = taps[k - lo];
        acc = exact ? __fadd_rn(acc, __fmul_rn(v, t)) : fmaf(v, t, acc);
    }
    dst[(size_t)y * cols + x] = acc;
}

// ---- 2x bilinear upsample (BASELINE config 3: "2x upsampled base octave") ------------------------------------------
// The reference has no upsample path (createInitialImage ignores doubleSize, src/sift.cpp:219-227); this is the front end the
// north star names, with cv::resize(INTER_LINEAR) semantics: half-pixel centres, sx = floor(x/2 - 0.25), weights 0.25/0.75,
// edge replicate (fx = 0 at sx < 0 and sx >= cols-1).  Horizontal pass then vertical pass, mul and add rounded separately.
__global__ void upsample2x_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols, size_t src_fs, size_t dst_fs) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= 2 * cols || y >= 2 * rows) return;
    float fx = (float)((x + 0.5) * 0.5 - 0.5), fy = (float)((y + 0.5) * 0.5 - 0.5);
    int sx = (int)floorf(fx), sy = (int)floorf(fy);
    fx -= sx; fy -= sy;
    if (sx < 0) { sx = 0; fx = 0.f; }
    if (sx >= cols - 1) { sx = cols - 1; fx = 0.f; }
    if (sy < 0) { sy = 0; fy = 0.f; }
    if (sy >= rows - 1) { sy = rows - 1; fy = 0.f; }
    const int sx1 = min(sx + 1, cols - 1), sy1 = min(sy + 1, rows - 1);
    const float* s = src + (size_t)blockIdx.z * src_fs;
    const float a0 = 1.f - fx, a1 = fx, b0 = 1.f - fy, b1 = fy;
    const float h0 = __fadd_rn(__fmul_rn(__ldg(s + (size_t)sy * cols + sx), a0), __fmul_rn(__ldg(s + (size_t)sy * cols + sx1), a1));
    const float h1 = __fadd_rn(__fmul_rn(__ldg(s + (size_t)sy1 * cols + sx), a0), __fmul_rn(__ldg(s + (size_t)sy1 * cols + sx1), a1));
    dst[(size_t)blockIdx.z * dst_fs + (size_t)y * (2 * cols) + x] = __fadd_rn(__fmul_rn(h0, b0), __fmul_rn(h1, b1));
}

// ---- driver's colour conversion (src/main.cpp:84): cvtColor(img, gray, COLOR_RGB2GRAY) applied to the BGR bytes imread returns,
// i.e. channel 0 gets the "R" weight.  cv2 4.13 fixed point: (9798*c0 + 19235*c1 + 3735*c2 + 16384) >> 15  (OpenCV
// 4.0 used the 14-bit table -- the tests pin the wheel that is available).  Interleaved 3-byte pixels in, dense u8 gray out.
__global__ void rgb2gray_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t n_pixels) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pixels) return;
    const uint8_t* p = src + 3 * i;
    dst[i] = (uint8_t)((9798u * p[0] + 19235u * p[1] + 3735u * p[2] + 16384u) >> 15);
}

__global__ void dog_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ d, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = b[i] - a[i];
}

}  // namespace

// per-device opt-in to > 48 KB dynamic shared memory; called from sift_b200_create after cudaSetDevice
void init_pyramid_kernels() {
    cudaFuncSetAttribute(base_blur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BASE_SMEM_BYTES);
    cudaFuncSetAttribute(octave_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OCT_SMEM_BYTES);
}

int launch_base_blur(const float* src, size_t src_frame_stride, int src_pitch, const uint8_t* src_u8, const OctaveView& o0, int n_frames, cudaStream_t st) {
    dim3 grid((o0.cols + TW - 1) / TW, (o0.rows + TH - 1) / TH, n_frames);
    const int vec = !src_u8 && src_pitch % 4 == 0 && src_frame_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    base_blur_kernel<<<grid, NT, BASE_SMEM_BYTES, st>>>(src_u8 ? nullptr : src, src_u8, src_frame_stride, src_pitch, o0.G[0], o0.frame_stride, o0.pitch, o0.rows,
                                                       o0.cols, vec);
    return 1;
}

int launch_octave(const PyrView& pv, int o, int n_frames, bool write_all_levels, cudaStream_t st) {
    const OctaveView& v = pv.oct[o];
    OctArgs a;
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
    dim3 grid((v.cols + TW - 1) / TW, (v.rows + TH - 1) / TH, n_frames);
    octave_kernel<<<grid, NT, OCT_SMEM_BYTES, st>>>(a);
    return 1;
}

// Separable blur with run-time taps: horizontal (rows >= rows-1 zeroed at the source) then vertical.
// taps_hi: last tap offset (radius for Gaussian_Blur, radius-1 for Gaussian_Blur_1D which also runs vertical first).
int launch_generic_blur(const float* src, float* dst, int rows, int cols, const float* d_taps, int radius, int taps_hi, cudaStream_t st) {
    float* tmp = nullptr;
    if (cudaMallocAsync((void**)&tmp, sizeof(float) * (size_t)rows * cols, st) != cudaSuccess) return -1;
    dim3 blk(32, 8), grid((cols + 31) / 32, (rows + 7) / 8);
    const bool one_d = taps_hi != radius;
    if (!one_d) {
        blur_pass_kernel<<<grid, blk, 0, st>>>(src, tmp, rows, cols, d_taps, -radius, taps_hi, 0, 0, 1);
        blur_pass_kernel<<<grid, blk, 0, st>>>(tmp, dst, rows, cols, d_taps, -radius, taps_hi, 1, 0, 0);
    } else {  // Gaussian_Blur_1D: vertical then horizontal, exact non-FMA arithmetic (src/sift.cpp:193-212)
        blur_pass_kernel<<<grid, blk, 0, st>>>(src, tmp, rows, cols, d_taps, -radius, taps_hi, 1, 1, 0);
        blur_pass_kernel<<<grid, blk, 0, st>>>(tmp, dst, rows, cols, d_taps, -radius, taps_hi, 0, 1, 0);
    }
    cudaFreeAsync(tmp, st);
    return 2;
}

int launch_upsample2x(const float* src, float* dst, int rows, int cols, int n_frames, cudaStream_t st) {
    dim3 blk(32, 8), grid((2 * cols + 31) / 32, (2 * rows + 7) / 8, n_frames);
    upsample2x_kernel<<<grid, blk, 0, st>>>(src, dst, rows, cols, (size_t)rows * cols, (size_t)rows * cols * 4);
    return 1;
}

int launch_rgb2gray_u8(const uint8_t* src, uint8_t* dst, size_t n_pixels, cudaStream_t st) {
    rgb2gray_u8_kernel<<<(unsigned)((n_pixels + 255) / 256), 256, 0, st>>>(src, dst, n_pixels);
    return 1;
}

int launch_dog(const PyrView& pv, int n_frames, cudaStream_t st) {
    int n = 0;
    for (int o = 0; o < pv.n_oct; ++o) {
        const OctaveView& v = pv.oct[o];
        size_t cnt = v.frame_stride * n_frames;
        for (int i = 0; i < kNumScales - 1; ++i) {
            dog_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(v.G[i], v.G[i + 1], v.D[i], cnt);
            ++n;
        }
    }
    return n;
}

}  // namespace siftb200
