// pyramid.cu -- Gaussian scale-space + DoG for sm_100a.
//
// Replaces Gaussian_Blur / buildGaussianPyramid / buildDoGPyramid (reference src/sift.cpp:95-153, 219-283).
// The reference blurs every scale of an octave from the OCTAVE BASE with an unnormalised, truncated
// (radius floor(3 sigma)) sampled 2-D Gaussian, zero padded, with source row rows-1 / col cols-1 read as
// zero (:116).  That kernel and that mask are both separable, so each scale is two 1-D passes here.
//
// octave_kernel: one CTA owns a 32x64 output tile of one frame.  All arithmetic is packed FFMA2 (fma.rn.f32x2, new on
// sm_100): the kernel is instruction-issue bound, and one FFMA2 issues two FMAs.  Both shared tiles are stored ROW-PAIR
// INTERLEAVED -- element (y, x) lives in float2 [y/2][x], component y&1 -- so that an aligned register pair always holds
// the same column of two adjacent rows.
//   phase 1  masked octave-base tile + 18-px halo -> shared (16-byte global loads)
//   phase 2  horizontal pass, 4 scales (radii 4/8/12/18): a thread produces 8 adjacent outputs for TWO rows from 8+2r
//            8-byte shared loads; taps are broadcast uniform-register operands; lanes run down row pairs (odd pitch)
//   phase 3  vertical pass, lane = column, 8 output rows (4 row pairs) per thread: even taps hit aligned input pairs
//            (broadcast tap), odd taps are applied as a tap PAIR {t[2d-1], t[2d+1]} to the same input pair and land in a
//            second, swapped accumulator; the two are added at the end.  All 4 scales stay in registers so the DoG
//            subtraction, the G1/G2 stores and the NEAREST-decimated next-octave base (src[2y][2x], :253-254) are fused
//            into the epilogue; every store is a full 128-byte row segment.
// HBM traffic per pixel: 4 B read (+halo re-reads served by L2) and 28 B written (G1,G2,D0..D3, 1/4 G0').
#include "sift_internal.cuh"

namespace siftb200 {

constexpr int kOddStride = 20;  // >= kMaxRadius + 1 tap pairs
__constant__ float c_taps[5][kTapStride];
__constant__ float2 c_odd[5][kOddStride];  // c_odd[s][d] = {t[2d-1], t[2d+1]} (zero outside the kernel)

namespace {
__host__ __device__ constexpr int rad_of(int s) { return s <= 1 ? 4 : s == 2 ? 8 : s == 3 ? 12 : 18; }
}  // namespace

void upload_taps(const float host_taps[5][kTapStride]) {
    cudaMemcpyToSymbol(c_taps, host_taps, sizeof(float) * 5 * kTapStride);
    float2 odd[5][kOddStride] = {};
    for (int s = 0; s < 5; ++s) {
        const int R = rad_of(s);
        for (int d = 0; d <= R; ++d) {
            odd[s][d].x = d >= 1 ? host_taps[s][2 * d - 1] : 0.f;
            odd[s][d].y = d < R ? host_taps[s][2 * d + 1] : 0.f;
        }
    }
    cudaMemcpyToSymbol(c_odd, odd, sizeof(odd));
}

namespace {

constexpr int TW = 32;   // tile width  (= warp width: one lane per column in the vertical pass)
#ifndef TILE_H
#define TILE_H 64
#endif
#ifndef OCT_CTAS
#define OCT_CTAS 3
#endif
#ifndef BASE_CTAS
#define BASE_CTAS 4
#endif
constexpr int TH = TILE_H;      // tile height
constexpr int NT = TILE_H * 4;  // threads per CTA: one warp per 8 output rows in the vertical pass
constexpr int GRP = 8;   // outputs per thread along the filter direction
constexpr int HP2 = TW + 1;  // pitch of the horizontal-pass results, in float2 (odd)

// element (y, x) of a row-pair-interleaved tile with pitch P2 (float2 per row pair)
__device__ __forceinline__ int il_index(int y, int x, int P2) { return ((y >> 1) * P2 + x) * 2 + (y & 1); }

// Horizontal: 8 adjacent outputs of a (2R+1)-tap FIR for two rows at once, from 8+2R float2 inputs.
template <int S>
__device__ __forceinline__ void fir8_rows2(const float2* __restrict__ in, float2 (&acc)[GRP]) {
    constexpr int R = rad_of(S);
#pragma unroll
    for (int k = 0; k < GRP; ++k) acc[k] = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < GRP + 2 * R; ++t) {
        const float2 v = in[t];
#pragma unroll
        for (int k = 0; k < GRP; ++k) {
            const int j = t - k;
            if (j >= 0 && j <= 2 * R) acc[k] = __ffma2_rn(v, make_float2(c_taps[S][j], c_taps[S][j]), acc[k]);
        }
    }
}

// Vertical: 8 consecutive output rows of one column.  `in` points at the input pair holding rows {r0 - R, r0 - R + 1} of the
// column (r0 = first output row, even), pairs are HP2 apart.  out[r0+2j] = P[j].x + Q[j].y, out[r0+2j+1] = P[j].y + Q[j].x.
template <int S>
__device__ __forceinline__ void fir8_col(const float2* __restrict__ in, float (&g)[GRP]) {
    constexpr int R = rad_of(S);
    float2 P[GRP / 2], Q[GRP / 2];
#pragma unroll
    for (int j = 0; j < GRP / 2; ++j) P[j] = Q[j] = make_float2(0.f, 0.f);
#pragma unroll
    for (int m = 0; m < GRP / 2 + R; ++m) {
        const float2 v = in[m * HP2];
#pragma unroll
        for (int j = 0; j < GRP / 2; ++j) {
            const int d = m - j;
            if (d >= 0 && d <= R) {
                P[j] = __ffma2_rn(v, make_float2(c_taps[S][2 * d], c_taps[S][2 * d]), P[j]);
                Q[j] = __ffma2_rn(v, c_odd[S][d], Q[j]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < GRP / 2; ++j) {
        g[2 * j] = P[j].x + Q[j].y;
        g[2 * j + 1] = P[j].y + Q[j].x;
    }
}

// Horizontal pass of scale S over the rows the vertical pass will need: [HALO-R, HALO+TH+R) of the input tile (HALO-R even).
template <int S, int HALO>
__device__ __forceinline__ void hpass(const float2* __restrict__ sIn, float2* __restrict__ sH, int tid) {
    constexpr int R = rad_of(S);
    static_assert((HALO - R) % 2 == 0, "row pairs of the pass must be row pairs of the tile");
    constexpr int IP2 = TW + 2 * HALO + 1;
    constexpr int NPAIRS = (TH + 2 * R) / 2;
    constexpr int NITEMS = NPAIRS * (TW / GRP);
    for (int id = tid; id < NITEMS; id += NT) {
        const int m = id % NPAIRS, g = id / NPAIRS;
        float2 acc[GRP];
        fir8_rows2<S>(sIn + ((HALO - R) / 2 + m) * IP2 + (HALO - R + GRP * g), acc);
        float2* out = sH + m * HP2 + GRP * g;
#pragma unroll
        for (int k = 0; k < GRP; ++k) out[k] = acc[k];
    }
}

// Masked tile load, scalar: thread = (column, row phase); 4-byte loads coalesced along the row.  Used for the base blur
// (arbitrary source pitch / u8 source).  Zero padding AND the reference's ">= rows-1 / cols-1 reads as zero" (src/sift.cpp:116).
template <int HALO>
__device__ __forceinline__ void load_tile(float* __restrict__ sIn, const float* __restrict__ src, const uint8_t* __restrict__ src8, int rows,
                                          int cols, int pitch, int ty0, int tx0, int tid) {
    constexpr int IW = TW + 2 * HALO, IH = TH + 2 * HALO, IP2 = IW + 1;
    constexpr int RPP = NT / IW;  // rows per pass
    const int x = tid % IW, y0 = tid / IW;
    if (y0 >= RPP) return;
    const int gx = tx0 - HALO + x;
    const bool col_ok = gx >= 0 && gx < cols - 1;
    for (int y = y0; y < IH; y += RPP) {
        const int gy = ty0 - HALO + y;
        float v = 0.f;
        if (col_ok && gy >= 0 && gy < rows - 1) v = src8 ? (float)src8[(size_t)gy * pitch + gx] : __ldg(src + (size_t)gy * pitch + gx);
        sIn[il_index(y, x, IP2)] = v;
    }
}

// Masked tile load for a uint8 source whose rows are 4-byte aligned (cols, pitch and the tile origin multiples of 4): a thread takes
// four columns of BOTH rows of a row pair -- two 4-byte loads, consecutive threads on consecutive quads of a row: coalesced -- and writes
// four complete float2 {row 2p, row 2p+1} entries.  Same mask as load_tile (zero padding, last row / last column read as zero); the
// byte-per-thread form costs twice the float path's time on the u8 front end of the host pipeline.
template <int HALO>
__device__ __forceinline__ void load_tile_u8x4(float2* __restrict__ sIn, const uint8_t* __restrict__ src8, int rows, int cols, int pitch, int ty0,
                                               int tx0, int tid) {
    constexpr int IW = TW + 2 * HALO, IH = TH + 2 * HALO, IP2 = IW + 1, NQ = IW / 4, NP = IH / 2;
    static_assert(IW % 4 == 0 && HALO % 4 == 0 && IH % 2 == 0, "quads of columns, pairs of rows");
    for (int id = tid; id < NP * NQ; id += NT) {
        const int p = id / NQ, q = id - p * NQ;
        const int gy = ty0 - HALO + 2 * p, gx = tx0 - HALO + 4 * q;
        const bool col_in = gx >= 0 && gx < cols;  // a quad is inside the image as a whole
        const uint8_t* ptr = src8 + (ptrdiff_t)gy * pitch + gx;
        uint32_t a = 0u, b = 0u;
        if (col_in && gy >= 0 && gy < rows - 1) a = __ldg(reinterpret_cast<const uint32_t*>(ptr));
        if (col_in && gy + 1 >= 0 && gy + 1 < rows - 1) b = __ldg(reinterpret_cast<const uint32_t*>(ptr + pitch));
        if (gx + 3 == cols - 1) { a &= 0x00ffffffu; b &= 0x00ffffffu; }  // the last column reads as zero (src/sift.cpp:116)
        float2* dst = sIn + p * IP2 + 4 * q;
#pragma unroll
        for (int k = 0; k < 4; ++k) dst[k] = make_float2((float)((a >> (8 * k)) & 255u), (float)((b >> (8 * k)) & 255u));
    }
}

// Masked tile load for a float source with even pitch and even tile origin, in two halves (global loads first, shared stores
// after).  Warp w owns row pairs w, w+8, ...; lane l owns column pair l (a thread reads the same two columns of BOTH rows of a
// pair: two 8-byte loads, coalesced along the row) -- so the column mask is computed once per thread, the row mask is warp
// uniform and the addresses advance by a constant.  Tiles wider than 64 columns put the few extra column pairs in one more slot.
// Every shared store writes complete float2 {row 2p, row 2p+1} entries: each store wavefront carries 128 useful bytes.
// Mask = zero padding AND the reference's ">= rows-1 / cols-1 reads as zero" window fetch (src/sift.cpp:116).
template <int HALO>
struct TileIO {
    static_assert(HALO % 2 == 0, "8-byte aligned window");
    static constexpr int IW = TW + 2 * HALO, IH = TH + 2 * HALO, IP2 = IW + 1;
    static constexpr int Q = IW / 2, NP = IH / 2, NW = NT / 32;  // column pairs per row, row pairs, warps
    static constexpr int NMAIN = (NP + NW - 1) / NW;              // main slots per thread
    static constexpr int XC = Q > 32 ? Q - 32 : 0;                // extra column pairs
    static constexpr int NX = (NP * XC + NT - 1) / NT;            // extra slots per thread
    static constexpr int NSLOT = NMAIN + NX;
    static_assert(NX <= 1, "one extra slot");

    static __device__ __forceinline__ float2 ld_row(const float* p, bool ok) { return ok ? __ldg(reinterpret_cast<const float2*>(p)) : make_float2(0.f, 0.f); }

    static __device__ __forceinline__ void fetch(float2 (&pre)[2 * NSLOT], const float* __restrict__ src, int rows, int pitch, int ty0, int tx0, int tid) {
        const int lane = tid & 31, w = tid >> 5;
        {
            const int gx = tx0 - HALO + 2 * lane;
            const bool col_ok = lane < Q && gx >= 0 && gx < pitch;
            int gy = ty0 - HALO + 2 * w;
            const float* ptr = src + (ptrdiff_t)gy * pitch + gx;
#pragma unroll
            for (int k = 0; k < NMAIN; ++k) {
                const bool pair_ok = col_ok && (w + NW * k < NP);
                pre[2 * k] = ld_row(ptr, pair_ok && gy >= 0 && gy < rows - 1);
                pre[2 * k + 1] = ld_row(ptr + pitch, pair_ok && gy + 1 >= 0 && gy + 1 < rows - 1);
                gy += 2 * NW;
                ptr += (ptrdiff_t)2 * NW * pitch;
            }
        }
        if (NX) {
            const int p = tid / (XC ? XC : 1), h = 32 + tid % (XC ? XC : 1);
            const int gx = tx0 - HALO + 2 * h, gy = ty0 - HALO + 2 * p;
            const bool ok = p < NP && gx >= 0 && gx < pitch;
            const float* ptr = src + (ptrdiff_t)gy * pitch + gx;
            pre[2 * NMAIN] = ld_row(ptr, ok && gy >= 0 && gy < rows - 1);
            pre[2 * NMAIN + 1] = ld_row(ptr + pitch, ok && gy + 1 >= 0 && gy + 1 < rows - 1);
        }
    }
    static __device__ __forceinline__ void put(float2* out, float2 a, float2 b, int gx, int cols) {
        out[0] = (gx >= 0 && gx < cols - 1) ? make_float2(a.x, b.x) : make_float2(0.f, 0.f);
        out[1] = (gx + 1 >= 0 && gx + 1 < cols - 1) ? make_float2(a.y, b.y) : make_float2(0.f, 0.f);
    }
    static __device__ __forceinline__ void stash(const float2 (&pre)[2 * NSLOT], float2* __restrict__ sIn, int cols, int tx0, int tid) {
        const int lane = tid & 31, w = tid >> 5;
        if (lane < Q) {
            const int gx = tx0 - HALO + 2 * lane;
#pragma unroll
            for (int k = 0; k < NMAIN; ++k)
                if (w + NW * k < NP) put(sIn + (w + NW * k) * IP2 + 2 * lane, pre[2 * k], pre[2 * k + 1], gx, cols);
        }
        if (NX) {
            const int p = tid / (XC ? XC : 1), h = 32 + tid % (XC ? XC : 1);
            if (p < NP) put(sIn + p * IP2 + 2 * h, pre[2 * NMAIN], pre[2 * NMAIN + 1], tx0 - HALO + 2 * h, cols);
        }
    }
};

// shared-memory offsets in float2 (one float2 = one column of a row pair)
constexpr int H_OFF1 = 0;
constexpr int H_OFF2 = H_OFF1 + (TH / 2 + 4) * HP2;
constexpr int H_OFF3 = H_OFF2 + (TH / 2 + 8) * HP2;
constexpr int H_OFF4 = H_OFF3 + (TH / 2 + 12) * HP2;
constexpr int H_END = H_OFF4 + (TH / 2 + 18) * HP2;
constexpr int OCT_HALO = kMaxRadius;
constexpr int OCT_IN = (TW + 2 * OCT_HALO + 1) * (TH / 2 + OCT_HALO);
constexpr int OCT_SMEM_BYTES = (OCT_IN + H_END) * 8;

struct OctArgs {
    const float* G0;
    float *G1, *G2, *G3, *G4, *D0, *D1, *D2, *D3;
    float* nextG0;
    int rows, cols, pitch;
    size_t frame_stride;
    int nrows, ncols, npitch;
    size_t nframe_stride;
};

#ifndef OCT_SPLIT_BARRIER
#define OCT_SPLIT_BARRIER 1
#endif
// mbarrier helpers (one phase per CTA: a tile is processed once)
__device__ __forceinline__ void bar_init(unsigned long long* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void warp_arrive(unsigned long long* b) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __forceinline__ void bar_wait(unsigned long long* b) {
    unsigned ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
    } while (!ok);
}

// One CTA per tile, 3 CTAs per SM.  Two variants were measured slower and dropped (profiles/README.md): a persistent CTA that
// kept the next tile's loads in flight in registers (128 registers -> 2 CTAs/SM, 46.7 us), and a CTA marching down four blocks
// re-using the last 2R horizontal-pass rows (13 % fewer FFMAs but spills, an extra barrier and a row shift per block: 46 us).
__global__ void __launch_bounds__(NT, OCT_CTAS) octave_kernel(const OctArgs a) {
    extern __shared__ float2 smem2[];
    float2* sIn = smem2;
    float2* sH = smem2 + OCT_IN;
    using IO = TileIO<OCT_HALO>;
    const int tid = threadIdx.x;
#if OCT_SPLIT_BARRIER
    __shared__ unsigned long long s_bar[4];
    if (tid == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) bar_init(&s_bar[q], NT / 32);
    }
#endif
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    const int f = blockIdx.z;
    const size_t foff = (size_t)f * a.frame_stride;
    {
        float2 pre[2 * IO::NSLOT];
        IO::fetch(pre, a.G0 + foff, a.rows, a.pitch, ty0, tx0, tid);
        IO::stash(pre, sIn, a.cols, tx0, tid);
    }
    __syncthreads();
    const int x = tid & 31, rg = tid >> 5;
    float g1[GRP], g2[GRP], g3[GRP], g4[GRP];
#if OCT_SPLIT_BARRIER
    // One mbarrier per scale instead of one CTA barrier between the passes: a warp signals each scale's horizontal results as it finishes
    // them, and the vertical pass of scale 4 (done first by everybody) starts as soon as ITS rows are complete -- warps with fewer
    // horizontal items no longer idle until the slowest warp has finished the last scale.
    hpass<4, OCT_HALO>(sIn, sH + H_OFF4, tid);
    warp_arrive(&s_bar[3]);
    hpass<3, OCT_HALO>(sIn, sH + H_OFF3, tid);
    warp_arrive(&s_bar[2]);
    hpass<2, OCT_HALO>(sIn, sH + H_OFF2, tid);
    warp_arrive(&s_bar[1]);
    hpass<1, OCT_HALO>(sIn, sH + H_OFF1, tid);
    warp_arrive(&s_bar[0]);
    bar_wait(&s_bar[3]);
    fir8_col<4>(sH + H_OFF4 + (rg * GRP / 2) * HP2 + x, g4);
    bar_wait(&s_bar[2]);
    fir8_col<3>(sH + H_OFF3 + (rg * GRP / 2) * HP2 + x, g3);
    bar_wait(&s_bar[1]);
    fir8_col<2>(sH + H_OFF2 + (rg * GRP / 2) * HP2 + x, g2);
    bar_wait(&s_bar[0]);
    fir8_col<1>(sH + H_OFF1 + (rg * GRP / 2) * HP2 + x, g1);
#else
    hpass<4, OCT_HALO>(sIn, sH + H_OFF4, tid);
    hpass<3, OCT_HALO>(sIn, sH + H_OFF3, tid);
    hpass<2, OCT_HALO>(sIn, sH + H_OFF2, tid);
    hpass<1, OCT_HALO>(sIn, sH + H_OFF1, tid);
    __syncthreads();
    fir8_col<1>(sH + H_OFF1 + (rg * GRP / 2) * HP2 + x, g1);
    fir8_col<2>(sH + H_OFF2 + (rg * GRP / 2) * HP2 + x, g2);
    fir8_col<3>(sH + H_OFF3 + (rg * GRP / 2) * HP2 + x, g3);
    fir8_col<4>(sH + H_OFF4 + (rg * GRP / 2) * HP2 + x, g4);
#endif

    const int gx = tx0 + x;
    constexpr int IP2 = TW + 2 * OCT_HALO + 1;
    const float* sInF = reinterpret_cast<const float*>(sIn);
    const int r0 = rg * GRP;
    if (ty0 + TH <= a.rows - 1 && tx0 + TW <= a.cols - 1) {
        // interior tile (no pixel on the last row/column, none outside): no per-pixel checks.  Every level is addressed as
        // (uniform 64-bit frame base) + (one 32-bit byte offset shared by all eight arrays), so a row costs one integer add.
        char *const G1 = reinterpret_cast<char*>(a.G1 + foff), *const G2 = reinterpret_cast<char*>(a.G2 + foff);
        char *const G3 = reinterpret_cast<char*>(a.G3 + foff), *const G4 = reinterpret_cast<char*>(a.G4 + foff);
        char *const D0 = reinterpret_cast<char*>(a.D0 + foff), *const D1 = reinterpret_cast<char*>(a.D1 + foff);
        char *const D2 = reinterpret_cast<char*>(a.D2 + foff), *const D3 = reinterpret_cast<char*>(a.D3 + foff);
        char* const NX = reinterpret_cast<char*>(a.nextG0 + (size_t)f * a.nframe_stride);
        const uint32_t pitch4 = (uint32_t)a.pitch * 4u, npitch4 = (uint32_t)a.npitch * 4u;
        uint32_t off = (uint32_t)(ty0 + r0) * pitch4 + (uint32_t)gx * 4u;
        uint32_t noff = (uint32_t)((ty0 + r0) >> 1) * npitch4 + (uint32_t)(gx >> 1) * 4u;
        const bool deep = a.G3 != nullptr, next = a.nextG0 != nullptr && !(x & 1);
        auto st = [](char* base, uint32_t o, float v) { *reinterpret_cast<float*>(base + o) = v; };
#pragma unroll
        for (int k = 0; k < GRP; ++k, off += pitch4) {
            const float g0 = sInF[il_index(OCT_HALO + r0 + k, OCT_HALO + x, IP2)];
            st(G1, off, g1[k]);
            st(G2, off, g2[k]);
            if (deep) { st(G3, off, g3[k]); st(G4, off, g4[k]); }
            st(D0, off, g1[k] - g0);
            st(D1, off, g2[k] - g1[k]);
            st(D2, off, g3[k] - g2[k]);
            st(D3, off, g4[k] - g3[k]);
            if (!(k & 1)) {
                if (next) st(NX, noff, g2[k]);
                noff += npitch4;
            }
        }
        return;
    }
    if (gx >= a.cols) return;
#pragma unroll
    for (int k = 0; k < GRP; ++k) {
        const int gy = ty0 + r0 + k;
        if (gy >= a.rows) break;
        const size_t p = foff + (size_t)gy * a.pitch + gx;
        // DoG level 0 uses the real base value; the masked copy in shared memory is zero on the last row/col.
        float g0 = sInF[il_index(OCT_HALO + r0 + k, OCT_HALO + x, IP2)];
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
constexpr int BASE_IN = (TW + 2 * BASE_HALO + 1) * (TH / 2 + BASE_HALO);
constexpr int BASE_SMEM_BYTES = (BASE_IN + (TH / 2 + 4) * HP2) * 8;

__global__ void __launch_bounds__(NT, BASE_CTAS)
    base_blur_kernel(const float* __restrict__ src, const uint8_t* __restrict__ src8, size_t src_frame_stride, int src_pitch, float* __restrict__ dst,
                     size_t dst_frame_stride, int dst_pitch, int rows, int cols, int vec) {
    extern __shared__ float2 smem2[];
    float2* sIn = smem2;
    float2* sH = smem2 + BASE_IN;
    float* sInF = reinterpret_cast<float*>(sIn);
    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    if (vec == 2) {  // uint8 source, 4-byte aligned rows
        load_tile_u8x4<BASE_HALO>(sIn, src8 + (size_t)blockIdx.z * src_frame_stride, rows, cols, src_pitch, ty0, tx0, tid);
    } else if (vec) {  // float source, aligned rows: 8-byte loads of row pairs
        using IO = TileIO<BASE_HALO>;
        float2 pre[2 * IO::NSLOT];
        IO::fetch(pre, src + (size_t)blockIdx.z * src_frame_stride, rows, src_pitch, ty0, tx0, tid);
        IO::stash(pre, sIn, cols, tx0, tid);
    } else {
        load_tile<BASE_HALO>(sInF, src ? src + (size_t)blockIdx.z * src_frame_stride : nullptr, src8 ? src8 + (size_t)blockIdx.z * src_frame_stride : nullptr,
                             rows, cols, src_pitch, ty0, tx0, tid);
    }
    __syncthreads();
    hpass<0, BASE_HALO>(sIn, sH, tid);
    __syncthreads();
    const int x = tid & 31, rg = tid >> 5;
    float g[GRP];
    fir8_col<0>(sH + (rg * GRP / 2) * HP2 + x, g);
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
    const int vec = src_u8 ? ((o0.cols % 4 == 0 && src_pitch % 4 == 0 && src_frame_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(src_u8) & 3) == 0) ? 2 : 0)
                           : (src_pitch % 4 == 0 && src_frame_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0);
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
