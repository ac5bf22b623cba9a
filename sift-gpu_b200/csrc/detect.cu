// detect.cu -- scale-space extrema, sub-pixel refinement, orientation assignment, canonical ordering.
//
// Replaces findScaleSpaceExtrema / findScaleSpaceExtremaComputer / adjustLocalExtrema / calcOrientationHist
// (reference src/sift.cpp:287-577).  Compiled with --fmad=false: every mul/add rounds separately, in the
// reference's expression order, so that given identical DoG/Gaussian levels the refinement arithmetic
// matches the CPU bit for bit (only expf/exp2 differ by library).
//
//   extrema_kernel      warp = 30 columns x 64 rows strip (+1 halo lane each side, +1 halo row above/below).  Per row
//                       the four DoG levels are loaded once (coalesced); a lane keeps the two previous rows of its column,
//                       so the max/min over the 3 rows x 3 levels of its own column are register operations, and the
//                       neighbouring columns are consulted (four shuffles) only when some lane of the warp is the
//                       extremum of its column neighbourhood beyond the threshold: "val >= all 26 neighbours" (:493-511)
//                       without divergent probing.  Survivors are appended as 32-bit scan-order keys (the compiler
//                       aggregates the atomic per warp); refine_kernel runs the Taylor refinement on them.
//   gradient_kernel     {magnitude, fastAtan2 orientation} of every pixel of G1/G2, once, for the two window consumers.
//   orientation_kernel  warp per refined point; lane = window column, the warp walks down the window rows (batches of
//                       four gathers, the next batch in flight) with a separable Gaussian weight, and votes into
//                       LANE-PRIVATE 36-bin histograms in shared memory ([bin][lane], conflict-free plain adds: shared
//                       float atomics are CAS loops on sm_100a), summed with rotated 16-byte reads; shuffle max, peak split.
//   order_scan_kernel   CTA per frame: bitonic sort of (scan-order key, index) + exclusive scan of peak counts
//                       => output slot of every (point, peak) in the reference's push_back order (:538).
#include "sift_internal.cuh"

namespace siftb200 {
namespace {

__device__ __forceinline__ int cv_round(float v) { return __float2int_rn(v); }  // cvRound: round half to even

// hal::fastAtan2 (degrees): branch-free min/max form of the scalar polynomial in oracle/oracle_prims.h (bit-identical).
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float s = (float)(180.0 / 3.1415926535897932384626433832795);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s, p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    const float ax = fabsf(x), ay = fabsf(y);
    const float mn = fminf(ax, ay), mx = fmaxf(ax, ay);
    const float c = mn / (mx + (float)2.2204460492503131e-16);
    const float c2 = c * c;
    float a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    a = ax >= ay ? a : 90.f - a;
    a = x < 0 ? 180.f - a : a;
    a = y < 0 ? 360.f - a : a;
    return a;
}

// Matx33f::solve(DECOMP_LU) = Cramer's rule with d = 1/det, zeros when det == 0 (OpenCV Matx_FastSolveOp<_,3,3,1>).
__device__ __forceinline__ void solve3(const float* a, const float* b, float* x) {
    float d = a[0] * (a[4] * a[8] - a[7] * a[5]) - a[1] * (a[3] * a[8] - a[6] * a[5]) + a[2] * (a[3] * a[7] - a[6] * a[4]);
    if (d == 0) { x[0] = x[1] = x[2] = 0; return; }
    d = 1 / d;
    x[0] = d * (b[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (b[1] * a[8] - a[5] * b[2]) + a[2] * (b[1] * a[7] - a[4] * b[2]));
    x[1] = d * (a[0] * (b[1] * a[8] - a[5] * b[2]) - b[0] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * b[2] - b[1] * a[6]));
    x[2] = d * (a[0] * (a[4] * b[2] - b[1] * a[7]) - a[1] * (a[3] * b[2] - b[1] * a[6]) + b[0] * (a[3] * a[7] - a[4] * a[6]));
}

struct Lv {
    const float* p;
    int pitch;
    __device__ __forceinline__ float at(int r, int c) const { return __ldg(p + (size_t)r * pitch + c); }
};

// adjustLocalExtrema, src/sift.cpp:287-388.  D[0..3] = DoG levels of this octave/frame.
__device__ bool adjust_local_extrema(const float* const* D, int rows, int cols, int pitch, int octv, int& layer, int& r, int& c, Refined& out) {
    const float img_scale = (float)(1. / 255);
    const float deriv_scale = img_scale * 0.5f;
    const float second_deriv_scale = img_scale;
    const float cross_deriv_scale = img_scale * 0.25f;
    float xi = 0, xr = 0, xc = 0, contr = 0;
    int i = 0;
    for (; i < kMaxInterpSteps; i++) {
        const Lv img{D[layer], pitch}, prev{D[layer - 1], pitch}, next{D[layer + 1], pitch};
        float dD[3] = {(img.at(r, c + 1) - img.at(r, c - 1)) * deriv_scale, (img.at(r + 1, c) - img.at(r - 1, c)) * deriv_scale,
                       (next.at(r, c) - prev.at(r, c)) * deriv_scale};
        float v2 = img.at(r, c) * 2;
        float dxx = (img.at(r, c + 1) + img.at(r, c - 1) - v2) * second_deriv_scale;
        float dyy = (img.at(r + 1, c) + img.at(r - 1, c) - v2) * second_deriv_scale;
        float dss = (next.at(r, c) + prev.at(r, c) - v2) * second_deriv_scale;
        float dxy = (img.at(r + 1, c + 1) - img.at(r + 1, c - 1) - img.at(r - 1, c + 1) + img.at(r - 1, c - 1)) * cross_deriv_scale;
        float dxs = (next.at(r, c + 1) - next.at(r, c - 1) - prev.at(r, c + 1) + prev.at(r, c - 1)) * cross_deriv_scale;
        float dys = (next.at(r + 1, c) - next.at(r - 1, c) - prev.at(r + 1, c) + prev.at(r - 1, c)) * cross_deriv_scale;
        float H[9] = {dxx, dxy, dxs, dxy, dyy, dys, dxs, dys, dss};
        float X[3];
        solve3(H, dD, X);
        xi = -X[2]; xr = -X[1]; xc = -X[0];
        if (fabsf(xi) < 0.5f && fabsf(xr) < 0.5f && fabsf(xc) < 0.5f) break;
        const float big = (float)(INT_MAX / 3);
        if (fabsf(xi) > big || fabsf(xr) > big || fabsf(xc) > big) return false;
        c += cv_round(xc);
        r += cv_round(xr);
        layer += cv_round(xi);
        if (layer < 1 || layer > kOctaveLayers || c < kImgBorder || c >= cols - kImgBorder || r < kImgBorder || r >= rows - kImgBorder) return false;
    }
    if (i >= kMaxInterpSteps) return false;
    {
        const Lv img{D[layer], pitch}, prev{D[layer - 1], pitch}, next{D[layer + 1], pitch};
        float dD[3] = {(img.at(r, c + 1) - img.at(r, c - 1)) * deriv_scale, (img.at(r + 1, c) - img.at(r - 1, c)) * deriv_scale,
                       (next.at(r, c) - prev.at(r, c)) * deriv_scale};
        float t = 0;
        t += dD[0] * xc; t += dD[1] * xr; t += dD[2] * xi;
        contr = img.at(r, c) * img_scale + t * 0.5f;
        if (fabsf(contr) * kOctaveLayers < 0.04f) return false;
        float v2 = img.at(r, c) * 2.f;
        float dxx = (img.at(r, c + 1) + img.at(r, c - 1) - v2) * second_deriv_scale;
        float dyy = (img.at(r + 1, c) + img.at(r - 1, c) - v2) * second_deriv_scale;
        float dxy = (img.at(r + 1, c + 1) - img.at(r + 1, c - 1) - img.at(r - 1, c + 1) + img.at(r - 1, c - 1)) * cross_deriv_scale;
        float tr = dxx + dyy;
        float det = dxx * dyy - dxy * dxy;
        const float edgeThreshold = 10.f;
        if (det <= 0 || tr * tr * edgeThreshold >= (edgeThreshold + 1) * (edgeThreshold + 1) * det) return false;
    }
    out.x = (c + xc) * (1 << octv);
    out.y = (r + xr) * (1 << octv);
    out.octave = octv + (layer << 8) + (__double2int_rn(((double)xi + 0.5) * 255) << 16);
    // powf(2.f, y): evaluated in double and rounded once so it agrees with a correctly rounded host powf
    out.size = 1.6f * (float)exp2((double)((layer + xi) / kOctaveLayers)) * (1 << octv) * 2;
    out.response = fabsf(contr);
    out.rc = ((uint32_t)r << 16) | (uint32_t)c;
    out.pad = 0;
    return true;
}

constexpr int EX_COLS = kExtremaCols;  // output columns per warp (lanes 1..30; lanes 0 and 31 are halo)
constexpr int EX_ROWS = kExtremaRows;  // output rows per warp
constexpr int EX_WARPS = 4;

#ifndef EX_MIN_CTAS
#define EX_MIN_CTAS 1
#endif
__global__ void __launch_bounds__(EX_WARPS * 32, EX_MIN_CTAS) extrema_kernel(const __grid_constant__ PyrView pv, const DetectBuf db) {
    const int lane = threadIdx.x & 31;
    const int strip = blockIdx.x * EX_WARPS + (threadIdx.x >> 5);
    if (strip >= pv.total_tiles) return;
    int o = 0;
#pragma unroll 1
    for (int k = 1; k < pv.n_oct; ++k)
        if (strip >= pv.oct[k].tile_base) o = k;
    const OctaveView& ov = pv.oct[o];
    const int t = strip - ov.tile_base;
    const int f = blockIdx.y;
    const int rows = ov.rows, cols = ov.cols, pitch = ov.pitch;
    const int c = kImgBorder + (t % ov.tiles_x) * EX_COLS - 1 + lane;
    const int r_begin = kImgBorder + (t / ov.tiles_x) * EX_ROWS;
    const int r_end = min(r_begin + EX_ROWS, rows - kImgBorder);  // exclusive
    const size_t foff = (size_t)f * ov.frame_stride;
    const bool col_in = c < cols;
    const bool col_out = lane >= 1 && lane <= EX_COLS && c < cols - kImgBorder;
    const float* p0 = ov.D[0] + foff + (size_t)(r_begin - 1) * pitch + (col_in ? c : 0);
    const float* p1 = ov.D[1] + foff + (size_t)(r_begin - 1) * pitch + (col_in ? c : 0);
    const float* p2 = ov.D[2] + foff + (size_t)(r_begin - 1) * pitch + (col_in ? c : 0);
    const float* p3 = ov.D[3] + foff + (size_t)(r_begin - 1) * pitch + (col_in ? c : 0);

    // Column first: the lane keeps the two previous rows of its column (raw values, four levels); per row the max/min over the three rows
    // of each level and then over the three levels of a layer are plain register operations.  Only a pixel that is the extremum of its
    // own column neighbourhood (9 values) can be one of all 27, so the neighbouring columns are consulted -- four shuffles -- only for
    // (row, layer) pairs where some lane of the warp passes that filter and the |val| > 8 test; about half of them on the benchmark frames.
    float ra[4], rb[4];  // rows y-2, y-1
#pragma unroll
    for (int l = 0; l < 4; ++l) ra[l] = rb[l] = 0.f;
    float nx[4] = {__ldg(p0), __ldg(p1), __ldg(p2), __ldg(p3)};  // row r_begin-1
#pragma unroll 1
    for (int y = r_begin - 1; y <= r_end; ++y) {
        const float v[4] = {nx[0], nx[1], nx[2], nx[3]};
        if (y < r_end) {  // prefetch row y+1 (<= rows-5) while row y is processed
            p0 += pitch; p1 += pitch; p2 += pitch; p3 += pitch;
            nx[0] = __ldg(p0); nx[1] = __ldg(p1); nx[2] = __ldg(p2); nx[3] = __ldg(p3);
        }
        if (y >= r_begin + 1) {  // rows y-2, y-1, y are in registers: test row y-1
            float cmx[4], cmn[4];
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                cmx[l] = fmaxf(ra[l], fmaxf(rb[l], v[l]));
                cmn[l] = fminf(ra[l], fminf(rb[l], v[l]));
            }
#pragma unroll
            for (int layer0 = 1; layer0 <= kOctaveLayers; ++layer0) {
                const float val = rb[layer0];
                const float A = fmaxf(cmx[layer0 - 1], fmaxf(cmx[layer0], cmx[layer0 + 1]));
                const float B = fminf(cmn[layer0 - 1], fminf(cmn[layer0], cmn[layer0 + 1]));
                // |val| > 8 (the literal threshold, :564) and val >= / <= all 26 neighbours (non-strict, :494-511)
                const bool own = col_out && ((val > 8.0f && val >= A) || (val < -8.0f && val <= B));
                if (!__any_sync(0xffffffffu, own)) continue;
                const float Al = __shfl_up_sync(0xffffffffu, A, 1), Ar = __shfl_down_sync(0xffffffffu, A, 1);
                const float Bl = __shfl_up_sync(0xffffffffu, B, 1), Br = __shfl_down_sync(0xffffffffu, B, 1);
                const bool hit = own && ((val > 8.0f && val >= fmaxf(Al, Ar)) || (val < -8.0f && val <= fminf(Bl, Br)));
                if (hit) {
                    const int slot = atomicAdd(db.n_cand + f, 1);
                    if (slot < db.cap_c)
                        db.cand[(size_t)f * db.cap_c + slot] = ((uint32_t)o << 27) | ((uint32_t)(layer0 - 1) << 26) | ((uint32_t)(y - 1) << 13) | (uint32_t)c;
                }
            }
        }
#pragma unroll
        for (int l = 0; l < 4; ++l) { ra[l] = rb[l]; rb[l] = v[l]; }
    }
}

// adjustLocalExtrema for every candidate: thread per candidate, survivors appended as 32-byte records.
__global__ void __launch_bounds__(128) refine_kernel(const __grid_constant__ PyrView pv, const DetectBuf db) {
    const int f = blockIdx.y;
    int n = db.n_cand[f];
    if (n > db.cap_c) n = db.cap_c;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t key = db.cand[(size_t)f * db.cap_c + i];
        const int o = key >> 27, layer0 = ((key >> 26) & 1) + 1, r = (key >> 13) & 8191, c = key & 8191;
        const OctaveView& ov = pv.oct[o];
        const size_t foff = (size_t)f * ov.frame_stride;
        const float* D[4] = {ov.D[0] + foff, ov.D[1] + foff, ov.D[2] + foff, ov.D[3] + foff};
        int r1 = r, c1 = c, layer = layer0;
        Refined rec;
        if (!adjust_local_extrema(D, ov.rows, ov.cols, ov.pitch, o, layer, r1, c1, rec)) continue;
        rec.key = key;
        const int slot = atomicAdd(db.n_refined + f, 1);
        if (slot < db.cap_r) db.refined[(size_t)f * db.cap_r + slot] = rec;
    }
}

// ---- gradient maps ------------------------------------------------------------------------------------------------
// Both calcOrientationHist (:413-426) and calcSIFTDescriptor (:623-633) evaluate, per window sample,
//   dx = I(y,x+1)-I(y,x-1), dy = I(y-1,x)-I(y+1,x), Mag = sqrt(dx*dx+dy*dy), Ori = fastAtan2(dy,dx)
// on a Gaussian level.  Windows of neighbouring keypoints overlap and a frame has more window samples (~9 M) than level
// pixels (2*sumP = 5.5 M), so the pair {Mag, Ori} is computed ONCE per pixel here -- same operations, bit-identical values --
// and the two consumers gather one float2 per sample instead of four floats plus the atan2/sqrt arithmetic.
// Warp = strip of 64 columns x GR_ROWS rows of ONE level, 256-byte aligned; a lane owns two adjacent columns: one 8-byte load per
// row, one 16-byte store per row ({Mag, Ori} of both pixels), and the horizontal neighbours of the pair come from two shuffles
// (the strip's outer neighbours are fetched by lanes 0 and 31).  All row loads are issued before any arithmetic.  Pixels of the
// first/last row and column get no meaningful value and are never sampled (0 < y < rows-1, 0 < x < cols-1 in both consumers);
// the 16-byte store may cover such a pixel or the row padding.
constexpr int GR_COLS = kGradCols, GR_ROWS = kGradRows, GR_WARPS = 4;
static_assert(GR_COLS == 64, "two columns per lane");

__device__ __forceinline__ float2 grad_px(float left, float right, float up, float down) {
    const float dx = right - left, dy = up - down;
    return make_float2(sqrtf(dx * dx + dy * dy), fast_atan2_deg(dy, dx));
}

// levels lv0 + blockIdx.z: {1, 2} in the fused pipeline, all five in the stage-level API
__global__ void __launch_bounds__(GR_WARPS * 32) gradient_kernel(const __grid_constant__ PyrView pv, int lv0) {
    const int lane = threadIdx.x & 31;
    const int strip = blockIdx.x * GR_WARPS + (threadIdx.x >> 5);
    if (strip >= pv.total_grad_tiles) return;
    int o = 0;
#pragma unroll 1
    for (int k = 1; k < pv.n_oct; ++k)
        if (strip >= pv.oct[k].grad_tile_base) o = k;
    const OctaveView& ov = pv.oct[o];
    const int t = strip - ov.grad_tile_base;
    const int rows = ov.rows, cols = ov.cols, pitch = ov.pitch;
    const int x = (t % ov.grad_tiles_x) * GR_COLS + 2 * lane;
    const int y0 = 1 + (t / ov.grad_tiles_x) * GR_ROWS;  // only 0 < y < rows-1, 0 < x < cols-1 is ever sampled
    const int y1 = min(y0 + GR_ROWS, rows - 1);
    const size_t foff = (size_t)blockIdx.y * ov.frame_stride;
    const int level = lv0 + blockIdx.z;
    const float* __restrict__ G = ov.G[level] + foff;
    float2* __restrict__ MO = ov.MO[level] + foff;
    const int xl = min(x, pitch - 2);  // loads stay inside the row (pitch is even); columns >= cols only feed pixels that are not stored
    const int xh = lane == 0 ? max(x - 1, 0) : min(x + 2, pitch - 1);  // outer neighbour column of the strip (lanes 0 and 31)
    const bool has_h = lane == 0 || lane == 31;
    float2 v[GR_ROWS + 2];
    float h[GR_ROWS];
#pragma unroll
    for (int k = 0; k < GR_ROWS + 2; ++k) v[k] = __ldg(reinterpret_cast<const float2*>(G + (size_t)min(y0 - 1 + k, rows - 1) * pitch + xl));
#pragma unroll
    for (int k = 0; k < GR_ROWS; ++k) h[k] = has_h ? __ldg(G + (size_t)min(y0 + k, rows - 1) * pitch + xh) : 0.f;
    float4* out = reinterpret_cast<float4*>(MO + (size_t)y0 * pitch + x);
    const bool col_ok = x < cols;
#pragma unroll
    for (int k = 0; k < GR_ROWS; ++k) {
        const float2 c = v[k + 1];
        float left = __shfl_up_sync(0xffffffffu, c.y, 1), right = __shfl_down_sync(0xffffffffu, c.x, 1);
        if (lane == 0) left = h[k];
        if (lane == 31) right = h[k];
        const float2 p0 = grad_px(left, c.y, v[k].x, v[k + 2].x);
        const float2 p1 = grad_px(c.x, right, v[k].y, v[k + 2].y);
        if (col_ok && y0 + k < y1) *out = make_float4(p0.x, p0.y, p1.x, p1.y);
        out += pitch / 2;
    }
}

// ---- orientation: calcOrientationHist + peak logic, src/sift.cpp:389-458, 518-541 ----------------------------
constexpr int ORI_WARPS = 8;
#ifndef ORI_BATCH_
#define ORI_BATCH_ 4
#endif
constexpr int ORI_BATCH = ORI_BATCH_;  // window rows gathered per batch (two batches in flight)

#ifndef ORI_MIN_CTAS
#define ORI_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(ORI_WARPS * 32, ORI_MIN_CTAS) orientation_kernel(const __grid_constant__ PyrView pv, const DetectBuf db) {
    __shared__ float s_tmp[ORI_WARPS][kOriBins + 4];
    __shared__ float s_hist[ORI_WARPS][kOriBins];
    __shared__ __align__(16) float s_priv[ORI_WARPS][kOriBins * 32];  // lane-private bins, [bin][lane]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.y;
    int n = db.n_refined[f];
    if (n > db.cap_r) n = db.cap_r;
    float* temphist = s_tmp[warp] + 2;
    float* hist = s_hist[warp];
    for (int i = blockIdx.x * ORI_WARPS + warp; i < n; i += gridDim.x * ORI_WARPS) {
        const Refined rec = db.refined[(size_t)f * db.cap_r + i];
        const int o = rec.octave & 255, layer = (rec.octave >> 8) & 255;
        const OctaveView& ov = pv.oct[o];
        const float2* mo = ov.MO[layer] + (size_t)f * ov.frame_stride;  // {magnitude, orientation} of G[layer]
        const int rows = ov.rows, cols = ov.cols, pitch = ov.pitch;
        const int py = rec.rc >> 16, px = rec.rc & 0xffff;
        const float scl_octv = rec.size * 0.5f / (1 << o);
        const int radius = cv_round(4.5f * scl_octv);
        const float sigma = 1.5f * scl_octv;
        const float expf_scale = -1.f / (2.f * sigma * sigma);
        float* priv = s_priv[warp] + lane;
#pragma unroll
        for (int b = 0; b < kOriBins; ++b) priv[b * 32] = 0.f;
        // window (2*radius+1)^2; rows/cols outside 0 < y < rows-1, 0 < x < cols-1 are skipped (:405,410).  Lane = window column (passes of 32
        // columns; a pipeline window has at most 35), the warp walks down the rows: the address advances by the pitch, the column part of
        // the weight is a per-lane constant and the row part is warp-uniform, so a sample costs a load, two multiplies, the bin and the
        // lane-private add.  The Gaussian weight exp((i^2 + j^2) * s) is taken as exp(i^2 s) * exp(j^2 s) from one 32-entry table held
        // across the lanes (radius < 32; a few ulp from the exp of the sum, far below what moves a peak); larger windows use the direct form.
        const int i_lo = max(-radius, 1 - py), i_hi = min(radius, rows - 2 - py);
        const int j_lo = max(-radius, 1 - px), j_hi = min(radius, cols - 2 - px);
        const bool use_tab = radius < 32;
        const float tab = expf((float)(lane * lane) * expf_scale);  // entry |k| = lane
        if (i_lo <= i_hi) {
            for (int cb = j_lo; cb <= j_hi; cb += 32) {
                const int j = cb + lane;
                const bool col_ok = j <= j_hi;
                const int jc = col_ok ? j : j_hi;  // clamped: a readable pixel, its vote is dropped
                const float wtab = __shfl_sync(0xffffffffu, tab, abs(jc) & 31);  // every lane takes part in the shuffle
                const float wcol = !col_ok ? 0.f : use_tab ? wtab : 1.f;          // a clamped lane votes zeros
                const float2* ptr = mo + (size_t)(py + i_lo) * pitch + (px + jc);
                auto vote = [&](const float2 g, int ii) {
                    const float wt = __shfl_sync(0xffffffffu, tab, abs(ii) & 31);
                    const float wrow = use_tab ? wt : expf((float)(ii * ii + jc * jc) * expf_scale);
                    int bin = cv_round((kOriBins / 360.f) * g.y);
                    if (bin >= kOriBins) bin -= kOriBins;
                    if (bin < 0) bin += kOriBins;
                    priv[bin * 32] += (wrow * wcol) * g.x;
                };
                // batches of ORI_BATCH rows; the gathers of the next batch are issued before the votes of the current one
                const int nrows_w = i_hi - i_lo + 1;
                float2 cur[ORI_BATCH], nxt[ORI_BATCH];
#pragma unroll
                for (int u = 0; u < ORI_BATCH; ++u) cur[u] = u < nrows_w ? __ldg(ptr + (size_t)u * pitch) : make_float2(0.f, 0.f);
                for (int r0 = 0; r0 < nrows_w; r0 += ORI_BATCH) {
                    ptr += (size_t)ORI_BATCH * pitch;
#pragma unroll
                    for (int u = 0; u < ORI_BATCH; ++u) nxt[u] = r0 + ORI_BATCH + u < nrows_w ? __ldg(ptr + (size_t)u * pitch) : make_float2(0.f, 0.f);
#pragma unroll
                    for (int u = 0; u < ORI_BATCH; ++u)
                        if (r0 + u < nrows_w) vote(cur[u], i_lo + r0 + u);
#pragma unroll
                    for (int u = 0; u < ORI_BATCH; ++u) cur[u] = nxt[u];
                }
            }
        }
        __syncwarp();
        for (int b = lane; b < kOriBins; b += 32) {  // 16-byte reads of the bin's 32 copies, rotated by the lane: all banks distinct
            const float* col = s_priv[warp] + b * 32;
            float acc = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 t = *reinterpret_cast<const float4*>(col + (((q + lane) & 7) << 2));
                acc += (t.x + t.y) + (t.z + t.w);
            }
            temphist[b] = acc;
        }
        __syncwarp();
        if (lane == 0) {
            temphist[-1] = temphist[kOriBins - 1];
            temphist[-2] = temphist[kOriBins - 2];
            temphist[kOriBins] = temphist[0];
            temphist[kOriBins + 1] = temphist[1];
        }
        __syncwarp();
        float mx = -3.4e38f;
        for (int b = lane; b < kOriBins; b += 32) {
            const float h = (temphist[b - 2] + temphist[b + 2]) * (1.f / 16.f) + (temphist[b - 1] + temphist[b + 1]) * (4.f / 16.f) + temphist[b] * (6.f / 16.f);
            hist[b] = h;
            mx = fmaxf(mx, h);
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
        __syncwarp();
        const float mag_thr = mx * 0.8f;
        int count = 0;
        float* ang_out = db.angles + ((size_t)f * db.cap_r + i) * kMaxPeaks;
        for (int base = 0; base < kOriBins; base += 32) {
            const int j = base + lane;
            bool pk = false;
            float angle = 0.f;
            if (j < kOriBins) {
                const int l = j > 0 ? j - 1 : kOriBins - 1;
                const int r2 = j < kOriBins - 1 ? j + 1 : 0;
                const float hj = hist[j], hl = hist[l], hr = hist[r2];
                if (hj > hl && hj > hr && hj >= mag_thr) {
                    float bin = j + 0.5f * (hl - hr) / (hl - 2 * hj + hr);
                    bin = bin < 0 ? kOriBins + bin : bin >= kOriBins ? bin - kOriBins : bin;
                    angle = 360.f - (360.f / kOriBins) * bin;
                    if (fabsf(angle - 360.f) < 1.1920928955078125e-7f) angle = 0.f;
                    pk = true;
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, pk);
            if (pk) ang_out[count + __popc(m & ((1u << lane) - 1))] = angle;
            count += __popc(m);
        }
        if (lane == 0) db.n_peaks[(size_t)f * db.cap_r + i] = count;
        __syncwarp();
    }
}

// ---- canonical order + output offsets -----------------------------------------------------------------------
constexpr int SORT_THREADS = 1024;
constexpr int SORT_SMEM_ELEMS = 8192;  // 64 KB of (key<<32 | index)

__global__ void __launch_bounds__(SORT_THREADS) order_scan_kernel(const DetectBuf db, int* __restrict__ counts_out) {
    extern __shared__ unsigned long long s_buf[];
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int f = blockIdx.x, tid = threadIdx.x;
    int n = db.n_refined[f];
    if (n > db.cap_r) n = db.cap_r;
    int npad = 1;
    while (npad < n) npad <<= 1;
    unsigned long long* buf = npad <= SORT_SMEM_ELEMS ? s_buf : db.sort_tmp + (size_t)f * db.cap_r_pow2;
    const Refined* rec = db.refined + (size_t)f * db.cap_r;
    for (int i = tid; i < npad; i += SORT_THREADS) buf[i] = i < n ? ((unsigned long long)rec[i].key << 32) | (unsigned)i : ~0ull;
    __syncthreads();
    for (int k = 2; k <= npad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < npad; i += SORT_THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = buf[i], b = buf[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { buf[i] = b; buf[ixj] = a; }
                }
            }
            __syncthreads();
        }
    // exclusive scan of peak counts in sorted order
    if (tid == 0) s_carry = 0;
    __syncthreads();
    const int* np = db.n_peaks + (size_t)f * db.cap_r;
    int* order = db.order + (size_t)f * db.cap_r;
    int* off = db.kp_offset + (size_t)f * db.cap_r;
    const int lane = tid & 31, warp = tid >> 5;
    for (int base = 0; base < n; base += SORT_THREADS) {
        const int p = base + tid;
        int idx = -1, v = 0;
        if (p < n) { idx = (int)(buf[p] & 0xffffffffu); v = np[idx]; }
        int incl = v;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, s);
            if (lane >= s) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, s);
                if (lane >= s) w += t;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const int carry = s_carry;
        const int excl = carry + (warp ? s_warp[warp - 1] : 0) + incl - v;
        if (p < n) { order[p] = idx; off[idx] = excl; }
        __syncthreads();
        if (tid == SORT_THREADS - 1) s_carry = carry + s_warp[31];
        __syncthreads();
    }
    if (tid == 0) {
        int total = s_carry;
        db.n_kp[f] = total;  // what the descriptor stage iterates over
        // a full candidate / refined list means keypoints were dropped and the true count is unknown: report the sentinel cap_r + 1 (above
        // any admissible cap); n_kp keeps the number of records that do exist
        if (db.n_refined[f] > db.cap_r || db.n_cand[f] > db.cap_c) total = db.cap_r + 1;
        counts_out[f] = total;
    }
}

}  // namespace

void init_detect_kernels() { cudaFuncSetAttribute(order_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_SMEM_ELEMS * 8); }

int launch_extrema(const PyrView& pv, const DetectBuf& db, int n_frames, cudaStream_t st) {
    cudaMemsetAsync(db.n_refined, 0, sizeof(int) * n_frames, st);
    cudaMemsetAsync(db.n_cand, 0, sizeof(int) * n_frames, st);
    dim3 grid((pv.total_tiles + EX_WARPS - 1) / EX_WARPS, n_frames);
    extrema_kernel<<<grid, EX_WARPS * 32, 0, st>>>(pv, db);
    refine_kernel<<<dim3(48, n_frames), 128, 0, st>>>(pv, db);
    return 2;
}

int launch_gradient(const PyrView& pv, int n_frames, cudaStream_t st) {
    const bool full = pv.oct[0].MO[0] != nullptr;
    const dim3 grid((pv.total_grad_tiles + GR_WARPS - 1) / GR_WARPS, n_frames, full ? kNumScales : kOctaveLayers);
    gradient_kernel<<<grid, GR_WARPS * 32, 0, st>>>(pv, full ? 0 : 1);
    return 1;
}

int launch_orientation(const PyrView& pv, const DetectBuf& db, int n_frames, cudaStream_t st) {
    // whole waves: about 64 CTAs per frame, rounded so that the launch is a multiple of the CTAs the device holds at once (the last,
    // partly filled wave of a 64 x 32 launch was 14 % of the kernel)
    static int g_ori_ctas = getenv("SIFT_B200_ORI_CTAS") ? atoi(getenv("SIFT_B200_ORI_CTAS")) : 0;
    const int resident = num_sms() * ORI_MIN_CTAS;
    const int waves = (64 * n_frames + resident - 1) / resident;
    const int per_frame = g_ori_ctas > 0 ? g_ori_ctas : (waves * resident + n_frames - 1) / n_frames;
    dim3 grid(per_frame, n_frames);
    orientation_kernel<<<grid, ORI_WARPS * 32, 0, st>>>(pv, db);
    return 1;
}

int launch_order_scan(const DetectBuf& db, int n_frames, int* d_counts, cudaStream_t st) {
    order_scan_kernel<<<n_frames, SORT_THREADS, SORT_SMEM_ELEMS * 8, st>>>(db, d_counts);
    return 1;
}

}  // namespace siftb200
