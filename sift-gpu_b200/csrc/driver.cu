// driver.cu -- the two stages of the reference's DRIVER (src/main.cpp) that sit either side of SIFT_NCL and are not matching:
//
//   resize_linear_u8_kernel   readImage's resize(img, img, Size(960,960)) (src/main.cpp:83): cv::resize(INTER_LINEAR) on 8-bit pixels,
//                             i.e. OpenCV's fixed-point bilinear -- 11-bit coefficients saturate_cast<short>(c * 2048), horizontal pass in
//                             int, vertical pass ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2 -- restated from OpenCV's
//                             published resize.cpp and pinned bit for bit against cv2.resize (tests).  One thread per output pixel.
//   homography_*              findHomography(obj, scene, RANSAC) (src/main.cpp:55): K hypotheses evaluated in parallel, one warp each --
//                             4 distinct correspondences from a counter-based generator, degenerate samples rejected, the 8x8 system of the
//                             4-point homography solved in double, inliers (squared reprojection error <= thresh^2, float arithmetic as in
//                             OpenCV's computeError) counted over all correspondences by the warp's lanes.  The consensus set of the best
//                             hypothesis (ties: lowest index, so the result is deterministic) is refitted on the host in double: normalised
//                             DLT over the inliers, then Levenberg-Marquardt on the reprojection error, as OpenCV does after its RANSAC loop.
//                             OpenCV's random sequence is not reproduced: parity is "same consensus set and the same refit", not bit identity.
#include <math.h>

#include <algorithm>
#include <vector>

#include "sift_internal.cuh"

namespace siftb200 {
namespace {

__global__ void __launch_bounds__(256) resize_linear_u8_kernel(const uint8_t* __restrict__ src, int srows, int scols, int cn, uint8_t* __restrict__ dst,
                                                               int drows, int dcols, const int* __restrict__ xofs, const short* __restrict__ ialpha,
                                                               const int* __restrict__ yofs, const short* __restrict__ ibeta) {
    const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
    if (dx >= dcols) return;
    const int sx0 = xofs[dx], sx1 = min(sx0 + 1, scols - 1);
    const int a0 = ialpha[2 * dx], a1 = ialpha[2 * dx + 1];
    const int sy = yofs[dy];
    const int y0 = min(max(sy, 0), srows - 1), y1 = min(max(sy + 1, 0), srows - 1);
    const int b0 = ibeta[2 * dy], b1 = ibeta[2 * dy + 1];
    const uint8_t* r0 = src + (size_t)y0 * scols * cn;
    const uint8_t* r1 = src + (size_t)y1 * scols * cn;
    uint8_t* out = dst + ((size_t)dy * dcols + dx) * cn;
    for (int c = 0; c < cn; ++c) {
        const int h0 = r0[sx0 * cn + c] * a0 + r0[sx1 * cn + c] * a1;
        const int h1 = r1[sx0 * cn + c] * a0 + r1[sx1 * cn + c] * a1;
        const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
        out[c] = (uint8_t)min(max(v, 0), 255);
    }
}

// cvRound(float) = round half to even; saturate_cast<short>
short sat_short(float v) {
    const long r = lrintf(v);
    return (short)(r < -32768 ? -32768 : r > 32767 ? 32767 : r);
}

// ---- RANSAC homography -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mix32(uint32_t x) {  // counter-based generator (lowbias32 finaliser)
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// H (h33 = 1) from 4 correspondences: 8x8 Gaussian elimination with partial pivoting in double.  false: singular.
__device__ bool homography_from4(const float2* s, const float2* d, double* H) {
    double A[8][9];
    for (int i = 0; i < 4; ++i) {
        const double x = s[i].x, y = s[i].y, X = d[i].x, Y = d[i].y;
        double r0[9] = {x, y, 1, 0, 0, 0, -x * X, -y * X, X};
        double r1[9] = {0, 0, 0, x, y, 1, -x * Y, -y * Y, Y};
        for (int j = 0; j < 9; ++j) { A[2 * i][j] = r0[j]; A[2 * i + 1][j] = r1[j]; }
    }
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {  // rolled on purpose: one lane per hypothesis runs this, code size matters more than speed
        int piv = c;
#pragma unroll 1
        for (int r = c + 1; r < 8; ++r)
            if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
        if (fabs(A[piv][c]) < 1e-12) return false;
        if (piv != c)
            for (int j = c; j < 9; ++j) { const double t = A[c][j]; A[c][j] = A[piv][j]; A[piv][j] = t; }
        const double inv = 1.0 / A[c][c];
#pragma unroll 1
        for (int r = c + 1; r < 8; ++r) {
            const double f = A[r][c] * inv;
#pragma unroll 1
            for (int j = c; j < 9; ++j) A[r][j] -= f * A[c][j];
        }
    }
#pragma unroll 1
    for (int r = 7; r >= 0; --r) {
        double v = A[r][8];
#pragma unroll 1
        for (int j = r + 1; j < 8; ++j) v -= A[r][j] * H[j];
        H[r] = v / A[r][r];
    }
    H[8] = 1.0;
    return true;
}

// OpenCV's sample check for homographies (haveCollinearPoints + the orientation test of checkSubset): reject a sample with three
// nearly collinear points in either image, or one whose two point quadruples are oriented differently.
__device__ bool sample_ok(const float2* s, const float2* d) {
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const float2* p = pass ? d : s;
#pragma unroll 1
        for (int i = 0; i < 4; ++i)
#pragma unroll 1
            for (int j = i + 1; j < 4; ++j)
#pragma unroll 1
                for (int k = j + 1; k < 4; ++k) {
                    const double dx1 = p[j].x - p[i].x, dy1 = p[j].y - p[i].y, dx2 = p[k].x - p[i].x, dy2 = p[k].y - p[i].y;
                    if (fabs(dx2 * dy1 - dy2 * dx1) <= 1.1920928955078125e-7 * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2))) return false;
                }
    }
    int neg = 0;
    for (int i = 0; i < 4; ++i) {  // sign of det[p_j p_k p_l] must agree between the two images for every triple
        const int j = (i + 1) & 3, k = (i + 2) & 3;
        const double a = ((double)s[j].x - s[i].x) * ((double)s[k].y - s[i].y) - ((double)s[j].y - s[i].y) * ((double)s[k].x - s[i].x);
        const double b = ((double)d[j].x - d[i].x) * ((double)d[k].y - d[i].y) - ((double)d[j].y - d[i].y) * ((double)d[k].x - d[i].x);
        neg += (a < 0) != (b < 0);
    }
    return neg == 0 || neg == 4;
}

// squared reprojection error in float, as OpenCV's HomographyEstimatorCallback::computeError
__device__ __forceinline__ float reproj_err(const float* H, float2 s, float2 d) {
    const float ww = 1.f / (H[6] * s.x + H[7] * s.y + 1.f);
    const float dx = (H[0] * s.x + H[1] * s.y + H[2]) * ww - d.x;
    const float dy = (H[3] * s.x + H[4] * s.y + H[5]) * ww - d.y;
    return dx * dx + dy * dy;
}

// one warp per hypothesis: models[k] = 9 floats (H as float, what OpenCV's error function sees), counts[k] = inliers (-1: no model)
__global__ void __launch_bounds__(128) homography_hypotheses_kernel(const float2* __restrict__ src, const float2* __restrict__ dst, int n, int n_hyp,
                                                                    float thresh2, uint32_t seed, float* __restrict__ models, int* __restrict__ counts) {
    const int k = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k >= n_hyp) return;
    float Hf[9];
    int ok = 0;
    if (lane == 0) {
        float2 s[4], d[4];
        int idx[4];
#pragma unroll 1
        for (int attempt = 0; attempt < 16 && !ok; ++attempt) {
            uint32_t ctr = seed ^ mix32((uint32_t)k * 64u + (uint32_t)attempt);
            bool distinct = true;
            for (int i = 0; i < 4; ++i) {
                ctr = mix32(ctr + 0x9e3779b9u * (uint32_t)(i + 1));
                idx[i] = (int)(((uint64_t)ctr * (uint64_t)n) >> 32);
                for (int j = 0; j < i; ++j) distinct = distinct && idx[j] != idx[i];
            }
            if (!distinct) continue;
            for (int i = 0; i < 4; ++i) { s[i] = src[idx[i]]; d[i] = dst[idx[i]]; }
            if (!sample_ok(s, d)) continue;
            double H[9];
            if (!homography_from4(s, d, H)) continue;
            for (int i = 0; i < 9; ++i) Hf[i] = (float)H[i];
            ok = 1;
        }
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
#pragma unroll
    for (int i = 0; i < 9; ++i) Hf[i] = __shfl_sync(0xffffffffu, Hf[i], 0);
    int cnt = 0;
    if (ok)
        for (int i = lane; i < n; i += 32) cnt += reproj_err(Hf, src[i], dst[i]) <= thresh2;
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, sft);
    if (lane == 0) {
        counts[k] = ok ? cnt : -1;
        for (int i = 0; i < 9; ++i) models[(size_t)k * 9 + i] = Hf[i];
    }
}

__global__ void homography_mask_kernel(const float2* __restrict__ src, const float2* __restrict__ dst, int n, const float* __restrict__ model, float thresh2,
                                       uint8_t* __restrict__ mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float H[9];
    for (int j = 0; j < 9; ++j) H[j] = model[j];
    mask[i] = reproj_err(H, src[i], dst[i]) <= thresh2 ? 1 : 0;
}

// ---- host refit (double) -----------------------------------------------------------------------------------------------------
// symmetric 9x9 eigen-decomposition by cyclic Jacobi rotations; returns the eigenvector of the smallest eigenvalue
void smallest_eigenvector9(double A[9][9], double v_out[9]) {
    double V[9][9] = {};
    for (int i = 0; i < 9; ++i) V[i][i] = 1;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0;
        for (int p = 0; p < 9; ++p)
            for (int q = p + 1; q < 9; ++q) off += A[p][q] * A[p][q];
        if (off < 1e-300) break;
        for (int p = 0; p < 9; ++p)
            for (int q = p + 1; q < 9; ++q) {
                if (fabs(A[p][q]) < 1e-300) continue;
                const double theta = (A[q][q] - A[p][p]) / (2 * A[p][q]);
                const double t = (theta >= 0 ? 1 : -1) / (fabs(theta) + sqrt(theta * theta + 1));
                const double c = 1 / sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < 9; ++k) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
                for (int k = 0; k < 9; ++k) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
                for (int k = 0; k < 9; ++k) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq; }
            }
    }
    int best = 0;
    for (int i = 1; i < 9; ++i)
        if (A[i][i] < A[best][best]) best = i;
    for (int i = 0; i < 9; ++i) v_out[i] = V[i][best];
}

// normalised DLT over m correspondences (OpenCV's HomographyEstimatorCallback::runKernel: centroid + mean-absolute-deviation scaling)
bool dlt_refit(const std::vector<float2>& s, const std::vector<float2>& d, double H[9]) {
    const int m = (int)s.size();
    if (m < 4) return false;
    double cs[2] = {0, 0}, cd[2] = {0, 0};
    for (int i = 0; i < m; ++i) { cs[0] += s[i].x; cs[1] += s[i].y; cd[0] += d[i].x; cd[1] += d[i].y; }
    cs[0] /= m; cs[1] /= m; cd[0] /= m; cd[1] /= m;
    double ss[2] = {0, 0}, sd[2] = {0, 0};
    for (int i = 0; i < m; ++i) { ss[0] += fabs(s[i].x - cs[0]); ss[1] += fabs(s[i].y - cs[1]); sd[0] += fabs(d[i].x - cd[0]); sd[1] += fabs(d[i].y - cd[1]); }
    if (fabs(ss[0]) < 1e-300 || fabs(ss[1]) < 1e-300 || fabs(sd[0]) < 1e-300 || fabs(sd[1]) < 1e-300) return false;
    ss[0] = m / ss[0]; ss[1] = m / ss[1]; sd[0] = m / sd[0]; sd[1] = m / sd[1];
    double L[9][9] = {};
    for (int i = 0; i < m; ++i) {
        const double x = (s[i].x - cs[0]) * ss[0], y = (s[i].y - cs[1]) * ss[1];
        const double X = (d[i].x - cd[0]) * sd[0], Y = (d[i].y - cd[1]) * sd[1];
        const double Lx[9] = {x, y, 1, 0, 0, 0, -X * x, -X * y, -X};
        const double Ly[9] = {0, 0, 0, x, y, 1, -Y * x, -Y * y, -Y};
        for (int j = 0; j < 9; ++j)
            for (int k = 0; k < 9; ++k) L[j][k] += Lx[j] * Lx[k] + Ly[j] * Ly[k];
    }
    double h[9];
    smallest_eigenvector9(L, h);
    // H = inv(T_dst) * H0 * T_src
    const double invD[9] = {1 / sd[0], 0, cd[0], 0, 1 / sd[1], cd[1], 0, 0, 1};
    const double Ts[9] = {ss[0], 0, -cs[0] * ss[0], 0, ss[1], -cs[1] * ss[1], 0, 0, 1};
    double t[9], r[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { t[i * 3 + j] = 0; for (int k = 0; k < 3; ++k) t[i * 3 + j] += invD[i * 3 + k] * h[k * 3 + j]; }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { r[i * 3 + j] = 0; for (int k = 0; k < 3; ++k) r[i * 3 + j] += t[i * 3 + k] * Ts[k * 3 + j]; }
    if (fabs(r[8]) < 1e-300) return false;
    for (int i = 0; i < 9; ++i) H[i] = r[i] / r[8];
    return true;
}

double reproj_cost(const std::vector<float2>& s, const std::vector<float2>& d, const double* h) {
    double c = 0;
    for (size_t i = 0; i < s.size(); ++i) {
        const double ww = 1.0 / (h[6] * s[i].x + h[7] * s[i].y + 1.0);
        const double ex = (h[0] * s[i].x + h[1] * s[i].y + h[2]) * ww - d[i].x, ey = (h[3] * s[i].x + h[4] * s[i].y + h[5]) * ww - d[i].y;
        c += ex * ex + ey * ey;
    }
    return c;
}

// Levenberg-Marquardt on the 8 free parameters (h33 = 1), reprojection error over the inliers (OpenCV's HomographyRefineCallback)
void lm_refine(const std::vector<float2>& s, const std::vector<float2>& d, double H[9], int iters) {
    double lambda = 1e-3, cost = reproj_cost(s, d, H);
    for (int it = 0; it < iters; ++it) {
        double JtJ[8][8] = {}, Jte[8] = {};
        for (size_t i = 0; i < s.size(); ++i) {
            const double x = s[i].x, y = s[i].y;
            const double ww = 1.0 / (H[6] * x + H[7] * y + 1.0);
            const double xi = (H[0] * x + H[1] * y + H[2]) * ww, yi = (H[3] * x + H[4] * y + H[5]) * ww;
            const double ex = xi - d[i].x, ey = yi - d[i].y;
            const double Jx[8] = {x * ww, y * ww, ww, 0, 0, 0, -x * ww * xi, -y * ww * xi};
            const double Jy[8] = {0, 0, 0, x * ww, y * ww, ww, -x * ww * yi, -y * ww * yi};
            for (int a = 0; a < 8; ++a) {
                Jte[a] += Jx[a] * ex + Jy[a] * ey;
                for (int b = 0; b < 8; ++b) JtJ[a][b] += Jx[a] * Jx[b] + Jy[a] * Jy[b];
            }
        }
        bool improved = false;
        for (int tries = 0; tries < 8 && !improved; ++tries) {
            double A[8][9];
            for (int a = 0; a < 8; ++a) {
                for (int b = 0; b < 8; ++b) A[a][b] = JtJ[a][b] + (a == b ? lambda * JtJ[a][a] : 0);
                A[a][8] = -Jte[a];
            }
            bool ok = true;
            for (int c = 0; c < 8 && ok; ++c) {
                int piv = c;
                for (int r = c + 1; r < 8; ++r)
                    if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
                if (fabs(A[piv][c]) < 1e-300) { ok = false; break; }
                if (piv != c) for (int j = 0; j < 9; ++j) std::swap(A[c][j], A[piv][j]);
                for (int r = c + 1; r < 8; ++r) {
                    const double f = A[r][c] / A[c][c];
                    for (int j = c; j < 9; ++j) A[r][j] -= f * A[c][j];
                }
            }
            double dh[8];
            if (ok)
                for (int r = 7; r >= 0; --r) {
                    double v = A[r][8];
                    for (int j = r + 1; j < 8; ++j) v -= A[r][j] * dh[j];
                    dh[r] = v / A[r][r];
                }
            if (ok) {
                double Hn[9];
                for (int a = 0; a < 8; ++a) Hn[a] = H[a] + dh[a];
                Hn[8] = 1;
                const double cn = reproj_cost(s, d, Hn);
                if (cn < cost) {
                    for (int a = 0; a < 9; ++a) H[a] = Hn[a];
                    improved = cost - cn > 1e-14 * cost;
                    cost = cn;
                    lambda = std::max(lambda * 0.1, 1e-12);
                    if (!improved) return;
                    break;
                }
            }
            lambda *= 10;
        }
        if (!improved) return;
    }
}

}  // namespace

// d_src [srows][scols][cn] u8 -> d_dst [drows][dcols][cn] u8 on `st`.  d_tab: device scratch of resize_tab_bytes(drows, dcols) bytes.
size_t resize_tab_bytes(int drows, int dcols) { return (size_t)(dcols + drows) * (sizeof(int) + 2 * sizeof(short)); }

int launch_resize_linear_u8(const uint8_t* d_src, int srows, int scols, int cn, uint8_t* d_dst, int drows, int dcols, void* d_tab, cudaStream_t st) {
    // coefficient tables exactly as OpenCV builds them (resize.cpp, INTER_LINEAR branch): scale = 1 / ((double)dst / src),
    // f = (float)((d + 0.5) * scale - 0.5), s = floor(f), f -= s; horizontally s < 0 -> (0, f = 0), s >= width-1 -> (width-1, f = 0);
    // vertically the row indices are clipped instead; coefficients saturate_cast<short>(c * 2048)
    std::vector<int> xofs(dcols), yofs(drows);
    std::vector<short> ia(2 * dcols), ib(2 * drows);
    const double scale_x = 1. / ((double)dcols / scols), scale_y = 1. / ((double)drows / srows);
    for (int dx = 0; dx < dcols; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)floorf(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= scols - 1) { fx = 0; sx = scols - 1; }
        xofs[dx] = sx;
        ia[2 * dx] = sat_short((1.f - fx) * 2048);
        ia[2 * dx + 1] = sat_short(fx * 2048);
    }
    for (int dy = 0; dy < drows; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        const int sy = (int)floorf(fy);
        fy -= sy;
        yofs[dy] = sy;
        ib[2 * dy] = sat_short((1.f - fy) * 2048);
        ib[2 * dy + 1] = sat_short(fy * 2048);
    }
    char* t = static_cast<char*>(d_tab);
    int* d_xofs = reinterpret_cast<int*>(t);
    int* d_yofs = d_xofs + dcols;
    short* d_ia = reinterpret_cast<short*>(d_yofs + drows);
    short* d_ib = d_ia + 2 * dcols;
    // the staging vectors die with this call: plain (synchronous w.r.t. the host) copies
    cudaMemcpyAsync(d_xofs, xofs.data(), dcols * sizeof(int), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_yofs, yofs.data(), drows * sizeof(int), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_ia, ia.data(), 2 * dcols * sizeof(short), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_ib, ib.data(), 2 * drows * sizeof(short), cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);
    resize_linear_u8_kernel<<<dim3((dcols + 255) / 256, drows), 256, 0, st>>>(d_src, srows, scols, cn, d_dst, drows, dcols, d_xofs, d_ia, d_yofs, d_ib);
    return 1;
}

// RANSAC homography over n device-resident correspondences.  d_work: homography_work_bytes(n, n_hyp) bytes of device scratch.
// Host outputs: H9 (row-major, h33 = 1), mask[n] (consensus set of the best hypothesis), *n_inliers.  Returns kernels launched, or -1
// when no model was found (fewer than 4 points, all samples degenerate).
size_t homography_work_bytes(int n, int n_hyp) { return (size_t)n_hyp * (9 * sizeof(float) + sizeof(int)) + (size_t)n; }

int run_homography_ransac(const float2* d_src, const float2* d_dst, const float2* h_src, const float2* h_dst, int n, float thresh, int n_hyp, uint32_t seed,
                          void* d_work, double* H9, uint8_t* mask, int* n_inliers, cudaStream_t st) {
    *n_inliers = 0;
    if (n < 4) return -1;
    float* d_models = static_cast<float*>(d_work);
    int* d_counts = reinterpret_cast<int*>(d_models + (size_t)n_hyp * 9);
    uint8_t* d_mask = reinterpret_cast<uint8_t*>(d_counts + n_hyp);
    const float thresh2 = thresh * thresh;
    homography_hypotheses_kernel<<<(n_hyp + 3) / 4, 128, 0, st>>>(d_src, d_dst, n, n_hyp, thresh2, seed, d_models, d_counts);
    std::vector<int> counts(n_hyp);
    cudaMemcpyAsync(counts.data(), d_counts, n_hyp * sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    int best = -1;
    for (int k = 0; k < n_hyp; ++k)
        if (counts[k] >= 4 && (best < 0 || counts[k] > counts[best])) best = k;
    if (best < 0) return -1;
    homography_mask_kernel<<<(n + 127) / 128, 128, 0, st>>>(d_src, d_dst, n, d_models + (size_t)best * 9, thresh2, d_mask);
    cudaMemcpyAsync(mask, d_mask, n, cudaMemcpyDeviceToHost, st);
    float Hf[9];
    cudaMemcpyAsync(Hf, d_models + (size_t)best * 9, sizeof(Hf), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    std::vector<float2> si, di;
    for (int i = 0; i < n; ++i)
        if (mask[i]) { si.push_back(h_src[i]); di.push_back(h_dst[i]); }
    *n_inliers = (int)si.size();
    for (int i = 0; i < 9; ++i) H9[i] = Hf[i];
    if (si.size() > 4) {  // OpenCV: least-squares refit on the consensus set, then 10 LM iterations
        double H[9];
        if (dlt_refit(si, di, H)) {
            lm_refine(si, di, H, 10);
            for (int i = 0; i < 9; ++i) H9[i] = H[i];
        }
    }
    return 2;
}

}  // namespace siftb200
