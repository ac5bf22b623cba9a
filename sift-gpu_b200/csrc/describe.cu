// describe.cu -- 4x4x8 SIFT descriptor with the reference's RootSIFT-of-quantised tail.
//
// Replaces calDescriptor / unpackOctave / calcSIFTDescriptor (reference src/sift.cpp:579-753).
// Compiled with --fmad=false so the sample arithmetic (rotation, bins, trilinear weights) rounds like the
// CPU expression order.
//
// The reference scatters every window sample into a (4+2)x(4+2)x(8+2) histogram.  A scatter into one shared histogram
// needs shared-memory float atomics, which on sm_100a are CAS loops (ATOMS.CAST.SPIN) that serialise badly because
// neighbouring samples hit the same bins.  This kernel is atomics-free and deterministic, one CTA (64 threads) per
// keypoint, one pass:
//   - the 4 cell-rows of the descriptor grid are split in two pairs p (a in {2p, 2p+1}); a 4-lane group owns (pair, window
//     row): its lanes walk the row's j-interval 2p-1 <= rbin < 2p+2, -1 < cbin < 4 (two slab inequalities rounded
//     outwards; the reference's exact test decides), one sample per lane per step with the next step's load already in
//     flight; lanes of a group read consecutive pixels (coalesced 32-byte segments);
//   - each sample reads {Mag, Ori} of its pixel from the level's gradient map (detect.cu: gradient_kernel, the reference's
//     own per-sample arithmetic done once per pixel), applies the Gaussian weight, and its trilinear votes that fall into the pair's cells go straight into THREAD-PRIVATE histograms
//     [2 cell-rows][4 cells][9 bins] in shared memory (layout [bin][thread]: conflict-free plain read-modify-write).
//     A sample is evaluated 1.2 times on average (twice only when its two cell-rows straddle the pairs);
//   - tail: sum the 32 private copies of each pair (rotated, conflict-free); a thread owns 2 of the 128 output elements;
//     fold the circular bin, then L2 -> clamp 0.2 -> x512 -> uchar (round half even) -> L1 -> sqrt with block reductions.
// Only the inner 4x4 cells are kept by the reference (:676-684), so the border cells are never formed.
#include "sift_internal.cuh"

namespace siftb200 {
namespace {

constexpr int DW = 4, DB = 8;  // SIFT_DESCR_WIDTH, SIFT_DESCR_HIST_BINS (src/sift.cpp:12,15)
#ifndef DESC_DT
#define DESC_DT 64
#endif
constexpr int DT = DESC_DT;    // threads per CTA.  Zeroing, interval tables, column sums and norms are per-warp costs paid once per keypoint:
                               // measured 128 / 64 / 32 threads: 66.5 / 60.3 / 61.4 us per frame (32: too few row groups per pair, lower occupancy)
constexpr int HALF = DT / 2;   // threads (= private histogram copies) per cell-row pair
constexpr int NEL = 128 / DT;  // output elements per thread
constexpr int CTAS_PER_SM = DT == 128 ? 6 : DT == 64 ? 10 : 18;  // what 227 KB of shared memory admits (38 KB / 21.6 KB per CTA)
#ifndef DESC_GL
#define DESC_GL 4
#endif
constexpr int GL = DESC_GL;    // lanes per group (a group walks one window row of one cell-row pair); measured 2/4/8: 69.7/66.3/69.1 us
constexpr int PRIV_BINS = 2 * DW * (DB + 1);      // private histogram of one thread: [2 cell-rows][4 cells][9 bins]
constexpr int TRASH = PRIV_BINS;                   // private trash bins that swallow votes for cells outside the pair / the 4x4 grid
constexpr int PRIV_FLOATS = (PRIV_BINS + 2) * DT;  // + two trash rows (a vote updates bins i and i+1)
constexpr int NB = 128;                            // window rows per band (intervals of one band live in shared memory)
// Interval tables are 32-bit on purpose: the same kernel with 16-bit tables measured 141 us instead of 58 us per frame (twice, in two
// different versions of this kernel; cause not found), and aliasing them with the column sums to fit an 11th CTA per SM gained nothing.
typedef int tab_t;
constexpr int TAB_BYTES = 2 * 2 * NB * (int)sizeof(tab_t), SUM_BYTES = (2 * PRIV_BINS + 8) * 4;
constexpr int DESC_SMEM_BYTES = PRIV_FLOATS * 4 + SUM_BYTES + TAB_BYTES;

__device__ __forceinline__ int cv_round(float v) { return __float2int_rn(v); }
__device__ __forceinline__ int cv_floor(float v) { return __float2int_rd(v); }

// sum over the CTA (DT/32 warps); every thread gets the result.  `red` = 4 floats of shared scratch.
__device__ __forceinline__ float block_sum(float v, float* red, int tid) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if (DT == 32) return v;
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    if (DT == 128) return (red[0] + red[1]) + (red[2] + red[3]);
    return red[0] + red[1];
}

// window radius of calcSIFTDescriptor (:587-590)
__device__ __forceinline__ int descr_radius(float scl, int rows, int cols) {
    const float hist_width = 3.f * scl;
    const int radius = cv_round(hist_width * 1.4142135623730951f * (DW + 1) * 0.5f);
    const int diag = (int)sqrt(((double)cols) * cols + ((double)rows) * rows);
    return min(radius, diag);
}

// conservative j-interval of {lo_v < j*k + off < hi_v}, intersected into [lo, hi]; false when the row misses the slab
__device__ __forceinline__ bool slab(float k, float inv_k, float off, float lo_v, float hi_v, float& lo, float& hi) {
    if (inv_k != 0.f) {
        const float u0 = (lo_v - off) * inv_k, u1 = (hi_v - off) * inv_k;
        lo = fmaxf(lo, fminf(u0, u1));
        hi = fminf(hi, fmaxf(u0, u1));
        return true;
    }
    return off > lo_v - 0.01f && off < hi_v + 0.01f;  // k ~ 0: the row is inside or outside as a whole
}

// calcSIFTDescriptor, src/sift.cpp:579-722, for one keypoint by one CTA.  dst: 128 floats in global memory.
// mo: the level's gradient map {Mag, Ori} (detect.cu gradient_kernel) -- the values the reference computes per sample (:623-633).
__device__ void calc_descriptor(const float2* __restrict__ mo, int rows, int cols, int pitch, float ptx, float pty, float ori, float scl,
                                float* __restrict__ smem, float* __restrict__ dst) {
    const int tid = threadIdx.x;
    const int px = cv_round(ptx), py = cv_round(pty);
    float cos_t = cosf(ori * (float)(3.1415926535897932384626433832795 / 180));
    float sin_t = sinf(ori * (float)(3.1415926535897932384626433832795 / 180));
    const float bins_per_rad = DB / 360.f;
    const float exp_scale = -1.f / (DW * DW * 0.5f);
    const float hist_width = 3.f * scl;
    const int radius = descr_radius(scl, rows, cols);
    cos_t /= hist_width;
    sin_t /= hist_width;
    float* s_priv = smem;                      // [PRIV_BINS + trash][DT]
    float* s_sum = smem + PRIV_FLOATS;         // [2 pairs][PRIV_BINS] column sums (tail only)
    float* s_red = s_sum + 2 * PRIV_BINS;      // 8 floats
    tab_t* s_jlo = reinterpret_cast<tab_t*>(s_red + 8);  // [2 pairs][NB]
    tab_t* s_jhi = s_jlo + 2 * NB;
    const float inv_s = fabsf(sin_t) > 1e-6f ? 1.f / sin_t : 0.f;
    const float inv_c = fabsf(cos_t) > 1e-6f ? 1.f / cos_t : 0.f;
    const int jmin = max(-radius, 1 - px), jmax = min(radius, cols - 2 - px);   // 0 < c < cols-1  (:621)
    const int imin = max(-radius, 1 - py), imax = min(radius, rows - 2 - py);   // 0 < r < rows-1

    for (int k = tid * 4; k < PRIV_FLOATS; k += DT * 4) *reinterpret_cast<float4*>(s_priv + k) = make_float4(0.f, 0.f, 0.f, 0.f);
    const int gl = tid & (GL - 1);
    const int p = tid / HALF, slot = (tid % HALF) / GL;  // cell-row pair (its HALF threads are contiguous), row slot (HALF/GL slots per pair)
    float* priv = s_priv + tid;
    for (int band0 = imin; band0 <= imax; band0 += NB) {
        const int nrows = min(NB, imax - band0 + 1);
        // accepted j-interval of every (pair, row) of the band: 2p-1 <= rbin < 2p+2 and -1 < cbin < 4, widened by a pixel
        for (int it = tid; it < 2 * nrows; it += DT) {
            const int pp = it >= nrows, i = band0 + it - pp * nrows;
            float lo = (float)jmin, hi = (float)jmax;
            bool ok = slab(sin_t, inv_s, i * cos_t + 1.5f, 2 * pp - 1.f, 2 * pp + 2.f, lo, hi);
            ok = ok && slab(cos_t, inv_c, -(i * sin_t) + 1.5f, -1.f, 4.f, lo, hi);
            // floor/ceil of the real-valued slab bounds already cover every sample the float test can accept (the bounds are
            // accurate to ~1e-5 px; a sample that close to the boundary carries a ~1e-6 share of its vote)
            s_jlo[pp * NB + it - pp * nrows] = (tab_t)(ok ? max(jmin, (int)floorf(lo)) : 1);
            s_jhi[pp * NB + it - pp * nrows] = (tab_t)(ok ? min(jmax, (int)ceilf(hi)) : 0);
        }
        __syncthreads();
        // flattened walk: a group advances through its rows (slot, slot+8, ...) one 8-sample step per iteration (one sample per
        // lane), so the four groups of a warp never wait for each other at row boundaries; the NEXT step's gradient-map load is
        // issued before the current step's arithmetic (software pipeline).
        int r = slot - HALF / GL, jb = 1, jhi = 0;
        const float2* rowp = mo;
        float isin = 0.f, icos = 0.f;
        // advance (r, jb, jhi, rowp, isin, icos) to the next step; false when the group has no more work in this band
        auto advance = [&]() -> bool {
            jb += GL;
            if (jb > jhi) {
                do {
                    r += HALF / GL;
                    if (r >= nrows) return false;
                    jb = s_jlo[p * NB + r];
                    jhi = s_jhi[p * NB + r];
                } while (jb > jhi);
                const int i = band0 + r;
                rowp = mo + (size_t)(py + i) * pitch + px;
                isin = i * sin_t;
                icos = i * cos_t;
            }
            return true;
        };
        // RING steps in flight: the gradient-map load of step k+RING-1 is issued before the arithmetic of step k.  A stage carries only
        // {j, i*sin, i*cos, Mag/Ori}: j = J_IDLE marks a lane past the end of its row (any real rotation sends it far outside the window,
        // so the reference's own range test rejects it), j = J_DONE a group that has run out of rows.
        struct Step { int j; float isin, icos; float2 mo; };
        constexpr int J_IDLE = 1 << 20, J_DONE = 0x7fffffff;
        auto issue = [&](Step& st) {
            st.mo = make_float2(0.f, 0.f);
            st.j = J_DONE;
            if (advance()) {
                const int j = jb + gl;
                st.isin = isin; st.icos = icos;
                st.mo = __ldg(rowp + min(j, jmax));  // clamped: always an interior pixel
                st.j = j <= jhi ? j : J_IDLE;
            }
        };
        // one step: the eight trilinear votes of the sample of this lane (position/metadata/gradient in st)
        auto vote = [&](const Step& st) {
            const int j = st.j;
            const float isin_c = st.isin, icos_c = st.icos;
            const float2 cur = st.mo;
            {
            const float c_rot = j * cos_t - isin_c;
            const float r_rot = j * sin_t + icos_c;
            float rbin = r_rot + DW / 2 - 0.5f;
            float cbin = c_rot + DW / 2 - 0.5f;
            const bool acc = rbin > -1 && rbin < DW && cbin > -1 && cbin < DW;  // (:620)
            const float w_ = expf((c_rot * c_rot + r_rot * r_rot) * exp_scale);
            float obin = (cur.y - ori) * bins_per_rad;
            const float mag = cur.x * w_;
            const int r0 = cv_floor(rbin), c0 = cv_floor(cbin);
            int o0 = cv_floor(obin);
            rbin -= r0; cbin -= c0; obin -= o0;
            if (o0 < 0) o0 += DB;
            if (o0 >= DB) o0 -= DB;
            const int la = r0 - 2 * p;  // local cell-row of the r0 vote; the r0+1 vote goes to la+1
            if (acc && mag != 0.f && la >= -1 && la <= 1) {
                // trilinear split in the reference's operation order (:656-662); votes for cells outside this pair or
                // outside the 4x4 grid land in the thread's trash bin, so the eight updates are branch-free
                const float v_r1 = mag * rbin, v_r0 = mag - v_r1;
                const float v_rc11 = v_r1 * cbin, v_rc10 = v_r1 - v_rc11;
                const float v_rc01 = v_r0 * cbin, v_rc00 = v_r0 - v_rc01;
                const bool r_lo = la >= 0, r_hi = la <= 0, c_lo = c0 >= 0, c_hi = c0 <= DW - 2;
                const int b00 = (la * DW + c0) * (DB + 1) + o0;  // bin of (r0, c0, o0)
                const int i00 = r_lo && c_lo ? b00 : TRASH, i01 = r_lo && c_hi ? b00 + (DB + 1) : TRASH;
                const int i10 = r_hi && c_lo ? b00 + DW * (DB + 1) : TRASH, i11 = r_hi && c_hi ? b00 + (DW + 1) * (DB + 1) : TRASH;
                float v1;
                v1 = v_rc00 * obin; priv[i00 * DT] += v_rc00 - v1; priv[(i00 + 1) * DT] += v1;
                v1 = v_rc01 * obin; priv[i01 * DT] += v_rc01 - v1; priv[(i01 + 1) * DT] += v1;
                v1 = v_rc10 * obin; priv[i10 * DT] += v_rc10 - v1; priv[(i10 + 1) * DT] += v1;
                v1 = v_rc11 * obin; priv[i11 * DT] += v_rc11 - v1; priv[(i11 + 1) * DT] += v1;
            }
        }
        };
        // ring of four stages, unrolled by four so that a stage is refilled in place (no register shuffling)
        Step s0, s1, s2, s3;
        issue(s0); issue(s1); issue(s2); issue(s3);
        for (;;) {
            if (s0.j == J_DONE) break;
            vote(s0); issue(s0);
            if (s1.j == J_DONE) break;
            vote(s1); issue(s1);
            if (s2.j == J_DONE) break;
            vote(s2); issue(s2);
            if (s3.j == J_DONE) break;
            vote(s3); issue(s3);
        }
        __syncthreads();
    }

    // ---- tail, stage A: the 2 x 72 column sums over the HALF private copies of each pair (rotated read: conflict-free) ----
    for (int sidx = tid; sidx < 2 * PRIV_BINS; sidx += DT) {
        const int sp = sidx >= PRIV_BINS, bin = sidx - sp * PRIV_BINS;
        const float* col = s_priv + bin * DT + sp * HALF;  // the HALF private copies of pair sp are contiguous
        float acc = 0.f;
#pragma unroll 16
        for (int g = 0; g < HALF; ++g) acc += col[(g + tid) & (HALF - 1)];
        s_sum[sidx] = acc;
    }
    __syncthreads();
    // ---- stage B: output element e = cell*8 + k; thread tid owns elements tid, tid + DT, ... ----
    float v[NEL];
    float part = 0.f;
#pragma unroll
    for (int q = 0; q < NEL; ++q) {
        const int e = tid + q * DT;
        const int e_cell = e >> 3, e_k = e & 7;
        const int e_a = e_cell >> 2, e_b = e_cell & 3;
        const float* ssum = s_sum + (e_a >> 1) * PRIV_BINS + ((e_a & 1) * DW + e_b) * (DB + 1);
        v[q] = ssum[e_k];
        if (e_k == 0) v[q] += ssum[DB];  // hist[idx] += hist[idx+n] (:680); hist[idx+n+1] is never written since o0 <= n-1
        part += v[q] * v[q];
    }
    float nrm2 = block_sum(part, s_red, tid);
    const float thr = sqrtf(nrm2) * 0.2f;
    part = 0.f;
#pragma unroll
    for (int q = 0; q < NEL; ++q) { v[q] = fminf(v[q], thr); part += v[q] * v[q]; }
    nrm2 = block_sum(part, s_red, tid);
    nrm2 = 512.f / fmaxf(sqrtf(nrm2), 1.1920928955078125e-7f);
    part = 0.f;
#pragma unroll
    for (int q = 0; q < NEL; ++q) {
        int u = __float2int_rn(v[q] * nrm2);  // saturate_cast<uchar>: round half to even, clamp to 0..255
        u = min(max(u, 0), 255);
        v[q] = (float)u * nrm2;
        part += v[q];
    }
    float nrm1 = block_sum(part, s_red, tid);
    nrm1 = 1.f / fmaxf(nrm1, 1.1920928955078125e-7f);
#pragma unroll
    for (int q = 0; q < NEL; ++q) dst[tid + q * DT] = sqrtf(v[q] * nrm1);
    __syncthreads();
}

__global__ void __maxnreg__(DT == 64 ? 88 : DT == 128 ? 80 : 112)
    describe_kernel(const __grid_constant__ PyrView pv, const DetectBuf db, SiftKeypoint* __restrict__ kp_out, float* __restrict__ desc_out, int cap) {
    extern __shared__ float smem[];
    const int f = blockIdx.y;
    int n = db.n_refined[f];
    if (n > db.cap_r) n = db.cap_r;
    for (int p = blockIdx.x; p < n; p += gridDim.x) {
        const int i = db.order[(size_t)f * db.cap_r + p];
        const Refined rec = db.refined[(size_t)f * db.cap_r + i];
        const int np = db.n_peaks[(size_t)f * db.cap_r + i];
        const int base = db.kp_offset[(size_t)f * db.cap_r + i];
        // unpackOctave (:724-731); firstOctave = 0 in SIFT_NCL (:86)
        const int octave = rec.octave & 255, layer = (rec.octave >> 8) & 255;
        const float scale = 1.f / (1 << octave);
        const OctaveView& ov = pv.oct[octave];
        const float2* img = ov.MO[layer] + (size_t)f * ov.frame_stride;
        const float size = rec.size * scale;
        for (int k = 0; k < np; ++k) {
            const int slot = base + k;
            if (slot >= cap) break;
            const float kp_angle = db.angles[((size_t)f * db.cap_r + i) * kMaxPeaks + k];
            float angle = 360.f - kp_angle;
            if (fabsf(angle - 360.f) < 1.1920928955078125e-7f) angle = 0.f;
            calc_descriptor(img, ov.rows, ov.cols, ov.pitch, rec.x * scale, rec.y * scale, angle, size * 0.5f, smem, desc_out + ((size_t)f * cap + slot) * 128);
            if (threadIdx.x == 0) {
                SiftKeypoint kp;
                kp.x = rec.x; kp.y = rec.y; kp.size = rec.size; kp.angle = kp_angle; kp.response = rec.response;
                kp.octave = rec.octave; kp.class_id = -1;
                kp_out[(size_t)f * cap + slot] = kp;
            }
        }
    }
}

// calDescriptor on caller-supplied keypoints (stage-level API): any octave/layer the reference's CV_Assert admits.
__global__ void __launch_bounds__(DT)
    describe_given_kernel(const __grid_constant__ PyrView pv, const SiftKeypoint* __restrict__ kps, int n, float* __restrict__ desc_out, int first_octave,
                          int* __restrict__ err) {
    extern __shared__ float smem[];
    for (int p = blockIdx.x; p < n; p += gridDim.x) {
        const SiftKeypoint kp = kps[p];
        int octave = kp.octave & 255;
        const int layer = (kp.octave >> 8) & 255;
        octave = octave < 128 ? octave : (-128 | octave);
        const float scale = octave >= 0 ? 1.f / (1 << octave) : (float)(1 << -octave);
        if (!(octave >= first_octave && layer <= kOctaveLayers + 2) || octave - first_octave >= pv.n_oct || layer >= kNumScales) {
            if (threadIdx.x == 0) atomicExch(err, 1);  // CV_Assert, src/sift.cpp:744
            continue;
        }
        const OctaveView& ov = pv.oct[octave - first_octave];
        float angle = 360.f - kp.angle;
        if (fabsf(angle - 360.f) < 1.1920928955078125e-7f) angle = 0.f;
        const float size = kp.size * scale;
        calc_descriptor(ov.MO[layer], ov.rows, ov.cols, ov.pitch, kp.x * scale, kp.y * scale, angle, size * 0.5f, smem, desc_out + (size_t)p * 128);
    }
}

}  // namespace

void init_describe_kernels() {
    cudaFuncSetAttribute(describe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DESC_SMEM_BYTES);
    cudaFuncSetAttribute(describe_given_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DESC_SMEM_BYTES);
}

int launch_describe(const PyrView& pv, const DetectBuf& db, int n_frames, SiftKeypoint* d_kp, float* d_desc, int cap, cudaStream_t st) {
    dim3 grid(148 * CTAS_PER_SM, n_frames);
    describe_kernel<<<grid, DT, DESC_SMEM_BYTES, st>>>(pv, db, d_kp, d_desc, cap);
    return 1;
}

int launch_describe_given(const PyrView& pv, const SiftKeypoint* d_kps, int n, float* d_desc, int first_octave, int* d_err, cudaStream_t st) {
    if (n <= 0) return 0;
    int blocks = n < 148 * 4 ? n : 148 * 4;
    describe_given_kernel<<<blocks, DT, DESC_SMEM_BYTES, st>>>(pv, d_kps, n, d_desc, first_octave, d_err);
    return 1;
}

}  // namespace siftb200
