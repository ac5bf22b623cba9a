// describe.cu -- 4x4x8 SIFT descriptor with the reference's RootSIFT-of-quantised tail.
//
// Replaces calDescriptor / unpackOctave / calcSIFTDescriptor (reference src/sift.cpp:579-753).
// Compiled with --fmad=false so the sample arithmetic (rotation, bins, trilinear weights) rounds like the
// CPU expression order.
//
// The reference scatters every window sample into a (4+2)x(4+2)x(8+2) histogram.  A scatter into one shared histogram
// needs shared-memory float atomics, which on sm_100a are CAS loops (ATOMS.CAST.SPIN) that serialise badly because
// neighbouring samples hit the same bins.  This kernel is atomics-free and deterministic, one CTA of five warps per
// keypoint, one pass ("slab warps"):
//   - the window is cut into the five unit slabs of the row coordinate, r0 <= rbin < r0+1 for r0 = -1..3; warp s owns slab
//     r0 = s-1, so every sample is evaluated exactly once, floor(rbin) is a warp constant (no floor, no per-sample cell-row
//     logic) and the two cell-rows a slab votes into (r0, r0+1) are static: the border warps (r0 = -1, 3) skip the half of the
//     votes that falls outside the 4x4 grid with a warp-uniform branch instead of sending them to trash bins;
//   - a group of GL lanes (GL = 4..32, chosen per keypoint from the mean run length of a slab row) walks a window row's j-interval
//     of the slab (two slab inequalities rounded outwards by 1e-3 px; the reference's own float test decides membership), one
//     sample per lane per step, the gradient-map load of the step three ahead already in flight;
//   - the Gaussian weight is separable in window coordinates (rotation preserves i^2+j^2): exp(-(i^2+j^2)/(8 hw^2)) =
//     wrow[i] * wcol[j], two small per-keypoint tables instead of an expf per sample (|relative difference| to the reference's
//     exp of the rotated, rounded coordinates ~1e-7);
//   - the trilinear votes (reference operation order, :656-672) go to THREAD-PRIVATE histograms [2 cell-rows][6 cells][9 bins] in
//     shared memory (layout [bin][thread]: conflict-free plain read-modify-write, one address register + immediates); the two
//     extra cells (c0 = -1 and c0+1 = 4) absorb the out-of-grid column votes without any select;
//   - tail: column sums over the 32 private copies of each warp (rotated, conflict-free), slab pairs added per cell-row,
//     circular bin fold, then L2 -> clamp 0.2 -> x512 -> uchar (round half even) -> L1 -> sqrt with block reductions.
// Only the inner 4x4 cells are kept by the reference (:676-684), so the border cells are never formed.
#include <type_traits>

#include "sift_internal.cuh"

namespace siftb200 {
namespace {

constexpr int DW = 4, DB = 8;  // SIFT_DESCR_WIDTH, SIFT_DESCR_HIST_BINS (src/sift.cpp:12,15)
constexpr int NSLAB = DW + 1;  // floor(rbin) in -1..3
constexpr int DT = NSLAB * 32; // threads per CTA: one warp per slab
constexpr int PC = DW + 2;     // private cells per cell-row: c0+1 in 0..5 (cells -1 and 4 are scratch)
constexpr int PRIV_BINS = 2 * PC * (DB + 1);   // private histogram of one thread: [2 cell-rows][6 cells][9 bins]
constexpr int PRIV_FLOATS = PRIV_BINS * DT;
constexpr int NB = 84;                         // window rows per band (tables of one band live in shared memory); radius <= 40 => one band
constexpr int WCOL = 96;                       // columns per block (column-weight table); radius <= 40 => one block
constexpr int NSUM = NSLAB * 2 * DW * (DB + 1);  // column sums of the tail: [slab][cell-row][4 cells][9 bins]
constexpr int CTAS_PER_SM = 3;
// shared memory: private histograms | per-(slab,row) j-intervals (int2) | per-row {i*sin, i*cos, wrow, -} | wcol | reduction scratch.
// The tail's column sums alias the interval table.
constexpr int TAB_BYTES = NSLAB * NB * 8, ROW_BYTES = NB * 16;
constexpr int DESC_SMEM_BYTES = PRIV_FLOATS * 4 + TAB_BYTES + ROW_BYTES + WCOL * 4 + 32;
static_assert(NSUM * 4 <= TAB_BYTES, "tail sums alias the interval table");

__device__ __forceinline__ int cv_round(float v) { return __float2int_rn(v); }
__device__ __forceinline__ int cv_floor(float v) { return __float2int_rd(v); }

// sum over the CTA (5 warps); every thread gets the result.  `red` = 8 floats of shared scratch.
__device__ __forceinline__ float block_sum(float v, float* red, int tid) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    return ((red[0] + red[1]) + (red[2] + red[3])) + red[4];
}

// window radius of calcSIFTDescriptor (:587-590)
__device__ __forceinline__ int descr_radius(float scl, int rows, int cols) {
    const float hist_width = 3.f * scl;
    const int radius = cv_round(hist_width * 1.4142135623730951f * (DW + 1) * 0.5f);
    const int diag = (int)sqrt(((double)cols) * cols + ((double)rows) * rows);
    return min(radius, diag);
}

// j-interval of {lo_v <= j*k + off <= hi_v} widened by a margin that covers the rounding of this very computation, intersected into
// [lo, hi]; false when the row misses the slab.  `flat`: k is so small that j*k moves by < 0.05 over the window: the row is inside or
// outside as a whole (tested with that slack).  The exact per-sample test decides membership; this only has to be a superset.
__device__ __forceinline__ bool slab(float k, float inv_k, float margin, bool flat, float off, float lo_v, float hi_v, float& lo, float& hi) {
    if (!flat) {
        const float u0 = (lo_v - off) * inv_k, u1 = (hi_v - off) * inv_k;
        lo = fmaxf(lo, fminf(u0, u1) - margin);
        hi = fminf(hi, fmaxf(u0, u1) + margin);
        return true;
    }
    return off > lo_v - 0.06f && off < hi_v + 0.06f;
}

// calcSIFTDescriptor, src/sift.cpp:579-722, for one keypoint by one CTA.  dst: 128 floats in global memory.
// mo: the level's gradient map {Mag, Ori} (detect.cu gradient_kernel) -- the values the reference computes per sample (:623-633).
__device__ void calc_descriptor(const float2* __restrict__ mo, int rows, int cols, int pitch, float ptx, float pty, float ori, float scl, float gl_div,
                                float* __restrict__ smem, float* __restrict__ dst) {
    const int tid = threadIdx.x, lane = tid & 31, s = tid >> 5;  // s: slab, floor(rbin) = s - 1
    const int px = cv_round(ptx), py = cv_round(pty);
    const float cos_u = cosf(ori * (float)(3.1415926535897932384626433832795 / 180));
    const float sin_u = sinf(ori * (float)(3.1415926535897932384626433832795 / 180));
    const float bins_per_rad = DB / 360.f;
    const float hist_width = 3.f * scl;
    const int radius = descr_radius(scl, rows, cols);
    const float cos_t = cos_u / hist_width;
    const float sin_t = sin_u / hist_width;
    // exponent of the separable weight: (c_rot^2 + r_rot^2) * (-1/8) = (i^2 + j^2) * (cos_t^2 + sin_t^2) * (-1/8)
    const float es = (float)(-0.125 * ((double)cos_t * cos_t + (double)sin_t * sin_t));
    float* s_priv = smem;                                              // [PRIV_BINS][DT]
    int2* s_tab = reinterpret_cast<int2*>(smem + PRIV_FLOATS);         // [NSLAB][NB] {jlo, jhi}
    float* s_sum = reinterpret_cast<float*>(s_tab);                    // tail only
    float4* s_row = reinterpret_cast<float4*>(s_tab + NSLAB * NB);     // [NB] {i*sin_t, i*cos_t, wrow, -}
    float* s_wcol = reinterpret_cast<float*>(s_row + NB);              // [WCOL] wcol[j - cb0]
    float* s_red = s_wcol + WCOL;                                      // 8 floats
    const int jmin = max(-radius, 1 - px), jmax = min(radius, cols - 2 - px);   // 0 < c < cols-1  (:621)
    const int imin = max(-radius, 1 - py), imax = min(radius, rows - 2 - py);   // 0 < r < rows-1

    for (int k = tid * 4; k < PRIV_FLOATS; k += DT * 4) *reinterpret_cast<float4*>(s_priv + k) = make_float4(0.f, 0.f, 0.f, 0.f);
    // slab geometry (per keypoint): interval margins, flat-direction flags, lanes per group
    const float wspan = (float)(radius > 1 ? radius : 1);
    const bool flat_s = fabsf(sin_t) * wspan < 0.05f, flat_c = fabsf(cos_t) * wspan < 0.05f;
    const float inv_s = flat_s ? 0.f : 1.f / sin_t, inv_c = flat_c ? 0.f : 1.f / cos_t;
    const float mar_s = 1e-3f + 4e-6f * fabsf(inv_s), mar_c = 1e-3f + 4e-6f * fabsf(inv_c);
    // mean run length of a slab row: (slab area 5 hw^2) / (rows it crosses, hw (|cos| + 5 |sin|)); lanes per group = largest power of
    // two <= run / gl_div, in [4, 32] (longer groups: fewer 128-byte lines per gather; shorter: fewer idle lanes at row ends)
    const float run = 5.f * hist_width / (fabsf(cos_u) + 5.f * fabsf(sin_u));
    int glsh = 2;
    while (glsh < 5 && (float)(2 << glsh) * gl_div <= run) ++glsh;
    const int GL = 1 << glsh, gl = lane & (GL - 1), slot = lane >> glsh, ng = 32 >> glsh;
    const float r0f = (float)(s - 1), r1f = (float)s;
    float* priv = s_priv + tid;
    // The tables are addressed relative to `priv` (an address the walk keeps in a register anyway); the opaque copy of tid keeps the
    // compiler from folding this back to the shared-window base, which it would re-materialise (S2R + 2 ops) at every use.
    int tid_o = tid;
    asm volatile("" : "+r"(tid_o));
    const float* wcol_p = priv + (PRIV_FLOATS + (TAB_BYTES + ROW_BYTES) / 4 - tid_o);       // == s_wcol
    const int2* tab_p = reinterpret_cast<const int2*>(priv + (PRIV_FLOATS - tid_o)) + s * NB;  // == s_tab + s * NB
    const float4* row_p = reinterpret_cast<const float4*>(priv + (PRIV_FLOATS + TAB_BYTES / 4 - tid_o));  // == s_row
    // The window is processed in blocks of NB rows x WCOL columns so that the interval and weight tables fit shared memory whatever the
    // keypoint size; a pipeline keypoint (radius <= 40) is one block.
    for (int band0 = imin; band0 <= imax; band0 += NB)
    for (int cb0 = jmin; cb0 <= jmax; cb0 += WCOL) {
        const int nrows = min(NB, imax - band0 + 1);
        const int cb1 = min(jmax, cb0 + WCOL - 1);
        __syncthreads();  // previous block's walk (and the zeroing above) done before the tables are rewritten
        for (int k = tid; k <= cb1 - cb0; k += DT) s_wcol[k] = expf((float)((cb0 + k) * (cb0 + k)) * es);
        // per row: the column-slab interval -1 < cbin < 4 (shared by the five row slabs), then the five row-slab intervals
        for (int r = tid; r < nrows; r += DT) {
            const int i = band0 + r;
            const float isin = i * sin_t, icos = i * cos_t;
            s_row[r] = make_float4(isin, icos, expf((float)(i * i) * es), 0.f);
            float lo0 = (float)cb0, hi0 = (float)cb1;
            const bool ok0 = slab(cos_t, inv_c, mar_c, flat_c, -isin + 1.5f, -1.f, 4.f, lo0, hi0);
#pragma unroll
            for (int q = 0; q < NSLAB; ++q) {
                float lo = lo0, hi = hi0;
                const bool ok = ok0 && slab(sin_t, inv_s, mar_s, flat_s, icos + 1.5f, q - 1.f, (float)q, lo, hi);
                s_tab[q * NB + r] = ok ? make_int2(max(cb0, (int)ceilf(lo)), min(cb1, (int)floorf(hi))) : make_int2(1, 0);
            }
        }
        __syncthreads();
        // first / last non-empty row of this warp's slab
        int rlo = nrows, rhi = -1;
        for (int rb = 0; rb < nrows; rb += 32) {
            const int r = rb + lane;
            bool ne = false;
            if (r < nrows) { const int2 t = s_tab[s * NB + r]; ne = t.x <= t.y; }
            const unsigned m = __ballot_sync(0xffffffffu, ne);
            if (m) { rlo = min(rlo, rb + __ffs(m) - 1); rhi = rb + 31 - __clz(m); }
        }
        // has_lo / has_hi: cell-rows r0 and r0+1 lie inside the 4x4 grid (compile-time per warp role: border warps carry half the votes)
        auto walk = [&](auto LO, auto HI) {
            constexpr bool has_lo = decltype(LO)::value, has_hi = decltype(HI)::value;
            // flattened walk: a group advances through its rows (rlo+slot, +ng, ...) one GL-sample step per iteration (one sample per
            // lane), so the groups of a warp never wait for each other at row boundaries.
            int r = rlo + slot - ng, jb = 1, jhi = 0;
            const float2* rowp = mo;
            float isin = 0.f, icos = 0.f, wrow = 0.f;
            auto advance = [&]() -> bool {
                jb += GL;
                if (jb > jhi) {
                    do {
                        r += ng;
                        if (r > rhi) return false;
                        const int2 t = tab_p[r];
                        jb = t.x; jhi = t.y;
                    } while (jb > jhi);
                    const float4 rv = row_p[r];
                    isin = rv.x; icos = rv.y; wrow = rv.z;
                    rowp = mo + (size_t)(py + band0 + r) * pitch + px;
                }
                return true;
            };
            // RING steps in flight: the gradient-map load of step k+RING-1 is issued before the votes of step k.  A stage carries the
            // sample's bin coordinates, weight and gradient; a lane past the end of its row gets cbin = 1e9 (rejected by the reference's
            // own range test), w < 0 marks a group that has run out of rows.
            struct Step { float rbin, cbin, w; float2 mo; };
            auto issue = [&](Step& st) {
                st.mo = make_float2(0.f, 0.f);
                st.w = -1.f;
                if (advance()) {
                    const int j = jb + gl;
                    const int jc = min(j, jhi);  // clamped: always a pixel of the interval / a table entry
                    st.mo = __ldg(rowp + jc);
                    st.w = wrow * wcol_p[jc - cb0];
                    const float jf = (float)j;
                    const float c_rot = jf * cos_t - isin;
                    const float r_rot = jf * sin_t + icos;
                    st.rbin = r_rot + DW / 2 - 0.5f;
                    const float cbin = c_rot + DW / 2 - 0.5f;
                    st.cbin = j <= jhi ? cbin : 1e9f;
                }
            };
            // one step: the eight trilinear votes of the sample of this lane, branch-free: a rejected sample votes zeros into cell 0
            auto vote = [&](const Step& st) {
                const float rbin = st.rbin;
                float cbin = st.cbin;
                // floor(rbin) == r0 (this warp's slab) and -1 < cbin < 4 (:620); rbin == -1 exactly (rejected by the reference) votes 0
                const bool acc = rbin >= r0f && rbin < r1f && cbin > -1 && cbin < DW;
                const float mag = acc ? st.mo.x * st.w : 0.f;
                cbin = acc ? cbin : 0.f;
                const float rf = rbin - r0f;
                float obin = (st.mo.y - ori) * bins_per_rad;
                const int c0 = cv_floor(cbin);
                const int o0 = cv_floor(obin);
                const float cf = cbin - c0;
                obin -= o0;
                // trilinear split in the reference's operation order (:656-662)
                const float v_r1 = mag * rf, v_r0 = mag - v_r1;
                float* b = priv + ((c0 + 1) * (DB + 1) + (o0 & (DB - 1))) * DT;  // o0 in [-8, 7]: the reference's two wrap tests == & 7
                if (has_lo) {
                    const float v_rc01 = v_r0 * cf, v_rc00 = v_r0 - v_rc01;
                    float v1;
                    v1 = v_rc00 * obin; b[0] += v_rc00 - v1; b[DT] += v1;
                    v1 = v_rc01 * obin; b[(DB + 1) * DT] += v_rc01 - v1; b[(DB + 2) * DT] += v1;
                }
                if (has_hi) {
                    const float v_rc11 = v_r1 * cf, v_rc10 = v_r1 - v_rc11;
                    float v1;
                    v1 = v_rc10 * obin; b[PC * (DB + 1) * DT] += v_rc10 - v1; b[(PC * (DB + 1) + 1) * DT] += v1;
                    v1 = v_rc11 * obin; b[(PC + 1) * (DB + 1) * DT] += v_rc11 - v1; b[((PC + 1) * (DB + 1) + 1) * DT] += v1;
                }
            };
            // ring of four stages, unrolled by four so that a stage is refilled in place (no register shuffling)
            Step s0, s1, s2, s3;
            issue(s0); issue(s1); issue(s2); issue(s3);
            for (;;) {
                if (s0.w < 0.f) break;
                vote(s0); issue(s0);
                if (s1.w < 0.f) break;
                vote(s1); issue(s1);
                if (s2.w < 0.f) break;
                vote(s2); issue(s2);
                if (s3.w < 0.f) break;
                vote(s3); issue(s3);
            }
        };
        if (s == 0) walk(std::false_type{}, std::true_type{});
        else if (s == DW) walk(std::true_type{}, std::false_type{});
        else walk(std::true_type{}, std::true_type{});
    }
    __syncthreads();

    // ---- tail, stage A: column sums over the 32 private copies of each (slab, cell-row, cell 0..3, bin) (rotated read: conflict-free) ----
    for (int sidx = tid; sidx < NSUM; sidx += DT) {
        const int q = sidx / (2 * DW * (DB + 1)), rem = sidx - q * (2 * DW * (DB + 1));
        const int lr = rem / (DW * (DB + 1)), cb = rem - lr * (DW * (DB + 1));  // cb = cell * 9 + bin
        const float* col = s_priv + ((lr * PC + 1) * (DB + 1) + cb) * DT + q * 32;  // private cell index = cell + 1
        float acc = 0.f;
#pragma unroll 16
        for (int g = 0; g < 32; ++g) acc += col[(g + tid) & 31];
        s_sum[sidx] = acc;
    }
    __syncthreads();
    // ---- stage B: output element e = (a*4 + b)*8 + k, one per thread; cell-row a = slab a+1's row r0 plus slab a's row r0+1 ----
    float v = 0.f;
    if (tid < 128) {
        const int e_cell = tid >> 3, e_k = tid & 7;
        const int e_a = e_cell >> 2, e_b = e_cell & 3;
        const float* lo = s_sum + ((e_a + 1) * 2 + 0) * (DW * (DB + 1)) + e_b * (DB + 1);
        const float* hi = s_sum + (e_a * 2 + 1) * (DW * (DB + 1)) + e_b * (DB + 1);
        v = lo[e_k] + hi[e_k];
        if (e_k == 0) v += lo[DB] + hi[DB];  // hist[idx] += hist[idx+n] (:680); hist[idx+n+1] is never written since o0 <= n-1
    }
    float nrm2 = block_sum(v * v, s_red, tid);
    const float thr = sqrtf(nrm2) * 0.2f;
    v = fminf(v, thr);
    nrm2 = block_sum(v * v, s_red, tid);
    nrm2 = 512.f / fmaxf(sqrtf(nrm2), 1.1920928955078125e-7f);
    int u = __float2int_rn(v * nrm2);  // saturate_cast<uchar>: round half to even, clamp to 0..255
    u = min(max(u, 0), 255);
    v = (float)u * nrm2;
    float nrm1 = block_sum(v, s_red, tid);
    nrm1 = 1.f / fmaxf(nrm1, 1.1920928955078125e-7f);
    if (tid < 128) dst[tid] = sqrtf(v * nrm1);
    __syncthreads();
}

__global__ void __launch_bounds__(DT, CTAS_PER_SM)
    describe_kernel(const __grid_constant__ PyrView pv, const DetectBuf db, SiftKeypoint* __restrict__ kp_out, float* __restrict__ desc_out, int cap,
                    float gl_div) {
    extern __shared__ __align__(16) float smem[];
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
            calc_descriptor(img, ov.rows, ov.cols, ov.pitch, rec.x * scale, rec.y * scale, angle, size * 0.5f, gl_div, smem,
                            desc_out + ((size_t)f * cap + slot) * 128);
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
__global__ void __launch_bounds__(DT, CTAS_PER_SM)
    describe_given_kernel(const __grid_constant__ PyrView pv, const SiftKeypoint* __restrict__ kps, int n, float* __restrict__ desc_out, int first_octave,
                          int* __restrict__ err, float gl_div) {
    extern __shared__ __align__(16) float smem[];
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
        calc_descriptor(ov.MO[layer], ov.rows, ov.cols, ov.pitch, kp.x * scale, kp.y * scale, angle, size * 0.5f, gl_div, smem, desc_out + (size_t)p * 128);
    }
}

float g_gl_div = 1.5f;

}  // namespace

void init_describe_kernels() {
    cudaFuncSetAttribute(describe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DESC_SMEM_BYTES);
    cudaFuncSetAttribute(describe_given_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DESC_SMEM_BYTES);
    if (const char* e = getenv("SIFT_B200_DESC_GLDIV")) g_gl_div = (float)atof(e);  // tuning probe: run length per group lane count
}

int launch_describe(const PyrView& pv, const DetectBuf& db, int n_frames, SiftKeypoint* d_kp, float* d_desc, int cap, cudaStream_t st) {
    dim3 grid(num_sms() * CTAS_PER_SM, n_frames);
    describe_kernel<<<grid, DT, DESC_SMEM_BYTES, st>>>(pv, db, d_kp, d_desc, cap, g_gl_div);
    return 1;
}

int launch_describe_given(const PyrView& pv, const SiftKeypoint* d_kps, int n, float* d_desc, int first_octave, int* d_err, cudaStream_t st) {
    if (n <= 0) return 0;
    const int blocks = n < num_sms() * CTAS_PER_SM ? n : num_sms() * CTAS_PER_SM;
    describe_given_kernel<<<blocks, DT, DESC_SMEM_BYTES, st>>>(pv, d_kps, n, d_desc, first_octave, d_err, g_gl_div);
    return 1;
}

}  // namespace siftb200
