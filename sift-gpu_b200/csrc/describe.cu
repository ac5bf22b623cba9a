// describe.cu -- 4x4x8 SIFT descriptor with the reference's RootSIFT-of-quantised tail.
//
// Replaces calDescriptor / unpackOctave / calcSIFTDescriptor (reference src/sift.cpp:579-753).
// Compiled with --fmad=false so the sample arithmetic (rotation, bins, trilinear weights) rounds like the
// CPU expression order.
//
// The reference scatters every window sample into a (4+2)x(4+2)x(8+2) histogram.  A scatter into one shared histogram
// needs shared-memory float atomics, which on sm_100a are CAS loops (ATOMS.CAST.SPIN) that serialise badly because
// neighbouring samples hit the same bins.  This path is atomics-free, barrier-free and deterministic:
//
//   describe_prep_kernel   thread per output keypoint: everything that depends on the keypoint alone (rounded centre, rotation
//                          scaled by the cell width, weight exponent, clipped window, slab-interval constants) is computed ONCE
//                          into a 64-byte DescParams record, and the 28-byte keypoint record is written.
//   describe_kernel        ONE WARP per keypoint (four independent warps per CTA, no __syncthreads anywhere).  The window is cut
//                          into the five unit slabs of the row coordinate, r0 <= rbin < r0+1 for r0 = -1..3, processed one after
//                          the other:
//     - floor(rbin) is a constant of the slab: no floor, no per-sample cell-row logic, every sample is evaluated exactly once, and
//       the border slabs (r0 = -1, 3) skip the half of the votes that falls outside the 4x4 grid with a warp-uniform branch;
//     - the slab's samples are enumerated as a compact list of 4-pixel chunks of window rows (row intervals from two slab
//       inequalities rounded outwards by 1e-3 px; the reference's own float test decides membership): step k gives chunk 8k+g
//       to lane group g, so all 32 lanes stay busy whatever the rotation, consecutive groups read consecutive pixels, and the
//       gradient-map load of the step three ahead is already in flight;
//     - the Gaussian weight is separable in window coordinates (rotation preserves i^2+j^2): exp(-(i^2+j^2)/(8 hw^2)) =
//       wrow[i] * wcol[j], two small per-keypoint tables instead of an expf per sample (|relative difference| to the reference's
//       exp of the rotated, rounded coordinates ~1e-7);
//     - the trilinear votes (reference operation order, :656-672) go to LANE-PRIVATE histograms in shared memory, layout
//       [cell-row buffer 0/1][6 cells][9 bins][32 lanes]: conflict-free plain read-modify-write, one address computation + immediates;
//       the two extra cells (c0 = -1 and c0+1 = 4) absorb the out-of-grid column votes without any select.  Slab r0 votes into
//       cell-rows r0 and r0+1, so the two buffers roll: after slab r0 cell-row r0 is complete, its 36 bins are summed over the 32
//       lanes (rotated 16-byte reads, conflict-free), the circular bin is folded, and lane l keeps output element (r0, l) in a register;
//     - tail in registers + shuffles: L2 -> clamp 0.2 -> x512 -> uchar (round half even) -> L1 -> sqrt.
// Only the inner 4x4 cells are kept by the reference (:676-684), so the border cells are never formed.
#include "sift_internal.cuh"

namespace siftb200 {
namespace {

constexpr int DW = 4, DB = 8;  // SIFT_DESCR_WIDTH, SIFT_DESCR_HIST_BINS (src/sift.cpp:12,15)
constexpr int NSLAB = DW + 1;  // floor(rbin) in -1..3
// Shape of a lane's private histogram (tuning knobs; the smaller it is, the more warps an SM holds -- the kernel is latency bound):
//   DESC_CELLS 6: cells c0+1 in 0..5, cells 0 and 5 are scratch for the out-of-grid column votes (no extra instruction)
//              5: one scratch cell (index 0) shared by c0 = -1 and c0+1 = 4 (one select)
//              4: no scratch: out-of-grid votes are clamped into the grid and their stores predicated off
//   DESC_BINS  9: bin o0+1 = 8 is folded onto bin 0 in the tail;  8: (o0+1) & 7 wraps at vote time (measured 41.3 vs 42.7 us/frame: the
//              second column-sum pass for the folded bin goes away and a tenth CTA fits an SM)
#ifndef DESC_CELLS
#define DESC_CELLS 4
#endif
#ifndef DESC_BINS
#define DESC_BINS 8
#endif
constexpr int PC = DESC_CELLS;  // private cells per cell-row
constexpr int PB = DESC_BINS;   // private bins per cell
constexpr int C1 = PC == 4 ? 0 : 1;            // private index of grid cell 0
constexpr int ROWBUF = PC * PB * 32;           // floats of one cell-row buffer: [cells][bins][32 lanes]
constexpr int NB = 84;         // window rows per block (row tables live in shared memory); radius <= 40 => one block
constexpr int WCOL = 96;       // window columns per block (column-weight table)
constexpr int NCHUNK = 248;    // chunk-list capacity per pass (a pipeline slab needs <= 244)
constexpr int NPAD = 8;        // dummy chunks behind the list: round up to a whole step
// shared memory per warp (floats): private histograms | rowA float4[NB+1] {i*sin, i*cos, wrow, row offset} (+1 dummy row) | wcol[WCOL] | chunk list
constexpr int WARP_FLOATS = 2 * ROWBUF + 4 * (NB + 1) + WCOL + NCHUNK + NPAD;
static_assert(WARP_FLOATS % 4 == 0, "16-byte alignment of every warp's region");
#ifndef DESC_RING
#define DESC_RING 3
#endif
#ifndef DESC_ROLES
#define DESC_ROLES 1  // 1: the walk is instantiated per slab role (border slabs skip half the votes); 0: one instantiation
#endif
constexpr int RING = DESC_RING;  // gradient-map loads in flight per lane
#ifndef DESC_WARPS_PER_CTA
#define DESC_WARPS_PER_CTA 2
#endif
constexpr int DESC_WARPS = DESC_WARPS_PER_CTA;  // independent warps per CTA
constexpr int DESC_SMEM_BYTES = DESC_WARPS * WARP_FLOATS * 4;
// register budget: the launch bound asks for one CTA less than shared memory admits when that avoids spills (DESC_MIN_CTAS)
constexpr int CTAS_PER_SM = (227 * 1024) / (DESC_SMEM_BYTES + 1024) < 32 / DESC_WARPS ? (227 * 1024) / (DESC_SMEM_BYTES + 1024) : 32 / DESC_WARPS;

#ifndef DESC_MIN_CTAS
#define DESC_MIN_CTAS CTAS_PER_SM
#endif

__device__ __forceinline__ int cv_round(float v) { return __float2int_rn(v); }
__device__ __forceinline__ int cv_floor(float v) { return __float2int_rd(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

// window radius of calcSIFTDescriptor (:587-590)
__device__ __forceinline__ int descr_radius(float scl, int rows, int cols) {
    const float hist_width = 3.f * scl;
    const int radius = cv_round(hist_width * 1.4142135623730951f * (DW + 1) * 0.5f);
    const int diag = (int)sqrt(((double)cols) * cols + ((double)rows) * rows);
    return min(radius, diag);
}

// Everything calcSIFTDescriptor derives from the keypoint alone (:579-612), for one keypoint of level (octave, layer).
__device__ DescParams make_params(int rows, int cols, int level, float ptx, float pty, float ori, float scl) {
    DescParams P;
    P.px = cv_round(ptx); P.py = cv_round(pty);
    const float cos_u = cosf(ori * (float)(3.1415926535897932384626433832795 / 180));
    const float sin_u = sinf(ori * (float)(3.1415926535897932384626433832795 / 180));
    const float hist_width = 3.f * scl;
    const int radius = descr_radius(scl, rows, cols);
    P.cos_t = cos_u / hist_width;
    P.sin_t = sin_u / hist_width;
    P.ori = ori;
    // exponent of the separable weight: (c_rot^2 + r_rot^2) * (-1/8) = (i^2 + j^2) * (cos_t^2 + sin_t^2) * (-1/8)
    P.es = (float)(-0.125 * ((double)P.cos_t * P.cos_t + (double)P.sin_t * P.sin_t));
    P.jmin = max(-radius, 1 - P.px); P.jmax = min(radius, cols - 2 - P.px);   // 0 < c < cols-1  (:621)
    P.imin = max(-radius, 1 - P.py); P.imax = min(radius, rows - 2 - P.py);   // 0 < r < rows-1
    // slab-interval constants: a direction whose coefficient moves the bin coordinate by < 0.05 over the whole window is "flat"
    // (inv = 0: rows are inside or outside as a whole); otherwise intervals are widened by a margin that covers their own rounding
    const float wspan = (float)(radius > 1 ? radius : 1);
    const bool flat_s = fabsf(P.sin_t) * wspan < 0.05f, flat_c = fabsf(P.cos_t) * wspan < 0.05f;
    P.inv_s = flat_s ? 0.f : 1.f / P.sin_t;
    P.inv_c = flat_c ? 0.f : 1.f / P.cos_t;
    P.mar_s = 1e-3f + 4e-6f * fabsf(P.inv_s);
    P.mar_c = 1e-3f + 4e-6f * fabsf(P.inv_c);
    P.level = level;
    return P;
}

// j-interval of {lo_v <= j*k + off <= hi_v} widened by `margin`, intersected into [lo, hi]; false when the row misses the slab.
// inv_k == 0 marks a flat direction (tested with slack).  The exact per-sample test decides membership; this only has to be a superset.
__device__ __forceinline__ bool slab(float inv_k, float margin, float off, float lo_v, float hi_v, float& lo, float& hi) {
    if (inv_k != 0.f) {
        const float u0 = (lo_v - off) * inv_k, u1 = (hi_v - off) * inv_k;
        lo = fmaxf(lo, fminf(u0, u1) - margin);
        hi = fminf(hi, fmaxf(u0, u1) + margin);
        return true;
    }
    return off > lo_v - 0.06f && off < hi_v + 0.06f;
}

// One slab's walk over its chunk list.  HAS_LO / HAS_HI: cell-rows r0 and r0+1 lie inside the 4x4 grid (the border slabs carry half the
// votes).  Step k hands chunk 8k + grp to lane group grp; the gradient-map load of step k+3 is issued before the votes of step k.
// A stage carries the sample's bin coordinates, weight and gradient; padding chunks sit on a dummy row whose row coordinate lies
// outside every slab, a lane past the end of its chunk gets cbin = 1e9: both are rejected by the reference's own range test (:620).
template <bool HAS_LO, bool HAS_HI>
__device__ __forceinline__ void walk_slab(const float2* __restrict__ mo, const float4* __restrict__ s_rowA, const float* __restrict__ s_wcol,
                                          const int* __restrict__ s_chunk, int nsteps, int cb0, float cos_t, float sin_t, float ori, float r0f, float r1f,
                                          float* __restrict__ priv, int grp, int gl) {
    const float bins_per_rad = DB / 360.f;
    auto issue = [&](float& st_rbin, float& st_cbin, float& st_w, float2& st_mo, int k) {
        const int e = s_chunk[k * 8 + grp];
        const int r = e & 255, joff = (e >> 8) & 255, cnt = e >> 16, j0 = joff + cb0;  // j0: window column of the chunk's first pixel
        const int d = min(gl, cnt - 1);  // clamped: always a pixel of the interval / a table entry
        const float4 rv = s_rowA[r];
        st_mo = __ldg(mo + (__float_as_int(rv.w) + j0 + d));
        st_w = rv.z * s_wcol[joff + d];
        const float jf = (float)(j0 + gl);
        const float c_rot = jf * cos_t - rv.x;
        const float r_rot = jf * sin_t + rv.y;
        st_rbin = r_rot + DW / 2 - 0.5f;
        const float cbin = c_rot + DW / 2 - 0.5f;
        st_cbin = gl < cnt ? cbin : 1e9f;
    };
    // one step: the trilinear votes of the sample of this lane, branch-free
    auto vote = [&](const float rbin, const float cbin, const float st_w, const float2 st_mo) {
        // floor(rbin) == r0 (this slab); rbin == -1 exactly (rejected by the reference) votes 0
        const bool racc = rbin >= r0f && rbin < r1f;
        const int c0 = cv_floor(cbin);  // 1e9 saturates to INT_MAX
        float obin = (st_mo.y - ori) * bins_per_rad;
        const int o0 = cv_floor(obin);
        obin -= o0;
        // private bin addresses: cells cA (c0) and cB (c0+1), bins oA (o0) and oB (o0+1); o0 in [-8, 7]: the reference's two wrap tests == & 7
        const int oA = (o0 & (DB - 1)) * 32;
        const int oB = PB == 9 ? oA + 32 : ((o0 + 1) & (DB - 1)) * 32;
        int cA, cB;
        bool stA = true, stB = true;
        float mag, cf;
        if (PC == 4) {
            // -1 < cbin < 4 (:620) <=> c0 in [-1, 3] (cbin == -1 exactly votes 0 into the grid); votes for cells outside the grid are formed
            // on a clamped address and not stored
            stA = (unsigned)c0 <= (unsigned)(DW - 1);
            stB = (unsigned)(c0 + 1) <= (unsigned)(DW - 1);
            cA = min(max(c0, 0), DW - 1) * (PB * 32);
            cB = min(max(c0 + 1, 0), DW - 1) * (PB * 32);
            mag = racc ? st_mo.x * st_w : 0.f;
            cf = cbin - (float)c0;
        } else {
            const bool acc = racc && cbin > -1 && cbin < DW;
            mag = acc ? st_mo.x * st_w : 0.f;
            const int c0s = acc ? c0 : 0;
            cf = acc ? cbin - (float)c0 : 0.f;
            cA = (c0s + 1) * (PB * 32);
            cB = (PC == 5 && c0s == DW - 1) ? 0 : cA + PB * 32;
        }
        const float rf = rbin - r0f;
        // trilinear split in the reference's operation order (:656-662)
        const float v_r1 = mag * rf, v_r0 = mag - v_r1;
        // buffer 0 = cell-row r0, buffer 1 = cell-row r0+1: every address is one of two registers plus an immediate.  All loads are issued before
        // the first store (one shared-memory latency per sample instead of four: the compiler cannot prove pA != pB and would otherwise keep
        // the read-modify-writes of the two cells in program order).  pA == pB only happens for clamped out-of-grid cells, and then at most
        // one of the two is stored, so loading everything first reads the same values.
        float* const pA = priv + cA + oA, * const pB = priv + cB + oA;
        constexpr int OB = PB == 9 ? 32 : 0;
        float* const pAo = PB == 9 ? pA : priv + cA + oB, * const pBo = PB == 9 ? pB : priv + cB + oB;
        float tA0 = 0.f, tA1 = 0.f, tB0 = 0.f, tB1 = 0.f, uA0 = 0.f, uA1 = 0.f, uB0 = 0.f, uB1 = 0.f;
        if (HAS_LO) { tA0 = pA[0]; tA1 = pAo[OB]; tB0 = pB[0]; tB1 = pBo[OB]; }
        if (HAS_HI) { uA0 = pA[ROWBUF]; uA1 = pAo[ROWBUF + OB]; uB0 = pB[ROWBUF]; uB1 = pBo[ROWBUF + OB]; }
        if (HAS_LO) {
            const float v_rc01 = v_r0 * cf, v_rc00 = v_r0 - v_rc01;
            const float vA = v_rc00 * obin, vB = v_rc01 * obin;
            tA0 += v_rc00 - vA; tA1 += vA; tB0 += v_rc01 - vB; tB1 += vB;
        }
        if (HAS_HI) {
            const float v_rc11 = v_r1 * cf, v_rc10 = v_r1 - v_rc11;
            const float vA = v_rc10 * obin, vB = v_rc11 * obin;
            uA0 += v_rc10 - vA; uA1 += vA; uB0 += v_rc11 - vB; uB1 += vB;
        }
        if (stA) {
            if (HAS_LO) { pA[0] = tA0; pAo[OB] = tA1; }
            if (HAS_HI) { pA[ROWBUF] = uA0; pAo[ROWBUF + OB] = uA1; }
        }
        if (stB) {
            if (HAS_LO) { pB[0] = tB0; pBo[OB] = tB1; }
            if (HAS_HI) { pB[ROWBUF] = uB0; pBo[ROWBUF + OB] = uB1; }
        }
    };
    // ring of RING stages (scalar arrays, fully unrolled: everything stays in registers), refilled in place; the list is padded to whole
    // steps only, so the refill is guarded by a warp-uniform bound
    float q_rbin[RING], q_cbin[RING], q_w[RING];
    float2 q_mo[RING];
#pragma unroll
    for (int q = 0; q < RING; ++q) {
        q_rbin[q] = 0.f; q_cbin[q] = 1e9f; q_w[q] = 0.f; q_mo[q] = make_float2(0.f, 0.f);
        if (q < nsteps) issue(q_rbin[q], q_cbin[q], q_w[q], q_mo[q], q);
    }
#pragma unroll 1
    for (int k = 0; k < nsteps; k += RING) {
#pragma unroll
        for (int q = 0; q < RING; ++q) {
            if (q == 0 || k + q < nsteps) {
                vote(q_rbin[q], q_cbin[q], q_w[q], q_mo[q]);
                if (k + q + RING < nsteps) issue(q_rbin[q], q_cbin[q], q_w[q], q_mo[q], k + q + RING);
            }
        }
    }
}

// calcSIFTDescriptor, src/sift.cpp:579-722, for one keypoint by one warp.  dst: 128 floats in global memory.
// mo: the level's gradient map {Mag, Ori} (detect.cu gradient_kernel) -- the values the reference computes per sample (:623-633).
// smem: this warp's WARP_FLOATS floats.
__device__ void calc_descriptor(const float2* __restrict__ mo, int pitch, const DescParams& P, float* __restrict__ smem, float* __restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const float cos_t = P.cos_t, sin_t = P.sin_t, es = P.es;
    float* s_priv = smem;                                            // [2][PC][PB][32]
    float4* s_rowA = reinterpret_cast<float4*>(smem + 2 * ROWBUF);   // [NB+1] {i*sin_t, i*cos_t, wrow, (int) element offset of the row}; [NB]: dummy
    float* s_wcol = reinterpret_cast<float*>(s_rowA + NB + 1);       // [WCOL] wcol[j - cb0]
    int* s_chunk = reinterpret_cast<int*>(s_wcol + WCOL);            // [NCHUNK + NPAD] row | (j0 - cb0) << 8 | pixels << 16
    float* priv = s_priv + lane;
    const int grp = lane >> 2, gl = lane & 3;
    const unsigned lt_mask = (1u << lane) - 1u;
    const bool single = P.imax - P.imin < NB && P.jmax - P.jmin < WCOL;  // the whole window is one block (every pipeline keypoint)
    {   // keep the level pointer as one opaque 64-bit value: a sample address is then a single multiply-add on a 32-bit index
        unsigned long long m = reinterpret_cast<unsigned long long>(mo);
        asm volatile("" : "+l"(m));
        mo = reinterpret_cast<const float2*>(m);
    }

    for (int k = lane * 4; k < 2 * ROWBUF; k += 128) *reinterpret_cast<float4*>(s_priv + k) = make_float4(0.f, 0.f, 0.f, 0.f);
    float v[DW];  // lane l: output element (cell-row a, l) = (a, cell l >> 3, bin l & 7)
#pragma unroll
    for (int a = 0; a < DW; ++a) v[a] = 0.f;

#pragma unroll 1
    for (int s = 0; s < NSLAB; ++s) {
        const float r0f = (float)(s - 1), r1f = (float)s;
        // The window is processed in blocks of NB rows x WCOL columns so that the tables fit shared memory whatever the keypoint size.
#pragma unroll 1
        for (int band0 = P.imin; band0 <= P.imax; band0 += NB)
#pragma unroll 1
        for (int cb0 = P.jmin; cb0 <= P.jmax; cb0 += WCOL) {
            const int nrows = min(NB, P.imax - band0 + 1);
            const int cb1 = min(P.jmax, cb0 + WCOL - 1);
            if (!single || s == 0) {
                __syncwarp();
                for (int k = lane; k <= cb1 - cb0; k += 32) s_wcol[k] = expf((float)((cb0 + k) * (cb0 + k)) * es);
                // dummy row for the padding chunks: a readable pixel, weight 0, row coordinate far outside every slab
                if (lane == 0) s_rowA[NB] = make_float4(0.f, 1e9f, 0.f, __int_as_float((P.py + band0) * pitch + P.px));
                for (int r = lane; r < nrows; r += 32) {
                    const int i = band0 + r;
                    s_rowA[r] = make_float4(i * sin_t, i * cos_t, expf((float)(i * i) * es), __int_as_float((P.py + i) * pitch + P.px));
                }
            }
            __syncwarp();
            // rows the slab can touch: i = hw^2 (cos_t (rbin - 1.5) - sin_t (cbin - 1.5)) over the slab's corners rbin in {r0, r0+1},
            // cbin in {-1, 4}, widened by a row each side (the exact tests below decide); only those 32-row rounds are visited
            int rb_first, rb_end;
            {
                const float hw2 = 1.f / (cos_t * cos_t + sin_t * sin_t);
                const float ar0 = cos_t * (r0f - 1.5f) * hw2, ar1 = cos_t * (r1f - 1.5f) * hw2;
                const float ac0 = sin_t * 2.5f * hw2, ac1 = -ac0;  // -sin_t * (cbin - 1.5) for cbin = -1 and 4
                const float i_lo = fminf(ar0, ar1) + fminf(ac0, ac1), i_hi = fmaxf(ar0, ar1) + fmaxf(ac0, ac1);
                const int r_lo = max(0, (int)floorf(i_lo) - 1 - band0), r_hi = min(nrows - 1, (int)ceilf(i_hi) + 1 - band0);
                rb_first = r_lo & ~31;
                rb_end = r_hi + 1;
            }
            int pass0 = 0, total;
            do {
                // ---- chunk list of this slab: every row interval cut into pieces of <= 4 pixels, in row-major order ----
                total = 0;
                for (int rb = rb_first; rb < rb_end; rb += 32) {
                    const int r = rb + lane;
                    int cnt = 0, jlo = 0, jhi = -1;
                    if (r < nrows) {
                        const float4 rv = s_rowA[r];
                        // the column-slab interval -1 < cbin < 4, then the row-slab interval r0 <= rbin < r0+1
                        float lo = (float)cb0, hi = (float)cb1;
                        const bool ok = slab(P.inv_c, P.mar_c, -rv.x + 1.5f, -1.f, 4.f, lo, hi) && slab(P.inv_s, P.mar_s, rv.y + 1.5f, r0f, r1f, lo, hi);
                        jlo = max(cb0, (int)ceilf(lo)); jhi = min(cb1, (int)floorf(hi));
                        if (ok && jlo <= jhi) cnt = (jhi - jlo + 4) >> 2;
                    }
                    // exclusive prefix sum of cnt (< 32) over the lanes: five independent ballots instead of a five-deep shuffle chain
                    int excl = 0;
#pragma unroll
                    for (int bit = 0; bit < 5; ++bit) excl += __popc(__ballot_sync(0xffffffffu, (cnt >> bit) & 1) & lt_mask) << bit;
                    // whole 4-pixel chunks (the code advances by four columns per chunk), then the row's last chunk with what is left
                    int idx = total + excl - pass0;
                    int code = r | ((jlo - cb0) << 8) | (4 << 16);
                    for (int c = 1; c < cnt; ++c, ++idx, code += 4 << 8)
                        if ((unsigned)idx < (unsigned)NCHUNK) s_chunk[idx] = code;
                    if (cnt > 0 && (unsigned)idx < (unsigned)NCHUNK) s_chunk[idx] = (code & 0xffff) | ((jhi - jlo - 4 * (cnt - 1) + 1) << 16);
                    total += __shfl_sync(0xffffffffu, excl + cnt, 31);
                }
                const int nchunks = min(NCHUNK, total - pass0);
                const int nsteps = (nchunks + 7) >> 3;
                // padding chunks (dummy row, one pixel at cb0) up to a whole step: the walk needs no per-lane bounds test
                if (nchunks + lane < nsteps * 8) s_chunk[nchunks + lane] = NB | (1 << 16);
                __syncwarp();
#if DESC_ROLES
                if (s == 0) walk_slab<false, true>(mo, s_rowA, s_wcol, s_chunk, nsteps, cb0, cos_t, sin_t, P.ori, r0f, r1f, priv, grp, gl);
                else if (s == DW) walk_slab<true, false>(mo, s_rowA, s_wcol, s_chunk, nsteps, cb0, cos_t, sin_t, P.ori, r0f, r1f, priv, grp, gl);
                else walk_slab<true, true>(mo, s_rowA, s_wcol, s_chunk, nsteps, cb0, cos_t, sin_t, P.ori, r0f, r1f, priv, grp, gl);
#else
                // one instantiation for all slabs (smaller code): the border slabs vote their out-of-grid half into the other buffer's
                // scratch -- buffer 0 of slab 0 and buffer 1 of slab 4 are never summed
                walk_slab<true, true>(mo, s_rowA, s_wcol, s_chunk, nsteps, cb0, cos_t, sin_t, P.ori, r0f, r1f, priv, grp, gl);
#endif
                __syncwarp();
                pass0 += NCHUNK;
            } while (pass0 < total);
        }
        // ---- slab done.  Buffer 0 holds cell-row a = s-1, now complete: sum its 4 x 9 bins over the 32 lane-private copies and fold the
        // circular bin.  Then buffer 1 (cell-row s, half done) becomes buffer 0 of the next slab and buffer 1 is cleared. ----
        if (s >= 1) {
            const int a = s - 1;
            const int cell = lane >> 3, k = lane & 7;
            // lane l reads copies 4q'..4q'+3 of ITS bin with q' = (q + l) & 7: the 8 lanes of a quarter-warp hit 8 different 16-byte bank groups
            auto colsum = [&](int bin) {
                const float* col = s_priv + ((cell + C1) * PB + bin) * 32;
                float acc = 0.f;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 t = *reinterpret_cast<const float4*>(col + (((q + lane) & 7) << 2));
                    acc += (t.x + t.y) + (t.z + t.w);
                }
                return acc;
            };
            float e = colsum(k);
            if (PB == 9 && k == 0) e += colsum(DB);  // hist[idx] += hist[idx+n] (:680); hist[idx+n+1] is never written since o0 <= n-1
#pragma unroll
            for (int q = 0; q < DW; ++q)
                if (q == a) v[q] = e;
            __syncwarp();
        }
        if (s < NSLAB - 1) {
            for (int i = lane * 4; i < ROWBUF; i += 128) {
                *reinterpret_cast<float4*>(s_priv + i) = *reinterpret_cast<const float4*>(s_priv + ROWBUF + i);
                *reinterpret_cast<float4*>(s_priv + ROWBUF + i) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __syncwarp();
        }
    }

    // ---- tail (:689-721): L2 norm, clamp at 0.2, renormalise to 512, uchar, L1, sqrt ----
    float part = 0.f;
#pragma unroll
    for (int a = 0; a < DW; ++a) part += v[a] * v[a];
    float nrm2 = warp_sum(part);
    const float thr = sqrtf(nrm2) * 0.2f;
    part = 0.f;
#pragma unroll
    for (int a = 0; a < DW; ++a) { v[a] = fminf(v[a], thr); part += v[a] * v[a]; }
    nrm2 = warp_sum(part);
    nrm2 = 512.f / fmaxf(sqrtf(nrm2), 1.1920928955078125e-7f);
    part = 0.f;
#pragma unroll
    for (int a = 0; a < DW; ++a) {
        int u = __float2int_rn(v[a] * nrm2);  // saturate_cast<uchar>: round half to even, clamp to 0..255
        u = min(max(u, 0), 255);
        v[a] = (float)u * nrm2;
        part += v[a];
    }
    float nrm1 = warp_sum(part);
    nrm1 = 1.f / fmaxf(nrm1, 1.1920928955078125e-7f);
#pragma unroll
    for (int a = 0; a < DW; ++a) dst[a * 32 + lane] = sqrtf(v[a] * nrm1);
}

// Per output keypoint: DescParams record + the cv::KeyPoint record (unpackOctave :724-731, calDescriptor :739-749).
__global__ void __launch_bounds__(128) describe_prep_kernel(const __grid_constant__ PyrView pv, const DetectBuf db, SiftKeypoint* __restrict__ kp_out, int cap) {
    const int f = blockIdx.y;
    int n = db.n_refined[f];
    if (n > db.cap_r) n = db.cap_r;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Refined rec = db.refined[(size_t)f * db.cap_r + i];
        const int np = db.n_peaks[(size_t)f * db.cap_r + i];
        const int base = db.kp_offset[(size_t)f * db.cap_r + i];
        // firstOctave = 0 in SIFT_NCL (:86)
        const int octave = rec.octave & 255, layer = (rec.octave >> 8) & 255;
        const float scale = 1.f / (1 << octave);
        const OctaveView& ov = pv.oct[octave];
        const float size = rec.size * scale;
        for (int k = 0; k < np; ++k) {
            const int slot = base + k;
            if (slot >= cap) break;
            const float kp_angle = db.angles[((size_t)f * db.cap_r + i) * kMaxPeaks + k];
            float angle = 360.f - kp_angle;
            if (fabsf(angle - 360.f) < 1.1920928955078125e-7f) angle = 0.f;
            db.dparams[(size_t)f * db.cap_r + slot] = make_params(ov.rows, ov.cols, octave | (layer << 8), rec.x * scale, rec.y * scale, angle, size * 0.5f);
            SiftKeypoint kp;
            kp.x = rec.x; kp.y = rec.y; kp.size = rec.size; kp.angle = kp_angle; kp.response = rec.response;
            kp.octave = rec.octave; kp.class_id = -1;
            kp_out[(size_t)f * cap + slot] = kp;
        }
    }
}

// calDescriptor on caller-supplied keypoints (stage-level API): any octave/layer the reference's CV_Assert admits.
__global__ void __launch_bounds__(128) describe_prep_given_kernel(const __grid_constant__ PyrView pv, const SiftKeypoint* __restrict__ kps, int n,
                                                                  DescParams* __restrict__ params, int first_octave, int* __restrict__ err) {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const SiftKeypoint kp = kps[p];
        int octave = kp.octave & 255;
        const int layer = (kp.octave >> 8) & 255;
        octave = octave < 128 ? octave : (-128 | octave);
        const float scale = octave >= 0 ? 1.f / (1 << octave) : (float)(1 << -octave);
        if (!(octave >= first_octave && layer <= kOctaveLayers + 2) || octave - first_octave >= pv.n_oct || layer >= kNumScales) {
            atomicExch(err, 1);  // CV_Assert, src/sift.cpp:744
            DescParams P{};
            P.level = -1;
            params[p] = P;
            continue;
        }
        const OctaveView& ov = pv.oct[octave - first_octave];
        float angle = 360.f - kp.angle;
        if (fabsf(angle - 360.f) < 1.1920928955078125e-7f) angle = 0.f;
        const float size = kp.size * scale;
        params[p] = make_params(ov.rows, ov.cols, (octave - first_octave) | (layer << 8), kp.x * scale, kp.y * scale, angle, size * 0.5f);
    }
}

// counts == nullptr: one list of n_fixed records (stage-level API); else frame f = blockIdx.y has min(counts[f], cap) records at stride pstride.
__global__ void __launch_bounds__(DESC_WARPS * 32, DESC_MIN_CTAS)
    describe_kernel(const __grid_constant__ PyrView pv, const DescParams* __restrict__ params, const int* __restrict__ counts, int n_fixed, int cap, int pstride,
                    float* __restrict__ desc_out) {
    extern __shared__ __align__(16) float smem_all[];
    const int warp = threadIdx.x >> 5;
    float* smem = smem_all + warp * WARP_FLOATS;
    const int f = blockIdx.y;
    const int n = counts ? min(counts[f], cap) : n_fixed;
    const DescParams* mine = params + (size_t)f * pstride;
    float* out = desc_out + (size_t)f * cap * 128;
    const int stride = gridDim.x * DESC_WARPS;
    int p = blockIdx.x * DESC_WARPS + warp;
    if (p >= n) return;
    DescParams P = mine[p];
    for (; p < n; p += stride) {
        const DescParams cur = P;
        if (p + stride < n) P = mine[p + stride];  // next record in flight while this keypoint is processed
        if (cur.level < 0) continue;
        const OctaveView& ov = pv.oct[cur.level & 255];
        const float2* img = ov.MO[cur.level >> 8] + (size_t)f * ov.frame_stride;
        calc_descriptor(img, ov.pitch, cur, smem, out + (size_t)p * 128);
    }
}

}  // namespace

namespace {
int g_grid_div = 1;  // CTAs per frame = resident CTAs of the device / g_grid_div (tuning probe, env SIFT_B200_DESC_GRIDDIV)
}

void init_describe_kernels() {
    cudaFuncSetAttribute(describe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DESC_SMEM_BYTES);
    if (const char* e = getenv("SIFT_B200_DESC_GRIDDIV")) g_grid_div = atoi(e) > 0 ? atoi(e) : 1;
}

int launch_describe(const PyrView& pv, const DetectBuf& db, int n_frames, SiftKeypoint* d_kp, float* d_desc, int cap, cudaStream_t st) {
    describe_prep_kernel<<<dim3(16, n_frames), 128, 0, st>>>(pv, db, d_kp, cap);
    const int per_frame = (num_sms() * CTAS_PER_SM + g_grid_div - 1) / g_grid_div;
    describe_kernel<<<dim3(per_frame, n_frames), DESC_WARPS * 32, DESC_SMEM_BYTES, st>>>(pv, db.dparams, db.n_kp, 0, cap, db.cap_r, d_desc);
    return 2;
}

int launch_describe_given(const PyrView& pv, const SiftKeypoint* d_kps, int n, float* d_desc, int first_octave, int* d_err, DescParams* d_params,
                          cudaStream_t st) {
    if (n <= 0) return 0;
    describe_prep_given_kernel<<<(n + 127) / 128, 128, 0, st>>>(pv, d_kps, n, d_params, first_octave, d_err);
    const int blocks = (n + DESC_WARPS - 1) / DESC_WARPS < num_sms() * CTAS_PER_SM ? (n + DESC_WARPS - 1) / DESC_WARPS : num_sms() * CTAS_PER_SM;
    describe_kernel<<<blocks, DESC_WARPS * 32, DESC_SMEM_BYTES, st>>>(pv, d_params, nullptr, n, n, 0, d_desc);
    return 2;
}

}  // namespace siftb200
