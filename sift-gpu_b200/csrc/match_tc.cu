// match_tc.cu -- brute-force L2 kNN(2) on the 5th-generation tensor cores (tcgen05 + TMEM), SURVEY section 8(f)-2.
//
// The matcher (reference driver src/main.cpp:25-40) is the one dense contraction on the path:
//   ||q - t||^2 = ||q||^2 + ||t||^2 - 2 q.t ,   Q [nq x 128] . T^T [128 x nt].
// Parity for this stage is IDENTICAL INDICES, so bf16 tensor-core products alone are not enough.  Scheme:
//   1. prep_kernel splits every descriptor into two bf16 terms x = hi + lo (hi = bf16(x), lo = bf16(x - hi)) and writes them as
//      128-row operand tiles already in the canonical K-major, no-swizzle UMMA layout (8-row x 16-byte core matrices; chunk c of
//      K = columns 8c..8c+7 at byte c*2048 + row*16), followed by the rows' exact fp32 squared norms: one contiguous 66 048-byte
//      blob per tile, so a tile moves with a single 1-D TMA bulk copy;
//   2. match_tc_kernel: one CTA = 128 queries x a strided subset of the train tiles, warp-specialised:
//        warp 4  TMA producer: the query tile once, then train tiles into a 2-stage shared-memory ring (mbarrier complete_tx);
//        warp 5  MMA issuer: per train tile hi.hi + hi.lo + lo.hi = 3 x 8 tcgen05.mma.kind::f16 (M = N = 128, K = 16, fp32
//                accumulation) into one of TWO 128-column TMEM accumulators, tcgen05.commit -> mbarriers;
//        warps 0-3  epilogue: thread = query row = TMEM lane; tcgen05.ld.32x32b.x32 the 128 accumulator columns of the tile
//                that finished while the next tile's MMAs run.  Selection is branch-free and warp-uniform: a column's approximate
//                squared distance becomes a sortable integer key (float bits, column in the low 5 bits); a 32-column chunk keeps
//                its two smallest keys with integer min/max chains, and the thread keeps the 4 chunks with the smallest minimum.
//                The two nearest train rows of a split are always among {min, second min} of those chunks: 8 shortlist entries;
//      dot products are good to ~2^-16 relative;
//   3. every shortlist entry carries its approximate squared distance d2 and an error bound e = 2^-15 (|q|^2 + |t|^2) (twice the
//      worst case of the split products, by Cauchy-Schwarz); rerank_kernel finds the second-smallest upper bound d2 + e of a query
//      with a warp reduction and evaluates exactly (fp64, like match.cu) only the entries whose lower bound d2 - e does not exceed
//      it -- usually two or three -- then emits the best two.
// With exact arithmetic the two nearest rows of a split are always in its shortlist.  With the ~3e-5 error of the split products a
// true top-2 neighbour could be lost if three or more rows of ONE 32-row chunk, or the minima of five or more chunks, sit within
// that error of it.  That case is DETECTED, not tolerated: every (query, split) also records the smallest approximate distance any
// row OUTSIDE its shortlist can have (min of the kept chunks' second minima and the fourth chunk minimum); if that bound minus the
// worst-case error reaches the query's second-smallest upper bound, the shortlist is not provably complete and rerank_kernel
// evaluates that query against every train row exactly.  The result is therefore always the exact kernel's (identical indices,
// distances and tie order); the tests compare with the exact matcher, the oracle and the reference fixture.  Every mbarrier wait is
// bounded and traps.
#include <cuda_bf16.h>

#include "sift_internal.cuh"

namespace siftb200 {
namespace {

constexpr int TM = 128, TN = 128, DK = 128;         // tile: queries x train rows x descriptor length
constexpr int CHUNK_BYTES = TM * 16;                 // one K-chunk (8 bf16) of 128 rows
constexpr int OPER_BYTES = (DK / 8) * CHUNK_BYTES;   // 32 KB per operand matrix (hi or lo)
constexpr int TILE_BYTES = 2 * OPER_BYTES + TN * 4;  // hi | lo | norms
constexpr int NCHUNK = 4;                            // chunks (32 train rows) kept per (query, train split)
constexpr int SHORT = 2 * NCHUNK;                    // shortlist entries per (query, train split): two per kept chunk
constexpr int SLOTS = SHORT + 1;                     // + one record {idx -1, d2 = lower bound of every approximate distance outside the shortlist}
struct Cand { int32_t idx; float d2, e; };           // train row, approximate squared distance, bound on its error
constexpr int NSTAGE = 2;
constexpr int TC_THREADS = 192;
constexpr int SM_A = 0, SM_B = 2 * OPER_BYTES, SM_BAR = SM_B + NSTAGE * TILE_BYTES;
constexpr int TC_SMEM_BYTES = SM_BAR + 128;
static_assert(TILE_BYTES % 16 == 0 && SM_B % 128 == 0, "bulk-copy alignment");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE matrix descriptor (cute::UMMA::SmemDescriptor): start address, LBO (between the two K chunks of one
// MMA), SBO (between 8-row groups), all in 16-byte units; version = 1 (Blackwell).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((CHUNK_BYTES >> 4) & 0x3FFF) << 16;  // leading byte offset: next K chunk
    d |= (uint64_t)((128 >> 4) & 0x3FFF) << 32;          // stride byte offset: next 8 rows
    d |= (uint64_t)1 << 46;                              // descriptor version
    return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = BF16, both K-major, N = 128, M = 128.
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((TN >> 3) << 17) | ((TM >> 4) << 24);

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (int spin = 0; spin < (1 << 24); ++spin) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        __nanosleep(32);  // the pollers share issue slots with the epilogue warps
    }
    __trap();  // never hang the GPU on a protocol mistake
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ---- 1. operand tiles -----------------------------------------------------------------------------------------------------------
// block = one 128-row tile, 512 threads: thread (row, quarter) converts 32 of the row's 128 floats; all 8 float4 loads are issued first.
// Blocks [0, q_blocks) convert the query matrix, the rest the train matrix (whose tiles follow the query tiles in `tiles`).
__global__ void __launch_bounds__(512) prep_kernel(const float* __restrict__ q, int nq, const float* __restrict__ t, int nt, int q_blocks,
                                                   uint8_t* __restrict__ tiles, int* __restrict__ tn_max_bits) {
    __shared__ float part[4][TM];
    const bool is_q = (int)blockIdx.x < q_blocks;
    const float* src = is_q ? q : t;
    const int n = is_q ? nq : nt;
    const int row = threadIdx.x & (TM - 1), qtr = threadIdx.x >> 7;
    const int grow = ((int)blockIdx.x - (is_q ? 0 : q_blocks)) * TM + row;
    uint8_t* hi = tiles + (size_t)blockIdx.x * TILE_BYTES;
    uint8_t* lo = hi + OPER_BYTES;
    float4 v[8];
    const float4* p = reinterpret_cast<const float4*>(src + (size_t)(grow < n ? grow : 0) * DK) + 8 * qtr;
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = grow < n ? __ldg(p + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float x[8] = {v[2 * k].x, v[2 * k].y, v[2 * k].z, v[2 * k].w, v[2 * k + 1].x, v[2 * k + 1].y, v[2 * k + 1].z, v[2 * k + 1].w};
        uint32_t ph[4], pl[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(x[2 * e]), h1 = __float2bfloat16_rn(x[2 * e + 1]);
            const __nv_bfloat16 l0 = __float2bfloat16_rn(x[2 * e] - __bfloat162float(h0)), l1 = __float2bfloat16_rn(x[2 * e + 1] - __bfloat162float(h1));
            ph[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            pl[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
            s += x[2 * e] * x[2 * e] + x[2 * e + 1] * x[2 * e + 1];
        }
        const int off = (4 * qtr + k) * CHUNK_BYTES + row * 16;
        *reinterpret_cast<uint4*>(hi + off) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
        *reinterpret_cast<uint4*>(lo + off) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    }
    part[qtr][row] = s;
    __syncthreads();
    // rows past the end of the matrix get norm = +inf: as train rows they can never enter a shortlist, so the epilogue needs no column mask
    if (qtr == 0) {
        const float nrm = (part[0][row] + part[1][row]) + (part[2][row] + part[3][row]);
        reinterpret_cast<float*>(hi + 2 * OPER_BYTES)[row] = grow < n ? nrm : INFINITY;
        // largest squared train norm (non-negative floats order like their bit patterns): bounds the error of rows outside a shortlist
        if (!is_q && grow < n) {
            int m = __float_as_int(nrm);
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, sft));
            if ((threadIdx.x & 31) == 0) atomicMax(tn_max_bits, m);
        }
    }
}

// ---- 2. tensor-core shortlist ---------------------------------------------------------------------------------------------------
// grid = (query tiles, train splits): CTA (x, y) scans train tiles y, y + gridDim.y, ... and writes shortlist slot y of its queries
__global__ void __launch_bounds__(TC_THREADS, 1) match_tc_kernel(const uint8_t* __restrict__ q_tiles, int nq, const uint8_t* __restrict__ t_tiles, int nt,
                                                                 Cand* __restrict__ cand) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
    // bars: [0] a_full, [1..2] b_full, [3..4] b_empty, [5..6] acc_full, [7..8] acc_empty; then the TMEM base slot
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * TM;
    const int n_tt = (nt + TN - 1) / TN;
    const int n_my = (n_tt - (int)blockIdx.y + (int)gridDim.y - 1) / (int)gridDim.y;  // train tiles of this CTA

    if (warp == 4) {  // one warp allocates 2 x 128 TMEM columns (two fp32 accumulators of 128 lanes x 128 columns)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * TN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        mbar_init(BAR(0), 1);
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(BAR(1 + s), 1);      // b_full: the producer's expect_tx arrive
            mbar_init(BAR(3 + s), 1 + 4);  // b_empty: tcgen05.commit + one lane of each epilogue warp (they read the tile's norms)
            mbar_init(BAR(5 + s), 1);      // acc_full: tcgen05.commit
            mbar_init(BAR(7 + s), 4);      // acc_empty: one lane of each epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {  // ---- TMA producer ----
            mbar_expect_tx(BAR(0), 2 * OPER_BYTES);
            bulk_g2s(smem_u32(smem + SM_A), q_tiles + (size_t)blockIdx.x * TILE_BYTES, 2 * OPER_BYTES, BAR(0));
            for (int it = 0; it < n_my; ++it) {
                const int s = it % NSTAGE, ph = (it / NSTAGE) & 1;
                mbar_wait(BAR(3 + s), ph ^ 1);  // stage free (passes at once on the first lap)
                mbar_expect_tx(BAR(1 + s), TILE_BYTES);
                bulk_g2s(smem_u32(smem + SM_B + s * TILE_BYTES), t_tiles + (size_t)(blockIdx.y + it * gridDim.y) * TILE_BYTES, TILE_BYTES, BAR(1 + s));
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {  // ---- MMA issuer ----
            mbar_wait(BAR(0), 0);
            const uint32_t ah = smem_u32(smem + SM_A), al = ah + OPER_BYTES;
            for (int it = 0; it < n_my; ++it) {
                const int s = it % NSTAGE, ph = (it / NSTAGE) & 1;
                mbar_wait(BAR(1 + s), ph);      // operands landed
                mbar_wait(BAR(7 + s), ph ^ 1);  // accumulator s drained by the epilogue
                asm volatile("tcgen05.fence::after_thread_sync;");
                const uint32_t bh = smem_u32(smem + SM_B + s * TILE_BYTES), bl = bh + OPER_BYTES;
                const uint32_t d_tmem = tmem_base + (uint32_t)(s * TN);
                uint32_t accumulate = 0;
#pragma unroll 1
                for (int term = 0; term < 3; ++term) {  // hi.hi, hi.lo, lo.hi
                    const uint32_t a0 = term == 2 ? al : ah, b0 = term == 1 ? bl : bh;
#pragma unroll 1
                    for (int ks = 0; ks < DK / 16; ++ks) {
                        const uint64_t da = umma_desc(a0 + ks * 2 * CHUNK_BYTES), db = umma_desc(b0 + ks * 2 * CHUNK_BYTES);
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
                            "l"(da), "l"(db), "r"(IDESC), "r"(accumulate)
                            : "memory");
                        accumulate = 1;
                    }
                }
                // both arrive when every MMA above has completed (the commit implies tcgen05.fence::before_thread_sync)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(BAR(3 + s)) : "memory");
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(BAR(5 + s)) : "memory");
            }
        }
    } else {  // ---- epilogue warps 0-3: thread = query row = TMEM lane (warp w may only touch lanes [32w, 32w+32)) ----
        const int row = tid;  // 0..127
        const float qn = __ldg(reinterpret_cast<const float*>(q_tiles + (size_t)blockIdx.x * TILE_BYTES + 2 * OPER_BYTES) + row);
        // key = float bits of (d2 + bias) with the column's position in its chunk in the low 5 bits; bias = 2^-10 |q|^2 exceeds the error
        // of d2 for any row close enough to matter, so keys of candidates are positive and integer order = distance order
        const float bias = 9.765625e-4f * qn + 1e-30f, qb = qn + bias;
        int Lk[NCHUNK], Ls[NCHUNK], Lc[NCHUNK];  // kept chunks, ascending by minimum: min key, second-min key, chunk index
#pragma unroll
        for (int k = 0; k < NCHUNK; ++k) { Lk[k] = 0x7FFFFFFF; Ls[k] = 0x7FFFFFFF; Lc[k] = -1; }
        for (int it = 0; it < n_my; ++it) {
            const int s = it % NSTAGE, ph = (it / NSTAGE) & 1;
            const int t0 = (blockIdx.y + it * gridDim.y) * TN;
            const float* t_norm = reinterpret_cast<const float*>(smem + SM_B + s * TILE_BYTES + 2 * OPER_BYTES);
            mbar_wait(BAR(1 + s), ph);  // the tile's norms (TMA-written) are visible to this thread
            mbar_wait(BAR(5 + s), ph);  // its accumulator is complete
            asm volatile("tcgen05.fence::after_thread_sync;");
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(s * TN);
#pragma unroll 1
            for (int c0 = 0; c0 < TN; c0 += 32) {
                uint32_t r[32];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
                    "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                      "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
                      "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
                      "=r"(r[31])
                    : "r"(taddr + c0));
                asm volatile("tcgen05.wait::ld.sync.aligned;");
                // two smallest keys of the chunk; four independent (min, second-min) chains keep the dependent depth at 8 columns
                int m1[4], m2[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) m1[j] = m2[j] = 0x7FFFFFFF;
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    const float4 tn = *reinterpret_cast<const float4*>(t_norm + c0 + 4 * k4);  // same address in every lane: broadcast
                    const float tnv[4] = {tn.x, tn.y, tn.z, tn.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float d = __fmaf_rn(-2.f, __uint_as_float(r[4 * k4 + j]), tnv[j] + qb);
                        const int x = (__float_as_int(d) & ~31) | (4 * k4 + j);
                        m2[j] = min(m2[j], max(m1[j], x));
                        m1[j] = min(m1[j], x);
                    }
                }
#pragma unroll
                for (int w = 2; w > 0; w >>= 1)
#pragma unroll
                    for (int j = 0; j < w; ++j) {
                        const int lo = min(m1[j], m1[j + w]), hi = max(m1[j], m1[j + w]);
                        m2[j] = min(hi, min(m2[j], m2[j + w]));
                        m1[j] = lo;
                    }
                // insert the chunk into the sorted list (branch-free compare-and-swap down the list; strict <: earlier chunk wins ties)
                int k = m1[0], sc = m2[0], c = (t0 + c0) >> 5;
#pragma unroll
                for (int j = 0; j < NCHUNK; ++j) {
                    const bool pr = k < Lk[j];
                    const int tk = pr ? Lk[j] : k, ts = pr ? Ls[j] : sc, tc = pr ? Lc[j] : c;
                    Lk[j] = pr ? k : Lk[j]; Ls[j] = pr ? sc : Ls[j]; Lc[j] = pr ? c : Lc[j];
                    k = tk; sc = ts; c = tc;
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(BAR(7 + s));  // accumulator s may be overwritten
                mbar_arrive(BAR(3 + s));  // the tile's norms are no longer needed: the stage may be refilled
            }
        }
        if (q0 + row < nq) {
#pragma unroll
            for (int j = 0; j < NCHUNK; ++j) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int key = h ? Ls[j] : Lk[j];
                    Cand cd;
                    cd.idx = -1; cd.d2 = 0.f; cd.e = 0.f;
                    const int col = Lc[j] * 32 + (key & 31);
                    if (Lc[j] >= 0 && key != 0x7FFFFFFF && col < nt) {
                        const float tn = __ldg(reinterpret_cast<const float*>(t_tiles + (size_t)(col / TN) * TILE_BYTES + 2 * OPER_BYTES) + (col % TN));
                        cd.idx = col;
                        cd.d2 = __int_as_float(key & ~31) - bias;
                        cd.e = 6.103515625e-5f * (qn + tn) + 1e-30f;  // 2^-14 (|q|^2 + |t|^2): split products + the 5 key bits
                    }
                    cand[((size_t)(q0 + row) * gridDim.y + blockIdx.y) * SLOTS + 2 * j + h] = cd;
                }
            }
            // every train row of this split that is NOT in the shortlist has an approximate key >= its chunk's second minimum (kept
            // chunks) or >= the largest kept chunk minimum (other chunks)
            int bkey = Lk[NCHUNK - 1];
#pragma unroll
            for (int j = 0; j < NCHUNK; ++j) bkey = min(bkey, Ls[j]);
            Cand bd;
            bd.idx = -1; bd.e = 0.f;
            bd.d2 = bkey == 0x7FFFFFFF ? INFINITY : __int_as_float(bkey & ~31) - bias;
            cand[((size_t)(q0 + row) * gridDim.y + blockIdx.y) * SLOTS + SHORT] = bd;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * TN));
    }
}

// ---- 3. exact fp64 re-rank: warp per query ----------------------------------------------------------------------------------------
// Pass 1 (lanes over the shortlist entries): U2 = second-smallest upper bound d2 + e.  Completeness check: a row outside the
// shortlists has approximate distance >= its split's bound record and error <= e_max = 2^-14 (|q|^2 + max |t|^2); if bound - e_max
// <= U2 for any split, such a row could still be one of the two nearest and the query is matched against ALL train rows instead.
// Pass 2: every entry with d2 - e <= U2 could still be one of the two nearest; those are evaluated exactly, lane = 4 of the 128
// components (same arithmetic as match.cu).
__global__ void __launch_bounds__(256) rerank_kernel(const float* __restrict__ q, int nq, const float* __restrict__ t, int nt, const Cand* __restrict__ cand,
                                                     int n_cand, const int* __restrict__ tn_max_bits, float* __restrict__ dist, int32_t* __restrict__ idx,
                                                     int* __restrict__ n_fallback) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + warp;
    if (i >= nq) return;
    const Cand* mine = cand + (size_t)i * n_cand;
    float u1 = 3.4e38f, u2 = 3.4e38f;  // the two smallest upper bounds seen by this lane
    float bound = INFINITY;            // smallest "outside the shortlist" bound over the splits
    for (int k = lane; k < n_cand; k += 32) {
        const Cand c = mine[k];
        if (k % SLOTS == SHORT) bound = fminf(bound, c.d2);
        if (c.idx < 0) continue;
        const float u = c.d2 + c.e;
        if (u < u1) { u2 = u1; u1 = u; } else if (u < u2) u2 = u;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const float v1 = __shfl_xor_sync(0xffffffffu, u1, s), v2 = __shfl_xor_sync(0xffffffffu, u2, s);
        const float n1 = fminf(u1, v1), n2 = fminf(fmaxf(u1, v1), fminf(u2, v2));
        u1 = n1; u2 = n2;
        bound = fminf(bound, __shfl_xor_sync(0xffffffffu, bound, s));
    }
    const float4 a = __ldg(reinterpret_cast<const float4*>(q + (size_t)i * 128) + lane);
    float qn = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) qn += __shfl_xor_sync(0xffffffffu, qn, s);
    const float e_max = 6.103515625e-5f * (qn * 1.0001f + __int_as_float(*tn_max_bits)) + 1e-30f;
    const bool complete = bound - e_max > u2;  // false also when fewer than two entries exist (u2 = 3.4e38)
    double b0 = INFINITY, b1 = INFINITY;
    int i0 = -1, i1 = -1;
    auto consider = [&](int j) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(t + (size_t)j * 128) + lane);
        const double e0 = (double)a.x - (double)b.x, e1 = (double)a.y - (double)b.y, e2 = (double)a.z - (double)b.z, e3 = (double)a.w - (double)b.w;
        double d = e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) d += __shfl_xor_sync(0xffffffffu, d, s);
        d = sqrt(d);
        // ascending distance, exact ties to the lower train index (BFMatcher order)
        if (d < b0 || (d == b0 && j < i0)) { b1 = b0; i1 = i0; b0 = d; i0 = j; }
        else if (d < b1 || (d == b1 && j < i1)) { b1 = d; i1 = j; }
    };
    if (complete) {
        for (int k0 = 0; k0 < n_cand; k0 += 32) {
            Cand c; c.idx = -1; c.d2 = 0.f; c.e = 0.f;
            if (k0 + lane < n_cand) c = mine[k0 + lane];
            unsigned todo = __ballot_sync(0xffffffffu, c.idx >= 0 && c.d2 - c.e <= u2);
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                consider(__shfl_sync(0xffffffffu, c.idx, src));
            }
        }
    } else {
        for (int j = 0; j < nt; ++j) consider(j);
        if (lane == 0 && n_fallback) atomicAdd(n_fallback, 1);
    }
    if (lane == 0) {
        dist[2 * i] = (float)b0; dist[2 * i + 1] = (float)b1;
        idx[2 * i] = i0; idx[2 * i + 1] = i1;
    }
}

}  // namespace

void init_match_tc_kernels() { cudaFuncSetAttribute(match_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES); }

// Train splits: one CTA per SM (193 KB of shared memory), so the grid runs in waves of kNumSMs CTAs; a CTA costs its train tiles
// plus ~2 tiles' worth of fixed work (query tile load, pipeline fill).  Pick the split count with the cheapest last wave.
int match_tc_splits(int nq, int nt) {
    const int qt = (nq + TM - 1) / TM, tt = (nt + TN - 1) / TN;
    int best = 1;
    double best_cost = 1e30;
    for (int s = 1; s <= tt && s <= 64; ++s) {
        const int waves = (qt * s + num_sms() - 1) / num_sms();
        const double cost = waves * ((tt + s - 1) / s + 2.0);
        if (cost < best_cost) { best_cost = cost; best = s; }
    }
    return best;
}

size_t match_tc_scratch_bytes(int nq, int nt) {
    const size_t qt = (nq + TM - 1) / TM, tt = (nt + TN - 1) / TN;
    return (qt + tt) * (size_t)TILE_BYTES + (size_t)nq * match_tc_splits(nq, nt) * SLOTS * sizeof(Cand) + 16;
}

// L2 only.  d_scratch: match_tc_scratch_bytes(nq, nt) bytes, 128-byte aligned (operand tiles, then the shortlists).
int launch_match_tc(const float* d_q, int nq, const float* d_t, int nt, void* d_scratch, float* d_dist, int32_t* d_idx, cudaStream_t st) {
    if (nq <= 0) return 0;
    const int qt = (nq + TM - 1) / TM, tt = (nt + TN - 1) / TN;
    const int splits = match_tc_splits(nq, nt);
    uint8_t* q_tiles = static_cast<uint8_t*>(d_scratch);
    uint8_t* t_tiles = q_tiles + (size_t)qt * TILE_BYTES;
    Cand* cand = reinterpret_cast<Cand*>(t_tiles + (size_t)tt * TILE_BYTES);
    int* words = reinterpret_cast<int*>(cand + (size_t)nq * splits * SLOTS);  // [0] max squared train norm (float bits), [1] queries matched exhaustively
    cudaMemsetAsync(words, 0, 8, st);
    prep_kernel<<<qt + tt, 512, 0, st>>>(d_q, nq, d_t, nt, qt, q_tiles, words);
    match_tc_kernel<<<dim3(qt, splits), TC_THREADS, TC_SMEM_BYTES, st>>>(q_tiles, nq, t_tiles, nt, cand);
    rerank_kernel<<<(nq + 7) / 8, 256, 0, st>>>(d_q, nq, d_t, nt, cand, splits * SLOTS, words, d_dist, d_idx, words + 1);
    return 3;
}

}  // namespace siftb200
