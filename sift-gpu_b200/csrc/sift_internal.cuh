// sift_internal.cuh -- shared declarations of the sm_100a SIFT path (not part of the public C ABI).
//
// Data layout in HBM (per handle, for up to max_batch frames F):
//   per octave o: G[0..2] (G[3..4] only for the stage-level pyramid API) and D[0..3], each
//   [F][rows_o][pitch_o] float32 with pitch_o = cols_o rounded up to 32 floats (128-byte rows so every
//   warp-wide row segment is a whole number of 128-byte lines); MO[1..2] = {gradient magnitude, orientation}
//   of G1/G2 as float2 with the same pitch (all five in the stage-level API).
//   refined-point records, orientation peaks and ordering arrays: [F][cap_refined].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sift_b200.h"

namespace siftb200 {

constexpr int kMaxOctaves = 8;
constexpr int kNumScales = 5;      // src/sift.cpp:5
constexpr int kOctaveLayers = 2;   // src/sift.cpp:4
constexpr int kImgBorder = 5;      // src/sift.cpp:21
constexpr int kMaxInterpSteps = 5; // src/sift.cpp:24
constexpr int kOriBins = 36;       // src/sift.cpp:27
constexpr int kMaxPeaks = 18;      // strict local maxima of a 36-bin circular histogram
constexpr int kMaxRadius = 18;     // floor(3*sig[4]) = floor(3*6.196774)
#ifndef GR_ROWS_
#define GR_ROWS_ 8
#endif
constexpr int kGradRows = GR_ROWS_; // rows of a gradient strip (kGradCols aligned columns x kGradRows rows per warp, detect.cu)
constexpr int kGradCols = 64;
#ifndef EX_ROWS_
#define EX_ROWS_ 64
#endif
constexpr int kExtremaRows = EX_ROWS_;  // output rows of an extrema strip (30 output columns x kExtremaRows rows per warp, detect.cu)
constexpr int kExtremaCols = 30;

struct OctaveView {
    float* G[kNumScales];     // Gaussian levels (G[3], G[4] may be null in the fused pipeline)
    float* D[kNumScales - 1]; // DoG levels
    float2* MO[kNumScales];   // per-pixel gradient {magnitude, fastAtan2 orientation in degrees} of G[i] (null if not built)
    int rows, cols, pitch;    // pitch in floats
    size_t frame_stride;      // floats between consecutive frames of one level
    int tile_base;            // first extrema strip (30x16 outputs) of this octave in the flattened strip index
    int tiles_x;
    int grad_tile_base;       // first gradient strip (kGradCols x kGradRows outputs) of this octave in the flattened strip index
    int grad_tiles_x;
};

struct PyrView {
    OctaveView oct[kMaxOctaves];
    int n_oct;
    int total_tiles;
    int total_grad_tiles;
};

// One record per extremum that survived adjustLocalExtrema (src/sift.cpp:287-388).
struct Refined {
    uint32_t key;     // scan-order key: o<<27 | (layer0-1)<<26 | r0<<13 | c0   (initial, un-refined position)
    int32_t octave;   // packed like KeyPoint::octave
    float x, y, size, response;
    uint32_t rc;      // refined integer position r1<<16 | c1 in octave coordinates
    uint32_t pad;
};

// Everything calcSIFTDescriptor derives from the keypoint alone (describe.cu: describe_prep_kernel writes one per output keypoint).
struct DescParams {
    float cos_t, sin_t;          // rotation divided by the cell width 3*scl (src/sift.cpp:594-596)
    float es;                    // exponent scale of the separable Gaussian weight
    float ori;                   // 360 - kpt.angle, degrees
    float inv_s, inv_c;          // 1/sin_t, 1/cos_t for the slab intervals (0: flat direction)
    float mar_s, mar_c;          // interval margins
    int px, py;                  // cvRound of the keypoint position in octave coordinates
    int jmin, jmax, imin, imax;  // window clipped to 0 < r < rows-1, 0 < c < cols-1
    int level;                   // octave index | layer << 8; -1: rejected by the reference's CV_Assert
    int pad;
};
static_assert(sizeof(DescParams) == 64, "DescParams is read as four 16-byte words");

struct DetectBuf {
    uint32_t* cand;      // [F][cap_c] scan-order keys of the 27-neighbour extrema (before refinement)
    int* n_cand;         // [F]
    int cap_c;
    Refined* refined;    // [F][cap_r]
    int* n_refined;      // [F]  (atomic counters, may exceed cap_r)
    float* angles;       // [F][cap_r][kMaxPeaks]
    int* n_peaks;        // [F][cap_r]
    int* order;          // [F][cap_r]  sorted position -> refined index
    int* kp_offset;      // [F][cap_r]  refined index -> first output slot
    unsigned long long* sort_tmp; // [F][cap_r_pow2] global scratch when the frame does not fit shared memory
    DescParams* dparams; // [F][cap_r]  per output keypoint slot
    int* n_kp;           // [F] output keypoints actually produced (counts_out may carry the overflow sentinel instead)
    int cap_r;
    int cap_r_pow2;
};

// Blur taps: c_taps[0] = base sigma sqrt(1.6^2+0.2^2), c_taps[1..4] = sig[1..4] (src/sift.cpp:237-245).
// 1-D factor of the reference's 2-D kernel: exp(-i^2/den)/sqrt(2 PI sigma^2), den = float(2*sigma*sigma).
constexpr int kTapStride = 40;
int num_sms();  // multiProcessorCount of the current device (148 on B200), cached per device
void upload_taps(const float host_taps[5][kTapStride]);
void init_pyramid_kernels();
void init_detect_kernels();
void init_describe_kernels();
void init_match_tc_kernels();

// ---- launchers (each returns the number of kernels it launched) ---------------------------------
int launch_base_blur(const float* src, size_t src_frame_stride, int src_pitch, const uint8_t* src_u8, const OctaveView& o0, int n_frames,
                     cudaStream_t st);
int launch_octave(const PyrView& pv, int o, int n_frames, bool write_all_levels, cudaStream_t st);
// exact-order pyramid (pyramid_exact.cu): same outputs as the two launchers above, bit-identical to the reference's loop
void upload_taps_2d(const float* host_k2d);
int k2d_total();
int k2d_offset(int s);
int launch_exact_base(const float* src, size_t src_frame_stride, int src_pitch, const uint8_t* src_u8, const OctaveView& o0, int n_frames, cudaStream_t st);
int launch_exact_octave(const PyrView& pv, int o, int n_frames, bool write_all_levels, cudaStream_t st);
int launch_generic_blur(const float* src, float* dst, int rows, int cols, const float* d_taps, int radius, int taps_hi, cudaStream_t st);
int launch_dog(const PyrView& pv, int n_frames, cudaStream_t st);
int launch_rgb2gray_u8(const uint8_t* src, uint8_t* dst, size_t n_pixels, cudaStream_t st);
int launch_upsample2x(const float* src, float* dst, int rows, int cols, int n_frames, cudaStream_t st);
int launch_gradient(const PyrView& pv, int n_frames, cudaStream_t st);
int launch_extrema(const PyrView& pv, const DetectBuf& db, int n_frames, cudaStream_t st);
int launch_orientation(const PyrView& pv, const DetectBuf& db, int n_frames, cudaStream_t st);
int launch_order_scan(const DetectBuf& db, int n_frames, int* d_counts, cudaStream_t st);
int launch_describe(const PyrView& pv, const DetectBuf& db, int n_frames, SiftKeypoint* d_kp, float* d_desc, int cap, cudaStream_t st);
int launch_describe_given(const PyrView& pv, const SiftKeypoint* d_kps, int n, float* d_desc, int first_octave, int* d_err, DescParams* d_params,
                          cudaStream_t st);
int launch_match(const float* d_q, int nq, const float* d_t, int nt, int norm, float* d_dist, int32_t* d_idx, cudaStream_t st);
// the driver's front and back ends (driver.cu): fixed-point bilinear resize of 8-bit images, RANSAC homography
size_t resize_tab_bytes(int drows, int dcols);
int launch_resize_linear_u8(const uint8_t* d_src, int srows, int scols, int cn, uint8_t* d_dst, int drows, int dcols, void* d_tab, cudaStream_t st);
size_t homography_work_bytes(int n, int n_hyp);
int run_homography_ransac(const float2* d_src, const float2* d_dst, const float2* h_src, const float2* h_dst, int n, float thresh, int n_hyp, uint32_t seed,
                          void* d_work, double* H9, uint8_t* mask, int* n_inliers, cudaStream_t st);
// tcgen05 L2 matcher: bf16 hi/lo operand tiles -> tensor-core shortlists -> exact fp64 re-rank (match_tc.cu)
size_t match_tc_scratch_bytes(int nq, int nt);
int launch_match_tc(const float* d_q, int nq, const float* d_t, int nt, void* d_scratch, float* d_dist, int32_t* d_idx, cudaStream_t st);

}  // namespace siftb200
