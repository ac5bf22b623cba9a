/*
 * sift_b200.h -- C ABI of the B200-native SIFT detect+describe path (libsiftb200.so).
 *
 * This is the drop-in boundary for the hot path of canhld94/SIFT-GPU: each entry point below is what a
 * binding for the reference's include/sift.hpp interface calls (the C++ shim that keeps the reference's
 * own signatures is include/sift.hpp + sift-gpu_b200/host/sift_dropin.cpp; INTEGRATION.md shows the
 * makefile change).  Plain pointers and sizes only; no C++ or torch types.  Every function returns a
 * status code (the reference returns void and exit()s / throws; the shim maps codes back to that).
 *
 * Conventions
 *   - images: float32, 0..255, row-major (reference input contract: src/main.cpp:84-85, src/sift.cpp:111).
 *   - "packed pyramid": levels concatenated in the reference's vector index order -- gpyr[o*5+i]
 *     (src/sift.cpp:248,253,257), dogpyr[o*4+i] (:275) -- each level dense rows_o x cols_o with
 *     rows_{o+1} = rows_o/2, cols_{o+1} = cols_o/2 (:254).
 *   - keypoints: SiftKeypoint == cv::KeyPoint's 28-byte layout; output order is the reference's scan
 *     order (octave, layer, row, col, orientation-peak bin; src/sift.cpp:556-557,487,491,525).
 *   - descriptors: n x 128 float32, row i belongs to keypoint i (src/sift.cpp:83-85,751).
 *   - there is NO CPU fallback: without a CUDA device every call fails with SIFT_B200_ERR_CUDA.
 */
#ifndef SIFT_B200_H_
#define SIFT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cv::KeyPoint as a POD: {Point2f pt; float size; float angle; float response; int octave; int class_id}.
 * octave packs  o | layer<<8 | cvRound((xi+0.5)*255)<<16  exactly as src/sift.cpp:383. */
typedef struct SiftKeypoint {
    float x, y;
    float size;
    float angle;
    float response;
    int32_t octave;
    int32_t class_id;
} SiftKeypoint;

typedef struct SiftB200 SiftB200; /* opaque handle: device workspace for up to max_batch frames */

enum {
    SIFT_B200_OK = 0,
    SIFT_B200_ERR_CAPACITY = 1, /* more keypoints than `cap`: the first cap records (reference order) are written, *n_out = true count.
                                 * If an INTERNAL list overflowed (more than max(16384, 4*max_kp, pixels/64) raw extrema or more than max_kp
                                 * refined points in one frame) the true count is unknown: *n_out = max_kp_per_frame + 1, and only the records
                                 * the kernels produced are written -- host buffers beyond them are left untouched */
    SIFT_B200_ERR_ARG = 2,
    SIFT_B200_ERR_CUDA = 3,      /* no device / CUDA failure: sift_b200_last_error() has the text */
    SIFT_B200_ERR_TOO_SMALL = 4, /* an octave would be empty (reference: cv::resize throws, src/sift.cpp:254) */
    SIFT_B200_ERR_ASSERT = 5     /* reference CV_Assert would fire (src/sift.cpp:744) */
};

enum { SIFT_B200_NORM_L1 = 2, SIFT_B200_NORM_L2 = 4 }; /* cv::NormTypes values */

const char* sift_b200_last_error(void);
const char* sift_b200_version(void);

/* Workspace for frames up to max_rows x max_cols, max_batch frames per internal pass, at most
 * max_kp_per_frame keypoints per frame.  device = CUDA ordinal.
 * Threading: a handle serves one call at a time; different handles are independent (sift_b200_last_error is per thread).  A streaming
 * caller keeps a device busy with two host threads, one handle each, on alternate batches: one call's pipeline fill and drain then run
 * under the other's steady state (bench.py's e2e: 9.3 k frames/s against 8.95 k from one thread). */
int sift_b200_create(SiftB200** out, int max_rows, int max_cols, int max_batch, int max_kp_per_frame, int device);
int sift_b200_destroy(SiftB200* h);

/* ---- whole path ------------------------------------------------------------------------------- */

/* SIFT_NCL (include/sift.hpp:41-43, src/sift.cpp:59-91) for ONE host image, synchronous.
 * row_stride_bytes lets a non-continuous Mat through (the reference silently assumes continuous, :111).
 * kp_out[cap], desc_out[cap*128] are host buffers; *n_out receives the keypoint count. */
int sift_b200_detect_describe(SiftB200* h, const float* img, int rows, int cols, size_t row_stride_bytes,
                              SiftKeypoint* kp_out, float* desc_out, int cap, int* n_out);

/* Same path over n_frames device-resident frames (dense, frame stride rows*cols floats), asynchronous on
 * `stream` (a cudaStream_t passed as void*).  Outputs are device buffers: d_kp[n_frames*cap],
 * d_desc[n_frames*cap*128], d_counts[n_frames] (true counts, may exceed cap -> truncated).
 * This is the throughput entry point: the reference calls SIFT_NCL once per image (src/main.cpp:23-24);
 * frames are independent, so a batch is the same computation n_frames times. */
int sift_b200_detect_describe_batch_dev(SiftB200* h, const float* d_imgs, int n_frames, int rows, int cols,
                                        SiftKeypoint* d_kp, float* d_desc, int* d_counts, int cap, void* stream);

/* Host-buffer batch: pinned-host frames in, host keypoints/descriptors/counts out, copies inside the call
 * (synchronous).  Used for the end-to-end measurement. */
int sift_b200_detect_describe_batch_host(SiftB200* h, const float* imgs, int n_frames, int rows, int cols,
                                         SiftKeypoint* kp_out, float* desc_out, int* counts_out, int cap);

/* Same with uint8 gray frames on the host (what src/main.cpp:84 holds before convertTo): a quarter of the PCIe traffic. */
int sift_b200_detect_describe_batch_host_u8(SiftB200* h, const uint8_t* imgs, int n_frames, int rows, int cols,
                                            SiftKeypoint* kp_out, float* desc_out, int* counts_out, int cap);

/* u8 front end (src/main.cpp:84-85 semantics: gray u8 -> float32 without scaling), fused into the base
 * blur; device frames, otherwise identical to the batch_dev call. */
int sift_b200_detect_describe_batch_dev_u8(SiftB200* h, const uint8_t* d_imgs, int n_frames, int rows, int cols,
                                           SiftKeypoint* d_kp, float* d_desc, int* d_counts, int cap, void* stream);

/* The driver's colour front end (src/main.cpp:84): cvtColor(COLOR_RGB2GRAY) applied to the interleaved BGR bytes imread
 * returns -- channel 0 takes the "R" weight -- in cv2 4.13's 15-bit fixed point.  d_bgr [n_frames][rows][cols][3] u8 ->
 * d_gray [n_frames][rows][cols] u8 (feed it to ..._batch_dev_u8), asynchronous on `stream`. */
int sift_b200_rgb2gray_u8_dev(SiftB200* h, const uint8_t* d_bgr, int n_frames, int rows, int cols, uint8_t* d_gray, void* stream);

/* The rest of the driver's readImage (src/main.cpp:79-87) on host buffers, synchronous:
 *   resize_linear_u8: resize(img, img, Size(960,960)) (:83) = cv::resize(INTER_LINEAR) on 8-bit pixels with 1..4 interleaved channels,
 *                     OpenCV's fixed-point bilinear (11-bit coefficients), bit-identical to cv2.resize;
 *   rgb2gray_u8:      cvtColor(img, gray, COLOR_RGB2GRAY) (:84) on the BGR bytes imread returns (host form of ..._rgb2gray_u8_dev). */
int sift_b200_resize_linear_u8(SiftB200* h, const uint8_t* src, int rows, int cols, int channels, uint8_t* dst, int drows, int dcols);
int sift_b200_rgb2gray_u8(SiftB200* h, const uint8_t* bgr, int rows, int cols, uint8_t* gray);

/* 2x bilinear upsample front end (BASELINE config 3; the reference ignores doubleSize, src/sift.cpp:219-227, so this is an
 * extension with cv::resize(INTER_LINEAR) semantics: half-pixel centres, edge replicate).
 * Device: d_src [n_frames][rows][cols] -> d_dst [n_frames][2*rows][2*cols], asynchronous on `stream`; feed d_dst to
 * sift_b200_detect_describe_batch_dev.  Host: one image, upsample + SIFT_NCL; keypoint coordinates are in UPSAMPLED pixels. */
int sift_b200_upsample2x_dev(SiftB200* h, const float* d_src, int n_frames, int rows, int cols, float* d_dst, void* stream);
int sift_b200_detect_describe_up2(SiftB200* h, const float* img, int rows, int cols, SiftKeypoint* kp_out, float* desc_out, int cap,
                                  int* n_out, float* upsampled_out /* optional 2rows x 2cols host copy */);

/* Exact-pyramid mode (off by default; env SIFT_B200_EXACT_PYRAMID=1 turns it on at create).  The default pyramid is the separable
 * form of Gaussian_Blur, equal to the reference's non-separable loop (src/sift.cpp:110-153) up to float rounding.  With on != 0 every
 * later call on this handle (whole path and build_gaussian_pyramid) replays that loop literally -- same 2-D taps, same row-major
 * summation order, separately rounded multiply and add -- so the Gaussian/DoG pyramids are bit-identical to the reference's.  About
 * 4x slower; meant for validation. */
int sift_b200_set_exact_pyramid(SiftB200* h, int on);

/* ---- sub-modules (include/sift.hpp:47-67), host buffers, synchronous ---------------------------- */

/* Gaussian_Blur (include/sift.hpp:47, src/sift.cpp:123-153): unnormalised truncated 2-D Gaussian,
 * radius floor(3*sigma), zero padding with source row rows-1 / col cols-1 read as zero. */
int sift_b200_gaussian_blur(SiftB200* h, const float* src, int rows, int cols, double sigma, float* dst);
/* Gaussian_Blur_1D (include/sift.hpp:49, src/sift.cpp:170-217): separable variant that drops tap +w. */
int sift_b200_gaussian_blur_1d(SiftB200* h, const float* src, int rows, int cols, double sigma, float* dst);
/* buildGaussianPyramid (include/sift.hpp:51-53, src/sift.cpp:229-263): gpyr = packed 5*n_octaves levels. */
int sift_b200_build_gaussian_pyramid(SiftB200* h, const float* img, int rows, int cols, int n_octaves, float* gpyr);
/* buildDoGPyramid (include/sift.hpp:55-57, src/sift.cpp:265-283): dogpyr = packed 4*n_octaves levels. */
int sift_b200_build_dog_pyramid(SiftB200* h, const float* gpyr, int rows, int cols, int n_octaves, float* dogpyr);
/* findScaleSpaceExtrema (include/sift.hpp:59-62, src/sift.cpp:547-577). */
int sift_b200_find_scale_space_extrema(SiftB200* h, const float* gpyr, const float* dogpyr, int rows, int cols,
                                       int n_octaves, SiftKeypoint* kp_out, int cap, int* n_out);
/* calDescriptor (include/sift.hpp:64-67, src/sift.cpp:733-753): desc = n x 128, pre-allocated by the caller. */
int sift_b200_cal_descriptor(SiftB200* h, const float* gpyr, int rows, int cols, int n_octaves,
                             const SiftKeypoint* kps, int n, float* desc, int first_octave);

/* ---- matcher (src/main.cpp:25-40) --------------------------------------------------------------- */
/* BFMatcher(norm).knnMatch(query, train, k=2) + ratio test m1.distance <= ratio*m2.distance.
 * idx_out / dist_out: nq x 2 (ascending; exact ties -> lowest train index; -1 / +inf if nt < 2);
 * good_out (optional): nq flags.  Host buffers, synchronous. */
int sift_b200_match_knn2(SiftB200* h, const float* query, int nq, const float* train, int nt, int norm, double ratio,
                         int32_t* idx_out, float* dist_out, uint8_t* good_out);
/* Same call with two extras.  tensor_cores != 0 (NORM_L2 only): distances come from split-bf16 tcgen05 MMAs (fp32 accumulation in
 * TMEM), a short list of candidates per query is re-ranked exactly in fp64.  A query whose short list is not provably complete
 * (three or more train rows within the ~3e-5 error of the split products of its second neighbour) is detected and matched against
 * every train row exactly, so indices, distances and tie order ALWAYS equal the exact kernel's.  kernel_ms (optional): CUDA-event
 * time of the kernels alone. */
int sift_b200_match_knn2_ex(SiftB200* h, const float* query, int nq, const float* train, int nt, int norm, double ratio,
                            int32_t* idx_out, float* dist_out, uint8_t* good_out, int tensor_cores, float* kernel_ms);

/* Device-resident form: d_query [nq x 128], d_train [nt x 128] float32 on the device, results d_idx / d_dist [nq x 2] on the device,
 * asynchronous on `stream`.  The tensor-core path keeps its scratch in the handle: calls on one handle must be ordered (same stream, or
 * synchronised by the caller).  The ratio test is left to the caller (two floats). */
int sift_b200_match_knn2_dev(SiftB200* h, const float* d_query, int nq, const float* d_train, int nt, int norm, int32_t* d_idx,
                             float* d_dist, int tensor_cores, void* stream);

/* ---- homography consumer (src/main.cpp:44-62) ------------------------------------------------------- */
/* findHomography(src, dst, RANSAC, ransac_thresh) over n correspondences given as interleaved (x, y) float pairs (the driver passes
 * the query / scene keypoint positions of the ratio-test survivors, :48-53).  max_iters hypotheses (<= 0: OpenCV's default 2000) are
 * evaluated in parallel on the device; the consensus set of the best one (ties: lowest index -- the call is deterministic) is refitted
 * in double: normalised DLT, then Levenberg-Marquardt on the reprojection error, as OpenCV does after its RANSAC loop.  Outputs:
 * H9_out row-major 3x3 with h33 = 1, mask_out[n] (optional) = consensus set, *n_inliers_out.  ERR_TOO_SMALL when n < 4 or no
 * non-degenerate sample exists (cv::findHomography returns an empty Mat).  OpenCV's random sequence is not reproduced: the parity
 * bar is "same consensus set, same refit" (tests), not bit identity. */
int sift_b200_find_homography(SiftB200* h, const float* src_xy, const float* dst_xy, int n, double ransac_thresh, int max_iters,
                              double* H9_out, uint8_t* mask_out, int* n_inliers_out);

/* Chunk schedule of the host-batch entry points for n_frames frames on a handle created with max_batch (pure host logic, no
 * device needed): writes up to plan_cap chunk sizes, returns the number of chunks (or -1 on a bad argument).  Every chunk is in
 * [1, max_batch] and the chunks sum to n_frames; with taper != 0 the first and last chunks are short (max_batch/8, /4, /2). */
int sift_b200_chunk_plan(int n_frames, int max_batch, int taper, int* plan_out, int plan_cap);

/* ---- introspection for the benchmark ------------------------------------------------------------ */
/* Kernel launches issued by this handle since creation (bench.py reports the delta as gpu_launches). */
long long sift_b200_launch_count(const SiftB200* h);
/* CUDA-event time (ms) of the last batch_dev call, split by stage: [0] base blur, [1] octave blur+DoG, [2] gradient maps,
 * [3] extrema scan + refine, [4] orientation, [5] order+scan, [6] descriptors, [7] total.  Only filled when stage timing was
 * enabled with sift_b200_set_stage_timing(h, 1) (adds event records, no syncs). */
int sift_b200_set_stage_timing(SiftB200* h, int on);
int sift_b200_get_stage_ms(SiftB200* h, float* ms8);

#ifdef __cplusplus
}
#endif
#endif
