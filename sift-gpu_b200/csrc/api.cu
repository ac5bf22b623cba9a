// api.cu -- the C ABI of include/sift_b200.h: handle, workspace layout, stage orchestration.
//
// One handle = one device workspace for up to max_batch frames of up to max_rows x max_cols.  The whole path
// (base blur -> 5 x octave blur+DoG -> gradient maps -> extrema+refine -> orientation -> order+scan -> descriptor prep + descriptors)
// is thirteen kernel launches per chunk of max_batch frames, all asynchronous on the caller's stream, no host sync and no
// CPU fallback anywhere: if CUDA is unavailable every entry point returns SIFT_B200_ERR_CUDA.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "sift_internal.cuh"

using namespace siftb200;

namespace {
thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CUDA_TRY(expr)                                                                                  \
    do {                                                                                                \
        cudaError_t e_ = (expr);                                                                        \
        if (e_ != cudaSuccess) return fail(SIFT_B200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)

int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }
int round_up(int v, int m) { return (v + m - 1) / m * m; }
}  // namespace

namespace siftb200 {
int num_sms() {
    static int cached[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        cached[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
    }
    return cached[dev];
}
}  // namespace siftb200

struct SiftB200 {
    int device = 0, max_rows = 0, max_cols = 0, max_batch = 0, cap_kp = 0;
    cudaStream_t stream = nullptr;
    float* ws = nullptr;        // fused-pipeline levels: 7 per octave (G0..G2, D0..D3) x max_batch frames
    size_t ws_floats = 0;
    float* ws_full = nullptr;   // stage-level API: 9 levels per octave, one frame (lazy)
    size_t ws_full_floats = 0;
    DetectBuf db{};
    // second lane: own workspace, detection buffers and stream, so that consecutive chunks of a batch overlap
    // (one chunk's latency-bound kernels fill the issue slots the other's leave idle).  Allocated on first use.
    float* ws2 = nullptr;
    DetectBuf db2{};
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join[2] = {};
    int lanes = 2;
    float* d_img = nullptr;     // staging for host entry points [max_batch][max_rows*max_cols]
    SiftKeypoint* d_kp = nullptr;
    float* d_desc = nullptr;
    int* d_counts = nullptr;
    int* h_counts = nullptr;    // pinned
    // second staging set + copy streams for the pipelined host-batch entry point
    float* d_img2 = nullptr;
    float* d_img3 = nullptr;  // third input buffer: lets the H2D copy run one chunk ahead of the two compute lanes
    void* match_scratch = nullptr;   // operand tiles + shortlists of the tensor-core matcher, grown on demand
    size_t match_scratch_bytes = 0;
    bool exact_pyramid = false;  // replay the reference's non-separable blur loop bit for bit (pyramid_exact.cu)
    bool taper = true;        // host batch: short chunks at both ends of a call (env SIFT_B200_TAPER=0 disables)
    SiftKeypoint* d_kp2 = nullptr;
    float* d_desc2 = nullptr;
    int* d_counts2 = nullptr;
    int* h_counts2 = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[3] = {}, ev_comp[3] = {}, ev_cnt[2] = {}, ev_out[2] = {};
    long long launches = 0;
    bool stage_timing = false;
    cudaEvent_t ev[9] = {};
    bool ev_valid = false;
};

namespace {

size_t frame_floats(int rows, int cols, int n_oct, int levels) {
    size_t n = 0;
    for (int o = 0; o < n_oct; ++o) {
        n += (size_t)round_up(cols, 32) * rows * levels;
        rows /= 2; cols /= 2;
    }
    return n;
}

// Lay the levels of every octave out in `base`: [level][frame][rows_o][pitch_o].
int make_view(float* base, int rows, int cols, int n_oct, int n_frames, bool full, PyrView* pv) {
    if (n_oct < 1 || n_oct > kMaxOctaves) return SIFT_B200_ERR_ARG;
    memset(pv, 0, sizeof(*pv));
    pv->n_oct = n_oct;
    int tiles = 0, gtiles = 0;
    float* p = base;
    for (int o = 0; o < n_oct; ++o) {
        if (rows < 1 || cols < 1) return SIFT_B200_ERR_TOO_SMALL;
        OctaveView& v = pv->oct[o];
        v.rows = rows; v.cols = cols; v.pitch = round_up(cols, 32);
        v.frame_stride = (size_t)v.pitch * rows;
        const size_t lvl = v.frame_stride * n_frames;
        for (int i = 0; i < kNumScales; ++i) {
            if (i < 3 || full) { v.G[i] = p; p += lvl; } else v.G[i] = nullptr;
        }
        for (int i = 0; i < kNumScales - 1; ++i) { v.D[i] = p; p += lvl; }
        for (int i = 0; i < kNumScales; ++i) {
            if (i == 1 || i == 2 || full) { v.MO[i] = reinterpret_cast<float2*>(p); p += 2 * lvl; } else v.MO[i] = nullptr;
        }
        // gradient strips: kGradCols aligned columns x kGradRows rows over rows [1, rows-1)  (detect.cu)
        v.grad_tiles_x = cols > 2 ? (cols + kGradCols - 1) / kGradCols : 0;
        v.grad_tile_base = gtiles;
        gtiles += (cols > 2 && rows > 2) ? v.grad_tiles_x * ((rows - 2 + kGradRows - 1) / kGradRows) : 0;
        // extrema strips: 30 x 16 outputs over the interior [5, rows-5) x [5, cols-5)  (detect.cu)
        const int in_c = cols - 2 * kImgBorder, in_r = rows - 2 * kImgBorder;
        v.tiles_x = in_c > 0 ? (in_c + kExtremaCols - 1) / kExtremaCols : 0;
        v.tile_base = tiles;
        tiles += (in_c > 0 && in_r > 0) ? v.tiles_x * ((in_r + kExtremaRows - 1) / kExtremaRows) : 0;
        rows /= 2; cols /= 2;
    }
    pv->total_tiles = tiles;
    pv->total_grad_tiles = gtiles;
    return SIFT_B200_OK;
}

// 1-D factor of the reference's 2-D tap K[i][j]/8192 = exp(-(i^2+j^2)/den)/(2 PI s^2), den = float(2*s*s)
// (src/sift.cpp:95-108): exp(-i^2/den)/sqrt(2 PI s^2), computed in double and rounded to float once.
int make_taps(float sigma, float* taps /* >= 2*radius+1 */, int max_radius) {
    const float t3 = 3 * sigma;
    const int w = (int)floor((double)t3);
    if (w > max_radius) return -1;
    const float den_f = 2 * sigma * sigma;
    const double PI = 3.14159265359;
    const double norm = sqrt(1. / (2 * PI * sigma * sigma));
    for (int i = -w; i <= w; ++i) taps[i + w] = (float)(norm * exp(-(double)(i * i) / (double)den_f));
    return w;
}

// The reference's 2-D taps (src/sift.cpp:95-108): den = 2*sigma*sigma in FLOAT, the rest in double, x8192, rounded to float once.
void make_taps_2d(float sigma, float* k /* (2w+1)^2, row-major */) {
    const float t3 = 3 * sigma;
    const int w = (int)floor((double)t3);
    const int size = 2 * w + 1;
    const float den_f = 2 * sigma * sigma;
    const double PI = 3.14159265359;
    for (int i = -w; i <= w; ++i)
        for (int j = -w; j <= w; ++j) {
            double dat = 1. / (2 * PI * sigma * sigma) * exp(-(i * i + j * j) * 1. / (double)den_f);
            dat = dat * 8192;
            k[(i + w) * size + (j + w)] = (float)dat;
        }
}

void pipeline_sigmas(float sig[5]) {
    const double Sigma = 1.6, k = pow(2.0, 1.0 / kOctaveLayers);
    sig[0] = (float)sqrt(Sigma * Sigma + 0.2 * 0.2);  // base (src/sift.cpp:237)
    for (int i = 1; i < kNumScales; ++i) {
        const double st = pow(k * 1.0, (double)i) * Sigma;
        sig[i] = (float)sqrt(st * st - Sigma * Sigma);  // src/sift.cpp:240-245
    }
}

int ensure_full(SiftB200* h, int rows, int cols, int n_oct) {
    const size_t need = frame_floats(rows, cols, n_oct, 19);  // 5 G + 4 D + 5 float2 gradient maps
    if (need > h->ws_full_floats) {
        if (h->ws_full) cudaFree(h->ws_full);
        h->ws_full = nullptr; h->ws_full_floats = 0;
        CUDA_TRY(cudaMalloc((void**)&h->ws_full, need * sizeof(float)));
        h->ws_full_floats = need;
    }
    return SIFT_B200_OK;
}

int check_dims(const SiftB200* h, int rows, int cols) {
    if (!h) return fail(SIFT_B200_ERR_ARG, "null handle");
    if (rows < 1 || cols < 1 || rows > h->max_rows || cols > h->max_cols || rows >= 8192 || cols >= 8192)
        return fail(SIFT_B200_ERR_ARG, "image size outside the handle's max_rows x max_cols (or >= 8192)");
    return SIFT_B200_OK;
}

// upload / download one packed level set (dense rows) <-> pitched workspace levels of frame 0
int copy_levels(const PyrView& pv, bool gauss, int per, float* packed, bool to_device, cudaStream_t st) {
    for (int o = 0; o < pv.n_oct; ++o) {
        const OctaveView& v = pv.oct[o];
        for (int i = 0; i < per; ++i) {
            float* lv = gauss ? v.G[i] : v.D[i];
            if (to_device) CUDA_TRY(cudaMemcpy2DAsync(lv, (size_t)v.pitch * 4, packed, (size_t)v.cols * 4, (size_t)v.cols * 4, v.rows, cudaMemcpyHostToDevice, st));
            else CUDA_TRY(cudaMemcpy2DAsync(packed, (size_t)v.cols * 4, lv, (size_t)v.pitch * 4, (size_t)v.cols * 4, v.rows, cudaMemcpyDeviceToHost, st));
            packed += (size_t)v.rows * v.cols;
        }
    }
    return SIFT_B200_OK;
}

int alloc_detectbuf(DetectBuf& db, size_t F, int cap_r, size_t max_pixels) {
    db.cap_r = cap_r;
    db.cap_r_pow2 = next_pow2(db.cap_r);
    const size_t C = db.cap_r;
    // extrema before refinement: 20-85 % survive, and a frame of faint texture has many more candidates than keypoints:
    // the list also scales with the image area (one candidate per 64 pixels, twenty times the densest frame seen)
    db.cap_c = 4 * db.cap_r < 16384 ? 16384 : 4 * db.cap_r;
    if ((size_t)db.cap_c < max_pixels / 64) db.cap_c = (int)(max_pixels / 64);
    CUDA_TRY(cudaMalloc((void**)&db.cand, F * (size_t)db.cap_c * sizeof(uint32_t)));
    CUDA_TRY(cudaMalloc((void**)&db.n_cand, F * sizeof(int)));
    CUDA_TRY(cudaMalloc((void**)&db.refined, F * C * sizeof(Refined)));
    CUDA_TRY(cudaMalloc((void**)&db.n_refined, F * sizeof(int)));
    CUDA_TRY(cudaMalloc((void**)&db.angles, F * C * kMaxPeaks * sizeof(float)));
    CUDA_TRY(cudaMalloc((void**)&db.n_peaks, F * C * sizeof(int)));
    CUDA_TRY(cudaMalloc((void**)&db.order, F * C * sizeof(int)));
    CUDA_TRY(cudaMalloc((void**)&db.kp_offset, F * C * sizeof(int)));
    CUDA_TRY(cudaMalloc((void**)&db.sort_tmp, F * (size_t)db.cap_r_pow2 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMalloc((void**)&db.dparams, F * C * sizeof(DescParams)));
    CUDA_TRY(cudaMalloc((void**)&db.n_kp, F * sizeof(int)));
    return SIFT_B200_OK;
}

void free_detectbuf(DetectBuf& db) {
    cudaFree(db.cand); cudaFree(db.n_cand); cudaFree(db.refined); cudaFree(db.n_refined); cudaFree(db.angles); cudaFree(db.n_peaks);
    cudaFree(db.order); cudaFree(db.kp_offset); cudaFree(db.sort_tmp); cudaFree(db.dparams); cudaFree(db.n_kp);
    db = DetectBuf{};
}

int ensure_lane2(SiftB200* h) {
    if (h->ws2) return SIFT_B200_OK;
    CUDA_TRY(cudaMalloc((void**)&h->ws2, h->ws_floats * sizeof(float)));
    int rc = alloc_detectbuf(h->db2, h->max_batch, h->cap_kp, (size_t)h->max_rows * h->max_cols);
    if (rc) return rc;
    CUDA_TRY(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    for (auto& e : h->ev_join) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return SIFT_B200_OK;
}

// The whole path for ONE chunk (nf <= max_batch frames) on stream st, in lane `lane`'s workspace: 12 launches.
int enqueue_chunk(SiftB200* h, int lane, const float* d_imgs, const uint8_t* d_imgs8, int nf, int rows, int cols, SiftKeypoint* d_kp, float* d_desc,
                  int* d_counts, int cap, cudaStream_t st, bool timing) {
    const int n_oct = 5;  // SIFT_NCL hard-codes 5 octaves (src/sift.cpp:67-68,78)
    PyrView pv;
    const int rc = make_view(lane ? h->ws2 : h->ws, rows, cols, n_oct, nf, false, &pv);
    if (rc) return fail(rc, "make_view");
    const DetectBuf& db = lane ? h->db2 : h->db;
    const size_t fs = (size_t)rows * cols;
    if (timing) cudaEventRecord(h->ev[0], st);
    h->launches += h->exact_pyramid ? launch_exact_base(d_imgs, fs, cols, d_imgs8, pv.oct[0], nf, st) : launch_base_blur(d_imgs, fs, cols, d_imgs8, pv.oct[0], nf, st);
    if (timing) cudaEventRecord(h->ev[1], st);
    for (int o = 0; o < n_oct; ++o) h->launches += h->exact_pyramid ? launch_exact_octave(pv, o, nf, false, st) : launch_octave(pv, o, nf, false, st);
    if (timing) cudaEventRecord(h->ev[2], st);
    h->launches += launch_gradient(pv, nf, st);
    if (timing) cudaEventRecord(h->ev[3], st);
    h->launches += launch_extrema(pv, db, nf, st);
    if (timing) cudaEventRecord(h->ev[4], st);
    h->launches += launch_orientation(pv, db, nf, st);
    if (timing) cudaEventRecord(h->ev[5], st);
    h->launches += launch_order_scan(db, nf, d_counts, st);
    if (timing) cudaEventRecord(h->ev[6], st);
    h->launches += launch_describe(pv, db, nf, d_kp, d_desc, cap, st);
    if (timing) { cudaEventRecord(h->ev[7], st); h->ev_valid = true; }
    return SIFT_B200_OK;
}

int check_run(SiftB200* h, int n_frames, int rows, int cols, int cap) {
    int rc = check_dims(h, rows, cols);
    if (rc) return rc;
    if (n_frames < 0 || cap < 1 || cap > h->cap_kp) return fail(SIFT_B200_ERR_ARG, "cap must be in [1, max_kp_per_frame]");
    if ((rows >> 4) < 1 || (cols >> 4) < 1) return fail(SIFT_B200_ERR_TOO_SMALL, "image smaller than 16 px: octave 4 would be empty (reference throws in cv::resize)");
    CUDA_TRY(cudaSetDevice(h->device));
    return SIFT_B200_OK;
}

// n_frames in chunks of max_batch, asynchronous with respect to the host, ordered after / before the caller's stream st.
// With two or more chunks the chunks alternate between two lanes (own workspace + stream, forked from and joined back to st).
int run_pipeline(SiftB200* h, const float* d_imgs, const uint8_t* d_imgs8, int n_frames, int rows, int cols, SiftKeypoint* d_kp, float* d_desc,
                 int* d_counts, int cap, cudaStream_t st) {
    int rc = check_run(h, n_frames, rows, cols, cap);
    if (rc) return rc;
    const size_t fs = (size_t)rows * cols;
    const int n_chunks = (n_frames + h->max_batch - 1) / h->max_batch;
    const bool two = n_chunks >= 2 && h->lanes >= 2 && !h->stage_timing;
    if (two) {
        if ((rc = ensure_lane2(h))) return rc;
        CUDA_TRY(cudaEventRecord(h->ev_fork, st));
        CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_fork, 0));
        CUDA_TRY(cudaStreamWaitEvent(h->stream2, h->ev_fork, 0));
    }
    for (int k = 0; k < n_chunks; ++k) {
        const int f0 = k * h->max_batch;
        const int nf = n_frames - f0 < h->max_batch ? n_frames - f0 : h->max_batch;
        const int lane = two ? (k & 1) : 0;
        cudaStream_t cs = two ? (lane ? h->stream2 : h->stream) : st;
        rc = enqueue_chunk(h, lane, d_imgs ? d_imgs + f0 * fs : nullptr, d_imgs8 ? d_imgs8 + f0 * fs : nullptr, nf, rows, cols, d_kp + (size_t)f0 * cap,
                           d_desc + (size_t)f0 * cap * 128, d_counts + f0, cap, cs, h->stage_timing && k == n_chunks - 1);
        if (rc) return rc;
    }
    if (two) {
        CUDA_TRY(cudaEventRecord(h->ev_join[0], h->stream));
        CUDA_TRY(cudaEventRecord(h->ev_join[1], h->stream2));
        CUDA_TRY(cudaStreamWaitEvent(st, h->ev_join[0], 0));
        CUDA_TRY(cudaStreamWaitEvent(st, h->ev_join[1], 0));
    }
    CUDA_TRY(cudaGetLastError());
    return SIFT_B200_OK;
}

}  // namespace

namespace {
// frees on every exit path of the synchronous entry points below
struct DevMem {
    void* p = nullptr;
    ~DevMem() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    template <class T> T* as() const { return static_cast<T*>(p); }
};
struct Event {
    cudaEvent_t e = nullptr;
    ~Event() { if (e) cudaEventDestroy(e); }
};
}  // namespace

extern "C" {

static int create_impl(SiftB200* h);
int sift_b200_destroy(SiftB200* h);

const char* sift_b200_last_error(void) { return g_err.c_str(); }
const char* sift_b200_version(void) { return "sift_b200 0.1 (sm_100a)"; }

int sift_b200_create(SiftB200** out, int max_rows, int max_cols, int max_batch, int max_kp_per_frame, int device) {
    if (!out || max_rows < 16 || max_cols < 16 || max_rows >= 8192 || max_cols >= 8192 || max_batch < 1 || max_kp_per_frame < 1)
        return fail(SIFT_B200_ERR_ARG, "sift_b200_create: bad argument (16 <= rows, cols < 8192; batch, cap >= 1)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return fail(SIFT_B200_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(SIFT_B200_ERR_ARG, "bad device ordinal");
    CUDA_TRY(cudaSetDevice(device));
    SiftB200* h = new SiftB200();
    h->device = device; h->max_rows = max_rows; h->max_cols = max_cols; h->max_batch = max_batch; h->cap_kp = max_kp_per_frame;
    *out = nullptr;
    const int rc = create_impl(h);
    if (rc != SIFT_B200_OK) {  // a partially built handle is torn down here: the caller gets no handle and nothing leaks
        const std::string keep = g_err;
        sift_b200_destroy(h);
        g_err = keep;
        return rc;
    }
    *out = h;
    return SIFT_B200_OK;
}

static int create_impl(SiftB200* h) {
    const int max_rows = h->max_rows, max_cols = h->max_cols, max_batch = h->max_batch, max_kp_per_frame = h->cap_kp;
    CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->ws_floats = frame_floats(max_rows, max_cols, 5, 11) * max_batch;  // G0..G2, D0..D3, 2 x float2 gradient maps
    CUDA_TRY(cudaMalloc((void**)&h->ws, h->ws_floats * sizeof(float)));
    const size_t F = max_batch;
    if (int rc_db = alloc_detectbuf(h->db, F, max_kp_per_frame, (size_t)max_rows * max_cols)) return rc_db;
    if (const char* e = getenv("SIFT_B200_LANES")) h->lanes = atoi(e);
    if (const char* e = getenv("SIFT_B200_TAPER")) h->taper = atoi(e) != 0;
    if (const char* e = getenv("SIFT_B200_EXACT_PYRAMID")) h->exact_pyramid = atoi(e) != 0;
    CUDA_TRY(cudaMalloc((void**)&h->d_counts, F * sizeof(int)));
    CUDA_TRY(cudaMallocHost((void**)&h->h_counts, 2 * F * sizeof(int)));  // [F] reported counts | [F] records actually written
    for (auto& e : h->ev) CUDA_TRY(cudaEventCreate(&e));
    float sig[5];
    pipeline_sigmas(sig);
    float taps[5][kTapStride];
    memset(taps, 0, sizeof(taps));
    for (int s = 0; s < 5; ++s)
        if (make_taps(sig[s], taps[s], kMaxRadius) < 0) return fail(SIFT_B200_ERR_ARG, "tap radius");
    upload_taps(taps);
    {
        std::vector<float> k2d(k2d_total());
        for (int s = 0; s < 5; ++s) make_taps_2d(sig[s], k2d.data() + k2d_offset(s));
        upload_taps_2d(k2d.data());
    }
    init_pyramid_kernels();
    init_detect_kernels();
    init_describe_kernels();
    init_match_tc_kernels();
    CUDA_TRY(cudaGetLastError());
    return SIFT_B200_OK;
}

int sift_b200_destroy(SiftB200* h) {
    if (!h) return SIFT_B200_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    cudaFree(h->ws); cudaFree(h->ws_full); cudaFree(h->ws2);
    free_detectbuf(h->db);
    free_detectbuf(h->db2);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    for (auto& e : h->ev_join) if (e) cudaEventDestroy(e);
    cudaFree(h->d_img); cudaFree(h->d_kp); cudaFree(h->d_desc); cudaFree(h->d_counts);
    cudaFreeHost(h->h_counts);
    cudaFree(h->match_scratch);
    cudaFree(h->d_img2); cudaFree(h->d_img3); cudaFree(h->d_kp2); cudaFree(h->d_desc2); cudaFree(h->d_counts2);
    if (h->h_counts2) cudaFreeHost(h->h_counts2);
    for (int b = 0; b < 3; ++b) for (cudaEvent_t e : {h->ev_in[b], h->ev_comp[b]}) if (e) cudaEventDestroy(e);
    for (int b = 0; b < 2; ++b) for (cudaEvent_t e : {h->ev_cnt[b], h->ev_out[b]}) if (e) cudaEventDestroy(e);
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    for (auto& e : h->ev) if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return SIFT_B200_OK;
}

int sift_b200_detect_describe_batch_dev(SiftB200* h, const float* d_imgs, int n_frames, int rows, int cols, SiftKeypoint* d_kp, float* d_desc,
                                        int* d_counts, int cap, void* stream) {
    if (!d_imgs || !d_kp || !d_desc || !d_counts) return fail(SIFT_B200_ERR_ARG, "null buffer");
    return run_pipeline(h, d_imgs, nullptr, n_frames, rows, cols, d_kp, d_desc, d_counts, cap, (cudaStream_t)stream);
}

int sift_b200_detect_describe_batch_dev_u8(SiftB200* h, const uint8_t* d_imgs, int n_frames, int rows, int cols, SiftKeypoint* d_kp, float* d_desc,
                                           int* d_counts, int cap, void* stream) {
    if (!d_imgs || !d_kp || !d_desc || !d_counts) return fail(SIFT_B200_ERR_ARG, "null buffer");
    return run_pipeline(h, nullptr, d_imgs, n_frames, rows, cols, d_kp, d_desc, d_counts, cap, (cudaStream_t)stream);
}

static int ensure_staging(SiftB200* h) {
    if (h->d_img) return SIFT_B200_OK;
    const size_t F = h->max_batch;
    CUDA_TRY(cudaMalloc((void**)&h->d_img, F * (size_t)h->max_rows * h->max_cols * sizeof(float)));
    CUDA_TRY(cudaMalloc((void**)&h->d_kp, F * (size_t)h->cap_kp * sizeof(SiftKeypoint)));
    CUDA_TRY(cudaMalloc((void**)&h->d_desc, F * (size_t)h->cap_kp * 128 * sizeof(float)));
    return SIFT_B200_OK;
}

static int ensure_pipeline(SiftB200* h) {
    if (h->d_img2) return SIFT_B200_OK;
    const size_t F = h->max_batch;
    CUDA_TRY(cudaMalloc((void**)&h->d_img2, F * (size_t)h->max_rows * h->max_cols * sizeof(float)));
    CUDA_TRY(cudaMalloc((void**)&h->d_img3, F * (size_t)h->max_rows * h->max_cols * sizeof(float)));
    CUDA_TRY(cudaMalloc((void**)&h->d_kp2, F * (size_t)h->cap_kp * sizeof(SiftKeypoint)));
    CUDA_TRY(cudaMalloc((void**)&h->d_desc2, F * (size_t)h->cap_kp * 128 * sizeof(float)));
    CUDA_TRY(cudaMalloc((void**)&h->d_counts2, F * sizeof(int)));
    CUDA_TRY(cudaMallocHost((void**)&h->h_counts2, 2 * F * sizeof(int)));
    CUDA_TRY(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
    for (int b = 0; b < 3; ++b) {
        CUDA_TRY(cudaEventCreateWithFlags(&h->ev_in[b], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&h->ev_comp[b], cudaEventDisableTiming));
    }
    for (int b = 0; b < 2; ++b) {
        CUDA_TRY(cudaEventCreateWithFlags(&h->ev_cnt[b], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&h->ev_out[b], cudaEventDisableTiming));
    }
    return SIFT_B200_OK;
}

// Chunk schedule of one host-batch call.  The copy engines and the SMs only overlap in the steady state: the first chunk's H2D and
// the last chunk's kernels + D2H run alone.  Short chunks at both ends (max_batch/8, /4, /2) shrink those two bubbles; the middle
// runs full max_batch chunks, where the kernels are most efficient.
static std::vector<int> chunk_plan(int n_frames, int max_batch, bool taper) {
    std::vector<int> head, plan;
    int rem = n_frames;
    if (taper && n_frames >= 3 * max_batch) {
        // a head size is used at both ends; keep at least one full chunk in the middle (rem stays >= max_batch > 0)
        int used = 0;
        for (int sz = max_batch / 8 > 0 ? max_batch / 8 : 1; sz < max_batch && 2 * (used + sz) + max_batch <= n_frames; sz *= 2) {
            head.push_back(sz);
            used += sz;
        }
    }
    for (int sz : head) { plan.push_back(sz); rem -= 2 * sz; }
    if (rem % max_batch) { plan.push_back(rem % max_batch); rem -= rem % max_batch; }
    for (; rem > 0; rem -= max_batch) plan.push_back(max_batch);
    for (size_t i = head.size(); i-- > 0;) plan.push_back(head[i]);
    return plan;
}

// Host batch, software-pipelined over the chunks of chunk_plan(): the H2D of chunk k+1 (stream s_in, three input buffers) and the
// exact-size D2H of chunk k-1 (stream s_out) overlap the kernels of chunk k, which alternate between the two compute lanes.  The
// host only ever waits for the tiny counts copy of the PREVIOUS chunk, after the next chunk's copy and kernels have been queued.
static int batch_host_run(SiftB200* h, const void* imgs_v, int elem, int n_frames, int rows, int cols, SiftKeypoint* kp_out, float* desc_out,
                          int* counts_out, int cap);

// On a failure in the middle of a batch, copies into the caller's buffers may still be queued: drain every stream the pipeline uses
// before the error is returned, so nothing writes into caller memory after the call.
static int batch_host_impl(SiftB200* h, const void* imgs_v, int elem, int n_frames, int rows, int cols, SiftKeypoint* kp_out, float* desc_out,
                           int* counts_out, int cap) {
    const int rc = batch_host_run(h, imgs_v, elem, n_frames, rows, cols, kp_out, desc_out, counts_out, cap);
    if (rc != SIFT_B200_OK && rc != SIFT_B200_ERR_CAPACITY && h) {
        const std::string keep = g_err;
        for (cudaStream_t st : {h->s_in, h->s_out, h->stream, h->stream2})
            if (st) cudaStreamSynchronize(st);
        g_err = keep;
    }
    return rc;
}

static int batch_host_run(SiftB200* h, const void* imgs_v, int elem, int n_frames, int rows, int cols, SiftKeypoint* kp_out, float* desc_out,
                          int* counts_out, int cap) {
    const unsigned char* imgs = static_cast<const unsigned char*>(imgs_v);  // elem = 4 (float32 frames) or 1 (uint8 frames)
    int rc = check_dims(h, rows, cols);
    if (rc) return rc;
    if (!imgs || !kp_out || !desc_out || !counts_out) return fail(SIFT_B200_ERR_ARG, "null buffer");
    CUDA_TRY(cudaSetDevice(h->device));
    if ((rc = ensure_staging(h))) return rc;
    if ((rc = ensure_pipeline(h))) return rc;
    if ((rc = check_run(h, n_frames, rows, cols, cap))) return rc;
    const bool two = h->lanes >= 2 && !h->stage_timing;
    if (two && (rc = ensure_lane2(h))) return rc;
    cudaStream_t cst[2] = {h->stream, two ? h->stream2 : h->stream};  // chunk k computes in lane k&1
    const size_t fs = (size_t)rows * cols;
    float* d_img[3] = {h->d_img, h->d_img2, h->d_img3};
    SiftKeypoint* d_kp[2] = {h->d_kp, h->d_kp2};
    float* d_desc[2] = {h->d_desc, h->d_desc2};
    int* d_cnt[2] = {h->d_counts, h->d_counts2};
    int* h_cnt[2] = {h->h_counts, h->h_counts2};
    int status = SIFT_B200_OK;
    const std::vector<int> plan = chunk_plan(n_frames, h->max_batch, h->taper);
    const int n_chunks = (int)plan.size();
    std::vector<int> first(n_chunks + 1, 0);
    for (int k = 0; k < n_chunks; ++k) first[k + 1] = first[k] + plan[k];
    auto copy_in = [&](int k) -> int {
        const int ib = k % 3;
        if (k >= 3) CUDA_TRY(cudaStreamWaitEvent(h->s_in, h->ev_comp[ib], 0));  // chunk k-3 has consumed this input buffer
        CUDA_TRY(cudaMemcpyAsync(d_img[ib], imgs + (size_t)first[k] * fs * elem, plan[k] * fs * elem, cudaMemcpyHostToDevice, h->s_in));
        CUDA_TRY(cudaEventRecord(h->ev_in[ib], h->s_in));
        return SIFT_B200_OK;
    };
    auto flush = [&](int k) -> int {  // exact-size D2H of chunk k once its counts are on the host
        const int b = k & 1, f0 = first[k], nf = plan[k];
        CUDA_TRY(cudaEventSynchronize(h->ev_cnt[b]));
        for (int f = 0; f < nf; ++f) {
            int n = h_cnt[b][h->max_batch + f];  // records the kernels wrote (differs from the reported count when an internal list overflowed)
            counts_out[f0 + f] = h_cnt[b][f];
            if (h_cnt[b][f] > cap) status = SIFT_B200_ERR_CAPACITY;
            if (n > cap) n = cap;
            if (n > 0) {
                CUDA_TRY(cudaMemcpyAsync(kp_out + (size_t)(f0 + f) * cap, d_kp[b] + (size_t)f * cap, n * sizeof(SiftKeypoint), cudaMemcpyDeviceToHost, h->s_out));
                CUDA_TRY(cudaMemcpyAsync(desc_out + (size_t)(f0 + f) * cap * 128, d_desc[b] + (size_t)f * cap * 128, (size_t)n * 128 * sizeof(float),
                                         cudaMemcpyDeviceToHost, h->s_out));
            }
        }
        CUDA_TRY(cudaEventRecord(h->ev_out[b], h->s_out));
        return SIFT_B200_OK;
    };
    if (n_chunks > 0 && (rc = copy_in(0))) return rc;
    for (int k = 0; k < n_chunks; ++k) {
        const int b = k & 1, ib = k % 3, nf = plan[k];
        if (k + 1 < n_chunks && (rc = copy_in(k + 1))) return rc;  // queued before the host blocks in flush() below
        CUDA_TRY(cudaStreamWaitEvent(cst[b], h->ev_in[ib], 0));
        if (k >= 2) CUDA_TRY(cudaStreamWaitEvent(cst[b], h->ev_out[b], 0)); // chunk k-2's results have left this output buffer
        rc = enqueue_chunk(h, two ? b : 0, elem == 4 ? d_img[ib] : nullptr, elem == 1 ? reinterpret_cast<const uint8_t*>(d_img[ib]) : nullptr, nf, rows,
                           cols, d_kp[b], d_desc[b], d_cnt[b], cap, cst[b], false);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(h->ev_comp[ib], cst[b]));
        // results of chunk k-1 go out on s_out WHILE chunk k computes (queued before the wait on chunk k below)
        if (k >= 1 && (rc = flush(k - 1))) return rc;
        CUDA_TRY(cudaStreamWaitEvent(h->s_out, h->ev_comp[ib], 0));
        CUDA_TRY(cudaMemcpyAsync(h_cnt[b], d_cnt[b], nf * sizeof(int), cudaMemcpyDeviceToHost, h->s_out));
        CUDA_TRY(cudaMemcpyAsync(h_cnt[b] + h->max_batch, (two && b ? h->db2 : h->db).n_kp, nf * sizeof(int), cudaMemcpyDeviceToHost, h->s_out));
        CUDA_TRY(cudaEventRecord(h->ev_cnt[b], h->s_out));
    }
    if (n_chunks > 0 && (rc = flush(n_chunks - 1))) return rc;
    CUDA_TRY(cudaStreamSynchronize(h->s_out));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (two) CUDA_TRY(cudaStreamSynchronize(h->stream2));
    if (status) g_err = "keypoint capacity exceeded: outputs truncated";
    return status;
}

int sift_b200_detect_describe_batch_host(SiftB200* h, const float* imgs, int n_frames, int rows, int cols, SiftKeypoint* kp_out, float* desc_out,
                                         int* counts_out, int cap) {
    return batch_host_impl(h, imgs, 4, n_frames, rows, cols, kp_out, desc_out, counts_out, cap);
}

int sift_b200_detect_describe_batch_host_u8(SiftB200* h, const uint8_t* imgs, int n_frames, int rows, int cols, SiftKeypoint* kp_out, float* desc_out,
                                            int* counts_out, int cap) {
    return batch_host_impl(h, imgs, 1, n_frames, rows, cols, kp_out, desc_out, counts_out, cap);
}

// count reported to the caller (true count, or the overflow sentinel cap_r + 1) and number of records the kernels really wrote for frame 0
// of lane 0; synchronises the handle's stream
static int fetch_counts(SiftB200* h, int* reported, int* written) {
    CUDA_TRY(cudaMemcpyAsync(h->h_counts, h->d_counts, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->h_counts + 1, h->db.n_kp, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    *reported = h->h_counts[0];
    *written = h->h_counts[1];
    return SIFT_B200_OK;
}

int sift_b200_detect_describe(SiftB200* h, const float* img, int rows, int cols, size_t row_stride_bytes, SiftKeypoint* kp_out, float* desc_out, int cap,
                              int* n_out) {
    int rc = check_dims(h, rows, cols);
    if (rc) return rc;
    if (!img || !kp_out || !desc_out || !n_out) return fail(SIFT_B200_ERR_ARG, "null buffer");
    if (row_stride_bytes == 0) row_stride_bytes = (size_t)cols * 4;
    if (row_stride_bytes < (size_t)cols * 4) return fail(SIFT_B200_ERR_ARG, "row stride smaller than a row");
    CUDA_TRY(cudaSetDevice(h->device));
    if ((rc = ensure_staging(h))) return rc;
    CUDA_TRY(cudaMemcpy2DAsync(h->d_img, (size_t)cols * 4, img, row_stride_bytes, (size_t)cols * 4, rows, cudaMemcpyHostToDevice, h->stream));
    rc = run_pipeline(h, h->d_img, nullptr, 1, rows, cols, h->d_kp, h->d_desc, h->d_counts, cap, h->stream);
    if (rc) return rc;
    int n_rep = 0, n = 0;
    if ((rc = fetch_counts(h, &n_rep, &n))) return rc;
    *n_out = n_rep;
    int status = SIFT_B200_OK;
    if (n_rep > cap) status = fail(SIFT_B200_ERR_CAPACITY, "keypoint capacity exceeded: outputs truncated");
    if (n > cap) n = cap;
    if (n > 0) {
        CUDA_TRY(cudaMemcpyAsync(kp_out, h->d_kp, n * sizeof(SiftKeypoint), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaMemcpyAsync(desc_out, h->d_desc, (size_t)n * 128 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    return status;
}

int sift_b200_rgb2gray_u8_dev(SiftB200* h, const uint8_t* d_bgr, int n_frames, int rows, int cols, uint8_t* d_gray, void* stream) {
    if (!h || !d_bgr || !d_gray || rows < 1 || cols < 1 || n_frames < 0) return fail(SIFT_B200_ERR_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(h->device));
    h->launches += launch_rgb2gray_u8(d_bgr, d_gray, (size_t)n_frames * rows * cols, (cudaStream_t)stream);
    CUDA_TRY(cudaGetLastError());
    return SIFT_B200_OK;
}

int sift_b200_resize_linear_u8(SiftB200* h, const uint8_t* src, int rows, int cols, int channels, uint8_t* dst, int drows, int dcols) {
    if (!h || !src || !dst || rows < 1 || cols < 1 || drows < 1 || dcols < 1 || channels < 1 || channels > 4) return fail(SIFT_B200_ERR_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t sb = (size_t)rows * cols * channels, db = (size_t)drows * dcols * channels;
    DevMem m_src, m_dst, m_tab;
    CUDA_TRY(m_src.alloc(sb));
    CUDA_TRY(m_dst.alloc(db));
    CUDA_TRY(m_tab.alloc(resize_tab_bytes(drows, dcols)));
    CUDA_TRY(cudaMemcpyAsync(m_src.p, src, sb, cudaMemcpyHostToDevice, h->stream));
    h->launches += launch_resize_linear_u8(m_src.as<uint8_t>(), rows, cols, channels, m_dst.as<uint8_t>(), drows, dcols, m_tab.p, h->stream);
    CUDA_TRY(cudaMemcpyAsync(dst, m_dst.p, db, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaGetLastError());
    return SIFT_B200_OK;
}

int sift_b200_rgb2gray_u8(SiftB200* h, const uint8_t* bgr, int rows, int cols, uint8_t* gray) {
    if (!h || !bgr || !gray || rows < 1 || cols < 1) return fail(SIFT_B200_ERR_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t n = (size_t)rows * cols;
    DevMem m_src, m_dst;
    CUDA_TRY(m_src.alloc(n * 3));
    CUDA_TRY(m_dst.alloc(n));
    CUDA_TRY(cudaMemcpyAsync(m_src.p, bgr, n * 3, cudaMemcpyHostToDevice, h->stream));
    h->launches += launch_rgb2gray_u8(m_src.as<uint8_t>(), m_dst.as<uint8_t>(), n, h->stream);
    CUDA_TRY(cudaMemcpyAsync(gray, m_dst.p, n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaGetLastError());
    return SIFT_B200_OK;
}

int sift_b200_find_homography(SiftB200* h, const float* src_xy, const float* dst_xy, int n, double ransac_thresh, int max_iters, double* H9_out,
                              uint8_t* mask_out, int* n_inliers_out) {
    if (!h || n < 0 || (n > 0 && (!src_xy || !dst_xy)) || !H9_out || !n_inliers_out) return fail(SIFT_B200_ERR_ARG, "bad argument");
    *n_inliers_out = 0;
    for (int i = 0; i < 9; ++i) H9_out[i] = 0;
    if (n < 4) return fail(SIFT_B200_ERR_TOO_SMALL, "a homography needs at least 4 correspondences");
    if (ransac_thresh <= 0) ransac_thresh = 3.0;  // cv::findHomography's default ransacReprojThreshold
    const int n_hyp = max_iters > 0 ? (max_iters < 65536 ? max_iters : 65536) : 2000;  // cv::findHomography's default maxIters
    CUDA_TRY(cudaSetDevice(h->device));
    DevMem m_src, m_dst, m_work;
    CUDA_TRY(m_src.alloc((size_t)n * 8));
    CUDA_TRY(m_dst.alloc((size_t)n * 8));
    CUDA_TRY(m_work.alloc(homography_work_bytes(n, n_hyp)));
    CUDA_TRY(cudaMemcpyAsync(m_src.p, src_xy, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(m_dst.p, dst_xy, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
    std::vector<uint8_t> mask(n);
    const int nl = run_homography_ransac(m_src.as<float2>(), m_dst.as<float2>(), reinterpret_cast<const float2*>(src_xy), reinterpret_cast<const float2*>(dst_xy), n,
                                         (float)ransac_thresh, n_hyp, 0x5eed1234u, m_work.p, H9_out, mask.data(), n_inliers_out, h->stream);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaGetLastError());
    if (nl < 0) return fail(SIFT_B200_ERR_TOO_SMALL, "no homography: every sampled quadruple was degenerate");
    h->launches += nl;
    if (mask_out) memcpy(mask_out, mask.data(), n);
    return SIFT_B200_OK;
}

int sift_b200_upsample2x_dev(SiftB200* h, const float* d_src, int n_frames, int rows, int cols, float* d_dst, void* stream) {
    if (!h || !d_src || !d_dst || rows < 1 || cols < 1 || n_frames < 0) return fail(SIFT_B200_ERR_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(h->device));
    h->launches += launch_upsample2x(d_src, d_dst, rows, cols, n_frames, (cudaStream_t)stream);
    CUDA_TRY(cudaGetLastError());
    return SIFT_B200_OK;
}

int sift_b200_detect_describe_up2(SiftB200* h, const float* img, int rows, int cols, SiftKeypoint* kp_out, float* desc_out, int cap, int* n_out,
                                  float* upsampled_out) {
    int rc = check_dims(h, 2 * rows, 2 * cols);
    if (rc) return rc;
    if (!img || !kp_out || !desc_out || !n_out) return fail(SIFT_B200_ERR_ARG, "null buffer");
    CUDA_TRY(cudaSetDevice(h->device));
    if ((rc = ensure_staging(h))) return rc;
    DevMem m_src;  // freed on every exit path (after the stream has been synchronised, or on an error before any use)
    CUDA_TRY(m_src.alloc((size_t)rows * cols * 4));
    float* d_src = m_src.as<float>();
    CUDA_TRY(cudaMemcpyAsync(d_src, img, (size_t)rows * cols * 4, cudaMemcpyHostToDevice, h->stream));
    h->launches += launch_upsample2x(d_src, h->d_img, rows, cols, 1, h->stream);
    if (upsampled_out) CUDA_TRY(cudaMemcpyAsync(upsampled_out, h->d_img, (size_t)rows * cols * 16, cudaMemcpyDeviceToHost, h->stream));
    rc = run_pipeline(h, h->d_img, nullptr, 1, 2 * rows, 2 * cols, h->d_kp, h->d_desc, h->d_counts, cap, h->stream);
    if (rc) { cudaStreamSynchronize(h->stream); return rc; }
    int n_rep = 0, n = 0;
    if ((rc = fetch_counts(h, &n_rep, &n))) return rc;
    *n_out = n_rep;
    int status = SIFT_B200_OK;
    if (n_rep > cap) status = fail(SIFT_B200_ERR_CAPACITY, "keypoint capacity exceeded: outputs truncated");
    if (n > cap) n = cap;
    if (n > 0) {
        CUDA_TRY(cudaMemcpy(kp_out, h->d_kp, n * sizeof(SiftKeypoint), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(desc_out, h->d_desc, (size_t)n * 128 * sizeof(float), cudaMemcpyDeviceToHost));
    }
    return status;
}

// ---- sub-modules --------------------------------------------------------------------------------------------

static int blur_any(SiftB200* h, const float* src, int rows, int cols, double sigma_d, float* dst, bool one_d) {
    if (!h || !src || !dst || rows < 1 || cols < 1) return fail(SIFT_B200_ERR_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(h->device));
    std::vector<float> taps;
    int radius, hi;
    if (!one_d) {
        const float sigma = (float)sigma_d;  // Gaussian_Blur passes sigma through a float parameter (src/sift.cpp:95,128)
        const float t3 = 3 * sigma;
        radius = (int)floor((double)t3);
        if (radius < 0 || radius > 4096) return fail(SIFT_B200_ERR_ARG, "sigma out of range");
        taps.resize(2 * radius + 1);
        make_taps(sigma, taps.data(), radius);
        hi = radius;
    } else {  // getGaussianKernel1D keeps sigma in double; the tap loop is k in [-w, w) (src/sift.cpp:157-168,196)
        radius = (int)floor(3 * sigma_d);
        if (radius < 0 || radius > 4096) return fail(SIFT_B200_ERR_ARG, "sigma out of range");
        taps.resize(2 * radius + 1);
        const double PI = 3.14159265359;
        for (int i = -radius; i <= radius; ++i) taps[i + radius] = (float)(1. / sqrt(2 * PI * sigma_d * sigma_d) * exp(-((double)i * i) * 1. / (2 * sigma_d * sigma_d)));
        hi = radius - 1;
    }
    const size_t n = (size_t)rows * cols;
    DevMem m_src, m_dst, m_taps;  // freed on every exit path
    CUDA_TRY(m_src.alloc(n * 4));
    CUDA_TRY(m_dst.alloc(n * 4));
    CUDA_TRY(m_taps.alloc(taps.size() * 4));
    float *d_src = m_src.as<float>(), *d_dst = m_dst.as<float>(), *d_taps = m_taps.as<float>();
    CUDA_TRY(cudaMemcpyAsync(d_src, src, n * 4, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(d_taps, taps.data(), taps.size() * 4, cudaMemcpyHostToDevice, h->stream));
    const int nl = launch_generic_blur(d_src, d_dst, rows, cols, d_taps, radius, hi, h->stream);
    if (nl < 0) return fail(SIFT_B200_ERR_CUDA, "scratch allocation failed");
    h->launches += nl;
    CUDA_TRY(cudaMemcpyAsync(dst, d_dst, n * 4, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaGetLastError());
    return SIFT_B200_OK;
}

int sift_b200_gaussian_blur(SiftB200* h, const float* src, int rows, int cols, double sigma, float* dst) { return blur_any(h, src, rows, cols, sigma, dst, false); }
int sift_b200_gaussian_blur_1d(SiftB200* h, const float* src, int rows, int cols, double sigma, float* dst) { return blur_any(h, src, rows, cols, sigma, dst, true); }

static int check_stage(SiftB200* h, int rows, int cols, int n_oct) {
    int rc = check_dims(h, rows, cols);
    if (rc) return rc;
    if (n_oct < 1 || n_oct > kMaxOctaves) return fail(SIFT_B200_ERR_ARG, "n_octaves must be in [1, 8]");
    if ((rows >> (n_oct - 1)) < 1 || (cols >> (n_oct - 1)) < 1) return fail(SIFT_B200_ERR_TOO_SMALL, "an octave would be empty");
    CUDA_TRY(cudaSetDevice(h->device));
    return ensure_full(h, rows, cols, n_oct);
}

int sift_b200_build_gaussian_pyramid(SiftB200* h, const float* img, int rows, int cols, int n_octaves, float* gpyr) {
    int rc = check_stage(h, rows, cols, n_octaves);
    if (rc) return rc;
    if (!img || !gpyr) return fail(SIFT_B200_ERR_ARG, "null buffer");
    if ((rc = ensure_staging(h))) return rc;
    PyrView pv;
    if ((rc = make_view(h->ws_full, rows, cols, n_octaves, 1, true, &pv))) return fail(rc, "make_view");
    CUDA_TRY(cudaMemcpyAsync(h->d_img, img, (size_t)rows * cols * 4, cudaMemcpyHostToDevice, h->stream));
    if (h->exact_pyramid) {
        h->launches += launch_exact_base(h->d_img, (size_t)rows * cols, cols, nullptr, pv.oct[0], 1, h->stream);
        for (int o = 0; o < n_octaves; ++o) h->launches += launch_exact_octave(pv, o, 1, true, h->stream);
    } else {
        h->launches += launch_base_blur(h->d_img, (size_t)rows * cols, cols, nullptr, pv.oct[0], 1, h->stream);
        for (int o = 0; o < n_octaves; ++o) h->launches += launch_octave(pv, o, 1, true, h->stream);
    }
    if ((rc = copy_levels(pv, true, 5, gpyr, false, h->stream))) return rc;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaGetLastError());
    return SIFT_B200_OK;
}

int sift_b200_build_dog_pyramid(SiftB200* h, const float* gpyr, int rows, int cols, int n_octaves, float* dogpyr) {
    int rc = check_stage(h, rows, cols, n_octaves);
    if (rc) return rc;
    if (!gpyr || !dogpyr) return fail(SIFT_B200_ERR_ARG, "null buffer");
    PyrView pv;
    if ((rc = make_view(h->ws_full, rows, cols, n_octaves, 1, true, &pv))) return fail(rc, "make_view");
    if ((rc = copy_levels(pv, true, 5, const_cast<float*>(gpyr), true, h->stream))) return rc;
    h->launches += launch_dog(pv, 1, h->stream);
    if ((rc = copy_levels(pv, false, 4, dogpyr, false, h->stream))) return rc;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaGetLastError());
    return SIFT_B200_OK;
}

int sift_b200_find_scale_space_extrema(SiftB200* h, const float* gpyr, const float* dogpyr, int rows, int cols, int n_octaves, SiftKeypoint* kp_out, int cap,
                                       int* n_out) {
    int rc = check_stage(h, rows, cols, n_octaves);
    if (rc) return rc;
    if (!gpyr || !dogpyr || !kp_out || !n_out || cap < 1 || cap > h->cap_kp) return fail(SIFT_B200_ERR_ARG, "bad argument");
    if ((rc = ensure_staging(h))) return rc;
    PyrView pv;
    if ((rc = make_view(h->ws_full, rows, cols, n_octaves, 1, true, &pv))) return fail(rc, "make_view");
    if ((rc = copy_levels(pv, true, 5, const_cast<float*>(gpyr), true, h->stream))) return rc;
    if ((rc = copy_levels(pv, false, 4, const_cast<float*>(dogpyr), true, h->stream))) return rc;
    h->launches += launch_gradient(pv, 1, h->stream);
    h->launches += launch_extrema(pv, h->db, 1, h->stream);
    h->launches += launch_orientation(pv, h->db, 1, h->stream);
    h->launches += launch_order_scan(h->db, 1, h->d_counts, h->stream);
    // the descriptor kernel also emits the keypoint records; descriptors land in staging and are dropped
    h->launches += launch_describe(pv, h->db, 1, h->d_kp, h->d_desc, cap, h->stream);
    int n_rep = 0, n = 0;
    if ((rc = fetch_counts(h, &n_rep, &n))) return rc;
    *n_out = n_rep;
    int status = SIFT_B200_OK;
    if (n_rep > cap) status = fail(SIFT_B200_ERR_CAPACITY, "keypoint capacity exceeded: outputs truncated");
    if (n > cap) n = cap;
    if (n > 0) CUDA_TRY(cudaMemcpy(kp_out, h->d_kp, n * sizeof(SiftKeypoint), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaGetLastError());
    return status;
}

int sift_b200_cal_descriptor(SiftB200* h, const float* gpyr, int rows, int cols, int n_octaves, const SiftKeypoint* kps, int n, float* desc, int first_octave) {
    int rc = check_stage(h, rows, cols, n_octaves);
    if (rc) return rc;
    if (!gpyr || n < 0 || (n > 0 && (!kps || !desc))) return fail(SIFT_B200_ERR_ARG, "bad argument");
    if (n == 0) return SIFT_B200_OK;
    PyrView pv;
    if ((rc = make_view(h->ws_full, rows, cols, n_octaves, 1, true, &pv))) return fail(rc, "make_view");
    if ((rc = copy_levels(pv, true, 5, const_cast<float*>(gpyr), true, h->stream))) return rc;
    DevMem m_k, m_d, m_err, m_par;  // freed on every exit path
    CUDA_TRY(m_k.alloc((size_t)n * sizeof(SiftKeypoint)));
    CUDA_TRY(m_d.alloc((size_t)n * 128 * 4));
    CUDA_TRY(m_err.alloc(4));
    CUDA_TRY(m_par.alloc((size_t)n * sizeof(DescParams)));
    SiftKeypoint* d_k = m_k.as<SiftKeypoint>(); float* d_d = m_d.as<float>(); int* d_err = m_err.as<int>();
    CUDA_TRY(cudaMemsetAsync(d_err, 0, 4, h->stream));
    CUDA_TRY(cudaMemcpyAsync(d_k, kps, (size_t)n * sizeof(SiftKeypoint), cudaMemcpyHostToDevice, h->stream));
    h->launches += launch_gradient(pv, 1, h->stream);
    h->launches += launch_describe_given(pv, d_k, n, d_d, first_octave, d_err, m_par.as<DescParams>(), h->stream);
    int err = 0;
    CUDA_TRY(cudaMemcpyAsync(desc, d_d, (size_t)n * 128 * 4, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(&err, d_err, 4, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaGetLastError());
    if (err) return fail(SIFT_B200_ERR_ASSERT, "octave >= firstOctave && layer <= nOctaveLayers+2 (src/sift.cpp:744)");
    return SIFT_B200_OK;
}

int sift_b200_match_knn2_dev(SiftB200* h, const float* d_query, int nq, const float* d_train, int nt, int norm, int32_t* d_idx, float* d_dist,
                             int tensor_cores, void* stream) {
    if (!h || nq < 0 || nt < 0 || (nq > 0 && (!d_query || !d_idx || !d_dist)) || (nt > 0 && !d_train)) return fail(SIFT_B200_ERR_ARG, "bad argument");
    if (norm != SIFT_B200_NORM_L1 && norm != SIFT_B200_NORM_L2) return fail(SIFT_B200_ERR_ARG, "norm must be NORM_L1 (2) or NORM_L2 (4)");
    if (tensor_cores && norm != SIFT_B200_NORM_L2) return fail(SIFT_B200_ERR_ARG, "the tensor-core matcher computes NORM_L2 only");
    if (nq == 0) return SIFT_B200_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    // a train set shorter than the shortlist goes through the exact kernel (its -1 / +inf padding rule)
    if (tensor_cores && nt >= 4) {
        const size_t need = match_tc_scratch_bytes(nq, nt);
        if (need > h->match_scratch_bytes) {  // cudaFree synchronises the device: earlier matcher calls have finished with the old buffer
            if (h->match_scratch) cudaFree(h->match_scratch);
            h->match_scratch = nullptr; h->match_scratch_bytes = 0;
            CUDA_TRY(cudaMalloc(&h->match_scratch, need));
            h->match_scratch_bytes = need;
        }
        h->launches += launch_match_tc(d_query, nq, d_train, nt, h->match_scratch, d_dist, d_idx, st);
    } else {
        h->launches += launch_match(d_query, nq, d_train, nt, norm, d_dist, d_idx, st);
    }
    CUDA_TRY(cudaGetLastError());
    return SIFT_B200_OK;
}

int sift_b200_match_knn2_ex(SiftB200* h, const float* query, int nq, const float* train, int nt, int norm, double ratio, int32_t* idx_out,
                            float* dist_out, uint8_t* good_out, int tensor_cores, float* kernel_ms) {
    if (!h || nq < 0 || nt < 0 || (nq > 0 && (!query || !idx_out || !dist_out)) || (nt > 0 && !train)) return fail(SIFT_B200_ERR_ARG, "bad argument");
    if (norm != SIFT_B200_NORM_L1 && norm != SIFT_B200_NORM_L2) return fail(SIFT_B200_ERR_ARG, "norm must be NORM_L1 (2) or NORM_L2 (4)");
    if (tensor_cores && norm != SIFT_B200_NORM_L2) return fail(SIFT_B200_ERR_ARG, "the tensor-core matcher computes NORM_L2 only");
    if (kernel_ms) *kernel_ms = 0.f;
    if (nq == 0) return SIFT_B200_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    DevMem q, t, dist, idx;
    CUDA_TRY(q.alloc((size_t)nq * 512));
    CUDA_TRY(t.alloc((size_t)nt * 512));
    CUDA_TRY(dist.alloc((size_t)nq * 8));
    CUDA_TRY(idx.alloc((size_t)nq * 8));
    CUDA_TRY(cudaMemcpyAsync(q.p, query, (size_t)nq * 512, cudaMemcpyHostToDevice, h->stream));
    if (nt) CUDA_TRY(cudaMemcpyAsync(t.p, train, (size_t)nt * 512, cudaMemcpyHostToDevice, h->stream));
    Event e0, e1;
    if (kernel_ms) {
        CUDA_TRY(cudaEventCreate(&e0.e)); CUDA_TRY(cudaEventCreate(&e1.e));
        CUDA_TRY(cudaEventRecord(e0.e, h->stream));
    }
    const int rc = sift_b200_match_knn2_dev(h, q.as<float>(), nq, t.as<float>(), nt, norm, idx.as<int32_t>(), dist.as<float>(), tensor_cores, h->stream);
    if (rc) return rc;
    if (kernel_ms) CUDA_TRY(cudaEventRecord(e1.e, h->stream));
    CUDA_TRY(cudaMemcpyAsync(dist_out, dist.p, (size_t)nq * 8, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(idx_out, idx.p, (size_t)nq * 8, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (kernel_ms) CUDA_TRY(cudaEventElapsedTime(kernel_ms, e0.e, e1.e));
    CUDA_TRY(cudaGetLastError());
    if (good_out)  // ratio test exactly as written in the driver: float distance vs double product (src/main.cpp:38)
        for (int i = 0; i < nq; ++i) good_out[i] = (idx_out[2 * i + 1] >= 0 && dist_out[2 * i] <= ratio * dist_out[2 * i + 1]) ? 1 : 0;
    return SIFT_B200_OK;
}

int sift_b200_match_knn2(SiftB200* h, const float* query, int nq, const float* train, int nt, int norm, double ratio, int32_t* idx_out, float* dist_out,
                         uint8_t* good_out) {
    return sift_b200_match_knn2_ex(h, query, nq, train, nt, norm, ratio, idx_out, dist_out, good_out, 0, nullptr);
}

long long sift_b200_launch_count(const SiftB200* h) { return h ? h->launches : 0; }

int sift_b200_chunk_plan(int n_frames, int max_batch, int taper, int* plan_out, int plan_cap) {
    if (n_frames < 0 || max_batch < 1 || (plan_cap > 0 && !plan_out)) return -1;
    const std::vector<int> plan = chunk_plan(n_frames, max_batch, taper != 0);
    for (size_t k = 0; k < plan.size() && (int)k < plan_cap; ++k) plan_out[k] = plan[k];
    return (int)plan.size();
}

int sift_b200_set_exact_pyramid(SiftB200* h, int on) {
    if (!h) return fail(SIFT_B200_ERR_ARG, "null handle");
    h->exact_pyramid = on != 0;
    return SIFT_B200_OK;
}

int sift_b200_set_stage_timing(SiftB200* h, int on) {
    if (!h) return fail(SIFT_B200_ERR_ARG, "null handle");
    h->stage_timing = on != 0;
    h->ev_valid = false;
    return SIFT_B200_OK;
}

int sift_b200_get_stage_ms(SiftB200* h, float* ms8) {
    if (!h || !ms8 || !h->ev_valid) return fail(SIFT_B200_ERR_ARG, "no stage timing recorded");
    CUDA_TRY(cudaEventSynchronize(h->ev[7]));
    for (int i = 0; i < 7; ++i) CUDA_TRY(cudaEventElapsedTime(&ms8[i], h->ev[i], h->ev[i + 1]));
    CUDA_TRY(cudaEventElapsedTime(&ms8[7], h->ev[0], h->ev[7]));
    return SIFT_B200_OK;
}

}  // extern "C"
