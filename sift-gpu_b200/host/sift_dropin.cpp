// sift_dropin.cpp -- the reference's include/sift.hpp entry points implemented over the C ABI (include/sift_b200.h).
//
// Same names, argument meaning and error behaviour as reference src/sift.cpp: void returns; cv::Exception where the
// reference throws (cv::resize on an empty octave :254, CV_Assert :744); printf + exit(0) on mismatching pyramid
// levels (:276-279); the three stage-timer lines of SIFT_NCL (:70,80,88).  One process-wide device workspace is
// created lazily and grown on demand; the reference's caller is single-threaded (src/main.cpp), and so is this shim.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "sift.hpp"
#include "sift_b200.h"

namespace {

static_assert(sizeof(KeyPoint) == sizeof(SiftKeypoint), "cv::KeyPoint must be the 28-byte POD the C ABI writes");

struct Workspace {
    SiftB200* h = nullptr;
    int rows = 0, cols = 0, cap = 0;
    ~Workspace() { if (h) sift_b200_destroy(h); }
};
Workspace g_ws;

[[noreturn]] void raise(const char* what) { CV_Error(cv::Error::StsError, std::string(what) + ": " + sift_b200_last_error()); }

int device_ordinal() {
    const char* e = std::getenv("SIFT_B200_DEVICE");
    return e ? std::atoi(e) : 0;
}

SiftB200* workspace(int rows, int cols, int cap) {
    if (g_ws.h && rows <= g_ws.rows && cols <= g_ws.cols && cap <= g_ws.cap) return g_ws.h;
    const int nr = std::max(std::max(rows, g_ws.rows), 64), nc = std::max(std::max(cols, g_ws.cols), 64), ncap = std::max(std::max(cap, g_ws.cap), 1 << 15);
    if (g_ws.h) { sift_b200_destroy(g_ws.h); g_ws.h = nullptr; }
    if (sift_b200_create(&g_ws.h, nr, nc, 1, ncap, device_ordinal()) != SIFT_B200_OK) {
        g_ws.h = nullptr; g_ws.rows = g_ws.cols = g_ws.cap = 0;
        raise("sift_b200_create");
    }
    g_ws.rows = nr; g_ws.cols = nc; g_ws.cap = ncap;
    return g_ws.h;
}

Mat as_float(const Mat& m) {
    CV_Assert(m.type() == CV_32FC1);  // the reference reinterprets the buffer as data_t without checking (src/sift.cpp:111)
    return m;
}

void octave_dims(int rows, int cols, int o, int& r, int& c) { r = rows; c = cols; for (int k = 0; k < o; ++k) { r /= 2; c /= 2; } }

size_t packed_floats(int rows, int cols, int nOctaves, int per) {
    size_t n = 0;
    for (int o = 0; o < nOctaves; ++o) { int r, c; octave_dims(rows, cols, o, r, c); n += (size_t)r * c * per; }
    return n;
}

// std::vector<Mat> (reference index order) -> packed buffer; checks each level against the octave geometry
std::vector<float> pack_levels(const std::vector<Mat>& v, int nOctaves, int per, int& rows, int& cols) {
    CV_Assert((int)v.size() >= nOctaves * per && !v.empty());
    rows = v[0].rows; cols = v[0].cols;
    std::vector<float> out(packed_floats(rows, cols, nOctaves, per));
    float* p = out.data();
    for (int o = 0; o < nOctaves; ++o) {
        int r, c; octave_dims(rows, cols, o, r, c);
        for (int i = 0; i < per; ++i) {
            const Mat& m = v[o * per + i];
            CV_Assert(m.rows == r && m.cols == c && m.type() == CV_32FC1);
            for (int y = 0; y < r; ++y) std::memcpy(p + (size_t)y * c, m.ptr<float>(y), sizeof(float) * c);
            p += (size_t)r * c;
        }
    }
    return out;
}

void unpack_levels(const std::vector<float>& packed, int rows, int cols, int nOctaves, int per, std::vector<Mat>& v) {
    v.resize(nOctaves * per);
    const float* p = packed.data();
    for (int o = 0; o < nOctaves; ++o) {
        int r, c; octave_dims(rows, cols, o, r, c);
        for (int i = 0; i < per; ++i) {
            Mat m(r, c, CV_32FC1);
            for (int y = 0; y < r; ++y) std::memcpy(m.ptr<float>(y), p + (size_t)y * c, sizeof(float) * c);
            p += (size_t)r * c;
            v[o * per + i] = m;
        }
    }
}

}  // namespace

void SIFT_NCL(InputArray image, std::vector<KeyPoint>& keypoints, OutputArray descriptors) {
    Mat img = as_float(image.getMat());
    keypoints.clear();
    int cap = std::max(g_ws.cap, 1 << 15);
    for (;;) {
        SiftB200* h = workspace(img.rows, img.cols, cap);
        sift_b200_set_stage_timing(h, 1);
        std::vector<SiftKeypoint> kp(cap);
        std::vector<float> desc((size_t)cap * 128);
        int n = 0;
        const int rc = sift_b200_detect_describe(h, img.ptr<float>(0), img.rows, img.cols, (size_t)img.step1() * sizeof(float), kp.data(), desc.data(), cap, &n);
        if (rc == SIFT_B200_ERR_CAPACITY) { cap = n + n / 8 + 64; continue; }  // grow and redo: the reference has no cap
        if (rc != SIFT_B200_OK) raise("SIFT_NCL");
        float ms[8];
        if (sift_b200_get_stage_ms(h, ms) == SIFT_B200_OK) {  // the reference's three timer lines (src/sift.cpp:70,80,88)
            printf("pyramid construction time: %g\n", ms[0] + ms[1]);
            printf("keypoint localization time: %g\n", ms[2] + ms[3] + ms[4] + ms[5]);
            printf("descriptor extraction time: %g\n", ms[6]);
        }
        keypoints.resize(n);
        if (n) std::memcpy((void*)keypoints.data(), kp.data(), sizeof(SiftKeypoint) * n);
        descriptors.create(n, 128, CV_32F);
        Mat d = descriptors.getMat();
        for (int i = 0; i < n; ++i) std::memcpy(d.ptr<float>(i), desc.data() + (size_t)i * 128, 128 * sizeof(float));
        return;
    }
}

void SITF_BuildIn_OpenCV(InputArray image, std::vector<KeyPoint>& keypoints, OutputArray descriptors) {
#if defined(SIFT_B200_WITH_XFEATURES2D) && defined(SIFT_B200_HAVE_XFEATURES2D)
    Ptr<SIFT> detector = SIFT::create();
    Mat mask;
    detector->detectAndCompute(image, mask, keypoints, descriptors, false);
#else
    (void)image; (void)keypoints; (void)descriptors;
    CV_Error(cv::Error::StsError, "SITF_BuildIn_OpenCV: third-party CPU SIFT (opencv_contrib xfeatures2d) is not part of this library; "
                                  "rebuild with -DSIFT_B200_WITH_XFEATURES2D against opencv_contrib to forward to it");
#endif
}

static void blur_common(Mat& src, Mat& dst, double sigma, bool one_d) {
    Mat s = as_float(src);
    CV_Assert(s.isContinuous());
    Mat out(s.rows, s.cols, DATATYPE);
    SiftB200* h = workspace(s.rows, s.cols, 1);
    const int rc = one_d ? sift_b200_gaussian_blur_1d(h, s.ptr<float>(0), s.rows, s.cols, sigma, out.ptr<float>(0))
                         : sift_b200_gaussian_blur(h, s.ptr<float>(0), s.rows, s.cols, sigma, out.ptr<float>(0));
    if (rc != SIFT_B200_OK) raise(one_d ? "Gaussian_Blur_1D" : "Gaussian_Blur");
    dst = out;
}

void Gaussian_Blur(Mat& src, Mat& dst, double sigma) { blur_common(src, dst, sigma, false); }
void Gaussian_Blur_1D(Mat& src, Mat& dst, double sigma) { blur_common(src, dst, sigma, true); }

void buildGaussianPyramid(Mat& image, std::vector<Mat>& gpyr, int nOctaves) {
    Mat img = as_float(image);
    CV_Assert(img.isContinuous());
    SiftB200* h = workspace(img.rows, img.cols, 1);
    std::vector<float> packed(packed_floats(img.rows, img.cols, nOctaves, 5));
    const int rc = sift_b200_build_gaussian_pyramid(h, img.ptr<float>(0), img.rows, img.cols, nOctaves, packed.data());
    if (rc != SIFT_B200_OK) raise("buildGaussianPyramid");  // TOO_SMALL: the reference throws from cv::resize (:254)
    unpack_levels(packed, img.rows, img.cols, nOctaves, 5, gpyr);
}

void buildDoGPyramid(std::vector<Mat>& gpyr, std::vector<Mat>& dogpyr, int nOctaves) {
    for (int o = 0; o < nOctaves; ++o)
        for (int i = 0; i < 4; ++i)
            if (gpyr[o * 5 + i].size != gpyr[o * 5 + i + 1].size) {  // reference behaviour, src/sift.cpp:276-279
                printf("Different input size at o = %d and i = %d, abort!\n", o, i);
                exit(0);
            }
    int rows, cols;
    std::vector<float> g = pack_levels(gpyr, nOctaves, 5, rows, cols);
    std::vector<float> d(packed_floats(rows, cols, nOctaves, 4));
    SiftB200* h = workspace(rows, cols, 1);
    if (sift_b200_build_dog_pyramid(h, g.data(), rows, cols, nOctaves, d.data()) != SIFT_B200_OK) raise("buildDoGPyramid");
    unpack_levels(d, rows, cols, nOctaves, 4, dogpyr);
}

void findScaleSpaceExtrema(std::vector<Mat>& gpyr, std::vector<Mat>& dogpyr, std::vector<KeyPoint>& keypoints, int nOctaves) {
    keypoints.clear();
    int rows, cols, r2, c2;
    std::vector<float> g = pack_levels(gpyr, nOctaves, 5, rows, cols);
    std::vector<float> d = pack_levels(dogpyr, nOctaves, 4, r2, c2);
    CV_Assert(rows == r2 && cols == c2);
    int cap = std::max(g_ws.cap, 1 << 15);
    for (;;) {
        SiftB200* h = workspace(rows, cols, cap);
        std::vector<SiftKeypoint> kp(cap);
        int n = 0;
        const int rc = sift_b200_find_scale_space_extrema(h, g.data(), d.data(), rows, cols, nOctaves, kp.data(), cap, &n);
        if (rc == SIFT_B200_ERR_CAPACITY) { cap = n + n / 8 + 64; continue; }
        if (rc != SIFT_B200_OK) raise("findScaleSpaceExtrema");
        keypoints.resize(n);
        if (n) std::memcpy((void*)keypoints.data(), kp.data(), sizeof(SiftKeypoint) * n);
        return;
    }
}

void calDescriptor(std::vector<Mat>& gpyr, std::vector<KeyPoint>& keypoints, Mat& descriptors, int firstOctave) {
    const int n = (int)keypoints.size();
    if (n == 0) return;
    CV_Assert(descriptors.rows >= n && descriptors.cols == 128 && descriptors.type() == CV_32F);
    const int nOctaves = (int)gpyr.size() / 5;
    int rows, cols;
    std::vector<float> g = pack_levels(gpyr, nOctaves, 5, rows, cols);
    std::vector<float> out((size_t)n * 128);
    SiftB200* h = workspace(rows, cols, 1);
    const int rc = sift_b200_cal_descriptor(h, g.data(), rows, cols, nOctaves, reinterpret_cast<const SiftKeypoint*>(keypoints.data()), n, out.data(), firstOctave);
    if (rc != SIFT_B200_OK) raise("calDescriptor");  // ASSERT: octave >= firstOctave && layer <= nOctaveLayers+2 (:744)
    for (int i = 0; i < n; ++i) std::memcpy(descriptors.ptr<float>(i), out.data() + (size_t)i * 128, 128 * sizeof(float));
}
