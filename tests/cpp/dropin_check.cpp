// tests/cpp/dropin_check.cpp -- drives the drop-in C++ API (include/sift.hpp over the GPU library) the way the
// reference's src/main.cpp does, and checks every entry point against the CPU oracle (TEST code: links liboracle.so).
// Usage: dropin_check   (needs a CUDA device).  Prints "DROPIN OK" and exits 0 on success.
#include <cmath>
#include <cstdio>
#include <vector>

#include "sift.hpp"
#include "oracle.h"

static Mat make_image(int rows, int cols, unsigned seed) {
    Mat m(rows, cols, DATATYPE);
    unsigned s = seed;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (s >> 8) * (1.0f / 16777216.0f); };
    for (int y = 0; y < rows; ++y) for (int x = 0; x < cols; ++x) m.at<float>(y, x) = 128.f + (rnd() - 0.5f) * 4.f;
    for (int b = 0; b < rows * cols / 350; ++b) {
        float cx = rnd() * cols, cy = rnd() * rows, sg = 1.2f + rnd() * 7.f, amp = (40.f + rnd() * 70.f) * (rnd() < 0.5f ? -1.f : 1.f);
        int rad = (int)std::ceil(3 * sg);
        for (int y = std::max(0, (int)cy - rad); y < std::min(rows, (int)cy + rad + 1); ++y)
            for (int x = std::max(0, (int)cx - rad); x < std::min(cols, (int)cx + rad + 1); ++x)
                m.at<float>(y, x) += amp * std::exp(-((x - cx) * (x - cx) + (y - cy) * (y - cy)) / (2 * sg * sg));
    }
    for (int y = 0; y < rows; ++y) for (int x = 0; x < cols; ++x) m.at<float>(y, x) = std::min(255.f, std::max(0.f, m.at<float>(y, x)));
    return m;
}

#define CHECK(c) do { if (!(c)) { printf("CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main() {
    const int rows = 240, cols = 320;
    Mat img = make_image(rows, cols, 7);
    // --- the call src/main.cpp makes (:23) ---
    std::vector<KeyPoint> kps;
    Mat desc;
    SIFT_NCL(img, kps, desc);
    std::vector<OracleKeypoint> okp(1 << 16);
    std::vector<float> odesc((size_t)(1 << 16) * 128);
    int on = 0;
    CHECK(oracle32_sift_ncl(img.ptr<float>(0), rows, cols, okp.data(), odesc.data(), 1 << 16, &on, nullptr, nullptr, nullptr) == 0);
    printf("SIFT_NCL: %zu keypoints (oracle %d), descriptors %d x %d\n", kps.size(), on, desc.rows, desc.cols);
    CHECK((int)kps.size() == on && on > 30 && desc.rows == on && desc.cols == 128);
    int close_rows = 0;
    for (int i = 0; i < on; ++i) {
        CHECK(kps[i].octave == okp[i].octave && kps[i].class_id == -1);
        CHECK(std::fabs(kps[i].pt.x - okp[i].x) <= 0.01f && std::fabs(kps[i].pt.y - okp[i].y) <= 0.01f);
        float da = std::fabs(kps[i].angle - okp[i].angle); da = std::min(da, 360.f - da);
        CHECK(da <= 1.0f);
        double d2 = 0;
        for (int k = 0; k < 128; ++k) { double e = desc.at<float>(i, k) - odesc[(size_t)i * 128 + k]; d2 += e * e; }
        close_rows += std::sqrt(d2) <= 1e-3;
    }
    CHECK(close_rows >= on * 85 / 100);
    // --- sub-modules, chained exactly like SIFT_NCL chains them (src/sift.cpp:67-86) ---
    std::vector<Mat> gpyr, dog;
    buildGaussianPyramid(img, gpyr, 5);
    buildDoGPyramid(gpyr, dog, 5);
    CHECK(gpyr.size() == 25 && dog.size() == 20 && gpyr[24].rows == rows / 16 && dog[19].cols == cols / 16);
    std::vector<KeyPoint> kps2;
    findScaleSpaceExtrema(gpyr, dog, kps2, 5);
    CHECK(kps2.size() == kps.size());
    Mat desc2((int)kps2.size(), 128, CV_32F);
    calDescriptor(gpyr, kps2, desc2, 0);
    for (size_t i = 0; i < kps2.size(); ++i) for (int k = 0; k < 128; ++k) CHECK(desc2.at<float>((int)i, k) == desc.at<float>((int)i, k));
    Mat b0, b1;
    Gaussian_Blur(img, b0, 1.6);
    Gaussian_Blur_1D(img, b1, 1.6);
    std::vector<float> ob((size_t)rows * cols);
    oracle32_gaussian_blur_1d(img.ptr<float>(0), rows, cols, 1.6, ob.data());
    for (int i = 0; i < rows * cols; ++i) CHECK(b1.ptr<float>(0)[i] == ob[i]);
    oracle32_gaussian_blur(img.ptr<float>(0), rows, cols, 1.6, ob.data());
    for (int i = 0; i < rows * cols; ++i) CHECK(std::fabs(b0.ptr<float>(0)[i] - ob[i]) <= 2e-3f);
    // --- error behaviour: the reference throws (cv::resize on an empty octave / CV_Assert) ---
    bool threw = false;
    try { Mat tiny(8, 8, DATATYPE); std::vector<KeyPoint> k; Mat d; for (int i = 0; i < 64; ++i) tiny.ptr<float>(0)[i] = 0; SIFT_NCL(tiny, k, d); } catch (const cv::Exception&) { threw = true; }
    CHECK(threw);
    threw = false;
    try { std::vector<KeyPoint> bad(1); bad[0].octave = 7 << 8; bad[0].pt.x = 5; bad[0].pt.y = 5; bad[0].size = 4; bad[0].angle = 0; Mat d(1, 128, CV_32F); calDescriptor(gpyr, bad, d, 0); }
    catch (const cv::Exception&) { threw = true; }
    CHECK(threw);
    threw = false;
    try { std::vector<KeyPoint> k; Mat d; SITF_BuildIn_OpenCV(img, k, d); } catch (const cv::Exception&) { threw = true; }
    CHECK(threw);
    printf("DROPIN OK\n");
    return 0;
}
