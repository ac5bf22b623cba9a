/*
 * oracle/ref_wrap.cpp -- extern "C" doorway into the UNMODIFIED reference (compiled from
 * /root/reference/src/sift.cpp next to this file into oracle/_ref/libsift_ref.so, see oracle/Makefile).
 *
 * TEST INFRASTRUCTURE ONLY.  Each wrapper builds cv::Mat headers (cvshim) over caller memory, calls the
 * reference entry point declared in /root/reference/include/sift.hpp, and copies results out in the
 * packed layout documented in oracle/oracle.h.  The reference prints timers with printf on every blur
 * (src/sift.cpp:70,80,88,151); ref_set_quiet(1) parks stdout on /dev/null around each call so bench.py
 * can keep its one-JSON-line contract.
 */
#include <fcntl.h>
#include <unistd.h>

#include "sift.hpp" /* the reference's own header, found via -I /root/reference/include */
#include "oracle.h"

static int g_quiet = 0;
namespace {
struct Quiet {
    int saved;
    Quiet() : saved(-1) {
        if (!g_quiet) return;
        fflush(stdout);
        saved = dup(1);
        int nul = open("/dev/null", O_WRONLY);
        dup2(nul, 1);
        close(nul);
    }
    ~Quiet() {
        if (saved < 0) return;
        fflush(stdout);
        dup2(saved, 1);
        close(saved);
    }
};
void dims(int rows, int cols, int o, int& r, int& c) { r = rows; c = cols; for (int k = 0; k < o; ++k) { r /= 2; c /= 2; } }
/* wrap packed levels as Mat headers (copying: the reference may reassign Mats) */
std::vector<Mat> unpack(const float* p, int rows, int cols, int nOctaves, int per) {
    std::vector<Mat> v(nOctaves * per);
    for (int o = 0; o < nOctaves; ++o) {
        int r, c; dims(rows, cols, o, r, c);
        for (int i = 0; i < per; ++i) { Mat m(r, c, CV_32FC1); memcpy(m.data, p, sizeof(float) * r * c); p += (size_t)r * c; v[o * per + i] = m; }
    }
    return v;
}
void pack(const std::vector<Mat>& v, float* p) {
    for (size_t k = 0; k < v.size(); ++k) { memcpy(p, v[k].data, sizeof(float) * v[k].total()); p += v[k].total(); }
}
int put_kps(const std::vector<KeyPoint>& kps, OracleKeypoint* out, int cap, int* n_out) {
    *n_out = (int)kps.size();
    static_assert(sizeof(KeyPoint) == sizeof(OracleKeypoint), "KeyPoint must be the 28-byte POD");
    for (int i = 0; i < (int)kps.size() && i < cap; ++i) memcpy(&out[i], &kps[i], sizeof(OracleKeypoint));
    return (int)kps.size() > cap ? ORACLE_ERR_CAPACITY : 0;
}
}  // namespace

extern "C" {
void ref_set_quiet(int q) { g_quiet = q; }

void ref_gaussian_blur(const float* src, int rows, int cols, double sigma, float* dst) {
    Quiet q;
    Mat s(rows, cols, CV_32FC1, (void*)src), d;
    Gaussian_Blur(s, d, sigma);
    memcpy(dst, d.data, sizeof(float) * d.total());
}
void ref_gaussian_blur_1d(const float* src, int rows, int cols, double sigma, float* dst) {
    Quiet q;
    Mat s(rows, cols, CV_32FC1, (void*)src), d;
    Gaussian_Blur_1D(s, d, sigma);
    memcpy(dst, d.data, sizeof(float) * d.total());
}
void ref_build_gaussian_pyramid(const float* image, int rows, int cols, int nOctaves, float* gpyr) {
    Quiet q;
    Mat img(rows, cols, CV_32FC1, (void*)image);
    std::vector<Mat> g;
    buildGaussianPyramid(img, g, nOctaves);
    pack(g, gpyr);
}
void ref_build_dog_pyramid(const float* gpyr, int rows, int cols, int nOctaves, float* dogpyr) {
    Quiet q;
    std::vector<Mat> g = unpack(gpyr, rows, cols, nOctaves, 5), d;
    buildDoGPyramid(g, d, nOctaves);
    pack(d, dogpyr);
}
int ref_find_scale_space_extrema(const float* gpyr, const float* dogpyr, int rows, int cols, int nOctaves, OracleKeypoint* kp_out, int cap, int* n_out) {
    Quiet q;
    std::vector<Mat> g = unpack(gpyr, rows, cols, nOctaves, 5), d = unpack(dogpyr, rows, cols, nOctaves, 4);
    std::vector<KeyPoint> kps;
    findScaleSpaceExtrema(g, d, kps, nOctaves);
    return put_kps(kps, kp_out, cap, n_out);
}
int ref_cal_descriptor(const float* gpyr, int rows, int cols, int nOctaves, const OracleKeypoint* kps, int n, float* desc, int firstOctave) {
    Quiet q;
    std::vector<Mat> g = unpack(gpyr, rows, cols, nOctaves, 5);
    std::vector<KeyPoint> v(n);
    if (n) memcpy((void*)v.data(), kps, sizeof(OracleKeypoint) * n);
    Mat D(n, 128, CV_32F);
    try { calDescriptor(g, v, D, firstOctave); } catch (const cv::Exception&) { return ORACLE_ERR_ASSERT; }
    if (n) memcpy(desc, D.data, sizeof(float) * 128 * n);
    return 0;
}
int ref_sift_ncl(const float* image, int rows, int cols, OracleKeypoint* kp_out, float* desc_out, int cap, int* n_out) {
    Quiet q;
    Mat img(rows, cols, CV_32FC1, (void*)image), D;
    std::vector<KeyPoint> kps;
    SIFT_NCL(img, kps, D);
    int rc = put_kps(kps, kp_out, cap, n_out);
    if (desc_out) memcpy(desc_out, D.data, sizeof(float) * 128 * std::min((int)kps.size(), cap));
    return rc;
}
int ref_omp_max_threads(void) { return omp_get_max_threads(); }
/* torchrun exports OMP_NUM_THREADS=1; the reference itself would use every core in calDescriptor (src/sift.cpp:738) */
void ref_set_threads(int n) { omp_set_num_threads(n < 1 ? 1 : n); }
}
