/*
 * oracle_prims.h -- restatement of the OpenCV 4.0.x primitives the reference's hot path calls.
 *
 * TEST INFRASTRUCTURE ONLY (used by the oracle and by third_party/cvshim, never by the product).
 *
 * OpenCV is a third-party dependency of the reference that is NOT under /root/reference: the makefile
 * links /usr/local/lib/libopencv_*.so.4.0 (makefile:28-29) and there is no lockfile, so the pin is
 * "OpenCV 4.0.x + opencv_contrib".  Each primitive below restates OpenCV's published scalar algorithm;
 * tests/test_oracle_prims.py checks them against the cv2 4.13 wheel where cv2 exposes the primitive
 * (fastAtan2 via cv2.phase, exp, magnitude, rounding via saturating convertTo, NEAREST resize, solve).
 *
 * Call sites in the reference: cvRound src/sift.cpp:340-342,383,431,521,582,588; cvFloor :643-645;
 * Matx33f::solve :326; hal::exp32f/fastAtan2/magnitude32f :424-426,:632-634; saturate_cast<uchar> :709.
 */
#ifndef ORACLE_PRIMS_H_
#define ORACLE_PRIMS_H_

#include <float.h>
#include <math.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cvRound: SSE2 cvtsd2si / cvtss2si under the default rounding mode = round half to even. */
static inline int oracle_cv_round(double v) { return (int)lrint(v); }
/* cvFloor(float): i = (int)v; return i - (i > v). */
static inline int oraclef_cv_floor(float v) { int i = (int)v; return i - (i > v); }
static inline int oracled_cv_floor(double v) { int i = (int)v; return i - (i > v); }
#define oraclef_cv_round(v) oracle_cv_round((double)(v))
#define oracled_cv_round(v) oracle_cv_round((double)(v))

/* saturate_cast<uchar>(float) = clamp(cvRound(v), 0, 255). */
static inline unsigned char oraclef_saturate_u8(float v) { int i = oracle_cv_round(v); return (unsigned char)(i < 0 ? 0 : i > 255 ? 255 : i); }
static inline unsigned char oracled_saturate_u8(double v) { int i = oracle_cv_round(v); return (unsigned char)(i < 0 ? 0 : i > 255 ? 255 : i); }

/* hal::fastAtan2 (degrees): OpenCV's scalar atan_f32 -- 7th-order odd polynomial in min/max, float
 * arithmetic, + (float)DBL_EPSILON in the denominator; octant fix-ups 90-a, 180-a, 360-a. */
static inline float oraclef_fast_atan2(float y, float x) {
    const float s = (float)(180.0 / 3.1415926535897932384626433832795);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s, p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    float ax = fabsf(x), ay = fabsf(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}
/* fp64 twin: same polynomial (its 0.01 deg model error is part of the algorithm), double arithmetic. */
static inline double oracled_fast_atan2(double y, double x) {
    const double s = 180.0 / 3.1415926535897932384626433832795;
    const double p1 = 0.9997878412794807 * s, p3 = -0.3258083974640975 * s, p5 = 0.1555786518463281 * s, p7 = -0.04432655554792128 * s;
    double ax = fabs(x), ay = fabs(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + DBL_EPSILON);
        c2 = c * c;
        a = 90. - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180. - a;
    if (y < 0) a = 360. - a;
    return a;
}

/* hal::exp32f is a table+polynomial expf accurate to a few ulp; hal::magnitude32f is sqrt(x*x+y*y). */
static inline float oraclef_exp(float v) { return expf(v); }
static inline double oracled_exp(double v) { return exp(v); }
static inline float oraclef_magnitude(float x, float y) { return sqrtf(x * x + y * y); }
static inline double oracled_magnitude(double x, double y) { return sqrt(x * x + y * y); }
static inline float oraclef_sqrt(float v) { return sqrtf(v); }
static inline double oracled_sqrt(double v) { return sqrt(v); }
static inline float oraclef_cos(float v) { return cosf(v); }
static inline double oracled_cos(double v) { return cos(v); }
static inline float oraclef_sin(float v) { return sinf(v); }
static inline double oracled_sin(double v) { return sin(v); }
static inline float oraclef_pow2(float v) { return powf(2.f, v); }
static inline double oracled_pow2(double v) { return pow(2., v); }

/* Matx33f::solve(b, DECOMP_LU): OpenCV's Matx_FastSolveOp<_Tp,3,3,1> = Cramer's rule in _Tp
 * arithmetic with d = 1/det; returns zeros when det == 0.  A row-major 3x3. */
#define ORACLE_SOLVE3(NAME, T)                                                                                              \
    static inline void NAME(const T *a, const T *b, T *x) {                                                                 \
        T d = a[0] * (a[4] * a[8] - a[7] * a[5]) - a[1] * (a[3] * a[8] - a[6] * a[5]) + a[2] * (a[3] * a[7] - a[6] * a[4]); \
        if (d == 0) { x[0] = x[1] = x[2] = 0; return; }                                                                     \
        d = 1 / d;                                                                                                          \
        x[0] = d * (b[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (b[1] * a[8] - a[5] * b[2]) + a[2] * (b[1] * a[7] - a[4] * b[2])); \
        x[1] = d * (a[0] * (b[1] * a[8] - a[5] * b[2]) - b[0] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * b[2] - b[1] * a[6])); \
        x[2] = d * (a[0] * (a[4] * b[2] - b[1] * a[7]) - a[1] * (a[3] * b[2] - b[1] * a[6]) + b[0] * (a[3] * a[7] - a[4] * a[6])); \
    }
ORACLE_SOLVE3(oraclef_solve3, float)
ORACLE_SOLVE3(oracled_solve3, double)

#ifdef __cplusplus
}
#endif
#endif
