/*
 * oracle.h -- C interface of the CPU oracle (TEST INFRASTRUCTURE ONLY; see sift_oracle.c).
 * oracle32_* = fp32-faithful restatement of /root/reference/src/sift.cpp, oracle64_* = fp64 twin.
 *
 * Packed pyramid layout: levels concatenated in the reference's index order (gpyr: o*5+i,
 * dogpyr: o*4+i), each level dense rows_o x cols_o with rows_{o+1} = rows_o/2, cols_{o+1} = cols_o/2.
 */
#ifndef ORACLE_H_
#define ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

/* cv::KeyPoint as a 28-byte POD (include/sift.hpp uses cv::KeyPoint). */
typedef struct OracleKeypoint {
    float x, y, size, angle, response;
    int octave, class_id;
} OracleKeypoint;

enum { ORACLE_OK = 0, ORACLE_ERR_CAPACITY = 1, ORACLE_ERR_ASSERT = 2, ORACLE_ERR_ARG = 3 };

void oracle_set_threads(int n);
int oracle_get_threads(void);

#define ORACLE_DECL(P, REAL)                                                                                                  \
    void P##gaussian_blur_naive(const REAL *src, int rows, int cols, double sigma, REAL *dst);                                \
    void P##gaussian_blur(const REAL *src, int rows, int cols, double sigma, REAL *dst);                                      \
    void P##gaussian_blur_1d(const REAL *src, int rows, int cols, double sigma, REAL *dst);                                   \
    void P##build_gaussian_pyramid(const REAL *image, int rows, int cols, int nOctaves, REAL *gpyr);                          \
    void P##build_dog_pyramid(const REAL *gpyr, int rows, int cols, int nOctaves, REAL *dogpyr);                              \
    int P##find_scale_space_extrema(const REAL *gpyr, const REAL *dogpyr, int rows, int cols, int nOctaves,                   \
                                    OracleKeypoint *kp_out, int cap, int *n_out, int *cand, int cand_cap, int *n_cand,       \
                                    int *refd, int refd_cap, int *n_refd, REAL *hists);                                       \
    int P##cal_descriptor(const REAL *gpyr, int rows, int cols, const OracleKeypoint *kps, int nkp, float *desc,              \
                          float *prequant, int firstOctave);                                                                  \
    int P##sift_ncl(const float *image, int rows, int cols, OracleKeypoint *kp_out, float *desc_out, int cap, int *n_out,     \
                    float *gpyr_out, float *dog_out, float *prequant_out);

ORACLE_DECL(oracle32_, float)
ORACLE_DECL(oracle64_, double)

int oracle_match_knn2(const float *q, int nq, const float *t, int nt, int norm, double ratio, int *idx_out, float *dist_out,
                      unsigned char *good_out);

#ifdef __cplusplus
}
#endif
#endif
