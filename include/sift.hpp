// include/sift.hpp -- drop-in replacement for the reference's include/sift.hpp (canhld94/SIFT-GPU).
//
// Same eight entry points, same argument meaning, same output layout (reference include/sift.hpp:36-67), so the
// reference's src/main.cpp and makefile drive this library unchanged: main.cpp calls SIFT_NCL twice
// (src/main.cpp:23-24) and nothing else.  The bodies live in sift-gpu_b200/host/sift_dropin.cpp and forward to
// the C ABI of include/sift_b200.h (hand-written sm_100a kernels); there is no CPU implementation behind them.
//
// Includes what the reference header includes (reference include/sift.hpp:11-25), because its includers rely on it: the unchanged
// src/main.cpp gets imread, resize, cvtColor, BFMatcher, findHomography, imshow and <iostream> through this header alone.  The
// contrib module xfeatures2d is optional here (only SITF_BuildIn_OpenCV uses it): without it the namespace is declared empty so
// that the reference's `using namespace cv::xfeatures2d;` still compiles.  In the build container, where OpenCV C++ is absent,
// third_party/cvshim stands in for compile- and run-checks.
#ifndef SIFT_HPP_
#define SIFT_HPP_

#include <stdio.h>
#include <stdlib.h>

#include <iostream>
#include <vector>

#include <opencv2/core.hpp>
#include <opencv2/core/utility.hpp>
#include <opencv2/imgcodecs.hpp>
#include <opencv2/imgproc.hpp>
#include <opencv2/features2d.hpp>
#include <opencv2/highgui.hpp>
#include <opencv2/calib3d/calib3d.hpp>
#if defined(__has_include)
#if __has_include(<opencv2/core/types_c.h>)
#include <opencv2/core/types_c.h>  // cvPoint (src/main.cpp:59-60)
#endif
#if __has_include(<opencv2/xfeatures2d.hpp>)
#include <opencv2/xfeatures2d.hpp>
#define SIFT_B200_HAVE_XFEATURES2D 1
#endif
#endif
namespace cv { namespace xfeatures2d {} }

// The reference header injects both namespaces and these two names into every includer; main.cpp relies on it.
using namespace cv;
using namespace cv::xfeatures2d;

typedef float data_t;      // pixel type of every pyramid level (reference include/sift.hpp:31)
#define DATATYPE CV_32FC1  // (reference include/sift.hpp:33)

// --- whole pipeline -------------------------------------------------------------------------------------------

// Pyramid + DoG + extrema/refinement/orientation + descriptors on the GPU.  `image`: CV_32FC1, 0..255.
// `keypoints` is cleared and filled in the reference's scan order; `descriptors` is created as N x 128 CV_32F.
// Reference: src/sift.cpp:59-91.
void SIFT_NCL(InputArray image, std::vector<KeyPoint>& keypoints, OutputArray descriptors);

// OpenCV's own xfeatures2d SIFT (reference src/sift.cpp:49-57).  A third-party CPU path, outside this library:
// forwarded to cv::xfeatures2d::SIFT when built with -DSIFT_B200_WITH_XFEATURES2D against opencv_contrib, otherwise
// it throws cv::Exception.  (The typo in the name is the reference's API.)
void SITF_BuildIn_OpenCV(InputArray image, std::vector<KeyPoint>& keypoints, OutputArray descriptors);

// --- sub-modules ------------------------------------------------------------------------------------------------

// dst = src blurred with the reference's unnormalised truncated 2-D Gaussian (src/sift.cpp:95-153).
void Gaussian_Blur(Mat& src, Mat& dst, double sigma);

// The exported-but-unused separable variant that drops tap +w (src/sift.cpp:157-217); reproduced bit for bit.
void Gaussian_Blur_1D(Mat& src, Mat& dst, double sigma);

// gpyr[o*5+i], 5 scales per octave, every scale blurred from the octave base (src/sift.cpp:229-263).
void buildGaussianPyramid(Mat& image, std::vector<Mat>& gpyr, int nOctaves);

// dogpyr[o*4+i] = gpyr[o*5+i+1] - gpyr[o*5+i] (src/sift.cpp:265-283).
void buildDoGPyramid(std::vector<Mat>& gpyr, std::vector<Mat>& dogpyr, int nOctaves);

// 27-neighbour extrema with the literal threshold 8, Taylor refinement, orientation peaks (src/sift.cpp:547-577).
void findScaleSpaceExtrema(std::vector<Mat>& gpyr, std::vector<Mat>& dogpyr, std::vector<KeyPoint>& keypoints, int nOctaves);

// Fills the PRE-ALLOCATED N x 128 CV_32F `descriptors`, row i for keypoints[i] (src/sift.cpp:733-753).
void calDescriptor(std::vector<Mat>& gpyr, std::vector<KeyPoint>& keypoints, Mat& descriptors, int firstOctave);

#endif  // SIFT_HPP_
