// DistanceCalculator.hpp — drop-in for P/DistanceCalculator.hpp:17-48. Same macros, same
// global, same function names and parameter lists; the bodies (csrc/host/DistanceCalculator.cpp)
// marshal the vectors into the C-ABI (usv_moving_object_distance / usv_coordinate_position),
// i.e. the arithmetic runs in the sm_100a kernels of usv_distance.cu.
#ifndef DistanceCalculator_HPP
#define DistanceCalculator_HPP

#include <chrono>
#include <math.h>
#include <stdio.h>
#include <vector>

#include "usv_cv_compat.hpp"

#ifndef USV_NO_GLOBAL_USING
// the reference header does this at global scope (P/DistanceCalculator.hpp:12-14); kept so that
// its translation units compile unchanged
using namespace cv;
using namespace std;
using namespace std::chrono;
#endif

#define LeftCam true
#define RightCam false

#define XYFOVangle 70
#define ZYFOVangle 70
#define XPixelDimensions 640
#define YPixelDimensions 480
#define CameraDistcm 20.16
#define PI 3.14159265

// Global control variables (P/DistanceCalculator.cpp:6)
extern bool CoordinateDisplay;

double deg2rad(double deg);

double rad2deg(double rad);

// P/DistanceCalculator.cpp:15-88. Inputs by value, `dist` appended to; nothing is produced unless
// the three other-camera histories are non-empty (:28). Re-entrant (one GPU context per thread).
void MovingObjectDistanceCalculator(bool CameraSide, std::chrono::steady_clock::time_point ImgTimeStampThisCamera,
                                    std::vector<cv::Point2f> VectorCenter_pointThisCamera,
                                    std::vector<cv::Point2f> VectorCenter_pointOtherCamera,
                                    std::vector<cv::Point2f> OldVectorCenter_pointOtherCamera,
                                    std::vector<cv::Point2f> OlderVectorCenter_pointOtherCamera,
                                    std::vector<cv::Point2f> InterpolatedVectorCenter_pointOtherCamera,
                                    std::vector<cv::Point3i> InterframeMatchIndexesCompleteOtherCamera,
                                    std::chrono::steady_clock::time_point ImgTimeStampOtherCamera,
                                    std::chrono::steady_clock::time_point OldImgTimeStampOtherCamera,
                                    std::chrono::steady_clock::time_point OlderImgTimeStampOtherCamera,
                                    std::vector<double>& dist);

// P/DistanceCalculator.cpp:90-141; a no-op while CoordinateDisplay is false (:92).
void CooridinatePositionCalculator(bool CameraSide, std::vector<double> dist,
                                   std::vector<cv::Point2f> VectorCenter_pointThisCamera,
                                   std::vector<cv::Point3d>& PoscmFromReferencePointVector);

// Not in the reference: the two closed-form disparity -> distance maps it applies inline
// (pinhole, P/Main.cpp:694; power law, P/DistanceCalculator.cpp:84), batched on the GPU.
void DisparityToDistance(const std::vector<int>& disp, bool PowerLaw, std::vector<double>& dist);

#endif /* DistanceCalculator_HPP */
