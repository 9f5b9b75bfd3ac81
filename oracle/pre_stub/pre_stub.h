// pre_stub.h -- what /root/reference/src/preprocess.cpp needs from Eigen / OpenCV / VelodyneCapture.h, so that THAT SOURCE
// FILE COMPILES UNCHANGED here (none of those libraries is installed).  TEST INFRASTRUCTURE (oracle/), never linked by the
// product.  The recipe (oracle/Makefile) defines the include guards of the reference's common_include.h and
// VelodyneCapture.h on the command line, so those headers expand to nothing, and force-includes this file instead.
// Reference code when oracle/_ref runs: all of Preprocessor (readFrame, removeGround, removeOccluded, writePointCloud:
// every threshold, every float / double conversion, the std::map ordering).  Not reference code: the 20 lines below
// (a 3-float vector with Eigen's constructor / operator[] / operator- / norm semantics, the Laser record, CV_PI).
#pragma once
#include <algorithm>
#include <bitset>
#include <cmath>
#include <ctime>
#include <fstream>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#ifndef CV_PI
#define CV_PI 3.1415926535897932384626433832795
#endif

struct Vector3f {  // Eigen::Vector3f: float storage, scalar constructor arguments converted to float
    float v[3];
    Vector3f() : v{0.f, 0.f, 0.f} {}
    Vector3f(double x, double y, double z) : v{(float)x, (float)y, (float)z} {}
    float& operator[](int i) { return v[i]; }
    const float& operator[](int i) const { return v[i]; }
    Vector3f operator-(const Vector3f& o) const { Vector3f r; r.v[0] = v[0] - o.v[0]; r.v[1] = v[1] - o.v[1]; r.v[2] = v[2] - o.v[2]; return r; }
    float norm() const { return std::sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]); }  // sqrt(squaredNorm()), float
};

namespace velodyne {
struct Laser {  // include/VelodyneCapture.h:43-60
    double azimuth;
    double vertical;
    unsigned short distance;
    unsigned char intensity;
    unsigned char id;
    long long time;
};
}  // namespace velodyne

using namespace std;  // common_include.h relies on `using namespace std` further down the include chain
