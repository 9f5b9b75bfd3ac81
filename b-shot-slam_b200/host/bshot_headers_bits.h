// bshot_headers_bits.h -- stand-ins for the PCL / Eigen types the reference's front end exposes.
//
// The reference header of the same name (include/bshot_headers_bits.h:12-27) pulls in 16 PCL
// headers.  The drop-in keeps only the value types that cross the boundary, with the member names
// and memory layouts of the originals, and nothing of PCL's algorithms:
//   pcl::PointXYZ 16 B (x,y,z,pad)      pcl::Normal 32 B (normal[3],pad,curvature,pad[3])
//   pcl::SHOT352  1444 B (descriptor[352], rf[9])     pcl::Correspondence (index_query,index_match,distance)
//   pcl::PointCloud<T> { points, width, height, is_dense }   Eigen-like Vector3f / Matrix4f
// A maintainer who builds against real PCL/Eigen defines BSHOT_B200_USE_PCL and gets the originals.
#ifndef BSHOT_B200_HOST_HEADERS_BITS_H
#define BSHOT_B200_HOST_HEADERS_BITS_H

#include <bitset>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#ifdef BSHOT_B200_USE_PCL
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/correspondence.h>
#include <Eigen/Core>
using Eigen::Matrix4f;
using Eigen::Vector3f;
#else

struct Vector3f {
    float v[3];
    Vector3f() : v{0, 0, 0} {}
    Vector3f(float x, float y, float z) : v{x, y, z} {}
    float& operator[](int i) { return v[i]; }
    float operator[](int i) const { return v[i]; }
    Vector3f operator-(const Vector3f& o) const { return Vector3f(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
    Vector3f operator+(const Vector3f& o) const { return Vector3f(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
    float dot(const Vector3f& o) const { return v[0] * o.v[0] + v[1] * o.v[1] + v[2] * o.v[2]; }
    float norm() const { return std::sqrt(dot(*this)); }
    float sum() const { return v[0] + v[1] + v[2]; }
    bool operator==(const Vector3f& o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2]; }
};
static_assert(sizeof(Vector3f) == 12, "Eigen::Vector3f layout");

struct Matrix4f {  // column major like Eigen
    float m[16];
    static Matrix4f Identity() {
        Matrix4f r;
        for (int i = 0; i < 16; ++i) r.m[i] = (i % 5 == 0) ? 1.0f : 0.0f;
        return r;
    }
    float& operator()(int r, int c) { return m[c * 4 + r]; }
    float operator()(int r, int c) const { return m[c * 4 + r]; }
    Vector3f translation() const { return Vector3f(m[12], m[13], m[14]); }  // topRightCorner<3,1>()
    Vector3f apply(const Vector3f& p) const {
        return Vector3f(m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12], m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13],
                        m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14]);
    }
};

namespace pcl {

struct alignas(16) PointXYZ {
    union { float data[4]; struct { float x, y, z; }; };
    PointXYZ() : data{0.f, 0.f, 0.f, 1.f} {}
    PointXYZ(float x_, float y_, float z_) : data{x_, y_, z_, 1.f} {}
};
static_assert(sizeof(PointXYZ) == 16, "pcl::PointXYZ layout");

struct alignas(16) Normal {
    union { float data_n[4]; float normal[3]; struct { float normal_x, normal_y, normal_z; }; };
    union { struct { float curvature; }; float data_c[4]; };
    Normal() : data_n{0.f, 0.f, 0.f, 0.f}, data_c{0.f, 0.f, 0.f, 0.f} {}
};
static_assert(sizeof(Normal) == 32, "pcl::Normal layout");

struct SHOT352 {
    float descriptor[352];
    float rf[9];
    static int descriptorSize() { return 352; }
};
static_assert(sizeof(SHOT352) == 1444, "pcl::SHOT352 layout");

struct Correspondence {
    int index_query;
    int index_match;
    float distance;
    Correspondence() : index_query(0), index_match(-1), distance(std::numeric_limits<float>::max()) {}
};
typedef std::vector<Correspondence> Correspondences;

template <typename T>
struct PointCloud {
    std::vector<T> points;
    uint32_t width = 0, height = 0;
    bool is_dense = true;
    size_t size() const { return points.size(); }
    void clear() { points.clear(); width = height = 0; }
    void push_back(const T& p) { points.push_back(p); width = (uint32_t)points.size(); height = 1; }
    T& operator[](size_t i) { return points[i]; }
    const T& operator[](size_t i) const { return points[i]; }
    PointCloud& operator+=(const PointCloud& o) {
        points.insert(points.end(), o.points.begin(), o.points.end());
        width = (uint32_t)points.size(); height = 1;
        return *this;
    }
};

inline bool isFinite(const PointXYZ& p) { return std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z); }

}  // namespace pcl
#endif  // BSHOT_B200_USE_PCL

#endif
