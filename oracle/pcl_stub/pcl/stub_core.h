// stub_core.h -- minimal stand-ins for the PCL / Eigen names that /root/reference/include/bshot_bits.h
// uses, so that THAT HEADER COMPILES UNCHANGED here (PCL, Eigen, FLANN are not installed).
//
// TEST INFRASTRUCTURE (oracle/): nothing in the product links or includes this.
// What is reference code and what is not, when oracle/_ref/libbshot_ref.so runs:
//   * reference, unchanged, compiled from /root/reference/include/bshot_bits.h: minVect (:6-20),
//     bshot_descriptor (:23-27), bshot::calculate_normals (:43-94, incl. the keypoint-ordinal placement
//     quirk and the NaN branch), bshot::calculate_SHOT (:113-135), bshot::compute_bshot /
//     compute_bshot_from_SHOT (:138-278, the whole binarisation arithmetic);
//   * NOT reference: the PCL calls inside those methods resolve to the classes below, which forward to
//     the oracle's restatement of the published PCL algorithms (oracle/bshot_oracle.cpp, SURVEY.md
//     Appendix A).  So compute_bshot_from_SHOT and minVect are pinned to reference-compiled code
//     outright; calculate_normals / calculate_SHOT pin the reference's control flow and data
//     placement around the oracle's PCL primitives.
#pragma once
#include <cmath>
#include <cstddef>
#include <limits>
#include <memory>
#include <vector>

#include "../../bshot_oracle.h"

namespace Eigen {
struct Vector4f {
    float v[4];
    Vector4f() : v{0, 0, 0, 0} {}
    float& operator[](int i) { return v[i]; }
    const float& operator[](int i) const { return v[i]; }
};
}  // namespace Eigen

namespace pcl {
namespace io {}
namespace console {}

struct PointXYZ {
    union {
        float data[4];
        struct { float x, y, z; };
    };
    PointXYZ() : data{0.f, 0.f, 0.f, 1.f} {}
    PointXYZ(float x_, float y_, float z_) : data{x_, y_, z_, 1.f} {}
};

struct Normal {
    union {
        float data_n[4];
        float normal[3];
        struct { float normal_x, normal_y, normal_z; };
    };
    union {
        struct { float curvature; };
        float data_c[4];
    };
    Normal() : data_n{0.f, 0.f, 0.f, 0.f}, data_c{0.f, 0.f, 0.f, 0.f} {}
};

struct SHOT352 {
    float descriptor[352];
    float rf[9];
};

template <typename T>
struct PointCloud {
    typedef std::shared_ptr<PointCloud<T>> Ptr;
    typedef std::shared_ptr<const PointCloud<T>> ConstPtr;
    std::vector<T> points;
    unsigned width = 0, height = 0;
    bool is_dense = true;
    size_t size() const { return points.size(); }
    void clear() { points.clear(); width = height = 0; }
    void resize(size_t n) { points.resize(n); width = (unsigned)n; height = 1; }
    void push_back(const T& p) { points.push_back(p); width = (unsigned)points.size(); height = 1; }
    T& operator[](size_t i) { return points[i]; }
    const T& operator[](size_t i) const { return points[i]; }
    Ptr makeShared() const { return Ptr(new PointCloud<T>(*this)); }
};

inline bool isFinite(const PointXYZ& p) { return std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z); }

// the surface cloud a search object was given (also what computePointNormal's `cloud` argument is)
struct StubSurface {
    orc_cloud* c = nullptr;
    const void* key = nullptr;
    size_t n = 0;
    ~StubSurface() { if (c) orc_cloud_destroy(c); }
    void set(const PointCloud<PointXYZ>& cloud) {
        if (c) orc_cloud_destroy(c);
        c = orc_cloud_create(cloud.points.empty() ? nullptr : cloud.points[0].data, cloud.size(), 4);
        n = cloud.size();
    }
};

namespace search {
template <typename PointT>
class KdTree {
public:
    typedef std::shared_ptr<KdTree<PointT>> Ptr;
    std::shared_ptr<StubSurface> surf;
    void setInputCloud(const typename PointCloud<PointT>::ConstPtr& cloud) {
        surf.reset(new StubSurface());
        surf->set(*cloud);
    }
    // pcl::KdTreeFLANN::radiusSearch -> oracle restatement
    int radiusSearch(const PointT& p, double radius, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances,
                     unsigned int max_nn = 0) const {
        const float q[3] = {p.x, p.y, p.z};
        const int cap = (int)surf->n;
        k_indices.resize(cap);
        k_sqr_distances.resize(cap);
        const int n = orc_radius_search(surf->c, q, (float)radius, (int)max_nn, k_indices.data(), k_sqr_distances.data(), cap);
        k_indices.resize(n);
        k_sqr_distances.resize(n);
        return n;
    }
};
}  // namespace search

template <typename PointT>
class KdTreeFLANN : public search::KdTree<PointT> {};

// pcl::computePointNormal(cloud, indices, plane_parameters, curvature) -> oracle restatement (Appendix A.3)
template <typename PointT>
inline bool computePointNormal(const PointCloud<PointT>& cloud, const std::vector<int>& indices, Eigen::Vector4f& plane_parameters,
                               float& curvature) {
    float out[4];
    orc_point_normal_indices(cloud.points.empty() ? nullptr : cloud.points[0].data, cloud.size(), 4, indices.data(), (int)indices.size(), out);
    plane_parameters[0] = out[0]; plane_parameters[1] = out[1]; plane_parameters[2] = out[2]; plane_parameters[3] = 0.0f;
    curvature = out[3];
    return std::isfinite(out[0]);
}

// pcl::flipNormalTowardsViewpoint (features/normal_3d.h): flip when the normal points away from the viewpoint
template <typename PointT>
inline void flipNormalTowardsViewpoint(const PointT& point, float vp_x, float vp_y, float vp_z, float& nx, float& ny, float& nz) {
    vp_x -= point.x; vp_y -= point.y; vp_z -= point.z;
    const float cos_theta = (vp_x * nx + vp_y * ny + vp_z * nz);
    if (cos_theta < 0) { nx *= -1; ny *= -1; nz *= -1; }
}

template <typename PointInT, typename PointOutT>
class NormalEstimationOMP {  // the reference only configures it (include/bshot_bits.h:46-48), never computes with it
public:
    void setRadiusSearch(double) {}
    void setNumberOfThreads(unsigned) {}
};

template <typename PointT>
class VoxelGrid {  // calculate_voxel_grid_keypoints (include/bshot_bits.h:97-110) is off the path: declared only
public:
    void setLeafSize(float, float, float);
    void setInputCloud(const typename PointCloud<PointT>::ConstPtr&);
    void filter(PointCloud<PointT>&);
};

// pcl::SHOTEstimationOMP<PointXYZ, Normal, SHOT352>::compute -> oracle restatement (Appendix A.4 / A.5)
template <typename PointInT, typename PointNT, typename PointOutT>
class SHOTEstimationOMP {
    double radius_ = 0;
    typename PointCloud<PointInT>::ConstPtr input_, surface_;
    typename PointCloud<PointNT>::ConstPtr normals_;
public:
    void setRadiusSearch(double r) { radius_ = r; }
    void setNumberOfThreads(unsigned) {}
    void setSearchMethod(const typename search::KdTree<PointInT>::Ptr&) {}
    void setInputCloud(const typename PointCloud<PointInT>::ConstPtr& c) { input_ = c; }
    void setSearchSurface(const typename PointCloud<PointInT>::ConstPtr& c) { surface_ = c; }
    void setInputNormals(const typename PointCloud<PointNT>::ConstPtr& c) { normals_ = c; }
    void compute(PointCloud<PointOutT>& out) {
        const size_t k = input_->size(), n = surface_->size();
        StubSurface s;
        s.set(*surface_);
        std::vector<float> kp(3 * k), nrm(4 * n, 0.0f), shot(352 * k), rf(9 * k);
        for (size_t i = 0; i < k; ++i) { kp[3 * i] = (*input_)[i].x; kp[3 * i + 1] = (*input_)[i].y; kp[3 * i + 2] = (*input_)[i].z; }
        for (size_t i = 0; i < n && i < normals_->size(); ++i) {
            nrm[4 * i] = (*normals_)[i].normal_x; nrm[4 * i + 1] = (*normals_)[i].normal_y; nrm[4 * i + 2] = (*normals_)[i].normal_z;
            nrm[4 * i + 3] = (*normals_)[i].curvature;
        }
        orc_shot(s.c, kp.data(), k, (float)radius_, nrm.data(), nullptr, shot.data(), rf.data(), nullptr, 0);
        out.resize(k);
        for (size_t i = 0; i < k; ++i) {
            for (int j = 0; j < 352; ++j) out[i].descriptor[j] = shot[352 * i + j];
            for (int j = 0; j < 9; ++j) out[i].rf[j] = rf[9 * i + j];
        }
    }
};

}  // namespace pcl
