// bshot_bits.h -- host C++ mirror of the reference's `class bshot` (include/bshot_bits.h:30-281)
// over the B200 C ABI (include/bshot_b200.h).  Same public members, same method names, same
// argument meaning, `void` methods that never throw: failures leave NaN SHOT / all-ones B-SHOT like
// PCL does and are readable through last_status()/bshot_last_error().
//
//   reference                                    here
//   minVect<T>            :6-20                   identical semantics (first minimum), header-only
//   bshot_descriptor      :23-27                  std::bitset<352> bits  (48 B, passed to the GPU as is)
//   calculate_normals(r)  :43-94                  bshot_set_cloud + bshot_set_keypoints + bshot_compute_normals
//   calculate_SHOT(r)     :113-135                bshot_compute_shot (fills cloud1_shot incl. rf[9])
//   compute_bshot()       :138-142                bits of the same call (already on the device)
//   compute_bshot_from_SHOT(cloud, out) :144-278  bshot_binarize
#ifndef BSHOT_B200_HOST_BSHOT_BITS_H
#define BSHOT_B200_HOST_BSHOT_BITS_H

#include "bshot_headers_bits.h"
#include "../../include/bshot_b200.h"

template <typename T>
T minVect(const T* v, int n, int* ind = NULL) {
    assert(n > 0);
    T best = v[0];
    int arg = 0;
    for (int i = 1; i < n; ++i)
        if (v[i] < best) { best = v[i]; arg = i; }
    if (ind != NULL) *ind = arg;
    return best;
}

class bshot_descriptor {
public:
    std::bitset<352> bits;
};
static_assert(sizeof(bshot_descriptor) == 48, "bshot_descriptor must be the 48-byte device record");

class bshot {
public:
    pcl::PointCloud<pcl::PointXYZ> cloud1, cloud2;
    pcl::PointCloud<pcl::Normal> cloud1_normals, cloud2_normals;
    pcl::PointCloud<pcl::PointXYZ> cloud1_keypoints, cloud2_keypoints;
    pcl::PointCloud<pcl::SHOT352> cloud1_shot, cloud2_shot;
    std::vector<bshot_descriptor> cloud1_bshot, cloud2_bshot;

    // B200 additions (defaults reproduce the reference): normals placement mode and capacities
    int normals_mode = BSHOT_NORMALS_REFERENCE;
    int normal_max_nn = 300;  // include/bshot_bits.h:68

    explicit bshot(int device = 0, size_t max_points = 1u << 18, size_t max_keypoints = 1u << 14, size_t max_targets = 1u << 20)
        : ctx_(nullptr), status_(BSHOT_OK) {
        status_ = bshot_ctx_create(&ctx_, device, max_points, max_keypoints, max_targets);
    }
    ~bshot() { bshot_ctx_destroy(ctx_); }
    bshot(const bshot&) = delete;
    bshot& operator=(const bshot&) = delete;

    bshot_ctx* context() { return ctx_; }
    int last_status() const { return status_; }

    void calculate_normals(float radius) {
        if (!upload()) return;
        cloud1_normals.is_dense = true;
        cloud1_normals.points.resize(cloud1.size());  // keeps old entries, new ones are (0,0,0)
        std::vector<float> n4(cloud1.size() * 4);
        status_ = bshot_compute_normals(ctx_, normals_mode, radius, normal_max_nn, n4.data());
        if (status_ != BSHOT_OK) return;
        for (size_t i = 0; i < cloud1.size(); ++i) {
            pcl::Normal& n = cloud1_normals.points[i];
            n.normal_x = n4[4 * i]; n.normal_y = n4[4 * i + 1]; n.normal_z = n4[4 * i + 2]; n.curvature = n4[4 * i + 3];
            if (std::isnan(n.normal_x)) cloud1_normals.is_dense = false;
        }
        normals_uploaded_ = true;
    }

    void calculate_SHOT(float radius) {
        if (!upload()) return;
        if (!normals_uploaded_) {  // caller filled cloud1_normals itself: hand them to the device
            if (cloud1_normals.size() != cloud1.size()) { status_ = BSHOT_E_STATE; return; }
            std::vector<float> n4(cloud1.size() * 4);
            for (size_t i = 0; i < cloud1.size(); ++i) {
                const pcl::Normal& n = cloud1_normals.points[i];
                n4[4 * i] = n.normal_x; n4[4 * i + 1] = n.normal_y; n4[4 * i + 2] = n.normal_z; n4[4 * i + 3] = n.curvature;
            }
            status_ = bshot_set_normals(ctx_, n4.data(), cloud1.size());
            if (status_ != BSHOT_OK) return;
        }
        const size_t k = cloud1_keypoints.size();
        std::vector<float> shot(k * 352), rf(k * 9);
        pending_bits_.resize(k);
        status_ = bshot_compute_shot(ctx_, radius, reinterpret_cast<uint64_t*>(pending_bits_.data()), shot.data(), rf.data(),
                                     nullptr, nullptr);
        if (status_ != BSHOT_OK) return;
        cloud1_shot.points.resize(k);
        cloud1_shot.width = (uint32_t)k; cloud1_shot.height = 1; cloud1_shot.is_dense = true;
        for (size_t i = 0; i < k; ++i) {
            std::memcpy(cloud1_shot.points[i].descriptor, &shot[352 * i], sizeof(float) * 352);
            std::memcpy(cloud1_shot.points[i].rf, &rf[9 * i], sizeof(float) * 9);
            if (std::isnan(shot[352 * i])) cloud1_shot.is_dense = false;
        }
        bits_valid_ = true;
    }

    void compute_bshot() {
        if (bits_valid_ && pending_bits_.size() == cloud1_shot.size()) cloud1_bshot = pending_bits_;  // fused on the device
        else compute_bshot_from_SHOT(cloud1_shot, cloud1_bshot);
    }

    void compute_bshot_from_SHOT(pcl::PointCloud<pcl::SHOT352>& shot_descriptors_here, std::vector<bshot_descriptor>& bshot_descriptors) {
        bshot_descriptors.resize(shot_descriptors_here.size());
        if (shot_descriptors_here.size() == 0) return;
        status_ = bshot_binarize(ctx_, shot_descriptors_here.points[0].descriptor, shot_descriptors_here.size(),
                                 sizeof(pcl::SHOT352) / sizeof(float), reinterpret_cast<uint64_t*>(bshot_descriptors.data()));
    }

private:
    bool upload() {
        if (!ctx_) { status_ = BSHOT_E_CUDA; return false; }
        // re-upload when the caller replaced the clouds (`cb.cloud1 = src_pcl_`, src/lidar_odometry.cpp:159-162)
        const void* key = cloud1.points.empty() ? nullptr : (const void*)cloud1.points.data();
        if (key != cloud_key_ || cloud1.size() != cloud_n_) {
            status_ = bshot_set_cloud(ctx_, cloud1.points.empty() ? nullptr : cloud1.points[0].data, cloud1.size(), sizeof(pcl::PointXYZ));
            if (status_ != BSHOT_OK) return false;
            cloud_key_ = key; cloud_n_ = cloud1.size();
            normals_uploaded_ = false; bits_valid_ = false; kp_key_ = nullptr;
        }
        const void* kkey = cloud1_keypoints.points.empty() ? nullptr : (const void*)cloud1_keypoints.points.data();
        if (kkey != kp_key_ || cloud1_keypoints.size() != kp_n_) {
            status_ = bshot_set_keypoints(ctx_, cloud1_keypoints.points.empty() ? nullptr : cloud1_keypoints.points[0].data,
                                          cloud1_keypoints.size(), sizeof(pcl::PointXYZ));
            if (status_ != BSHOT_OK) return false;
            kp_key_ = kkey; kp_n_ = cloud1_keypoints.size();
            bits_valid_ = false;
        }
        return true;
    }

    bshot_ctx* ctx_;
    int status_;
    const void* cloud_key_ = nullptr;
    const void* kp_key_ = nullptr;
    size_t cloud_n_ = 0, kp_n_ = 0;
    bool normals_uploaded_ = false, bits_valid_ = false;
    std::vector<bshot_descriptor> pending_bits_;
};

#endif
