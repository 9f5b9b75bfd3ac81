// ref_shim.cpp -- C entry points around the REFERENCE's own header, compiled unchanged from
// /root/reference/include/bshot_bits.h (PCL names resolve to oracle/pcl_stub, see stub_core.h for exactly
// which arithmetic is reference code and which is the oracle's PCL restatement).
//
// TEST INFRASTRUCTURE.  Built only where /root/reference exists (oracle/Makefile target _ref) into
// oracle/_ref/libbshot_ref.so; the .so travels to the GPU box, the reference sources do not.  Used to pin
// orc_bshot / orc_match / orc_compute_descriptors and the GPU kernels to reference-compiled code
// (tests/test_ref_pin.py) and to generate tests/golden/ref_pin.npz (tests/golden/make_ref_pin.py).
#include <assert.h>
#include <stdint.h>
#include <string.h>

#include <bshot_bits.h>  // the reference header itself (-I/root/reference/include)

static_assert(sizeof(bshot_descriptor) == 48, "bshot_descriptor must be the 48-byte std::bitset<352> record");

extern "C" {

// bshot::compute_bshot_from_SHOT (include/bshot_bits.h:144-278) on n x 352 floats -> n x 48 bytes
void ref_bshot_from_shot(const float* shot352, size_t n, uint64_t* bits6_out) {
    pcl::PointCloud<pcl::SHOT352> shots;
    shots.resize(n);
    for (size_t i = 0; i < n; ++i) memcpy(shots[i].descriptor, shot352 + 352 * i, sizeof(float) * 352);
    std::vector<bshot_descriptor> out;
    bshot cb;
    cb.compute_bshot_from_SHOT(shots, out);
    for (size_t i = 0; i < n; ++i) memcpy(bits6_out + 6 * i, &out[i], 48);
}

// minVect<int> (include/bshot_bits.h:6-20)
int ref_minvect_int(const int* v, int n, int* ind) { return minVect(v, n, ind); }

// The matching loops of LidarOdometry::featureMatching (src/lidar_odometry.cpp:212-242) cannot be compiled
// (that file needs Sophus / g2o / OpenCV); they are re-stated here line for line around the reference's own
// minVect and bshot_descriptor (std::bitset<352> XOR + count()).  pairs_out may be NULL; returns #pairs.
int ref_feature_matching(const uint64_t* q, size_t nq, const uint64_t* t, size_t nt, int* left_nn, int* right_nn, int* pairs_out) {
    std::vector<bshot_descriptor> c1(nq), c2(nt);
    for (size_t i = 0; i < nq; ++i) memcpy(&c1[i], q + 6 * i, 48);
    for (size_t i = 0; i < nt; ++i) memcpy(&c2[i], t + 6 * i, 48);
    std::vector<int> dist(std::max(nq, nt) + 1);
    int min_ix;
    for (int i = 0; i < (int)nq; ++i) {
        for (int k = 0; k < (int)nt; ++k) dist[k] = (int)(c1[i].bits ^ c2[k].bits).count();
        minVect(dist.data(), (int)nt, &min_ix);
        left_nn[i] = min_ix;
    }
    for (int i = 0; i < (int)nt; ++i) {
        for (int k = 0; k < (int)nq; ++k) dist[k] = (int)(c2[i].bits ^ c1[k].bits).count();
        minVect(dist.data(), (int)nq, &min_ix);
        right_nn[i] = min_ix;
    }
    int n = 0;
    for (int i = 0; i < (int)nq; ++i)
        if (right_nn[left_nn[i]] == i) {
            if (pairs_out) { pairs_out[2 * n] = i; pairs_out[2 * n + 1] = left_nn[i]; }
            ++n;
        }
    return n;
}

// One `bshot cb` object living across frames like LidarOdometry::cb (include/lidar_odometry.h:57), driven the
// way LidarOdometry::extractKeypoints / computeDescriptors do (src/lidar_odometry.cpp:159-162,173-176):
// cloud1 = surface, cloud1_keypoints = keypoints, calculate_normals(r), calculate_SHOT(r), compute_bshot().
void* ref_cb_create() { return new bshot(); }
void ref_cb_destroy(void* h) { delete static_cast<bshot*>(h); }

void ref_cb_compute_descriptors(void* h, const float* xyz, size_t n, const float* kp_xyz, size_t k, float radius, uint64_t* bits6_out,
                                float* shot352_out, float* rf9_out, float* normals4_out /* n x 4 */) {
    bshot& cb = *static_cast<bshot*>(h);
    pcl::PointCloud<pcl::PointXYZ> cloud, kps;
    cloud.resize(n);
    for (size_t i = 0; i < n; ++i) cloud[i] = pcl::PointXYZ(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    kps.resize(k);
    for (size_t i = 0; i < k; ++i) kps[i] = pcl::PointXYZ(kp_xyz[3 * i], kp_xyz[3 * i + 1], kp_xyz[3 * i + 2]);
    cb.cloud1 = cloud;             // src/lidar_odometry.cpp:159
    cb.cloud1_keypoints = kps;     // :161
    cb.calculate_normals(radius);  // :174
    cb.calculate_SHOT(radius);     // :175
    cb.compute_bshot();            // :176
    assert(cb.cloud1_bshot.size() == k);
    for (size_t i = 0; i < k; ++i) {
        if (bits6_out) memcpy(bits6_out + 6 * i, &cb.cloud1_bshot[i], 48);
        if (shot352_out) memcpy(shot352_out + 352 * i, cb.cloud1_shot[i].descriptor, sizeof(float) * 352);
        if (rf9_out) memcpy(rf9_out + 9 * i, cb.cloud1_shot[i].rf, sizeof(float) * 9);
    }
    if (normals4_out)
        for (size_t i = 0; i < n; ++i) {
            const pcl::Normal& nn = cb.cloud1_normals[i];
            normals4_out[4 * i] = nn.normal_x; normals4_out[4 * i + 1] = nn.normal_y; normals4_out[4 * i + 2] = nn.normal_z;
            normals4_out[4 * i + 3] = nn.curvature;
        }
}

}  // extern "C"

// ---- the reference's preprocessor (src/preprocess.cpp), compiled unchanged into this library by oracle/Makefile -----------
// (its translation unit is built with the include guards of common_include.h / VelodyneCapture.h pre-defined and
// oracle/pre_stub/pre_stub.h force-included; here only the class declaration is needed, through the same route)
#include "ref_preprocess_decl.h"

extern "C" {
// Preprocessor::run() on one rotation: lasers (azimuth / vertical in degrees, distance in 2 mm units) -> kept points.
// returns the number of points (written up to cap)
size_t ref_preprocess_select(const double* azimuth_deg, const double* vertical_deg, const unsigned short* distance, size_t n,
                             const double* vert_angles_deg, size_t nv, double vert_init_rad, double lowpt_th, const int* select_list, size_t n_select,
                             int have_select_list, int save_selected, float* xyz_out, size_t cap) {
    std::vector<velodyne::Laser> lasers(n);
    for (size_t i = 0; i < n; ++i) {
        lasers[i].azimuth = azimuth_deg[i];
        lasers[i].vertical = vertical_deg[i];
        lasers[i].distance = distance[i];
        lasers[i].intensity = 0;
        lasers[i].id = (unsigned char)(i % 256);
        lasers[i].time = 0;
    }
    std::vector<double> va(vert_angles_deg, vert_angles_deg + nv);
    std::sort(va.begin(), va.end());                       // test/odometry_test.cpp:111-112
    auto pc = std::make_shared<std::vector<Vector3f>>();
    myslam::Preprocessor pre;
    pre.setVerticalAngles(va);                              // :115-117
    pre.setVerticalInitial(vert_init_rad);
    pre.setLowPtThreshold(lowpt_th);
    pre.setPointCloud(pc);                                  // :125
    pre.setLasers(lasers);                                  // :143
    if (have_select_list) {                                 // :144-157
        std::vector<int> list(select_list, select_list + n_select);
        pre.haveSelectList(true);
        pre.saveSelectPoints(save_selected != 0);
        pre.setSelectedPoints(list);
    } else {
        pre.haveSelectList(false);
        pre.saveSelectPoints(save_selected != 0);
    }
    pre.run();
    for (size_t i = 0; i < pc->size() && i < cap; ++i) { xyz_out[3 * i] = (*pc)[i][0]; xyz_out[3 * i + 1] = (*pc)[i][1]; xyz_out[3 * i + 2] = (*pc)[i][2]; }
    return pc->size();
}

size_t ref_preprocess(const double* azimuth_deg, const double* vertical_deg, const unsigned short* distance, size_t n,
                      const double* vert_angles_deg, size_t nv, double vert_init_rad, double lowpt_th, float* xyz_out, size_t cap) {
    return ref_preprocess_select(azimuth_deg, vertical_deg, distance, n, vert_angles_deg, nv, vert_init_rad, lowpt_th, nullptr, 0, 0, 1, xyz_out, cap);
}
}
