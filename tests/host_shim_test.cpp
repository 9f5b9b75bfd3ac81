// host_shim_test.cpp -- drives the reference-named C++ shims the way test/odometry_test.cpp:174-180
// drives the reference (setSrcFrame -> extractKeypoints -> computeDescriptors -> featureMatching),
// on a cloud read from a raw float32 xyz file.  Prints a digest that tests/test_host_shim.py checks
// against the oracle.  usage: host_shim_test cloud0.bin cloud1.bin out.bin
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <fstream>

#include "../b-shot-slam_b200/host/lidar_odometry.h"

static myslam::Frame::PCPtr load(const char* path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    const size_t bytes = (size_t)f.tellg();
    f.seekg(0);
    std::vector<float> raw(bytes / 4);
    f.read(reinterpret_cast<char*>(raw.data()), (std::streamsize)bytes);
    auto pc = std::make_shared<std::vector<Vector3f>>();
    for (size_t i = 0; i + 2 < raw.size(); i += 3) pc->push_back(Vector3f(raw[i], raw[i + 1], raw[i + 2]));
    return pc;
}

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    // minVect: first minimum wins (include/bshot_bits.h:6-20)
    int v[6] = {5, 3, 9, 3, 7, 3}, ind = -1;
    if (minVect(v, 6, &ind) != 3 || ind != 1) { std::printf("minVect broken\n"); return 1; }

    myslam::LidarOdometry lo(0);
    if (lo.last_status() != BSHOT_OK) { std::printf("ctx: %s\n", bshot_last_error()); return 1; }

    // direct `bshot` use (the reference's kp_test.cpp pattern) with SAME-SIZE clouds assigned one after the other:
    // vector copy-assignment reuses the buffer, so the shim must notice the new CONTENT, not a new address
    {
        myslam::Frame::PCPtr a = load(argv[1]), b = load(argv[2]);
        const size_t n = std::min(a->size(), b->size()) / 3, k = 200;
        auto fill = [&](bshot& cb, const myslam::Frame::PCPtr& pc) {
            pcl::PointCloud<pcl::PointXYZ> c, kp;
            for (size_t i = 0; i < n; ++i) c.push_back(pcl::PointXYZ((*pc)[3 * i][0], (*pc)[3 * i][1], (*pc)[3 * i][2]));
            for (size_t i = 0; i < k; ++i) kp.push_back(c.points[(i * 37) % n]);
            cb.cloud1 = c;               // copy-assign into the member
            cb.cloud1_keypoints = kp;
            cb.calculate_normals(3000);
            cb.calculate_SHOT(3000);
            cb.compute_bshot();
        };
        bshot reused(0, 1u << 17, 1u << 10, 1u << 10), fresh(0, 1u << 17, 1u << 10, 1u << 10);
        fill(reused, a);
        const void* addr_a = reused.cloud1.points.data();
        fill(reused, b);
        if (addr_a != (const void*)reused.cloud1.points.data()) std::printf("note: the vector buffer moved, address reuse not exercised\n");
        fill(fresh, b);
        if (reused.last_status() != BSHOT_OK || fresh.last_status() != BSHOT_OK) { std::printf("bshot: %s\n", bshot_last_error()); return 1; }
        if (reused.cloud1_bshot.size() != k || std::memcmp(reused.cloud1_bshot.data(), fresh.cloud1_bshot.data(), k * 48) != 0) {
            std::printf("stale device cloud: a same-size cloud was not re-uploaded\n");
            return 1;
        }
        std::printf("same-size reassignment ok\n");
    }
    std::ofstream out(argv[3], std::ios::binary);
    for (int fidx = 0; fidx < 2; ++fidx) {
        myslam::Frame::Ptr f = myslam::Frame::createFrame();
        f->setPointCloud(load(argv[1 + fidx]));
        if (fidx > 0) lo.passSrc2Ref();
        lo.setSrcFrame(f);
        lo.extractKeypoints();
        lo.computeDescriptors();
        lo.featureMatching();
        if (lo.last_status() != BSHOT_OK) { std::printf("frame %d: %s\n", fidx, bshot_last_error()); return 1; }
        const int k = (int)lo.cb.cloud1_bshot.size(), nc = (int)lo.corresp.size();
        out.write(reinterpret_cast<const char*>(&k), 4);
        out.write(reinterpret_cast<const char*>(&nc), 4);
        for (auto& p : *f->getKeypoints()) out.write(reinterpret_cast<const char*>(p.v), 12);
        out.write(reinterpret_cast<const char*>(lo.cb.cloud1_bshot.data()), (std::streamsize)k * 48);
        for (auto& s : lo.cb.cloud1_shot.points) out.write(reinterpret_cast<const char*>(s.rf), 36);
        for (auto& c : lo.corresp) { out.write(reinterpret_cast<const char*>(&c.index_query), 4); out.write(reinterpret_cast<const char*>(&c.index_match), 4); }
        // the target set featureMatching assembled (:187-206): map keypoints within 100 m, then the reference frame
        const int nt = (int)lo.cb.cloud2_bshot.size();
        out.write(reinterpret_cast<const char*>(&nt), 4);
        out.write(reinterpret_cast<const char*>(lo.cb.cloud2_bshot.data()), (std::streamsize)nt * 48);
        // the reference's updateMap (src/lidar_odometry.cpp:344-376): frame keypoints enter the global map
        for (int i = 0; i < k; ++i) {
            Vector3f pos = (*f->getKeypoints())[i];
            lo.map().addKeypoint(myslam::Keypoint::createKeypoint(pos, lo.seg_ratios_[i], lo.cb.cloud1_bshot[i]));
        }
        lo.setRun();
        std::printf("frame %d: %d keypoints, %d correspondences, map %d\n", fidx, k, nc, lo.map().size());
    }
    return 0;
}
