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
#include "../b-shot-slam_b200/host/preprocess.h"

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
    if (argc >= 6) {  // lasers.bin (n x {double az, double vert, u16 dist}) -> cloud.bin through the Preprocessor shim
        std::ifstream lf(argv[4], std::ios::binary | std::ios::ate);
        const size_t n = (size_t)lf.tellg() / 18;
        lf.seekg(0);
        std::vector<double> az(n), ve(n);
        std::vector<unsigned short> di(n);
        lf.read(reinterpret_cast<char*>(az.data()), (std::streamsize)(8 * n));
        lf.read(reinterpret_cast<char*>(ve.data()), (std::streamsize)(8 * n));
        lf.read(reinterpret_cast<char*>(di.data()), (std::streamsize)(2 * n));
        std::vector<velodyne::Laser> lasers(n);
        std::vector<double> ring;
        for (size_t i = 0; i < n; ++i) {
            lasers[i] = velodyne::Laser{az[i], ve[i], di[i], 0, (unsigned char)(i % 32), 0};
            if (std::find(ring.begin(), ring.end(), ve[i]) == ring.end()) ring.push_back(ve[i]);
        }
        auto pc = std::make_shared<std::vector<Vector3f>>();
        myslam::Preprocessor pre(lo.cb.context());       // test/odometry_test.cpp:114-125
        pre.setVerticalAngles(ring);
        pre.setVerticalInitial(-0.6);
        pre.setLowPtThreshold(-1950);
        pre.setPointCloud(pc);
        pre.setLasers(lasers);
        pre.run();
        if (pre.last_status() != BSHOT_OK) { std::printf("preprocess: %s\n", bshot_last_error()); return 1; }
        std::ofstream po(argv[5], std::ios::binary);
        for (auto& p : *pc) po.write(reinterpret_cast<const char*>(p.v), 12);
        std::printf("preprocessor: %zu returns -> %zu points\n", n, pc->size());
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
        // ... and its positions, the correspondences RANSAC kept, the gate's verdict and the pose (test/odometry_test.cpp:178-181)
        for (auto& p : lo.cb.cloud2_keypoints.points) out.write(reinterpret_cast<const char*>(p.data), 12);
        lo.evaluateEstimation();
        lo.poseEstimation();
        lo.updateMap();
        if (lo.last_status() != BSHOT_OK) { std::printf("frame %d (estimation): %s\n", fidx, bshot_last_error()); return 1; }
        const int nr = (int)lo.corr.size(), upd = lo.shouldUpdateMap ? 1 : 0;
        out.write(reinterpret_cast<const char*>(&nr), 4);
        for (auto& c : lo.corr) { out.write(reinterpret_cast<const char*>(&c.index_query), 4); out.write(reinterpret_cast<const char*>(&c.index_match), 4); }
        out.write(reinterpret_cast<const char*>(&upd), 4);
        float Tr[16], Tb[16];
        myslam::LidarOdometry::to_row_major(lo.getRansacTransformation(), Tr);
        myslam::LidarOdometry::to_row_major(f->getPose(), Tb);
        out.write(reinterpret_cast<const char*>(Tr), 64);
        out.write(reinterpret_cast<const char*>(Tb), 64);
        std::printf("frame %d: %d keypoints, %d correspondences, %d after RANSAC, map %d\n", fidx, k, nc, nr, lo.map().size());
    }
    return 0;
}
