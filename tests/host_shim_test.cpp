// host_shim_test.cpp -- drives the reference-named C++ shims the way test/odometry_test.cpp:174-180
// drives the reference (setSrcFrame -> extractKeypoints -> computeDescriptors -> featureMatching),
// on a cloud read from a raw float32 xyz file.  Prints a digest that tests/test_host_shim.py checks
// against the oracle.  usage: host_shim_test cloud0.bin cloud1.bin out.bin
#include <cstdio>
#include <cstdlib>
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
