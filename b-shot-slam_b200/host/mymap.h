// mymap.h -- mirror of the reference's global keypoint map (include/mymap.h:8-51, src/mymap.cpp).
// Host-side storage only (SURVEY 8a row a9 stays on the host; a GPU-resident map is 8f "next" #2).
// Same admission rule (800 mm / seg-ratio, src/mymap.cpp:15-24), same 10 m blocks, same range
// gather.  getKeypoints() hands out descriptors as contiguous 48-byte records that go to
// bshot_match / bshot_map_append unchanged.  Iteration order inside a block is insertion order
// here (the reference's unordered_map order is implementation defined).
#ifndef BSHOT_B200_HOST_MYMAP_H
#define BSHOT_B200_HOST_MYMAP_H

#include <unordered_map>

#include "keypoint.h"

namespace myslam {

class Map {
public:
    typedef std::shared_ptr<Map> Ptr;
    typedef std::vector<Vector3f> KPointCloud;
    struct Block {
        std::vector<Keypoint::Ptr> kps;
        Keypoint::Ptr* find(const Vector3f& p) {
            for (auto& k : kps) if (k->getPosition() == p) return &k;
            return nullptr;
        }
        size_t size() const { return kps.size(); }
    };
    typedef std::unordered_map<unsigned long, Block> BlockMap;

    Map() {}

    void addKeypoint(Keypoint::Ptr keypoint) {
        const unsigned long id = getBlockID(keypoint->getPosition());
        auto it = keypoints_.find(id);
        if (it == keypoints_.end()) {
            keypoints_[id].kps.push_back(keypoint);
            return;
        }
        for (auto& kp : it->second.kps)  // rejected if a stored neighbour within 800 mm is at least as salient
            if ((keypoint->getPosition() - kp->getPosition()).norm() < 800 && keypoint->getSegRatio() <= kp->getSegRatio()) return;
        if (Keypoint::Ptr* same = it->second.find(keypoint->getPosition())) *same = keypoint;  // overwrite at equal key
        else it->second.kps.push_back(keypoint);
    }

    void getKeypoints(Vector3f pos, float range, pcl::PointCloud<pcl::PointXYZ>& kpts_pos, std::vector<bshot_descriptor>& descriptors) {
        kpts_pos.clear();
        descriptors.clear();
        int lo[3], hi[3];
        for (int a = 0; a < 3; ++a) {
            lo[a] = int(std::round((pos[a] - range) / prec)) * prec;
            hi[a] = int(std::round((pos[a] + range) / prec)) * prec;
        }
        for (int x = lo[0]; x <= hi[0]; x += prec)
            for (int y = lo[1]; y <= hi[1]; y += prec)
                for (int z = lo[2]; z <= hi[2]; z += prec) {
                    auto it = keypoints_.find(getBlockID(Vector3f((float)x, (float)y, (float)z)));
                    if (it == keypoints_.end()) continue;
                    for (auto& kp : it->second.kps) {
                        kpts_pos.points.push_back(eigenPt2PclPt(kp->getPosition()));
                        descriptors.push_back(kp->getDescriptor());
                    }
                }
        kpts_pos.width = (uint32_t)kpts_pos.points.size();
        kpts_pos.height = 1;
    }

    void getAllKeypoints(std::vector<Vector3f>& vec) {
        vec.clear();
        for (auto& b : keypoints_) for (auto& kp : b.second.kps) vec.push_back(kp->getPosition());
    }
    void getBlockKeypoints(std::vector<KPointCloud>& kpc) {
        for (auto& b : keypoints_) {
            KPointCloud t;
            for (auto& kp : b.second.kps) t.push_back(kp->getPosition());
            kpc.push_back(t);
        }
    }
    // 64-bit block key = 21 bits per axis of the position rounded to the 10 m lattice (src/mymap.cpp:103-112)
    unsigned long getBlockID(Vector3f pos) {
        unsigned long key = 0;
        for (int a = 0; a < 3; ++a) {
            const int g = int(std::round(pos[a] / prec)) * prec;
            key = (key << 21) | ((unsigned long)(long)g & 0x1FFFFFul);
        }
        return key;
    }
    int size() {
        int n = 0;
        for (auto& b : keypoints_) n += (int)b.second.size();
        return n;
    }
    inline pcl::PointXYZ eigenPt2PclPt(Vector3f pt) { return pcl::PointXYZ(pt[0], pt[1], pt[2]); }

private:
    BlockMap keypoints_;
    int prec = 10000;  // block edge (mm)
};

}  // namespace myslam
#endif
