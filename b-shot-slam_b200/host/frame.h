// frame.h -- mirror of the reference's Frame holder (include/frame.h:9-51, src/frame.cpp)
#ifndef BSHOT_B200_HOST_FRAME_H
#define BSHOT_B200_HOST_FRAME_H

#include "bshot_headers_bits.h"

namespace myslam {

class Frame {
public:
    typedef std::shared_ptr<Frame> Ptr;
    typedef std::shared_ptr<std::vector<Vector3f>> PCPtr;
    typedef std::shared_ptr<std::vector<std::bitset<352>>> DCPPtr;  // 48 B records == device layout
    unsigned long id_;
    long long timestamp_;
    Matrix4f T_c_w_;
    PCPtr pointcloud_;
    PCPtr keypoints_;
    DCPPtr descriptors_;
    bool is_key_frame_;

    Frame() : id_((unsigned long)-1), timestamp_(-1), T_c_w_(Matrix4f::Identity()), is_key_frame_(false) {}
    Frame(long id, double time_stamp = 0, Matrix4f T_c_w = Matrix4f::Identity(), PCPtr pc = nullptr, PCPtr kps = nullptr,
          DCPPtr dcpts = nullptr, bool isKeyframe = false)
        : id_(id), timestamp_((long long)time_stamp), T_c_w_(T_c_w), pointcloud_(pc), keypoints_(kps), descriptors_(dcpts),
          is_key_frame_(isKeyframe) {}

    static Frame::Ptr createFrame() { static long factory_id = 0; return Frame::Ptr(new Frame(factory_id++)); }

    void setTimestamp(const long long timestamp) { timestamp_ = timestamp; }
    void setPose(const Matrix4f& T_c_w) { T_c_w_ = T_c_w; }
    void setPointCloud(PCPtr pc) { pointcloud_ = pc; }
    void setKeypoints(PCPtr kps) { keypoints_ = kps; }
    void setDescriptors(DCPPtr dcpts) { descriptors_ = dcpts; }
    unsigned long getID() { return id_; }
    long long getTimestamp() { return timestamp_; }
    Matrix4f getPose() { return T_c_w_; }
    PCPtr getPointCloud() { return pointcloud_; }
    PCPtr getKeypoints() { return keypoints_; }
    DCPPtr getDescriptors() { return descriptors_; }
    bool isKeyframe() { return is_key_frame_; }
};

}  // namespace myslam
#endif
