// keypoint.h -- mirror of the reference's Keypoint (include/keypoint.h:8-32, src/keypoint.cpp:23-32)
#ifndef BSHOT_B200_HOST_KEYPOINT_H
#define BSHOT_B200_HOST_KEYPOINT_H

#include "bshot_bits.h"

namespace myslam {

class Keypoint {
public:
    typedef std::shared_ptr<Keypoint> Ptr;
    Keypoint() : id_((unsigned long)-1), pos_(0, 0, 0), seg_ratio_(0) {}
    Keypoint(unsigned long id, Vector3f& position, float& seg_ratio, bshot_descriptor& descriptor)
        : id_(id), pos_(position), seg_ratio_(seg_ratio), descriptor_(descriptor) {}

    inline Vector3f getPosition() const { return pos_; }
    inline bshot_descriptor getDescriptor() const { return descriptor_; }
    inline unsigned long getId() const { return id_; }
    inline float getSegRatio() const { return seg_ratio_; }

    // positions are snapped to a 10 mm lattice by truncation (src/keypoint.cpp:25-29)
    static Keypoint::Ptr createKeypoint(Vector3f& pos, float seg_ratio, bshot_descriptor descriptor) {
        const int prec = 10;
        Vector3f snapped((float)(int(std::trunc(pos[0] / prec)) * prec), (float)(int(std::trunc(pos[1] / prec)) * prec),
                         (float)(int(std::trunc(pos[2] / prec)) * prec));
        return std::make_shared<Keypoint>(next_id()++, snapped, seg_ratio, descriptor);
    }

private:
    static unsigned long& next_id() { static unsigned long id = 0; return id; }
    unsigned long id_;
    Vector3f pos_;
    float seg_ratio_;
    bshot_descriptor descriptor_;
};

}  // namespace myslam
#endif
