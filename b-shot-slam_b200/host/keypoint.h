// keypoint.h -- host-side map entry with the reference's public surface (include/keypoint.h:8-32,
// src/keypoint.cpp:23-32): id, position snapped to the 10 mm map lattice, seg-ratio, 352-bit descriptor.
// The descriptor member is the 48-byte record the device kernels read, so a std::vector of descriptors taken from
// keypoints can be handed to bshot_map_append() without repacking (descriptor_words()).
#ifndef BSHOT_B200_HOST_KEYPOINT_H
#define BSHOT_B200_HOST_KEYPOINT_H

#include "bshot_bits.h"

namespace myslam {

class Keypoint {
public:
    using Ptr = std::shared_ptr<Keypoint>;

    // lattice the reference snaps map keypoints to (src/keypoint.cpp:25-29: truncation towards zero, 10 mm)
    static constexpr int kLatticeMm = 10;
    static float snap(float v) { return (float)(int(std::trunc(v / kLatticeMm)) * kLatticeMm); }

    Keypoint() : id_(kNoId), pos_(0, 0, 0), seg_ratio_(0) {}

    // same argument list as the reference constructor (include/keypoint.h:12)
    Keypoint(unsigned long id, Vector3f& position, float& seg_ratio, bshot_descriptor& descriptor)
        : id_(id), pos_(position), seg_ratio_(seg_ratio), descriptor_(descriptor) {}

    // factory of the reference (src/keypoint.cpp:23-32): consecutive ids, snapped position
    static Ptr createKeypoint(Vector3f& pos, float seg_ratio, bshot_descriptor descriptor) {
        Vector3f on_lattice(snap(pos[0]), snap(pos[1]), snap(pos[2]));
        const unsigned long id = next_id()++;
        return std::make_shared<Keypoint>(id, on_lattice, seg_ratio, descriptor);
    }

    // ---- accessors, names as in the reference (include/keypoint.h:17-26) ----
    unsigned long getId() const { return id_; }
    Vector3f getPosition() const { return pos_; }
    float getSegRatio() const { return seg_ratio_; }
    bshot_descriptor getDescriptor() const { return descriptor_; }

    // ---- device-facing view: the 6 little-endian u64 words of the std::bitset<352> ----
    const uint64_t* descriptor_words() const { return reinterpret_cast<const uint64_t*>(&descriptor_); }
    bool has_id() const { return id_ != kNoId; }

private:
    static constexpr unsigned long kNoId = (unsigned long)-1;
    static unsigned long& next_id() {
        static unsigned long counter = 0;
        return counter;
    }

    unsigned long id_;
    Vector3f pos_;
    float seg_ratio_;
    bshot_descriptor descriptor_;
};

}  // namespace myslam
#endif
