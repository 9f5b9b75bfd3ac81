// preprocess.h -- the reference's scan preprocessor class (include/preprocess.h:8-60) over the B200 C ABI: same setters, same
// run(); the range image, the ground / self-car / occlusion passes and the point writer of src/preprocess.cpp:38-227 run on the
// device (bshot_preprocess_select).  getRangeImage / getRemoveMap / getSelMap (debug views of the std::map range image) are not
// carried over: the device keeps columns, not maps.
#ifndef BSHOT_B200_HOST_PREPROCESS_H
#define BSHOT_B200_HOST_PREPROCESS_H

#include <algorithm>
#include <memory>
#include <vector>

#include "bshot_headers_bits.h"

namespace velodyne {
// the capture class' return record (include/VelodyneCapture.h:43-60)
struct Laser {
    double azimuth;
    double vertical;
    unsigned short distance;
    unsigned char intensity;
    unsigned char id;
    long long time;
    bool operator<(const Laser& o) const { return azimuth == o.azimuth ? id < o.id : azimuth < o.azimuth; }
};
}  // namespace velodyne

namespace myslam {

class Preprocessor {
public:
    typedef std::shared_ptr<Preprocessor> Ptr;

    explicit Preprocessor(bshot_ctx* ctx = nullptr) : ctx_(ctx) {}
    Preprocessor(bshot_ctx* ctx, std::vector<velodyne::Laser>& lasers, std::vector<double>& vertAngle, std::shared_ptr<std::vector<Vector3f>> pc)
        : ctx_(ctx), vertAngle_(vertAngle), pc_(std::move(pc)), lasers_(lasers) {
        std::sort(vertAngle_.begin(), vertAngle_.end());
    }

    void setContext(bshot_ctx* ctx) { ctx_ = ctx; }
    void setLasers(std::vector<velodyne::Laser>& lasers) { lasers_ = lasers; }
    void setSelectedPoints(std::vector<int>& selptlist) { selpts_ = selptlist; std::sort(selpts_.begin(), selpts_.end()); }
    void saveSelectPoints(bool savesel) { save_sel_ = savesel; }
    void haveSelectList(bool havesellist) { have_sel_list_ = havesellist; }
    void setVerticalAngles(std::vector<double>& vertAngle) { vertAngle_ = vertAngle; std::sort(vertAngle_.begin(), vertAngle_.end()); }
    void setVerticalInitial(double vertinit) { vert_init_ = vertinit; }
    void setLowPtThreshold(double lowptth) { lowpt_th = lowptth; }
    void setPointCloud(std::shared_ptr<std::vector<Vector3f>> pc) { pc_ = std::move(pc); }

    // src/preprocess.cpp:213-223; an empty laser list leaves an empty cloud (the reference would spin in readFrame)
    void run() {
        pc_->clear();
        const size_t n = lasers_.size();
        az_.resize(n); vert_.resize(n); dist_.resize(n); xyz_.resize(3 * std::max<size_t>(n, 1));
        for (size_t i = 0; i < n; ++i) { az_[i] = lasers_[i].azimuth; vert_[i] = lasers_[i].vertical; dist_[i] = lasers_[i].distance; }
        size_t kept = 0;
        last_status_ = bshot_preprocess_select(ctx_, az_.data(), vert_.data(), dist_.data(), n, vertAngle_.data(), vertAngle_.size(), vert_init_, lowpt_th,
                                               selpts_.data(), selpts_.size(), have_sel_list_ ? 1 : 0, save_sel_ ? 1 : 0, xyz_.data(), n, &kept);
        if (last_status_ != BSHOT_OK) return;
        pc_->reserve(kept);
        for (size_t i = 0; i < kept; ++i) pc_->push_back(Vector3f(xyz_[3 * i], xyz_[3 * i + 1], xyz_[3 * i + 2]));
    }
    int last_status() const { return last_status_; }

private:
    bshot_ctx* ctx_;
    std::vector<double> vertAngle_;
    double vert_init_ = -0.6;   // radian (src/preprocess.cpp:7)
    double lowpt_th = -2000;    // include/preprocess.h:44
    std::shared_ptr<std::vector<Vector3f>> pc_;
    std::vector<velodyne::Laser> lasers_;
    std::vector<int> selpts_;
    bool save_sel_ = true, have_sel_list_ = false;
    std::vector<double> az_, vert_;
    std::vector<unsigned short> dist_;
    std::vector<float> xyz_;
    int last_status_ = BSHOT_OK;
};

}  // namespace myslam
#endif
