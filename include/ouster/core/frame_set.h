// frame_set.h -- the part of FrameSet the scan->pointcloud path touches
// (mirrors ouster_core/include/ouster/core/frame_set.h:218-346): a vector of shared LidarFrames in which
// empty slots are allowed; valid_indices() walks the slots that hold a frame.  Everything else of the
// reference class (collation bookkeeping, frame-set sources) is outside SURVEY 8's scope.
#pragma once
#include <initializer_list>
#include <memory>
#include <vector>

#include "ouster/core/lidar_frame.h"

namespace ouster {
namespace sdk {
namespace core {

class FrameSet {
   public:
    FrameSet() = default;
    FrameSet(const std::vector<std::shared_ptr<LidarFrame>>& frames) : frames_(frames) {}
    FrameSet(std::vector<std::shared_ptr<LidarFrame>>&& frames) : frames_(std::move(frames)) {}
    FrameSet(std::initializer_list<std::shared_ptr<LidarFrame>> frames) : frames_(frames) {}

    size_t size() const { return frames_.size(); }
    const std::shared_ptr<LidarFrame>& operator[](size_t index) const { return frames_.at(index); }
    std::shared_ptr<LidarFrame>& operator[](size_t index) { return frames_.at(index); }

    /// indices of the slots that hold a frame (frame_set.h:308-325)
    std::vector<size_t> valid_indices() const {
        std::vector<size_t> out;
        for (size_t i = 0; i < frames_.size(); ++i)
            if (frames_[i]) out.push_back(i);
        return out;
    }

   private:
    std::vector<std::shared_ptr<LidarFrame>> frames_;
};

}  // namespace core
}  // namespace sdk
}  // namespace ouster
