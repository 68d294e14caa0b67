// frame_pipeline.h -- FramePipeline: a ring of LidarFrames driven through one FrameBatcher with
// several frames in flight (B200 extension; no counterpart in the reference, which decodes on the
// calling thread inside FrameBatcher::batch, ouster_core/src/lidar_frame.cpp:1698-1959).
//
// The host state machine of frame k+1 (header bookkeeping, column map) overlaps the H2D, fused
// kernel and D2H of frame k; consecutive frames alternate between two streams, so the upload of
// one frame and the download of the previous one use both PCIe directions at once.  Frames come
// out in order, `depth` frames after they went in (drain() at end of stream).
#pragma once
#include <deque>
#include <memory>
#include <vector>

#include "ouster/core/lidar_frame.h"

namespace ouster {
namespace sdk {
namespace core {

class FramePipeline {
   public:
    /// One slot of the ring: the frame and, when a fused cloud was requested, its XYZ /
    /// destaggered-range products.  A returned slot stays valid until the next slot is returned.
    struct Slot {
        LidarFrame frame;
        FusedCloud cloud;
    };

    /// depth: frames in flight (>= 1).  fused: optional prototype (lut + pixel_shift_by_row) copied
    /// into every slot; nullptr = decode only.
    FramePipeline(const std::shared_ptr<SensorInfo>& info, size_t depth = 3, const FusedCloud* fused = nullptr)
        : batcher_(info), depth_(depth < 1 ? 1 : depth) {
        batcher_.set_pipeline_depth(depth_ + 1);  // one more job than frames in flight: never blocks on reuse
        slots_.reserve(depth_ + 2);
        for (size_t i = 0; i < depth_ + 2; ++i) {  // 1 filling + depth in flight + 1 held by the caller
            slots_.emplace_back(new Slot{LidarFrame(info), FusedCloud{}});
            if (fused) {
                slots_.back()->cloud.lut = fused->lut;
                slots_.back()->cloud.lut_is_f64 = fused->lut_is_f64;
                slots_.back()->cloud.pixel_shift_by_row = fused->pixel_shift_by_row;
            }
            free_.push_back(i);
        }
        fused_ = fused != nullptr;
        next_fill();
    }
    FramePipeline(const SensorInfo& info, size_t depth = 3, const FusedCloud* fused = nullptr)
        : FramePipeline(std::make_shared<SensorInfo>(info), depth, fused) {}
    ~FramePipeline() {
        try {
            batcher_.wait_all();
        } catch (...) {
        }
    }
    FramePipeline(const FramePipeline&) = delete;
    FramePipeline& operator=(const FramePipeline&) = delete;

    FrameBatcher& batcher() { return batcher_; }
    size_t depth() const { return depth_; }
    size_t in_flight() const { return flight_.size(); }

    /// Feed one packet.  Returns a finished slot when this packet completed a frame and `depth`
    /// frames were already in flight, else nullptr.
    const Slot* push(const uint8_t* buf, size_t size, uint64_t host_timestamp) {
        if (!batcher_.batch(buf, size, host_timestamp, slots_[fill_]->frame)) return nullptr;
        return submitted();
    }
    const Slot* push(const Packet& packet) {
        if (!batcher_.batch(packet, slots_[fill_]->frame)) return nullptr;
        return submitted();
    }
    /// Burst form (see FrameBatcher::batch_burst): consumes packets until a frame completes.
    size_t push_burst(const uint8_t* packets, size_t n, size_t stride, size_t size,
                      const uint64_t* host_timestamps, const Slot** out) {
        bool complete = false;
        const size_t used =
            batcher_.batch_burst(packets, n, stride, size, host_timestamps, slots_[fill_]->frame, complete);
        *out = complete ? submitted() : nullptr;
        return used;
    }
    /// Oldest frame in flight, waited for; nullptr when nothing is in flight.  The partially
    /// batched frame (if any) is not flushed -- use batcher().flush() on fill_slot() for that.
    const Slot* drain() {
        if (flight_.empty()) return nullptr;
        return retire();
    }
    Slot& fill_slot() { return *slots_[fill_]; }

   private:
    const Slot* submitted() {
        flight_.push_back(fill_);
        const Slot* done = flight_.size() > depth_ ? retire() : nullptr;  // frees the slot held so far
        next_fill();
        return done;
    }
    const Slot* retire() {
        const size_t i = flight_.front();
        flight_.pop_front();
        batcher_.wait(slots_[i]->frame);
        if (held_ != kNone) free_.push_back(held_);
        held_ = i;
        return slots_[i].get();
    }
    void next_fill() {
        fill_ = free_.front();
        free_.pop_front();
        slots_[fill_]->frame.frame_id = -1;  // a fresh frame for the batcher (lidar_frame.cpp:1700-1703)
        if (fused_) batcher_.set_fused_cloud(&slots_[fill_]->cloud);
    }

    static constexpr size_t kNone = static_cast<size_t>(-1);
    FrameBatcher batcher_;
    size_t depth_;
    bool fused_{false};
    std::vector<std::unique_ptr<Slot>> slots_;
    std::deque<size_t> free_, flight_;
    size_t fill_{0};
    size_t held_{kNone};
};

}  // namespace core
}  // namespace sdk
}  // namespace ouster
