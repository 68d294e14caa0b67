// pose_util.h -- per-column pose application to point clouds
// (mirrors ouster_core/include/ouster/core/pose_util.h:24-160).  The loops run on the
// GPU (ob_dewarp).
#pragma once
#include <cstring>
#include <stdexcept>
#include <vector>

#include "ouster/core/b200_runtime.h"
#include "ouster/core/frame_set.h"
#include "ouster/core/lidar_frame.h"
#include "ouster/core/typedefs.h"
#include "ouster/core/xyzlut.h"

namespace ouster {
namespace sdk {
namespace core {

template <typename T>
using MatrixX16R = DenseArray<T>;  ///< W x 16: one flattened row-major 4x4 pose per row
using Poses = MatrixX16R<double>;

/// dewarp(dewarped, points, poses): dewarped[i*W + w] = R_w * points[i*W + w] + t_w
/// (pose_util.h:37-59).  points/dewarped are (H*W) x 3, poses W x 16.
template <typename T>
void dewarp(ArrayRef<T> dewarped, const ArrayRef<const T>& points, const ArrayRef<const T>& poses) {
    if (poses.cols() != 16 || points.cols() != 3 || dewarped.rows() != points.rows() ||
        poses.rows() == 0 || points.rows() % poses.rows() != 0)
        throw std::runtime_error("Number of points per set must match number of poses");
    b200::check(ob_dewarp(impl::lut_dtype<T>(), points.data(), poses.data(), points.rows(), poses.rows(),
                          dewarped.data(), b200::thread_stream()));
    b200::synchronize();
}

template <typename T>
PointCloudXYZ<T> dewarp(const PointCloudXYZ<T>& points, const MatrixX16R<T>& poses) {
    PointCloudXYZ<T> out(points.rows(), points.cols());
    dewarp<T>(ArrayRef<T>(out), ArrayRef<const T>(points), ArrayRef<const T>(poses));
    return out;
}

/// dewarp(lut, range, poses) == dewarp(lut(range), poses) in ONE pass over memory (B200 extension:
/// the fusion the reference asks for in impl/dewarp_impl.h:27-29): point (row, col) of the staggered
/// cloud becomes R_col * p + t_col.  range is H x W, poses W x 16.
template <typename T>
PointCloudXYZ<T> dewarp(const XYZLutT<T>& lut, const ArrayRef<const uint32_t>& range,
                        const MatrixX16R<T>& poses) {
    const size_t n = static_cast<size_t>(lut.h) * lut.w;
    if (static_cast<size_t>(range.rows()) * range.cols() != n)
        throw std::invalid_argument("unexpected image dimensions");
    if (poses.cols() != 16 || static_cast<size_t>(poses.rows()) != lut.w)
        throw std::runtime_error("Number of points per set must match number of poses");
    PointCloudXYZ<T> out(n, 3);
    ob_cloud_io io{};
    io.n_frames = 1;
    io.n_returns = 1;
    io.range = range.data();
    io.xyz = out.data();
    io.poses = poses.data();
    b200::check(ob_scan_to_cloud(lut.device_lut().get(), nullptr, 0, &io, b200::thread_stream()));
    b200::synchronize();
    return out;
}

/// dewarp(lidar_frame, xyzlut, min_range, max_range) (pose_util.h:456-485, impl/dewarp_impl.h:22-76):
/// world-frame points of the first return, columns between the first and last valid one in order,
/// status == 0 columns skipped, min_range <= r <= max_range (metres), each point posed with its
/// column's body_to_world.  Returns an n x 3 array (the reference's std::vector<Eigen::Vector3<T>>
/// has the same memory layout).  One fused GPU pass (ob_dewarp_frame).  Optional per-point
/// provenance like impl::dewarp_impl: column index and column timestamp.
template <typename T>
PointCloudXYZ<T> dewarp(const LidarFrame& lidar_frame, const XYZLutT<T>& xyzlut, double min_range,
                        double max_range, std::vector<uint32_t>* col_idxs = nullptr,
                        std::vector<uint64_t>* timestamps_ns = nullptr) {
    auto range = lidar_frame.field<uint32_t>(ChanField::RANGE);
    const size_t n = static_cast<size_t>(xyzlut.h) * xyzlut.w;
    if (static_cast<size_t>(range.rows()) * range.cols() != n || lidar_frame.w != xyzlut.w)
        throw std::invalid_argument("unexpected image dimensions");
    PointCloudXYZ<T> all(n, 3);
    std::vector<uint32_t> ci(col_idxs ? n : 0);
    std::vector<uint64_t> ts(timestamps_ns ? n : 0);
    ob_dewarp_frame_io io{};
    io.range = range.data();
    io.poses = lidar_frame.body_to_world().template get<double>();
    io.status = lidar_frame.status().data();
    io.timestamps = lidar_frame.timestamp().data();
    io.min_range = min_range;
    io.max_range = max_range;
    io.points = all.data();
    io.col_idx = col_idxs ? ci.data() : nullptr;
    io.timestamps_out = timestamps_ns ? ts.data() : nullptr;
    io.capacity = n;
    size_t count = 0;
    b200::check(ob_dewarp_frame(xyzlut.device_lut().get(), &io, &count, b200::thread_stream()));
    PointCloudXYZ<T> out(count, 3);
    if (count) std::memcpy(out.data(), all.data(), count * 3 * sizeof(T));
    if (col_idxs) col_idxs->insert(col_idxs->end(), ci.begin(), ci.begin() + static_cast<std::ptrdiff_t>(count));
    if (timestamps_ns)
        timestamps_ns->insert(timestamps_ns->end(), ts.begin(), ts.begin() + static_cast<std::ptrdiff_t>(count));
    return out;
}

/// dewarp(frame_set, xyzluts, min_range, max_range) (pose_util.h:475, impl/dewarp_impl.h:84-117): the
/// dewarped points of every frame of the set, frame after frame, each frame with its own LUT.  One batched
/// GPU pass for the whole set (ob_dewarp_frames: one launch).  Optional per-point
/// provenance like impl::dewarp_impl: frame index, column index, column timestamp (appended).
template <typename T>
PointCloudXYZ<T> dewarp(const FrameSet& frame_set, const std::vector<XYZLutT<T>>& xyzluts, double min_range,
                        double max_range, std::vector<uint32_t>* frame_idxs = nullptr,
                        std::vector<uint32_t>* col_idxs = nullptr, std::vector<uint64_t>* timestamps_ns = nullptr) {
    if (frame_set.size() != xyzluts.size())
        throw std::invalid_argument("Number of frames and number of XYZLuts must be the same");
    std::vector<ob_dewarp_frames_io> ios(frame_set.size());
    size_t cap = 0;
    for (size_t idx : frame_set.valid_indices()) {
        const LidarFrame& f = *frame_set[idx];
        const XYZLutT<T>& lut = xyzluts[idx];
        auto range = f.field<uint32_t>(ChanField::RANGE);
        const size_t n = static_cast<size_t>(lut.h) * lut.w;
        if (static_cast<size_t>(range.rows()) * range.cols() != n || f.w != lut.w)
            throw std::invalid_argument("unexpected image dimensions");
        ios[idx].lut = lut.device_lut().get();
        ios[idx].range = range.data();
        ios[idx].poses = f.body_to_world().template get<double>();
        ios[idx].status = f.status().data();
        ios[idx].timestamps = f.timestamp().data();
        cap += n;
    }
    PointCloudXYZ<T> all(cap, 3);
    const bool prov = frame_idxs || col_idxs || timestamps_ns;
    std::vector<uint32_t> fi(frame_idxs ? cap : 0), ci(col_idxs ? cap : 0);
    std::vector<uint64_t> ts(timestamps_ns ? cap : 0);
    (void)prov;
    size_t count = 0;
    if (cap)
        b200::check(ob_dewarp_frames(ios.data(), ios.size(), min_range, max_range, all.data(), cap,
                                     frame_idxs ? fi.data() : nullptr, col_idxs ? ci.data() : nullptr,
                                     timestamps_ns ? ts.data() : nullptr, nullptr, &count, b200::thread_stream()));
    PointCloudXYZ<T> out(count, 3);
    if (count) std::memcpy(out.data(), all.data(), count * 3 * sizeof(T));
    const auto n = static_cast<std::ptrdiff_t>(count);
    if (frame_idxs) frame_idxs->insert(frame_idxs->end(), fi.begin(), fi.begin() + n);
    if (col_idxs) col_idxs->insert(col_idxs->end(), ci.begin(), ci.begin() + n);
    if (timestamps_ns) timestamps_ns->insert(timestamps_ns->end(), ts.begin(), ts.begin() + n);
    return out;
}

/// transform(points, pose): one 4x4 pose (16 values, row-major) for every point (pose_util.h:118-160).
template <typename T>
PointCloudXYZ<T> transform(const PointCloudXYZ<T>& points, const T* pose16) {
    PointCloudXYZ<T> out(points.rows(), points.cols());
    if (points.rows() == 0) return out;
    b200::check(ob_dewarp(impl::lut_dtype<T>(), points.data(), pose16, points.rows(), 1, out.data(),
                          b200::thread_stream()));
    b200::synchronize();
    return out;
}

}  // namespace core
}  // namespace sdk
}  // namespace ouster
