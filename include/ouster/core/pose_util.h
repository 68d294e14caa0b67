// pose_util.h -- per-column pose application to point clouds
// (mirrors ouster_core/include/ouster/core/pose_util.h:24-160).  The loops run on the
// GPU (ob_dewarp).
#pragma once
#include <stdexcept>

#include "ouster/core/b200_runtime.h"
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
