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
