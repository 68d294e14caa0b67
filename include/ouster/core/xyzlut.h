// xyzlut.h -- XYZLutT<T>: per-pixel direction/offset lookup table and range -> XYZ projection
// (mirrors ouster_core/include/ouster/core/xyzlut.h:54-189).
//
// The tables live on the GPU (built there from the beam intrinsics); `direction`/`offset` are
// host copies kept for API compatibility (public const members in the reference).
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <vector>

#include "ouster/core/b200_runtime.h"
#include "ouster/core/impl/cartesian.h"
#include "ouster/core/lidar_frame.h"
#include "ouster/core/sensor_info.h"
#include "ouster/core/typedefs.h"

namespace ouster {
namespace sdk {
namespace core {

template <typename T>
class XYZLutT;
using XYZLut = XYZLutT<double>;
using XYZLutFloat = XYZLutT<float>;

namespace impl {
/// make_xyz_lut(w, h, range_unit, beam_to_lidar, transform, az, alt) -- xyzlut.cpp:11-89.
/// Throws std::invalid_argument("lut dimensions must be greater than zero") /
/// ("unexpected frame dimensions").
OUSTER_API_FUNCTION XYZLut make_xyz_lut(size_t w, size_t h, double range_unit,
                                        const mat4d& beam_to_lidar_transform, const mat4d& transform,
                                        const std::vector<double>& azimuth_angles_deg,
                                        const std::vector<double>& altitude_angles_deg);
/// make_xyz_lut(sensor, use_extrinsics) -- xyzlut.cpp:91-106.
OUSTER_API_FUNCTION XYZLut make_xyz_lut(const SensorInfo& sensor, bool use_extrinsics);
OUSTER_API_FUNCTION mat4d lut_transform(const SensorInfo& sensor, bool use_extrinsics);
}  // namespace impl

template <typename T>
class XYZLutT {
   public:
    const ArrayX3R<T> direction;
    const ArrayX3R<T> offset;
    const size_t h = 0;
    const size_t w = 0;

    template <typename>
    friend class XYZLutT;

    XYZLutT() = default;

    /// From sensor metadata: tables computed on the device in double, cast to T (xyzlut.h:111-124).
    XYZLutT(const SensorInfo& sensor, bool use_extrinsics = true)
        : XYZLutT(from_intrinsics(sensor.format.columns_per_frame, sensor.format.pixels_per_column,
                                  RANGE_UNIT, sensor.beam_to_lidar_transform,
                                  impl::lut_transform(sensor, use_extrinsics),
                                  sensor.beam_azimuth_angles, sensor.beam_altitude_angles)) {}

    /// Converting constructor (element-wise cast of the tables).
    template <typename OldT>
    XYZLutT(const XYZLutT<OldT>& other)
        : direction(other.direction.template cast<T>()),
          offset(other.offset.template cast<T>()),
          h(other.h),
          w(other.w) {
        upload();
    }

    XYZLutT(ArrayX3R<T> direction_, ArrayX3R<T> offset_, size_t h_, size_t w_)
        : direction(std::move(direction_)), offset(std::move(offset_)), h(h_), w(w_) {
        upload();
    }

    static XYZLutT from_intrinsics(size_t w, size_t h, double range_unit, const mat4d& b2l,
                                   const mat4d& transform, const std::vector<double>& az,
                                   const std::vector<double>& alt) {
        ob_lut* raw = nullptr;
        b200::check(ob_lut_from_intrinsics(impl::lut_dtype<T>(), w, h, range_unit, b2l.data(),
                                           transform.data(), az.data(), az.size(), alt.data(),
                                           alt.size(), b200::device(), &raw));
        std::shared_ptr<ob_lut> dev(raw, [](ob_lut* p) { ob_lut_destroy(p); });
        ArrayX3R<T> d(w * h, 3), o(w * h, 3);
        b200::check(ob_lut_download(raw, d.data(), o.data()));
        return XYZLutT(std::move(d), std::move(o), h, w, std::move(dev));
    }

    /// lut(range): no dimension check beyond the pixel count, like the reference operator()
    /// (xyzlut.h:139-143); the count mismatch surfaces as "unexpected image dimensions".
    PointCloudXYZ<T> operator()(const ArrayRef<const uint32_t>& range) const {
        PointCloudXYZ<T> points(range.rows() * range.cols(), 3);
        project(range.data(), range.size(), points.data());
        return points;
    }
    PointCloudXYZ<T> operator()(const LidarFrame& frame) const {
        return (*this)(frame.field<uint32_t>(ChanField::RANGE));
    }

    /// Device handle (shared with FusedCloud / ob_scan_to_cloud callers).
    const std::shared_ptr<ob_lut>& device_lut() const { return dev_; }

   private:
    XYZLutT(ArrayX3R<T> d, ArrayX3R<T> o, size_t h_, size_t w_, std::shared_ptr<ob_lut> dev)
        : direction(std::move(d)), offset(std::move(o)), h(h_), w(w_), dev_(std::move(dev)) {}

    void upload() {
        if (h == 0 || w == 0) return;
        if (direction.rows() != h * w || offset.rows() != h * w)
            throw std::invalid_argument("unexpected image dimensions");
        ob_lut* raw = nullptr;
        b200::check(ob_lut_create(impl::lut_dtype<T>(), direction.data(), offset.data(), h, w,
                                  b200::device(), &raw));
        dev_ = std::shared_ptr<ob_lut>(raw, [](ob_lut* p) { ob_lut_destroy(p); });
    }
    void project(const uint32_t* range, size_t n, T* points) const {
        if (!dev_) throw std::invalid_argument("unexpected image dimensions");
        b200::check(ob_cartesian(dev_.get(), range, n, points, b200::thread_stream()));
        b200::synchronize();
    }
    std::shared_ptr<ob_lut> dev_;
};

/// cartesian(frame, lut) / cartesian(range, lut): deprecated free functions kept by the reference
/// (xyzlut.h:173-189, xyzlut.cpp:111-124); throw std::invalid_argument("unexpected image dimensions").
OUSTER_API_FUNCTION PointCloudXYZd cartesian(const LidarFrame& frame, const XYZLut& lut);
OUSTER_API_FUNCTION PointCloudXYZd cartesian(const ArrayRef<const uint32_t>& range, const XYZLut& lut);

}  // namespace core
}  // namespace sdk
}  // namespace ouster
