// impl/cartesian.h -- range image -> Cartesian points through the LUT
// (mirrors ouster_core/include/ouster/core/impl/cartesian.h:36-108; the loop itself runs on the
// GPU: ob_lut_create + ob_cartesian).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <type_traits>

#include "ouster/core/b200_runtime.h"
#include "ouster/core/lidar_frame.h"
#include "ouster/core/typedefs.h"

namespace ouster {
namespace sdk {
namespace core {
namespace impl {

template <typename T>
constexpr ob_dtype lut_dtype() {
    static_assert(std::is_same<T, float>::value || std::is_same<T, double>::value,
                  "XYZ LUTs are float or double");
    return std::is_same<T, double>::value ? OB_F64 : OB_F32;
}

/// RAII handle of a device LUT built from host direction/offset tables.
template <typename T>
struct TransientLut {
    ob_lut* h{nullptr};
    TransientLut(const ArrayX3R<T>& direction, const ArrayX3R<T>& offset, size_t rows) {
        // rows x 1 geometry: the projection is per pixel, the image shape is irrelevant here
        b200::check(ob_lut_create(lut_dtype<T>(), direction.data(), offset.data(), rows, 1,
                                  b200::device(), &h));
    }
    ~TransientLut() { ob_lut_destroy(h); }
};

/// cartesianT(points, range, direction, offset): points must be pre-allocated (n x 3).
template <typename T>
void cartesianT(ArrayRef<T> points, const ArrayRef<const uint32_t>& range,
                const ArrayX3R<T>& direction, const ArrayX3R<T>& offset) {
    // the reference only asserts here (impl/cartesian.h:39-41); mismatches are reported
    if (points.rows() != direction.rows() || points.rows() != offset.rows() ||
        points.rows() != range.size())
        throw std::invalid_argument("unexpected image dimensions");
    TransientLut<T> lut(direction, offset, direction.rows());
    b200::check(ob_cartesian(lut.h, range.data(), range.size(), points.data(), b200::thread_stream()));
    b200::synchronize();
}

/// cartesianT(range, direction, offset) -> points; throws "unexpected image dimensions".
template <typename T>
PointCloudXYZ<T> cartesianT(const ArrayRef<const uint32_t>& range, const ArrayX3R<T>& direction,
                            const ArrayX3R<T>& offset) {
    if (range.cols() * range.rows() != direction.rows())
        throw std::invalid_argument("unexpected image dimensions");
    PointCloudXYZ<T> points(direction.rows(), 3);
    cartesianT<T>(ArrayRef<T>(points), range, direction, offset);
    return points;
}

template <typename T>
PointCloudXYZ<T> cartesianT(const LidarFrame& frame, const ArrayX3R<T>& direction,
                            const ArrayX3R<T>& offset) {
    return cartesianT<T>(frame.field<uint32_t>(ChanField::RANGE), direction, offset);
}

}  // namespace impl
}  // namespace core
}  // namespace sdk
}  // namespace ouster
