// impl/lidar_frame_impl.h -- destagger/stagger templates and frame_to_packets
// (mirrors ouster_core/include/ouster/core/impl/lidar_frame_impl.h:435-989).
// The row rotation itself runs on the GPU (ob_destagger).
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "ouster/core/b200_runtime.h"
#include "ouster/core/typedefs.h"

namespace ouster {
namespace sdk {
namespace core {

class LidarFrame;
class PacketFormat;
struct Packet;
struct SensorInfo;

namespace impl {
OUSTER_API_FUNCTION void destagger_raw(size_t elem_size, size_t k, const void* img,
                                       const std::vector<int>& pixel_shift_by_row, size_t h, size_t w,
                                       bool inverse, void* out);
OUSTER_API_FUNCTION void check_resolution(const SensorInfo& info, size_t rows, size_t cols);

/// LidarFrame -> lidar packets (inverse of batching), impl/lidar_frame_impl.h:435-531.
/// Host-side; used to synthesise packet streams exactly the way the reference's tests do.
OUSTER_API_FUNCTION std::vector<Packet> frame_to_packets(const LidarFrame& frame,
                                                         const PacketFormat& pf, uint32_t init_id,
                                                         uint64_t prod_sn);
template <typename OutputItT>
void frame_to_packets(const LidarFrame& frame, std::shared_ptr<PacketFormat> pf, OutputItT iter,
                      uint32_t init_id, uint64_t prod_sn);
/// The same packets produced on the GPU (K4, ob_encode_frames): the host only writes the packet-level
/// headers with the PacketFormat setters; column headers, set_block of every field and the CRC64 run
/// in one launch.  Byte-identical to frame_to_packets().
OUSTER_API_FUNCTION std::vector<Packet> frame_to_packets_device(const LidarFrame& frame,
                                                                const PacketFormat& pf, uint32_t init_id,
                                                                uint64_t prod_sn);
}  // namespace impl

/// destagger_into(img, shifts, inverse, destaggered) -- :733-760.
/// Throws "image height does not match shifts size" / "image and destaggered must have the same shape".
template <typename T>
inline void destagger_into(const ArrayRef<const T>& img, const std::vector<int>& pixel_shift_by_row,
                           bool inverse, ArrayRef<T> destaggered) {
    const size_t h = img.rows(), w = img.cols();
    if (pixel_shift_by_row.size() != h)
        throw std::invalid_argument{"image height does not match shifts size"};
    if (h != destaggered.rows() || w != destaggered.cols())
        throw std::invalid_argument{"image and destaggered must have the same shape"};
    impl::destagger_raw(sizeof(T), 1, img.data(), pixel_shift_by_row, h, w, inverse,
                        destaggered.data());
}

/// N-D form: img is h x w x k (k = product of trailing dims) -- :776-811.
template <typename T>
inline void destagger_into(const T* img, size_t h, size_t w, size_t k,
                           const std::vector<int>& pixel_shift_by_row, bool inverse, T* destaggered) {
    if (pixel_shift_by_row.size() != h)
        throw std::invalid_argument{"image height does not match shifts size"};
    impl::destagger_raw(sizeof(T), k, img, pixel_shift_by_row, h, w, inverse, destaggered);
}

template <typename T>
inline img_t<T> destagger(const ArrayRef<const T>& img, const std::vector<int>& pixel_shift_by_row,
                          bool inverse = false) {
    img_t<T> destaggered{img.rows(), img.cols()};
    destagger_into<T>(img, pixel_shift_by_row, inverse, ArrayRef<T>(destaggered));
    return destaggered;
}

/// SensorInfo forms -- :875-959; throw "Image resolution must match SensorInfo."
template <typename T>
inline void destagger_into(const SensorInfo& info, const ArrayRef<const T>& img, bool inverse,
                           ArrayRef<T> destaggered) {
    impl::check_resolution(info, img.rows(), img.cols());
    destagger_into<T>(img, info.format.pixel_shift_by_row, inverse, destaggered);
}
template <typename T>
inline img_t<T> destagger(const SensorInfo& info, const ArrayRef<const T>& img, bool inverse = false) {
    impl::check_resolution(info, img.rows(), img.cols());
    return destagger<T>(img, info.format.pixel_shift_by_row, inverse);
}
template <typename T>
inline img_t<T> stagger(const SensorInfo& info, const ArrayRef<const T>& img) {
    return destagger<T>(info, img, true);
}
// convenience overloads for owning arrays (Eigen::Ref converts implicitly in the reference)
template <typename T>
inline img_t<T> destagger(const img_t<T>& img, const std::vector<int>& pixel_shift_by_row,
                          bool inverse = false) {
    return destagger<T>(ArrayRef<const T>(img), pixel_shift_by_row, inverse);
}
template <typename T>
inline img_t<T> destagger(const SensorInfo& info, const img_t<T>& img, bool inverse = false) {
    return destagger<T>(info, ArrayRef<const T>(img), inverse);
}
template <typename T>
inline img_t<T> stagger(const SensorInfo& info, const img_t<T>& img) {
    return destagger<T>(info, ArrayRef<const T>(img), true);
}

}  // namespace core
}  // namespace sdk
}  // namespace ouster
