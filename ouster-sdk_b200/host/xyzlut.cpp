// xyzlut.cpp -- host glue of the LUT / destagger / frame_to_packets entry points and the
// per-thread runtime (device, stream) behind the C++ replacement headers.
#include <cstdlib>
#include <mutex>

#include "ouster/core/b200_runtime.h"
#include "ouster/core/lidar_frame.h"
#include "ouster/core/xyzlut.h"

namespace ouster {
namespace sdk {
namespace core {

// ---- runtime ----
namespace b200 {
namespace {
int initial_device() {
    const char* e = std::getenv("OUSTER_B200_DEVICE");
    return (e && *e) ? std::atoi(e) : 0;
}
thread_local int t_device = initial_device();
struct StreamHolder {
    ob_stream* s{nullptr};
    int device{-1};
    ~StreamHolder() {
        if (s) ob_stream_destroy(s);
    }
};
thread_local StreamHolder t_stream;
}  // namespace

void set_device(int device) { t_device = device; }
int device() { return t_device; }

ob_stream* thread_stream() {
    if (t_stream.s == nullptr || t_stream.device != t_device) {
        if (t_stream.s) ob_stream_destroy(t_stream.s);
        t_stream.s = nullptr;
        check(ob_stream_create(t_device, &t_stream.s));
        t_stream.device = t_device;
    }
    return t_stream.s;
}

void synchronize() { check(ob_stream_sync(thread_stream())); }
}  // namespace b200

// ---- LUT ----
namespace impl {

mat4d lut_transform(const SensorInfo& sensor, bool use_extrinsics) {
    if (!use_extrinsics) return sensor.lidar_to_sensor_transform;
    // extrinsics are applied after lidar_to_sensor; their translation is in metres
    mat4d ext = sensor.sensor_to_body;
    for (int r = 0; r < 3; ++r) ext(r, 3) /= RANGE_UNIT;
    return ext * sensor.lidar_to_sensor_transform;
}

XYZLut make_xyz_lut(size_t w, size_t h, double range_unit, const mat4d& beam_to_lidar_transform,
                    const mat4d& transform, const std::vector<double>& azimuth_angles_deg,
                    const std::vector<double>& altitude_angles_deg) {
    return XYZLut::from_intrinsics(w, h, range_unit, beam_to_lidar_transform, transform,
                                   azimuth_angles_deg, altitude_angles_deg);
}

XYZLut make_xyz_lut(const SensorInfo& sensor, bool use_extrinsics) {
    return make_xyz_lut(sensor.format.columns_per_frame, sensor.format.pixels_per_column, RANGE_UNIT,
                        sensor.beam_to_lidar_transform, lut_transform(sensor, use_extrinsics),
                        sensor.beam_azimuth_angles, sensor.beam_altitude_angles);
}

// ---- destagger ----
void destagger_raw(size_t elem_size, size_t k, const void* img, const std::vector<int>& shifts,
                   size_t h, size_t w, bool inverse, void* out) {
    static_assert(sizeof(int) == sizeof(int32_t), "pixel_shift_by_row is int32");
    b200::check(ob_destagger(elem_size, k, img, reinterpret_cast<const int32_t*>(shifts.data()),
                             shifts.size(), h, w, inverse ? 1 : 0, out, b200::thread_stream()));
    b200::synchronize();
}

void check_resolution(const SensorInfo& info, size_t rows, size_t cols) {
    if (rows != info.format.pixels_per_column || cols != info.format.columns_per_frame ||
        rows != info.format.pixel_shift_by_row.size())
        throw std::invalid_argument{"Image resolution must match SensorInfo."};
}

// ---- frame_to_packets (lidar packets) ----
std::vector<Packet> frame_to_packets(const LidarFrame& frame, const PacketFormat& pf,
                                     uint32_t init_id, uint64_t prod_sn) {
    const size_t cpp = static_cast<size_t>(pf.columns_per_packet);
    const size_t n_packets = frame.packet_timestamp().size();
    if (frame.w / cpp != n_packets)
        throw std::invalid_argument(
            "Mismatch between expected number of packets and PacketFormat.columns_per_packet");
    std::vector<Packet> out;
    out.reserve(n_packets);
    const bool with_crc =
        pf.udp_profile_lidar != UDPProfileLidar::LEGACY && pf.header_type == HeaderType::STANDARD;
    for (size_t pid = 0; pid < n_packets; ++pid) {
        LidarPacket pkt(pf.lidar_packet_size + 8);  // +8: FieldDecodeInfo::set touches 8 bytes
        uint8_t* b = pkt.buf.data();
        pkt.host_timestamp = frame.packet_timestamp()[pid];
        pf.set_shutdown(b, static_cast<uint8_t>(frame.thermal_shutdown()));
        pf.set_shot_limiting(b, static_cast<uint8_t>(frame.shot_limiting()));
        pf.set_shutdown_countdown(b, frame.shutdown_countdown);
        pf.set_shot_limiting_countdown(b, frame.shot_limiting_countdown);
        pf.set_frame_id(b, static_cast<uint32_t>(frame.frame_id));
        pf.set_init_id(b, init_id);
        pf.set_prod_sn(b, prod_sn);
        pf.set_packet_type(b, 0x1);
        pf.set_alert_flags(b, frame.alert_flags()[pid]);
        bool any_valid = false;
        for (size_t c = 0; c < cpp; ++c) {
            uint8_t* col = pf.nth_col(c, b);
            const size_t id = pid * cpp + c;
            pf.set_col_status(col, frame.status()[id]);
            pf.set_col_measurement_id(col, static_cast<uint16_t>(id));
            pf.set_col_timestamp(col, frame.timestamp()[id]);
            any_valid = any_valid || (frame.status()[id] & 0x01);
        }
        if (!any_valid && pkt.host_timestamp == 0) continue;  // nothing to send
        for (auto it = pf.begin(); it != pf.end(); ++it) {
            const std::string& name = it->first;
            if (!frame.has_field(name)) continue;
            const Field& fld = frame.field(name);
            size_t elem = fld.element_size();
            for (size_t d = 2; d < fld.shape().size(); ++d) elem *= fld.shape()[d];
            const int cols = static_cast<int>(frame.w);
            switch (elem) {
                case 1: pf.set_block(fld.get<uint8_t>(), cols, name, b); break;
                case 2: pf.set_block(fld.get<uint16_t>(), cols, name, b); break;
                case 4: pf.set_block(fld.get<uint32_t>(), cols, name, b); break;
                case 8: pf.set_block(fld.get<uint64_t>(), cols, name, b); break;
                case 6: pf.set_block(fld.get<impl::float3x16_t>(), cols, name, b); break;
                default: throw std::invalid_argument("unsupported field element size");
            }
        }
        pkt.buf.resize(pf.lidar_packet_size);
        if (with_crc) {
            const uint64_t crc = pf.calculate_crc(pkt.buf.data(), pkt.buf.size());
            std::memcpy(pkt.buf.data() + pkt.buf.size() - sizeof(crc), &crc, sizeof(crc));
        }
        out.push_back(std::move(pkt));
    }
    return out;
}

// ---- frame_to_packets on the device (K4) ----
std::vector<Packet> frame_to_packets_device(const LidarFrame& frame, const PacketFormat& pf, uint32_t init_id,
                                            uint64_t prod_sn) {
    const size_t cpp = static_cast<size_t>(pf.columns_per_packet);
    const size_t n_packets = frame.packet_timestamp().size();
    if (frame.w / cpp != n_packets)
        throw std::invalid_argument(
            "Mismatch between expected number of packets and PacketFormat.columns_per_packet");
    const bool legacy = pf.udp_profile_lidar == UDPProfileLidar::LEGACY;
    const bool with_crc = !legacy && pf.header_type == HeaderType::STANDARD;
    const size_t psz = pf.lidar_packet_size;
    // packet-level headers by the reference's setters; LEGACY keeps frame-level words inside its column
    // headers, so its template is the whole packet
    const size_t hb = legacy ? psz : pf.packet_header_size;
    std::vector<uint8_t> headers(n_packets * hb + 16, 0);
    std::vector<uint8_t> tmp(psz + 16);
    std::vector<bool> emit(n_packets, false);
    for (size_t pid = 0; pid < n_packets; ++pid) {
        std::fill(tmp.begin(), tmp.end(), 0);
        uint8_t* b = tmp.data();
        pf.set_shutdown(b, static_cast<uint8_t>(frame.thermal_shutdown()));
        pf.set_shot_limiting(b, static_cast<uint8_t>(frame.shot_limiting()));
        pf.set_shutdown_countdown(b, frame.shutdown_countdown);
        pf.set_shot_limiting_countdown(b, frame.shot_limiting_countdown);
        pf.set_frame_id(b, static_cast<uint32_t>(frame.frame_id));
        pf.set_init_id(b, init_id);
        pf.set_prod_sn(b, prod_sn);
        pf.set_packet_type(b, 0x1);
        pf.set_alert_flags(b, frame.alert_flags()[pid]);
        bool any_valid = false;
        for (size_t c = 0; c < cpp; ++c) {
            const size_t id = pid * cpp + c;
            if (legacy) {
                uint8_t* col = pf.nth_col(static_cast<int>(c), b);
                pf.set_col_status(col, frame.status()[id]);
                pf.set_col_measurement_id(col, static_cast<uint16_t>(id));
                pf.set_col_timestamp(col, frame.timestamp()[id]);
            }
            any_valid = any_valid || (frame.status()[id] & 0x01);
        }
        emit[pid] = any_valid || frame.packet_timestamp()[pid] != 0;
        std::memcpy(headers.data() + pid * hb, b, hb);
    }
    // decoder table: the fields the frame shares with the profile, profile order (foreach_channel_field)
    std::vector<ob_field_desc> descs;
    std::vector<const void*> srcs;
    for (auto it = pf.begin(); it != pf.end(); ++it) {
        const std::string& name = it->first;
        if (!frame.has_field(name) || name == ChanField::RAW_HEADERS) continue;
        const Field& fld = frame.field(name);
        const FieldDecodeInfo& info = pf.field_decode_info(name);
        size_t elem = fld.element_size();
        for (size_t d = 2; d < fld.shape().size(); ++d) elem *= fld.shape()[d];
        ob_field_desc d{};
        d.offset = static_cast<uint32_t>(info.offset);
        d.elem_size = static_cast<uint32_t>(elem);
        d.mask = info.mask;
        d.shift = info.shift;
        d.range_return = -1;
        descs.push_back(d);
        srcs.push_back(fld.get());
    }
    if (descs.size() > OB_MAX_FIELDS) throw std::invalid_argument("too many fields to encode");
    auto conv = [](const FieldDecodeInfo& fi) {
        ob_field_desc d{};
        d.offset = static_cast<uint32_t>(fi.offset);
        d.elem_size = 8;
        d.mask = fi.mask;
        d.shift = fi.shift;
        d.range_return = -1;
        return d;
    };
    ob_packet_layout L{};
    L.packet_header_size = static_cast<uint32_t>(pf.packet_header_size);
    L.col_header_size = static_cast<uint32_t>(pf.col_header_size);
    L.channel_data_size = static_cast<uint32_t>(pf.channel_data_size);
    L.col_size = static_cast<uint32_t>(pf.col_size);
    L.packet_size = static_cast<uint32_t>(psz);
    L.columns_per_packet = static_cast<uint32_t>(pf.columns_per_packet);
    L.pixels_per_column = static_cast<uint32_t>(pf.pixels_per_column);
    L.columns_per_frame = static_cast<uint32_t>(frame.w);
    L.col_timestamp = conv(pf.col_timestamp_info());
    L.col_measurement_id = conv(pf.col_measurement_id_info());
    L.col_status = conv(pf.col_status_info());
    ob_decoder* dec = nullptr;
    b200::check(ob_decoder_create(&L, descs.data(), descs.size(), b200::device(), &dec));
    std::shared_ptr<ob_decoder> guard_dec(dec, [](ob_decoder* d) { ob_decoder_destroy(d); });
    std::vector<uint8_t> wire(n_packets * psz);
    ob_encode_io io{};
    for (size_t k = 0; k < srcs.size(); ++k) io.fields[k] = srcs[k];
    io.timestamp = frame.timestamp().data();
    io.status = frame.status().data();
    io.packet_headers = headers.data();
    io.packet_header_bytes = hb;
    io.packets = wire.data();
    io.packet_stride = psz;
    b200::check(ob_encode_frames(dec, &io, 1, with_crc ? 1 : 0, b200::thread_stream()));
    b200::synchronize();
    std::vector<Packet> out;
    out.reserve(n_packets);
    for (size_t pid = 0; pid < n_packets; ++pid) {
        if (!emit[pid]) continue;  // nothing to send (lidar_frame_impl.h:497-500)
        LidarPacket pkt(psz);
        std::memcpy(pkt.buf.data(), wire.data() + pid * psz, psz);
        pkt.host_timestamp = frame.packet_timestamp()[pid];
        out.push_back(std::move(pkt));
    }
    return out;
}

}  // namespace impl

Field destagger(const SensorInfo& info, const Field& field, bool inverse) {
    Field result(field.tag(), field.shape());
    const auto& shp = field.shape();
    switch (field.tag()) {  // visit_field_2d: the arithmetic types only
        case ChanFieldType::UINT8: case ChanFieldType::UINT16: case ChanFieldType::UINT32:
        case ChanFieldType::UINT64: case ChanFieldType::INT8: case ChanFieldType::INT16:
        case ChanFieldType::INT32: case ChanFieldType::INT64: case ChanFieldType::FLOAT32:
        case ChanFieldType::FLOAT64:
            break;
        default:
            return result;  // FLOAT16 / CHAR / ZONE_STATE: silently zero-filled in the reference
    }
    if (shp.size() != 2) return result;
    if (info.format.pixel_shift_by_row.size() != shp[0])
        throw std::invalid_argument{"image height does not match shifts size"};
    impl::destagger_raw(field.element_size(), 1, field.get(), info.format.pixel_shift_by_row, shp[0],
                        shp[1], inverse, result.get());
    return result;
}

uint64_t column_timestamp_at_destaggered_pixel(size_t row, size_t col,
                                               const std::vector<int>& pixel_shift_by_row,
                                               const HeaderRef<const uint64_t>& column_timestamps) {
    const size_t width = column_timestamps.size();
    if (row >= pixel_shift_by_row.size() || col >= width)
        throw std::invalid_argument("row or column is out of range");
    const int w = static_cast<int>(width);
    const int offset = (w + pixel_shift_by_row[row] % w) % w;  // int arithmetic, as the reference
    const int staggered_col = (static_cast<int>(col) - offset + w) % w;
    return column_timestamps[static_cast<size_t>(staggered_col)];
}

PointCloudXYZd cartesian(const ArrayRef<const uint32_t>& range, const XYZLut& lut) {
    if (range.cols() * range.rows() != lut.direction.rows())
        throw std::invalid_argument("unexpected image dimensions");
    return lut(range);
}

PointCloudXYZd cartesian(const LidarFrame& frame, const XYZLut& lut) {
    return cartesian(frame.field<uint32_t>(ChanField::RANGE), lut);
}

}  // namespace core
}  // namespace sdk
}  // namespace ouster
