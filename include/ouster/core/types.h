// types.h -- PacketFormat: table-driven description of the lidar UDP packet layout
// (mirrors the hot-path part of ouster_core/include/ouster/core/types.h:109-1116).
//
// Host-side accessors decode headers (frame_id, column headers, ...) and single fields for
// tests/tools; the per-pixel decode of whole frames is done by the GPU from the same tables
// (PacketFormat::device_layout() / device_fields(), consumed by ob_decoder_create()).
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <memory>
#include <optional>
#include <string>
#include <utility>
#include <vector>

#include "ouster/core/chanfield.h"
#include "ouster/core/data_format.h"
#include "ouster/core/field_decode_info.h"
#include "ouster/core/sensor_info.h"
#include "ouster/core/visibility.h"

namespace ouster {
namespace sdk {
namespace core {

enum class ThermalShutdownStatus : uint8_t { NORMAL = 0x00, IMMINENT = 0x01 };
enum class ShotLimitingStatus : uint8_t { NORMAL = 0x00, IMMINENT = 0x01, REDUCTION_0_10 = 0x02 };

class OUSTER_API_CLASS PacketFormat {
   public:
    using FieldTypeEntry = std::pair<std::string, std::pair<ChanFieldType, int>>;
    using FieldIter = std::vector<FieldTypeEntry>::const_iterator;

    OUSTER_API_FUNCTION explicit PacketFormat(const DataFormat& format);
    OUSTER_API_FUNCTION explicit PacketFormat(const SensorInfo& info);

    UDPProfileLidar udp_profile_lidar;
    UDPProfileIMU udp_profile_imu;
    HeaderType header_type;
    size_t lidar_packet_size;
    int columns_per_packet;
    int pixels_per_column;
    size_t packet_header_size;
    size_t col_header_size;
    size_t col_footer_size;
    size_t col_size;
    size_t packet_footer_size;
    size_t channel_data_size;
    uint32_t max_frame_id;

    // ---- packet headers (parsing.cpp:736-791) ----
    OUSTER_API_FUNCTION uint16_t packet_type(const uint8_t* packet_buf) const;
    OUSTER_API_FUNCTION uint32_t frame_id(const uint8_t* packet_buf) const;
    OUSTER_API_FUNCTION uint32_t init_id(const uint8_t* packet_buf) const;
    OUSTER_API_FUNCTION uint64_t prod_sn(const uint8_t* packet_buf) const;
    OUSTER_API_FUNCTION uint8_t alert_flags(const uint8_t* lidar_buf) const;
    OUSTER_API_FUNCTION uint16_t countdown_thermal_shutdown(const uint8_t* lidar_buf) const;
    OUSTER_API_FUNCTION uint16_t countdown_shot_limiting(const uint8_t* lidar_buf) const;
    OUSTER_API_FUNCTION ThermalShutdownStatus thermal_shutdown(const uint8_t* lidar_buf) const;
    OUSTER_API_FUNCTION ShotLimitingStatus shot_limiting(const uint8_t* lidar_buf) const;
    OUSTER_API_FUNCTION const uint8_t* footer(const uint8_t* lidar_buf) const;

    // ---- measurement blocks (parsing.cpp:793-842) ----
    OUSTER_API_FUNCTION const uint8_t* nth_col(size_t col_idx, const uint8_t* lidar_buf) const;
    OUSTER_API_FUNCTION uint8_t* nth_col(size_t col_idx, uint8_t* lidar_buf) const;
    OUSTER_API_FUNCTION uint32_t col_status(const uint8_t* col_buf) const;
    OUSTER_API_FUNCTION uint64_t col_timestamp(const uint8_t* col_buf) const;
    OUSTER_API_FUNCTION uint16_t col_measurement_id(const uint8_t* col_buf) const;
    OUSTER_API_FUNCTION const uint8_t* nth_px(size_t px_idx, const uint8_t* col_buf) const;

    // ---- channel fields ----
    OUSTER_API_FUNCTION ChanFieldType field_type(const std::string& f) const;
    OUSTER_API_FUNCTION FieldIter begin() const;  ///< std::map (name) order, parsing.cpp:619-623
    OUSTER_API_FUNCTION FieldIter end() const;
    OUSTER_API_FUNCTION uint64_t field_value_mask(const std::string& f) const;
    OUSTER_API_FUNCTION int field_bitness(const std::string& f) const;
    OUSTER_API_FUNCTION int block_parsable() const;
    OUSTER_API_FUNCTION const FieldDecodeInfo& field_decode_info(const std::string& f) const;
    OUSTER_API_FUNCTION bool has_field(const std::string& f) const;
    /// Decode infos of the per-column headers, relative to the column start (parsing.cpp:499-538).
    OUSTER_API_FUNCTION const FieldDecodeInfo& col_timestamp_info() const;
    OUSTER_API_FUNCTION const FieldDecodeInfo& col_measurement_id_info() const;
    OUSTER_API_FUNCTION const FieldDecodeInfo& col_status_info() const;

    /// Decode one column of one field (host; throws "Dest type too small for specified field").
    template <typename T>
    void col_field(const uint8_t* col_buf, const std::string& f, T* dst, int dst_stride = 1) const {
        const FieldDecodeInfo& info = checked_field(f, sizeof(T));
        for (int px = 0; px < pixels_per_column; px++)
            dst[static_cast<size_t>(px) * dst_stride] =
                info.get<T>(col_buf + col_header_size + px * channel_data_size);
    }
    /// Decode one whole packet of one field into a row-major image with `cols` columns (host).
    template <typename T, int BlockDim>
    void block_field(T* data, int cols, const std::string& f, const uint8_t* lidar_buf) const {
        const FieldDecodeInfo& info = checked_field(f, sizeof(T));
        for (int icol = 0; icol < columns_per_packet; icol += BlockDim) {
            const uint16_t m_id = col_measurement_id(nth_col(icol, lidar_buf));
            for (int x = 0; x < BlockDim; ++x) {
                const uint8_t* col = nth_col(icol + x, lidar_buf) + col_header_size;
                for (int px = 0; px < pixels_per_column; ++px)
                    data[static_cast<ptrdiff_t>(cols) * px + m_id + x] =
                        info.get<T>(col + px * channel_data_size);
            }
        }
    }

    // ---- writers (inverse path, parsing.cpp:1007-1090) ----
    OUSTER_API_FUNCTION void set_col_status(uint8_t* col_buf, uint32_t status) const;
    OUSTER_API_FUNCTION void set_col_timestamp(uint8_t* col_buf, uint64_t ts) const;
    OUSTER_API_FUNCTION void set_col_measurement_id(uint8_t* col_buf, uint16_t m_id) const;
    OUSTER_API_FUNCTION void set_frame_id(uint8_t* lidar_buf, uint32_t frame_id) const;
    OUSTER_API_FUNCTION void set_init_id(uint8_t* lidar_buf, uint32_t init_id) const;
    OUSTER_API_FUNCTION void set_packet_type(uint8_t* packet_buf, uint16_t packet_type) const;
    OUSTER_API_FUNCTION void set_prod_sn(uint8_t* lidar_buf, uint64_t sn) const;
    OUSTER_API_FUNCTION void set_alert_flags(uint8_t* lidar_buf, uint8_t alert_flags) const;
    OUSTER_API_FUNCTION void set_shutdown(uint8_t* lidar_buf, uint8_t status) const;
    OUSTER_API_FUNCTION void set_shot_limiting(uint8_t* lidar_buf, uint8_t status) const;
    OUSTER_API_FUNCTION void set_shutdown_countdown(uint8_t* lidar_buf, uint8_t v) const;
    OUSTER_API_FUNCTION void set_shot_limiting_countdown(uint8_t* lidar_buf, uint8_t v) const;
    /// Encode a field of a whole packet from a row-major image (skips invalid columns).
    template <typename T>
    void set_block(const T* data, int cols, const std::string& f, uint8_t* lidar_buf) const {
        const FieldDecodeInfo& info = field_decode_info(f);
        const uint16_t m_id = col_measurement_id(nth_col(0, lidar_buf));
        for (int x = 0; x < columns_per_packet; ++x) {
            uint8_t* col = nth_col(x, lidar_buf);
            if (!(col_status(col) & 0x01)) continue;
            for (int px = 0; px < pixels_per_column; ++px)
                info.set(col + col_header_size + px * channel_data_size,
                         data[static_cast<ptrdiff_t>(cols) * px + m_id + x]);
        }
    }

    OUSTER_API_FUNCTION uint64_t calculate_crc(const uint8_t* buffer, size_t buffer_size) const;
    /// CRC64 stored in the last 8 bytes of the packet; empty for LEGACY and FUSA packets, which
    /// carry none (parsing.cpp:1219-1228).
    std::optional<uint64_t> crc(const uint8_t* buffer, size_t buffer_size) const {
        if (udp_profile_lidar == UDPProfileLidar::LEGACY ||
            udp_profile_lidar == UDPProfileLidar::FUSA_RNG15_RFL8_NIR8_DUAL || header_type == HeaderType::FUSA)
            return std::nullopt;
        uint64_t v = 0;
        for (int i = 7; i >= 0; --i) v = (v << 8) | buffer[buffer_size - 8 + static_cast<size_t>(i)];
        return v;
    }
    OUSTER_API_FUNCTION int frame_id_difference(uint32_t current, uint32_t other) const;

    /// Replace the channel-field table (the effect add_custom_profile() has on a profile,
    /// ouster_core/src/profile_extension.cpp:134-163); masks of 0 become the type mask.
    OUSTER_API_FUNCTION void set_custom_fields(
        const std::vector<std::pair<std::string, FieldDecodeInfo>>& fields, size_t chan_data_size);

   private:
    const FieldDecodeInfo& checked_field(const std::string& f, size_t dest_size) const;
    struct Impl;
    std::shared_ptr<Impl> impl_;
    std::vector<FieldTypeEntry> field_types_;
    void rebuild_field_types();
    friend class FrameBatcher;
};

/// Cached PacketFormat for a DataFormat (parsing.cpp:979-997).
OUSTER_API_FUNCTION const PacketFormat& get_format(const DataFormat& format);
OUSTER_API_FUNCTION const PacketFormat& get_format(const SensorInfo& info);

}  // namespace core
}  // namespace sdk
}  // namespace ouster
