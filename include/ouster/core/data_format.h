// data_format.h -- lidar data format description
// (mirrors ouster_core/include/ouster/core/data_format.h:27-137 for the fields the hot path reads).
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include "ouster/core/visibility.h"

namespace ouster {
namespace sdk {
namespace core {

/// Lidar UDP profile (values identical to the reference enum, data_format.h:27-72).
enum class UDPProfileLidar {
    UNKNOWN = 0,
    LEGACY,
    RNG19_RFL8_SIG16_NIR16_DUAL,
    RNG19_RFL8_SIG16_NIR16,
    RNG15_RFL8_NIR8,
    FIVE_WORD_PIXEL,
    FUSA_RNG15_RFL8_NIR8_DUAL,
    RNG15_RFL8_NIR8_DUAL,
    RNG15_RFL8_NIR8_ZONE16,
    RNG19_RFL8_SIG16_NIR16_ZONE16,
    RNG15_RFL8_WIN8,
    RNG19_RFL8_SIG16_ZONE16_DUAL,
    RNG19_RFL8_SIG16_NIR16_RGB16,
    RNG19_RFL8_SIG16_NIR16_RGB16_DUAL,
    OFF = 100,
};

enum class UDPProfileIMU { LEGACY = 0, ACCEL32_GYRO32_NMEA = 1, OFF = 100 };
enum class HeaderType { STANDARD = 0, FUSA = 1 };

using ColumnWindow = std::pair<uint16_t, uint16_t>;

/// Lidar mode: columns per frame and frame rate.
struct LidarMode {
    uint32_t columns{0};
    uint32_t fps{10};
};

struct OUSTER_API_CLASS DataFormat {
    uint32_t pixels_per_column{0};
    uint32_t columns_per_packet{0};
    uint32_t columns_per_frame{0};
    uint32_t imu_measurements_per_packet{0};
    uint32_t imu_packets_per_frame{0};
    std::vector<int> pixel_shift_by_row;
    ColumnWindow column_window{0, 0};
    UDPProfileLidar udp_profile_lidar{UDPProfileLidar::LEGACY};
    UDPProfileIMU udp_profile_imu{UDPProfileIMU::LEGACY};
    HeaderType header_type{HeaderType::STANDARD};
    uint16_t fps{10};
    bool zone_monitoring_enabled{false};

    OUSTER_API_FUNCTION int valid_columns_per_frame() const;
    OUSTER_API_FUNCTION int lidar_packets_per_frame() const;
    OUSTER_API_FUNCTION uint32_t max_frame_id() const;
};

OUSTER_API_FUNCTION bool operator==(const DataFormat& lhs, const DataFormat& rhs);
OUSTER_API_FUNCTION bool operator!=(const DataFormat& lhs, const DataFormat& rhs);

/// default_data_format(mode): 64 rows, 16 columns per packet, LEGACY (data_format.cpp:79-126)
OUSTER_API_FUNCTION DataFormat default_data_format(LidarMode mode);

OUSTER_API_FUNCTION std::string to_string(UDPProfileLidar profile);
OUSTER_API_FUNCTION UDPProfileLidar udp_profile_lidar_of_string(const std::string& s);
OUSTER_API_FUNCTION std::string to_string(HeaderType t);

constexpr uint32_t DEFAULT_COLUMNS_PER_PACKET = 16;  ///< defaults.h:5

}  // namespace core
}  // namespace sdk
}  // namespace ouster
