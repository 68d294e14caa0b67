// sensor_info.h -- the subset of SensorInfo the scan->pointcloud path reads
// (mirrors ouster_core/include/ouster/core/sensor_info.h:171-244; JSON metadata parsing is out
// of scope -- DESIGN.md -- callers fill the members directly or through from_default()).
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "ouster/core/data_format.h"
#include "ouster/core/typedefs.h"
#include "ouster/core/visibility.h"

namespace ouster {
namespace sdk {
namespace core {

constexpr double RANGE_UNIT = 0.001;  ///< metres per range unit (types.h:46)

struct Version {
    uint16_t major{0}, minor{0}, patch{0};
    bool operator<(const Version& o) const {
        if (major != o.major) return major < o.major;
        if (minor != o.minor) return minor < o.minor;
        return patch < o.patch;
    }
};

struct OUSTER_API_CLASS SensorInfo {
    uint64_t sn{0};
    std::string fw_rev{"UNKNOWN"};
    std::string prod_line{"OS-1-64"};
    DataFormat format{};
    std::vector<double> beam_azimuth_angles;
    std::vector<double> beam_altitude_angles;
    double lidar_origin_to_beam_origin_mm{0};
    mat4d beam_to_lidar_transform = mat4d::Identity();
    mat4d imu_to_sensor_transform = mat4d::Identity();
    mat4d lidar_to_sensor_transform = mat4d::Identity();
    mat4d sensor_to_body = mat4d::Identity();
    uint32_t init_id{0};

    /// Columns per frame / pixels per column.
    uint32_t w() const { return format.columns_per_frame; }
    uint32_t h() const { return format.pixels_per_column; }

    /// Firmware version parsed from fw_rev ("v3.2.0..." or "3.2.0"); 0.0.0 when unknown.
    OUSTER_API_FUNCTION Version get_version() const;

    /// Default OS-1-64 gen1 sensor for the given mode (sensor_info.cpp:163-222).
    OUSTER_API_FUNCTION static std::shared_ptr<SensorInfo> from_default(LidarMode mode);
};

OUSTER_API_FUNCTION double default_lidar_origin_to_beam_origin(const std::string& prod_line);
OUSTER_API_FUNCTION mat4d default_beam_to_lidar_transform(const std::string& prod_line);

}  // namespace core
}  // namespace sdk
}  // namespace ouster
