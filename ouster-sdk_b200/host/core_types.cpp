// core_types.cpp -- element types, DataFormat helpers, SensorInfo defaults
// (host mirror; reference: chanfield.h:136-170, data_format.cpp:79-168, sensor_info.cpp:89-222).
#include <cstdio>
#include <stdexcept>

#include "ouster/core/chanfield.h"
#include "ouster/core/data_format.h"
#include "ouster/core/sensor_info.h"

namespace ouster {
namespace sdk {
namespace core {

size_t field_type_size(ChanFieldType ft) {
    switch (ft) {
        case ChanFieldType::INT8:
        case ChanFieldType::UINT8:
        case ChanFieldType::CHAR: return 1;
        case ChanFieldType::INT16:
        case ChanFieldType::UINT16:
        case ChanFieldType::FLOAT16: return 2;
        case ChanFieldType::INT32:
        case ChanFieldType::UINT32:
        case ChanFieldType::FLOAT32: return 4;
        case ChanFieldType::INT64:
        case ChanFieldType::UINT64:
        case ChanFieldType::FLOAT64: return 8;
        default: return 0;
    }
}

uint64_t field_type_mask(ChanFieldType ft) {
    switch (field_type_size(ft)) {
        case 1: return 0xffull;
        case 2: return 0xffffull;
        case 4: return 0xffffffffull;
        case 8: return ~0ull;
        default: return 0;
    }
}

std::string to_string(ChanFieldType ft) {
    switch (ft) {
        case ChanFieldType::VOID: return "VOID";
        case ChanFieldType::UINT8: return "UINT8";
        case ChanFieldType::UINT16: return "UINT16";
        case ChanFieldType::UINT32: return "UINT32";
        case ChanFieldType::UINT64: return "UINT64";
        case ChanFieldType::INT8: return "INT8";
        case ChanFieldType::INT16: return "INT16";
        case ChanFieldType::INT32: return "INT32";
        case ChanFieldType::INT64: return "INT64";
        case ChanFieldType::FLOAT32: return "FLOAT32";
        case ChanFieldType::FLOAT64: return "FLOAT64";
        case ChanFieldType::CHAR: return "CHAR";
        case ChanFieldType::FLOAT16: return "FLOAT16";
        case ChanFieldType::ZONE_STATE: return "ZONE_STATE";
        default: return "UNKNOWN";
    }
}

// ---- DataFormat ----
int DataFormat::valid_columns_per_frame() const {
    const int a = column_window.first, b = column_window.second;
    return a <= b ? b - a + 1 : b + static_cast<int>(columns_per_frame) - a + 1;
}

int DataFormat::lidar_packets_per_frame() const {
    if (udp_profile_lidar == UDPProfileLidar::OFF) return 0;
    const int first_packet = static_cast<int>(column_window.first / columns_per_packet);
    const int last_packet = static_cast<int>(column_window.second / columns_per_packet);
    if (column_window.second >= column_window.first) return last_packet - first_packet + 1;
    // the azimuth window wraps through column 0
    const int all_packets = static_cast<int>(columns_per_frame / columns_per_packet) +
                            ((columns_per_frame % columns_per_packet) ? 1 : 0);
    if (first_packet == last_packet) return all_packets;
    return (all_packets - first_packet) + 1 + last_packet;
}

uint32_t DataFormat::max_frame_id() const {
    const bool wide = header_type == HeaderType::FUSA && udp_profile_lidar != UDPProfileLidar::LEGACY;
    return wide ? 0xffffffffu : 0xffffu;
}

bool operator==(const DataFormat& l, const DataFormat& r) {
    return l.pixels_per_column == r.pixels_per_column && l.columns_per_packet == r.columns_per_packet &&
           l.columns_per_frame == r.columns_per_frame &&
           l.imu_measurements_per_packet == r.imu_measurements_per_packet &&
           l.imu_packets_per_frame == r.imu_packets_per_frame &&
           l.pixel_shift_by_row == r.pixel_shift_by_row && l.column_window == r.column_window &&
           l.udp_profile_lidar == r.udp_profile_lidar && l.udp_profile_imu == r.udp_profile_imu &&
           l.header_type == r.header_type && l.fps == r.fps &&
           l.zone_monitoring_enabled == r.zone_monitoring_enabled;
}
bool operator!=(const DataFormat& l, const DataFormat& r) { return !(l == r); }

DataFormat default_data_format(LidarMode mode) {
    int unit;  // per-row stagger step of the 64-beam gen1 layout
    switch (mode.columns) {
        case 512: unit = 3; break;
        case 1024: unit = 6; break;
        case 2048: unit = 12; break;
        case 4096: unit = 24; break;
        default: throw std::invalid_argument{"default_data_format"};
    }
    DataFormat f;
    f.pixels_per_column = 64;
    f.columns_per_packet = DEFAULT_COLUMNS_PER_PACKET;
    f.columns_per_frame = mode.columns;
    f.pixel_shift_by_row.reserve(64);
    for (int i = 0; i < 16; ++i)
        for (int k = 3; k >= 0; --k) f.pixel_shift_by_row.push_back(k * unit);
    f.column_window = {0, static_cast<uint16_t>(mode.columns - 1)};
    f.udp_profile_lidar = UDPProfileLidar::LEGACY;
    f.udp_profile_imu = UDPProfileIMU::LEGACY;
    f.header_type = HeaderType::STANDARD;
    f.fps = static_cast<uint16_t>(mode.fps);
    return f;
}

namespace {
struct ProfileName {
    UDPProfileLidar p;
    const char* name;
};
const ProfileName kProfileNames[] = {
    {UDPProfileLidar::LEGACY, "LEGACY"},
    {UDPProfileLidar::RNG19_RFL8_SIG16_NIR16_DUAL, "RNG19_RFL8_SIG16_NIR16_DUAL"},
    {UDPProfileLidar::RNG19_RFL8_SIG16_NIR16, "RNG19_RFL8_SIG16_NIR16"},
    {UDPProfileLidar::RNG15_RFL8_NIR8, "RNG15_RFL8_NIR8"},
    {UDPProfileLidar::FIVE_WORD_PIXEL, "FIVE_WORD_PIXEL"},
    {UDPProfileLidar::FUSA_RNG15_RFL8_NIR8_DUAL, "FUSA_RNG15_RFL8_NIR8_DUAL"},
    {UDPProfileLidar::RNG15_RFL8_NIR8_DUAL, "RNG15_RFL8_NIR8_DUAL"},
    {UDPProfileLidar::RNG15_RFL8_NIR8_ZONE16, "RNG15_RFL8_NIR8_ZONE16"},
    {UDPProfileLidar::RNG19_RFL8_SIG16_NIR16_ZONE16, "RNG19_RFL8_SIG16_NIR16_ZONE16"},
    {UDPProfileLidar::RNG15_RFL8_WIN8, "RNG15_RFL8_WIN8"},
    {UDPProfileLidar::RNG19_RFL8_SIG16_ZONE16_DUAL, "RNG19_RFL8_SIG16_ZONE16_DUAL"},
    {UDPProfileLidar::RNG19_RFL8_SIG16_NIR16_RGB16, "RNG19_RFL8_SIG16_NIR16_RGB16"},
    {UDPProfileLidar::RNG19_RFL8_SIG16_NIR16_RGB16_DUAL, "RNG19_RFL8_SIG16_NIR16_RGB16_DUAL"},
    {UDPProfileLidar::OFF, "OFF"},
};
}  // namespace

std::string to_string(UDPProfileLidar profile) {
    for (const auto& e : kProfileNames)
        if (e.p == profile) return e.name;
    return "UNKNOWN";
}

UDPProfileLidar udp_profile_lidar_of_string(const std::string& s) {
    for (const auto& e : kProfileNames)
        if (s == e.name) return e.p;
    return UDPProfileLidar::UNKNOWN;
}

std::string to_string(HeaderType t) { return t == HeaderType::FUSA ? "FUSA" : "STANDARD"; }

// ---- SensorInfo ----
double default_lidar_origin_to_beam_origin(const std::string& prod_line) {
    if (prod_line.rfind("OS-0-", 0) == 0) return 27.67;
    if (prod_line.rfind("OS-1-", 0) == 0) return 15.806;
    if (prod_line.rfind("OS-2-", 0) == 0) return 13.762;
    return 12.163;  // gen 1
}

mat4d default_beam_to_lidar_transform(const std::string& prod_line) {
    mat4d m = mat4d::Identity();
    m(0, 3) = default_lidar_origin_to_beam_origin(prod_line);
    return m;
}

Version SensorInfo::get_version() const {
    Version v;
    const char* s = fw_rev.c_str();
    while (*s && (*s < '0' || *s > '9')) ++s;
    unsigned a = 0, b = 0, c = 0;
    if (std::sscanf(s, "%u.%u.%u", &a, &b, &c) >= 2) {
        v.major = static_cast<uint16_t>(a);
        v.minor = static_cast<uint16_t>(b);
        v.patch = static_cast<uint16_t>(c);
    }
    return v;
}

std::shared_ptr<SensorInfo> SensorInfo::from_default(LidarMode mode) {
    auto info = std::make_shared<SensorInfo>();
    info->sn = 0;
    info->fw_rev = "UNKNOWN";
    info->prod_line = "OS-1-64";
    info->format = default_data_format(mode);
    // gen-1 OS-1-64 beam table: 64 altitudes from +16.611 deg in steps of ~0.527 deg, azimuths
    // cycling {3.164, 1.055, -1.055, -3.164} (sensor_info.cpp:194-216)
    static const double alt_top[32] = {16.611, 16.084, 15.557, 15.029, 14.502, 13.975, 13.447, 12.920,
                                       12.393, 11.865, 11.338, 10.811, 10.283, 9.756,  9.229,  8.701,
                                       8.174,  7.646,  7.119,  6.592,  6.064,  5.537,  5.010,  4.482,
                                       3.955,  3.428,  2.900,  2.373,  1.846,  1.318,  0.791,  0.264};
    static const double az4[4] = {3.164, 1.055, -1.055, -3.164};
    info->beam_altitude_angles.resize(64);
    info->beam_azimuth_angles.resize(64);
    for (int i = 0; i < 32; ++i) {
        info->beam_altitude_angles[i] = alt_top[i];
        info->beam_altitude_angles[63 - i] = -alt_top[i];
    }
    for (int i = 0; i < 64; ++i) info->beam_azimuth_angles[i] = az4[i % 4];
    info->lidar_origin_to_beam_origin_mm = default_lidar_origin_to_beam_origin(info->prod_line);
    info->beam_to_lidar_transform = default_beam_to_lidar_transform(info->prod_line);
    mat4d imu = mat4d::Identity();
    imu(0, 3) = 6.253;
    imu(1, 3) = -11.775;
    imu(2, 3) = 7.645;
    info->imu_to_sensor_transform = imu;
    mat4d l2s = mat4d::Identity();
    l2s(0, 0) = -1;
    l2s(1, 1) = -1;
    l2s(2, 3) = 36.18;
    info->lidar_to_sensor_transform = l2s;
    info->sensor_to_body = mat4d::Identity();
    info->init_id = 0;
    return info;
}

}  // namespace core
}  // namespace sdk
}  // namespace ouster
