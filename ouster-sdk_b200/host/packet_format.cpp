// packet_format.cpp -- PacketFormat: profile tables, packet geometry, header accessors, CRC
// (host mirror of ouster_core/src/parsing.cpp:57-122, 170-363, 453-626, 736-842, 958-1090,
// 1183-1234, 1312-1321).
#include <algorithm>
#include <map>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <tuple>

#include "ouster/core/types.h"

namespace ouster {
namespace sdk {
namespace core {

FieldDecodeInfo field_info(size_t bit_start, size_t bit_size, size_t upshift, size_t max_length,
                           size_t num_elements) {
    const size_t value_bits = bit_size + upshift;
    if (value_bits > 64)
        throw std::invalid_argument(
            "failed creating FieldDecodeInfo: value cannot store more than 64 bits");
    FieldDecodeInfo out{};
    out.offset = bit_start >> 3;
    const size_t lsb = bit_start & 7;
    out.mask = bit_size >= 64 ? ~0ull : ((1ull << bit_size) - 1ull);
    out.mask = lsb ? (out.mask << lsb) : out.mask;
    out.shift = static_cast<int>(lsb) - static_cast<int>(upshift);
    out.num_elements = static_cast<int>(num_elements);

    size_t nbytes = ((value_bits + 7) / 8) / num_elements;
    if (nbytes == 1) out.ty_tag = ChanFieldType::UINT8;
    else if (nbytes == 2) out.ty_tag = ChanFieldType::UINT16;
    else if (nbytes == 3 || nbytes == 4) out.ty_tag = ChanFieldType::UINT32;
    else if (nbytes >= 5 && nbytes <= 8) out.ty_tag = ChanFieldType::UINT64;
    else out.ty_tag = ChanFieldType::VOID;

    if (max_length > 0) {
        if (out.offset + nbytes > max_length)
            throw std::invalid_argument(
                "failed creating FieldDecodeInfo: asked to read past end of packet");
        // an 8-byte load at `offset` must stay inside the buffer: slide the window back
        const long slide = static_cast<long>(out.offset) + 8 - static_cast<long>(max_length);
        if (slide > 0) {
            out.offset -= static_cast<size_t>(slide);
            out.mask <<= slide * 8;
            out.shift += static_cast<int>(slide * 8);
        }
    }
    return out;
}

namespace impl {
uint64_t get_value_mask(const FieldDecodeInfo& f) {
    const uint64_t tm = field_type_mask(f.ty_tag);
    uint64_t m = f.mask ? f.mask : tm;
    if (f.shift > 0) m >>= f.shift;
    if (f.shift < 0) m <<= -f.shift;
    return m & tm;
}
int get_bitness(const FieldDecodeInfo& f) {
    uint64_t m = get_value_mask(f);
    int n = 0;
    for (; m; m &= m - 1) ++n;
    return n;
}
}  // namespace impl

namespace {

// Profile tables as "NAME:first_bit:bits[:upshift[:elements]]" lists (bit positions inside one
// pixel's channel data block), plus the channel data size in bytes.
struct ProfileSpec {
    UDPProfileLidar profile;
    size_t chan_data_size;
    const char* fields;
};

const char* const kRaw3 = "RAW32_WORD1:0:32,RAW32_WORD2:32:32,RAW32_WORD3:64:32";
const char* const kDualCore =
    "RANGE:0:19,FLAGS:19:5,REFLECTIVITY:24:8,RANGE2:32:19,FLAGS2:51:5,REFLECTIVITY2:56:8,"
    "SIGNAL:64:16,SIGNAL2:80:16";

const ProfileSpec kProfiles[] = {
    {UDPProfileLidar::LEGACY, 12,
     "RANGE:0:20,FLAGS:28:4,REFLECTIVITY:32:8,SIGNAL:48:16,NEAR_IR:64:16,$RAW3"},
    {UDPProfileLidar::RNG19_RFL8_SIG16_NIR16_DUAL, 16,
     "$DUAL,NEAR_IR:96:16,WINDOW:120:8,$RAW3,RAW32_WORD4:96:32"},
    {UDPProfileLidar::RNG19_RFL8_SIG16_NIR16, 12,
     "RANGE:0:19,FLAGS:19:5,REFLECTIVITY:32:8,SIGNAL:48:16,NEAR_IR:64:16,WINDOW:88:8,$RAW3"},
    {UDPProfileLidar::RNG15_RFL8_NIR8, 4,
     "RANGE:0:15:3,FLAGS:15:1,REFLECTIVITY:16:8,NEAR_IR:24:8:4,RAW32_WORD1:0:32"},
    {UDPProfileLidar::FIVE_WORD_PIXEL, 20,
     "$DUAL,NEAR_IR:96:16,$RAW3,RAW32_WORD4:96:32,RAW32_WORD5:128:32"},
    {UDPProfileLidar::FUSA_RNG15_RFL8_NIR8_DUAL, 8,
     "RANGE:0:15:3,FLAGS:15:1,REFLECTIVITY:16:8,NEAR_IR:24:8:4,RANGE2:32:15:3,FLAGS2:47:1,"
     "REFLECTIVITY2:48:8,WINDOW:56:8,RAW32_WORD1:0:32,RAW32_WORD2:32:32"},
    {UDPProfileLidar::RNG15_RFL8_NIR8_DUAL, 8,
     "RANGE:0:15:3,FLAGS:15:1,REFLECTIVITY:16:8,NEAR_IR:24:8:4,RANGE2:32:15:3,FLAGS2:47:1,"
     "REFLECTIVITY2:48:8,WINDOW:56:8,RAW32_WORD1:0:32,RAW32_WORD2:32:32"},
    {UDPProfileLidar::OFF, 0, ""},
    {UDPProfileLidar::RNG15_RFL8_NIR8_ZONE16, 8,
     "RANGE:0:15:3,FLAGS:15:1,REFLECTIVITY:16:8,NEAR_IR:24:8:4,ZONE_MASK:32:16,WINDOW:48:8,"
     "RAW32_WORD1:0:32,RAW32_WORD2:32:32"},
    {UDPProfileLidar::RNG19_RFL8_SIG16_NIR16_ZONE16, 12,
     "RANGE:0:19,FLAGS:19:5,REFLECTIVITY:32:8,WINDOW:40:8,SIGNAL:48:16,NEAR_IR:64:16,"
     "ZONE_MASK:80:16,$RAW3"},
    {UDPProfileLidar::RNG15_RFL8_WIN8, 4,
     "RANGE:0:15:3,FLAGS:15:1,REFLECTIVITY:16:8,WINDOW:24:8,RAW32_WORD1:0:32"},
    {UDPProfileLidar::RNG19_RFL8_SIG16_ZONE16_DUAL, 16,
     "$DUAL,ZONE_MASK:96:16,WINDOW:120:8,$RAW3,RAW32_WORD4:96:32"},
    {UDPProfileLidar::RNG19_RFL8_SIG16_NIR16_RGB16, 16,
     "RANGE:0:19,FLAGS:19:5,REFLECTIVITY:24:8,SIGNAL:32:16,NEAR_IR:48:16,R:64:16,G:80:16,B:96:16,"
     "RGB:64:48:0:3,$RAW3,RAW32_WORD4:96:32"},
    {UDPProfileLidar::RNG19_RFL8_SIG16_NIR16_RGB16_DUAL, 20,
     "$DUAL,NEAR_IR:96:16,R:112:16,G:128:16,B:144:16,RGB:112:48:0:3,$RAW3,RAW32_WORD4:96:32,"
     "RAW32_WORD5:128:32"},
};

std::string expand(std::string s) {
    const std::pair<const char*, const char*> subs[] = {{"$RAW3", kRaw3}, {"$DUAL", kDualCore}};
    for (const auto& kv : subs) {
        size_t pos;
        while ((pos = s.find(kv.first)) != std::string::npos)
            s.replace(pos, std::string(kv.first).size(), kv.second);
    }
    return s;
}

std::map<std::string, FieldDecodeInfo> parse_fields(const char* spec) {
    std::map<std::string, FieldDecodeInfo> out;
    std::stringstream ss(expand(spec));
    std::string item;
    while (std::getline(ss, item, ',')) {
        if (item.empty()) continue;
        std::stringstream is(item);
        std::string tok;
        std::vector<std::string> parts;
        while (std::getline(is, tok, ':')) parts.push_back(tok);
        const size_t bit = std::stoul(parts.at(1)), bits = std::stoul(parts.at(2));
        const size_t up = parts.size() > 3 ? std::stoul(parts[3]) : 0;
        const size_t nel = parts.size() > 4 ? std::stoul(parts[4]) : 1;
        out[parts[0]] = field_info(bit, bits, up, 0, nel);
    }
    return out;
}

const ProfileSpec& find_profile(UDPProfileLidar p) {
    for (const auto& e : kProfiles)
        if (e.profile == p) return e;
    throw std::invalid_argument("Unknown lidar udp profile");
}

uint64_t crc64(const uint8_t* buf, size_t len) {
    // ECMA-182 polynomial, reflected, byte-at-a-time table (parsing.cpp:1187-1216)
    static const std::array<uint64_t, 256> table = [] {
        std::array<uint64_t, 256> t{};
        for (uint32_t b = 0; b < 256; ++b) {
            uint64_t r = b;
            for (int k = 0; k < 8; ++k) r = (r & 1) ? (r >> 1) ^ 0xC96C5795D7870F42ull : (r >> 1);
            t[b] = r;
        }
        return t;
    }();
    uint64_t crc = ~0ull;
    for (size_t i = 0; i < len; ++i) crc = table[(buf[i] ^ crc) & 0xff] ^ (crc >> 8);
    return ~crc;
}

}  // namespace

struct PacketFormat::Impl {
    std::map<std::string, FieldDecodeInfo> fields;
    FieldDecodeInfo packet_type, frame_id, init_id, prod_sn, alert_flags;
    FieldDecodeInfo countdown_thermal, countdown_shot, thermal_shutdown, shot_limiting;
    FieldDecodeInfo col_status, col_timestamp, col_measurement_id;
};

PacketFormat::PacketFormat(const DataFormat& format)
    : udp_profile_lidar(format.udp_profile_lidar),
      udp_profile_imu(format.udp_profile_imu),
      header_type(format.header_type),
      columns_per_packet(static_cast<int>(format.columns_per_packet)),
      pixels_per_column(static_cast<int>(format.pixels_per_column)),
      impl_(std::make_shared<Impl>()) {
    const bool legacy = udp_profile_lidar == UDPProfileLidar::LEGACY;
    const bool fusa = header_type == HeaderType::FUSA && !legacy;
    const ProfileSpec& spec = find_profile(udp_profile_lidar);

    packet_header_size = legacy ? 0 : 32;
    col_header_size = legacy ? 16 : 12;
    channel_data_size = spec.chan_data_size;
    col_footer_size = legacy ? 4 : 0;
    packet_footer_size = legacy ? 0 : 32;
    col_size = col_header_size + format.pixels_per_column * channel_data_size + col_footer_size;
    lidar_packet_size = packet_header_size + format.columns_per_packet * col_size + packet_footer_size;
    if (lidar_packet_size > 65535) throw std::invalid_argument("lidar_packet_size cannot exceed 65535");
    max_frame_id = format.max_frame_id();

    Impl& d = *impl_;
    d.fields = parse_fields(spec.fields);
    const FieldDecodeInfo none = field_info(0, 0);
    if (legacy) {
        // parsing.cpp:480-490: LEGACY lidar packets cannot be combined with the newer IMU / zone-monitor packet
        // formats (their headers and footers differ); the reference rejects the configuration with this text
        if (format.udp_profile_imu == UDPProfileIMU::ACCEL32_GYRO32_NMEA || format.zone_monitoring_enabled)
            throw std::runtime_error(
                "Invalid sensor configuration. Mixing LEGACY lidar packets and non-LEGACY IMU and ZONE packets is not "
                "possible or supported by the SDK. Your udp_profile_lidar may be incorrect for this data.");
        d.packet_type = d.init_id = d.prod_sn = d.alert_flags = none;
        d.countdown_thermal = d.countdown_shot = d.thermal_shutdown = d.shot_limiting = none;
        d.frame_id = field_info(80, 16);  // inside the first column header
        // the status word is the column footer: read it through a window that ends with the column
        const size_t bit = 8 * (col_size - col_footer_size);
        d.col_status = field_info(bit, 32, 0, (bit + 32) / 8);
    } else if (fusa) {
        d.packet_type = field_info(0, 8);
        d.init_id = field_info(8, 24);
        d.frame_id = field_info(32, 32);
        d.alert_flags = field_info(64, 8);
        d.prod_sn = field_info(88, 40);
        d.countdown_thermal = field_info(128, 8);
        d.countdown_shot = field_info(136, 8);
        d.thermal_shutdown = field_info(144, 4);
        d.shot_limiting = field_info(152, 4);
        d.col_status = field_info(80, 16);
    } else {
        d.packet_type = field_info(0, 16);
        d.frame_id = field_info(16, 16);
        d.init_id = field_info(32, 24);
        d.prod_sn = field_info(56, 40);
        d.alert_flags = field_info(96, 8);
        d.countdown_thermal = field_info(128, 8);
        d.countdown_shot = field_info(136, 8);
        d.thermal_shutdown = field_info(144, 4);
        d.shot_limiting = field_info(152, 4);
        d.col_status = field_info(80, 16);
    }
    d.col_timestamp = field_info(0, 64);
    d.col_measurement_id = field_info(64, 16);
    rebuild_field_types();
}

PacketFormat::PacketFormat(const SensorInfo& info) : PacketFormat(info.format) {}

void PacketFormat::rebuild_field_types() {
    field_types_.clear();
    for (const auto& kv : impl_->fields)
        field_types_.emplace_back(kv.first, std::make_pair(kv.second.ty_tag, kv.second.num_elements));
}

void PacketFormat::set_custom_fields(
    const std::vector<std::pair<std::string, FieldDecodeInfo>>& fields, size_t chan_data_size) {
    auto fresh = std::make_shared<Impl>(*impl_);
    fresh->fields.clear();
    for (const auto& kv : fields) {
        FieldDecodeInfo f = kv.second;
        if (f.mask == 0) f.mask = field_type_mask(f.ty_tag);
        fresh->fields[kv.first] = f;
    }
    impl_ = fresh;
    channel_data_size = chan_data_size;
    col_size = col_header_size + pixels_per_column * channel_data_size + col_footer_size;
    lidar_packet_size = packet_header_size + columns_per_packet * col_size + packet_footer_size;
    rebuild_field_types();
}

// ---- packet headers ----
uint16_t PacketFormat::packet_type(const uint8_t* b) const { return impl_->packet_type.get<uint16_t>(b); }
uint32_t PacketFormat::frame_id(const uint8_t* b) const { return impl_->frame_id.get<uint32_t>(b); }
uint32_t PacketFormat::init_id(const uint8_t* b) const { return impl_->init_id.get<uint32_t>(b); }
uint64_t PacketFormat::prod_sn(const uint8_t* b) const { return impl_->prod_sn.get<uint64_t>(b); }
uint8_t PacketFormat::alert_flags(const uint8_t* b) const { return impl_->alert_flags.get<uint8_t>(b); }
uint16_t PacketFormat::countdown_thermal_shutdown(const uint8_t* b) const {
    return impl_->countdown_thermal.get<uint16_t>(b);
}
uint16_t PacketFormat::countdown_shot_limiting(const uint8_t* b) const {
    return impl_->countdown_shot.get<uint16_t>(b);
}
ThermalShutdownStatus PacketFormat::thermal_shutdown(const uint8_t* b) const {
    return static_cast<ThermalShutdownStatus>(impl_->thermal_shutdown.get<uint8_t>(b));
}
ShotLimitingStatus PacketFormat::shot_limiting(const uint8_t* b) const {
    return static_cast<ShotLimitingStatus>(impl_->shot_limiting.get<uint8_t>(b));
}
const uint8_t* PacketFormat::footer(const uint8_t* lidar_buf) const {
    if (packet_footer_size == 0) return nullptr;
    return lidar_buf + packet_header_size + columns_per_packet * col_size;
}

// ---- measurement blocks ----
const uint8_t* PacketFormat::nth_col(size_t i, const uint8_t* lidar_buf) const {
    return lidar_buf + packet_header_size + i * col_size;
}
uint8_t* PacketFormat::nth_col(size_t i, uint8_t* lidar_buf) const {
    return lidar_buf + packet_header_size + i * col_size;
}
uint32_t PacketFormat::col_status(const uint8_t* c) const { return impl_->col_status.get<uint32_t>(c); }
uint64_t PacketFormat::col_timestamp(const uint8_t* c) const { return impl_->col_timestamp.get<uint64_t>(c); }
uint16_t PacketFormat::col_measurement_id(const uint8_t* c) const {
    return impl_->col_measurement_id.get<uint16_t>(c);
}
const uint8_t* PacketFormat::nth_px(size_t px, const uint8_t* col_buf) const {
    return col_buf + col_header_size + px * channel_data_size;
}

const FieldDecodeInfo& PacketFormat::col_timestamp_info() const { return impl_->col_timestamp; }
const FieldDecodeInfo& PacketFormat::col_measurement_id_info() const { return impl_->col_measurement_id; }
const FieldDecodeInfo& PacketFormat::col_status_info() const { return impl_->col_status; }

// ---- channel fields ----
bool PacketFormat::has_field(const std::string& f) const { return impl_->fields.count(f) != 0; }
const FieldDecodeInfo& PacketFormat::field_decode_info(const std::string& f) const {
    return impl_->fields.at(f);
}
const FieldDecodeInfo& PacketFormat::checked_field(const std::string& f, size_t dest_size) const {
    const FieldDecodeInfo& info = impl_->fields.at(f);
    if (dest_size < field_type_size(info.ty_tag) * static_cast<size_t>(info.num_elements))
        throw std::invalid_argument("Dest type too small for specified field");
    return info;
}
ChanFieldType PacketFormat::field_type(const std::string& f) const {
    return has_field(f) ? impl_->fields.at(f).ty_tag : ChanFieldType::VOID;
}
PacketFormat::FieldIter PacketFormat::begin() const { return field_types_.cbegin(); }
PacketFormat::FieldIter PacketFormat::end() const { return field_types_.cend(); }
uint64_t PacketFormat::field_value_mask(const std::string& f) const {
    return impl::get_value_mask(impl_->fields.at(f));
}
int PacketFormat::field_bitness(const std::string& f) const {
    return impl::get_bitness(impl_->fields.at(f));
}
int PacketFormat::block_parsable() const {
    for (int dim : {16, 8, 4})
        if (pixels_per_column % dim == 0 && columns_per_packet % dim == 0) return dim;
    return 0;
}

// ---- writers ----
void PacketFormat::set_col_status(uint8_t* c, uint32_t v) const { impl_->col_status.set(c, v); }
void PacketFormat::set_col_timestamp(uint8_t* c, uint64_t v) const { impl_->col_timestamp.set(c, v); }
void PacketFormat::set_col_measurement_id(uint8_t* c, uint16_t v) const {
    impl_->col_measurement_id.set(c, v);
}
void PacketFormat::set_frame_id(uint8_t* b, uint32_t v) const { impl_->frame_id.set(b, v); }
void PacketFormat::set_init_id(uint8_t* b, uint32_t v) const { impl_->init_id.set(b, v); }
void PacketFormat::set_packet_type(uint8_t* b, uint16_t v) const { impl_->packet_type.set(b, v); }
void PacketFormat::set_prod_sn(uint8_t* b, uint64_t v) const { impl_->prod_sn.set(b, v); }
void PacketFormat::set_alert_flags(uint8_t* b, uint8_t v) const { impl_->alert_flags.set(b, v); }
void PacketFormat::set_shutdown(uint8_t* b, uint8_t v) const { impl_->thermal_shutdown.set(b, v); }
void PacketFormat::set_shot_limiting(uint8_t* b, uint8_t v) const { impl_->shot_limiting.set(b, v); }
void PacketFormat::set_shutdown_countdown(uint8_t* b, uint8_t v) const {
    impl_->countdown_thermal.set(b, v);
}
void PacketFormat::set_shot_limiting_countdown(uint8_t* b, uint8_t v) const {
    impl_->countdown_shot.set(b, v);
}

uint64_t PacketFormat::calculate_crc(const uint8_t* buffer, size_t buffer_size) const {
    return crc64(buffer, buffer_size - 8);
}

int PacketFormat::frame_id_difference(uint32_t current, uint32_t other) const {
    const int64_t span = static_cast<int64_t>(max_frame_id) + 1;
    const int64_t half = max_frame_id >> 1;
    int64_t delta = static_cast<int64_t>(other) - static_cast<int64_t>(current);
    if (delta > half) delta -= span;
    else if (delta < -half) delta += span;
    return static_cast<int>(delta);
}

// ---- format cache ----
namespace {
auto key_of(const DataFormat& f) {
    return std::tie(f.pixels_per_column, f.columns_per_packet, f.columns_per_frame,
                    f.imu_measurements_per_packet, f.pixel_shift_by_row, f.column_window,
                    f.udp_profile_lidar, f.udp_profile_imu, f.header_type);
}
struct FormatLess {
    bool operator()(const DataFormat& a, const DataFormat& b) const { return key_of(a) < key_of(b); }
};
}  // namespace

const PacketFormat& get_format(const DataFormat& format) {
    static std::map<DataFormat, std::unique_ptr<PacketFormat>, FormatLess> cache;
    static std::mutex mx;
    std::lock_guard<std::mutex> lk(mx);
    auto it = cache.find(format);
    if (it == cache.end()) it = cache.emplace(format, std::make_unique<PacketFormat>(format)).first;
    return *it->second;
}
const PacketFormat& get_format(const SensorInfo& info) { return get_format(info.format); }

}  // namespace core
}  // namespace sdk
}  // namespace ouster
