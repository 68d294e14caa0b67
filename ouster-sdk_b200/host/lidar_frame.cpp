// lidar_frame.cpp -- Field / LidarFrame host containers and the default field sets
// (host mirror of ouster_core/src/lidar_frame.cpp:73-446, 1038-1117 and field.cpp:247-275).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <stdexcept>

#include "ouster/core/lidar_frame.h"
#include "ouster_b200.h"

namespace ouster {
namespace sdk {
namespace core {

// ---- HostBuffer: pooled page-locked blocks ----
namespace {
struct PinnedPool {
    std::mutex mx;
    std::multimap<size_t, void*> free_blocks;  // capacity -> block
    bool use_cuda = ob_device_count() > 0;
    static size_t round_up(size_t n) { return (n + 65535) & ~static_cast<size_t>(65535); }
    void* acquire(size_t bytes, size_t* cap, bool* pinned) {
        const size_t want = round_up(std::max<size_t>(bytes, 1));
        if (use_cuda) {
            {
                std::lock_guard<std::mutex> lk(mx);
                auto it = free_blocks.lower_bound(want);
                if (it != free_blocks.end() && it->first <= want * 2) {
                    void* p = it->second;
                    *cap = it->first;
                    free_blocks.erase(it);
                    *pinned = true;
                    return p;
                }
            }
            void* p = nullptr;
            if (ob_host_alloc(want, &p) == OB_OK && p) {
                *cap = want;
                *pinned = true;
                return p;
            }
        }
        void* p = std::malloc(want);
        if (!p) throw std::bad_alloc();
        *cap = want;
        *pinned = false;
        return p;
    }
    void give_back(void* p, size_t cap, bool pinned) {
        if (!p) return;
        if (!pinned) {
            std::free(p);
            return;
        }
        std::lock_guard<std::mutex> lk(mx);
        free_blocks.emplace(cap, p);  // kept for reuse for the life of the process
    }
};
PinnedPool& pool() {
    static PinnedPool* p = new PinnedPool();  // intentionally leaked: outlives static destructors
    return *p;
}
}  // namespace

HostBuffer::HostBuffer(size_t bytes) { resize(bytes); }
HostBuffer::HostBuffer(const HostBuffer& o) {
    resize(o.n_);
    if (o.n_) std::memcpy(p_, o.p_, o.n_);
}
HostBuffer::HostBuffer(HostBuffer&& o) noexcept
    : p_(o.p_), n_(o.n_), cap_(o.cap_), pinned_(o.pinned_), arena_(std::move(o.arena_)) {
    o.p_ = nullptr;
    o.n_ = o.cap_ = 0;
}
HostBuffer& HostBuffer::operator=(const HostBuffer& o) {
    if (this != &o) {
        resize(o.n_);
        if (o.n_) std::memcpy(p_, o.p_, o.n_);
    }
    return *this;
}
HostBuffer& HostBuffer::operator=(HostBuffer&& o) noexcept {
    if (this != &o) {
        release();
        p_ = o.p_;
        n_ = o.n_;
        cap_ = o.cap_;
        pinned_ = o.pinned_;
        arena_ = std::move(o.arena_);
        o.p_ = nullptr;
        o.n_ = o.cap_ = 0;
    }
    return *this;
}
HostBuffer::~HostBuffer() { release(); }
void HostBuffer::release() {
    arena_.reset();  // the block goes back to the pool with its last user (buffers of a carve(), jobs in flight)
    p_ = nullptr;
    n_ = cap_ = 0;
}

std::vector<HostBuffer> HostBuffer::carve(const std::vector<size_t>& sizes) {
    auto align = [](size_t n) { return (n + 255) & ~static_cast<size_t>(255); };
    size_t total = 0;
    for (size_t n : sizes) total += align(n);
    std::vector<HostBuffer> out(sizes.size());
    if (total == 0) return out;
    size_t cap = 0;
    bool pinned = false;
    void* block = pool().acquire(total, &cap, &pinned);
    std::memset(block, 0, total);
    std::shared_ptr<void> arena(block, [cap, pinned](void* p) { pool().give_back(p, cap, pinned); });
    size_t off = 0;
    for (size_t i = 0; i < sizes.size(); ++i) {
        if (sizes[i] == 0) continue;
        out[i].p_ = static_cast<uint8_t*>(block) + off;
        out[i].n_ = sizes[i];
        out[i].cap_ = align(sizes[i]);
        out[i].pinned_ = pinned;
        out[i].arena_ = arena;
        off += align(sizes[i]);
    }
    return out;
}
void HostBuffer::resize(size_t bytes) {
    if (bytes > cap_ || p_ == nullptr) {
        release();
        if (bytes == 0) return;
        p_ = static_cast<uint8_t*>(pool().acquire(bytes, &cap_, &pinned_));
        const size_t cap = cap_;
        const bool pinned = pinned_;
        arena_ = std::shared_ptr<void>(p_, [cap, pinned](void* p) { pool().give_back(p, cap, pinned); });
    }
    n_ = bytes;
    if (n_) std::memset(p_, 0, n_);
}
bool HostBuffer::operator==(const HostBuffer& o) const {
    return n_ == o.n_ && (n_ == 0 || std::memcmp(p_, o.p_, n_) == 0);
}

// ---- Field ----
Field::Field(ChanFieldType tag, const std::vector<size_t>& shape) : tag_(tag), shape_(shape) {
    size_t n = field_type_size(tag);
    for (size_t d : shape) n *= d;
    buf_.resize(n);
}

Field::Field(ChanFieldType tag, const std::vector<size_t>& shape, HostBuffer&& storage)
    : tag_(tag), shape_(shape), buf_(std::move(storage)) {
    size_t n = field_type_size(tag);
    for (size_t d : shape) n *= d;
    if (buf_.size() != n) buf_.resize(n);
}

void Field::set_zero() {
    if (buf_.size()) std::memset(buf_.data(), 0, buf_.size());
}

// ---- default field sets ----
namespace {
using T = ChanFieldType;
struct Slot {
    const char* name;
    T type;
};
struct SlotSet {
    UDPProfileLidar profile;
    std::vector<Slot> slots;
};
const std::vector<SlotSet>& slot_sets() {
    using P = UDPProfileLidar;
    namespace F = ChanField;
    static const std::vector<SlotSet> sets = {
        {P::LEGACY,
         {{F::RANGE, T::UINT32}, {F::SIGNAL, T::UINT16}, {F::NEAR_IR, T::UINT16},
          {F::REFLECTIVITY, T::UINT8}, {F::FLAGS, T::UINT8}}},
        {P::RNG19_RFL8_SIG16_NIR16_DUAL,
         {{F::RANGE, T::UINT32}, {F::RANGE2, T::UINT32}, {F::SIGNAL, T::UINT16},
          {F::SIGNAL2, T::UINT16}, {F::REFLECTIVITY, T::UINT8}, {F::REFLECTIVITY2, T::UINT8},
          {F::FLAGS, T::UINT8}, {F::FLAGS2, T::UINT8}, {F::NEAR_IR, T::UINT16}, {F::WINDOW, T::UINT8}}},
        {P::RNG19_RFL8_SIG16_NIR16,
         {{F::RANGE, T::UINT32}, {F::SIGNAL, T::UINT16}, {F::REFLECTIVITY, T::UINT8},
          {F::FLAGS, T::UINT8}, {F::NEAR_IR, T::UINT16}, {F::WINDOW, T::UINT8}}},
        {P::RNG15_RFL8_NIR8,
         {{F::RANGE, T::UINT32}, {F::REFLECTIVITY, T::UINT8}, {F::NEAR_IR, T::UINT16}, {F::FLAGS, T::UINT8}}},
        {P::RNG15_RFL8_WIN8,
         {{F::RANGE, T::UINT32}, {F::REFLECTIVITY, T::UINT8}, {F::WINDOW, T::UINT8}, {F::FLAGS, T::UINT8}}},
        {P::FIVE_WORD_PIXEL,
         {{F::RAW32_WORD1, T::UINT32}, {F::RAW32_WORD2, T::UINT32}, {F::RAW32_WORD3, T::UINT32},
          {F::RAW32_WORD4, T::UINT32}, {F::RAW32_WORD5, T::UINT32}}},
        {P::FUSA_RNG15_RFL8_NIR8_DUAL,
         {{F::RANGE, T::UINT32}, {F::REFLECTIVITY, T::UINT8}, {F::NEAR_IR, T::UINT16},
          {F::RANGE2, T::UINT32}, {F::REFLECTIVITY2, T::UINT8}, {F::FLAGS, T::UINT8},
          {F::FLAGS2, T::UINT8}, {F::WINDOW, T::UINT8}}},
        {P::RNG15_RFL8_NIR8_DUAL,
         {{F::RANGE, T::UINT32}, {F::REFLECTIVITY, T::UINT8}, {F::NEAR_IR, T::UINT16},
          {F::RANGE2, T::UINT32}, {F::REFLECTIVITY2, T::UINT8}, {F::FLAGS, T::UINT8},
          {F::FLAGS2, T::UINT8}, {F::WINDOW, T::UINT8}}},
        {P::OFF, {}},
        {P::RNG15_RFL8_NIR8_ZONE16,
         {{F::RANGE, T::UINT32}, {F::REFLECTIVITY, T::UINT8}, {F::NEAR_IR, T::UINT16},
          {F::FLAGS, T::UINT8}, {F::ZONE_MASK, T::UINT16}, {F::WINDOW, T::UINT8}}},
        {P::RNG19_RFL8_SIG16_NIR16_ZONE16,
         {{F::RANGE, T::UINT32}, {F::SIGNAL, T::UINT16}, {F::REFLECTIVITY, T::UINT8},
          {F::FLAGS, T::UINT8}, {F::NEAR_IR, T::UINT16}, {F::ZONE_MASK, T::UINT16}, {F::WINDOW, T::UINT8}}},
        {P::RNG19_RFL8_SIG16_ZONE16_DUAL,
         {{F::RANGE, T::UINT32}, {F::RANGE2, T::UINT32}, {F::SIGNAL, T::UINT16},
          {F::SIGNAL2, T::UINT16}, {F::REFLECTIVITY, T::UINT8}, {F::REFLECTIVITY2, T::UINT8},
          {F::FLAGS, T::UINT8}, {F::FLAGS2, T::UINT8}, {F::ZONE_MASK, T::UINT16}, {F::WINDOW, T::UINT8}}},
        {P::RNG19_RFL8_SIG16_NIR16_RGB16,
         {{F::RANGE, T::UINT32}, {F::SIGNAL, T::UINT16}, {F::REFLECTIVITY, T::UINT8},
          {F::NEAR_IR, T::UINT16}, {F::RGB, T::FLOAT16}, {F::FLAGS, T::UINT8}}},
        {P::RNG19_RFL8_SIG16_NIR16_RGB16_DUAL,
         {{F::RANGE, T::UINT32}, {F::RANGE2, T::UINT32}, {F::SIGNAL, T::UINT16},
          {F::SIGNAL2, T::UINT16}, {F::REFLECTIVITY, T::UINT8}, {F::REFLECTIVITY2, T::UINT8},
          {F::NEAR_IR, T::UINT16}, {F::RGB, T::FLOAT16}, {F::FLAGS, T::UINT8}, {F::FLAGS2, T::UINT8}}},
    };
    return sets;
}
}  // namespace

LidarFrameFieldTypes get_field_types(UDPProfileLidar profile) {
    for (const auto& s : slot_sets()) {
        if (s.profile != profile) continue;
        LidarFrameFieldTypes out;
        for (const auto& sl : s.slots) {
            FieldType ft(sl.name, sl.type, {}, FieldClass::PIXEL_FIELD);
            if (ft.name == ChanField::RGB) ft.extra_dims.push_back(3);  // H x W x 3
            out.push_back(ft);
        }
        return out;
    }
    throw std::invalid_argument("Unknown lidar udp profile");
}

LidarFrameFieldTypes get_field_types(const DataFormat& format, const Version& fw) {
    LidarFrameFieldTypes out = get_field_types(format.udp_profile_lidar);
    const bool zone = format.udp_profile_lidar == UDPProfileLidar::RNG19_RFL8_SIG16_NIR16_ZONE16 ||
                      format.udp_profile_lidar == UDPProfileLidar::RNG15_RFL8_NIR8_ZONE16;
    // WINDOW only exists from firmware 3.2.0 (3.2.1 for the zone profiles)
    if (fw < Version{3, 2, 0} || (zone && fw < Version{3, 2, 1})) {
        auto it = std::find_if(out.begin(), out.end(),
                               [](const FieldType& f) { return f.name == ChanField::WINDOW; });
        if (it != out.end()) out.erase(it);
    }
    return out;
}

LidarFrameFieldTypes get_field_types(const SensorInfo& info) {
    return get_field_types(info.format, info.get_version());
}

// ---- LidarFrame ----
LidarFrame::LidarFrame() = default;

void LidarFrame::init_headers(size_t columns_per_packet) {
    timestamp_.assign(w, 0);
    measurement_id_.assign(w, 0);
    status_.assign(w, 0);
    const size_t np = columns_per_packet ? w / columns_per_packet : 0;
    packet_timestamp_.assign(np, 0);
    alert_flags_.assign(np, 0);
    body_to_world_ = Field(ChanFieldType::FLOAT64, {w, 4, 4});
    double* p = body_to_world_.get<double>();
    for (size_t c = 0; c < w; ++c)
        for (int i = 0; i < 4; ++i) p[c * 16 + static_cast<size_t>(i) * 5] = 1.0;  // identity
}

void LidarFrame::set_column_pose(int index, const mat4d& pose) {
    if (index < 0 || index >= static_cast<int>(w)) throw std::out_of_range("Column index out of range");
    std::memcpy(body_to_world_.get<double>() + static_cast<size_t>(index) * 16, pose.data(), 16 * sizeof(double));
}

mat4d LidarFrame::get_column_pose(int index) const {
    if (index < 0 || index >= static_cast<int>(w)) throw std::out_of_range("Column index out of range");
    mat4d out;
    std::memcpy(out.m.data(), body_to_world_.get<double>() + static_cast<size_t>(index) * 16, 16 * sizeof(double));
    return out;
}

namespace {
// mode 0: first, 1: last, 2: min, 3: max over the packets with at least one valid column
bool valid_packet_ts(const LidarFrame& f, int mode, uint64_t* out) {
    const auto status = f.status();
    const auto pts = f.packet_timestamp();
    const size_t np = pts.size();
    if (np == 0) return false;
    const size_t cpp = f.w / np;
    bool have = false;
    uint64_t t = 0;
    for (size_t k = 0; k < np; ++k) {
        const size_t i = mode == 1 ? np - 1 - k : k;
        bool any = false;
        for (size_t c = 0; c < cpp && !any; ++c) any = (status[i * cpp + c] & 1u) != 0;
        if (!any) continue;
        if (mode < 2) {
            *out = pts[i];
            return true;
        }
        t = !have ? pts[i] : (mode == 2 ? std::min(t, pts[i]) : std::max(t, pts[i]));
        have = true;
    }
    *out = t;
    return have;
}
uint64_t ts_or_throw(bool ok, uint64_t t) {
    if (!ok) throw std::runtime_error("No valid packets in LidarFrame");
    return t;
}
}  // namespace

uint64_t LidarFrame::get_first_valid_packet_timestamp() const {
    uint64_t t = 0;
    const bool ok = valid_packet_ts(*this, 0, &t);  // sequenced before t is read
    return ts_or_throw(ok, t);
}
uint64_t LidarFrame::get_last_valid_packet_timestamp() const {
    uint64_t t = 0;
    const bool ok = valid_packet_ts(*this, 1, &t);  // sequenced before t is read
    return ts_or_throw(ok, t);
}
uint64_t LidarFrame::get_min_valid_packet_timestamp() const {
    uint64_t t = 0;
    const bool ok = valid_packet_ts(*this, 2, &t);  // sequenced before t is read
    return ts_or_throw(ok, t);
}
uint64_t LidarFrame::get_max_valid_packet_timestamp() const {
    uint64_t t = 0;
    const bool ok = valid_packet_ts(*this, 3, &t);  // sequenced before t is read
    return ts_or_throw(ok, t);
}
uint64_t LidarFrame::get_first_valid_lidar_packet_timestamp() const {
    uint64_t t = 0;
    return valid_packet_ts(*this, 0, &t) ? t : 0;
}
uint64_t LidarFrame::get_last_valid_lidar_packet_timestamp() const {
    uint64_t t = 0;
    return valid_packet_ts(*this, 1, &t) ? t : 0;
}

int LidarFrame::get_first_valid_column() const {
    for (size_t i = 0; i < status_.size(); ++i)
        if ((status_[i] & 1u) > 0) return static_cast<int>(i);
    throw std::runtime_error("No valid columns in LidarFrame");
}

int LidarFrame::get_last_valid_column() const {
    for (int i = static_cast<int>(status_.size()) - 1; i >= 0; --i)
        if ((status_[static_cast<size_t>(i)] & 1u) > 0) return i;
    throw std::runtime_error("No valid columns in LidarFrame");
}

LidarFrame::LidarFrame(size_t h_, size_t w_, const LidarFrameFieldTypes& field_types,
                       size_t columns_per_packet)
    : w(w_), h(h_) {
    init_headers(columns_per_packet);
    // the initial fields share one page-locked block (see HostBuffer::carve)
    std::vector<std::vector<size_t>> shapes;
    std::vector<size_t> sizes;
    for (const auto& ft : field_types) {
        shapes.push_back(field_shape(ft.field_class, ft.extra_dims));
        size_t n = field_type_size(ft.element_type);
        for (size_t d : shapes.back()) n *= d;
        sizes.push_back(n);
    }
    std::vector<HostBuffer> bufs = HostBuffer::carve(sizes);
    for (size_t i = 0; i < field_types.size(); ++i) {
        const FieldType& ft = field_types[i];
        if (has_field(ft.name)) throw std::invalid_argument("Duplicated field '" + ft.name + "'");
        field_class_[ft.name] = ft.field_class;
        fields_.emplace(ft.name, Field(ft.element_type, shapes[i], std::move(bufs[i])));
    }
}

LidarFrame::LidarFrame(size_t h_, size_t w_, UDPProfileLidar profile, size_t columns_per_packet)
    : LidarFrame(h_, w_, get_field_types(profile), columns_per_packet) {}

LidarFrame::LidarFrame(std::shared_ptr<SensorInfo> info, const std::vector<FieldType>& field_types)
    : LidarFrame(info->format.pixels_per_column, info->format.columns_per_frame, field_types,
                 info->format.columns_per_packet) {
    sensor_info = std::move(info);
}

LidarFrame::LidarFrame(std::shared_ptr<SensorInfo> info)
    : LidarFrame(info, get_field_types(*info)) {}

LidarFrame::LidarFrame(const SensorInfo& info) : LidarFrame(std::make_shared<SensorInfo>(info)) {}

ThermalShutdownStatus LidarFrame::thermal_shutdown() const {
    return static_cast<ThermalShutdownStatus>(frame_status & 0x0f);
}
ShotLimitingStatus LidarFrame::shot_limiting() const {
    return static_cast<ShotLimitingStatus>((frame_status & 0xf0) >> 4);
}

bool LidarFrame::has_field(const std::string& name) const { return fields_.count(name) != 0; }

Field& LidarFrame::field(const std::string& name) {
    auto it = fields_.find(name);
    if (it == fields_.end()) throw std::invalid_argument("Invalid field for LidarFrame");
    return it->second;
}
const Field& LidarFrame::field(const std::string& name) const {
    auto it = fields_.find(name);
    if (it == fields_.end()) throw std::invalid_argument("Invalid field for LidarFrame");
    return it->second;
}

Field& LidarFrame::checked(const std::string& name, ChanFieldType tag) {
    Field& f = field(name);
    if (f.tag() != tag)
        throw std::invalid_argument("Accessed field at wrong type. Field is " + to_string(f.tag()) +
                                    " but was accessed as " + to_string(tag));
    return f;
}

Field& LidarFrame::add_field(const std::string& name, ChanFieldType type,
                             const std::vector<size_t>& extra_dims, FieldClass field_class) {
    if (has_field(name)) throw std::invalid_argument("Duplicated field '" + name + "'");
    field_class_[name] = field_class;
    return fields_.emplace(name, Field(type, field_shape(field_class, extra_dims))).first->second;
}

std::vector<size_t> LidarFrame::field_shape(FieldClass field_class, const std::vector<size_t>& extra_dims) const {
    std::vector<size_t> shape;
    switch (field_class) {
        case FieldClass::PIXEL_FIELD: shape = {h, w}; break;
        case FieldClass::COLUMN_FIELD: shape = {w}; break;
        case FieldClass::PACKET_FIELD: shape = {packet_timestamp_.size()}; break;
        default: break;
    }
    shape.insert(shape.end(), extra_dims.begin(), extra_dims.end());
    return shape;
}

void FusedCloud::reserve(size_t fh, size_t fw, int n_returns) {
    const size_t n = fh * fw;
    const size_t xb = n * 3 * (lut_is_f64 ? 8 : 4);
    const size_t rb = pixel_shift_by_row.empty() ? 0 : n * 4;
    bool ok = true;
    for (int r = 0; r < 2; ++r) {
        const bool on = r < n_returns;
        ok &= xyz[r].size() == (on ? xb : 0) && range_destaggered[r].size() == (on ? rb : 0);
    }
    if (ok) return;
    std::vector<size_t> sizes;
    for (int r = 0; r < 2; ++r) {
        sizes.push_back(r < n_returns ? xb : 0);
        sizes.push_back(r < n_returns ? rb : 0);
    }
    std::vector<HostBuffer> bufs = HostBuffer::carve(sizes);
    for (int r = 0; r < 2; ++r) {
        xyz[r] = std::move(bufs[2 * r]);
        range_destaggered[r] = std::move(bufs[2 * r + 1]);
    }
}

Field& LidarFrame::add_field(const FieldType& t) {
    return add_field(t.name, t.element_type, t.extra_dims, t.field_class);
}

Field LidarFrame::del_field(const std::string& name) {
    auto it = fields_.find(name);
    if (it == fields_.end())
        throw std::invalid_argument("Attempted deleting non existing field '" + name + "'");
    Field out = std::move(it->second);
    fields_.erase(it);
    field_class_.erase(name);
    return out;
}

// FieldType of one field: element type, trailing dimensions, class (lidar_frame.cpp:545-547)
FieldType LidarFrame::field_type(const std::string& name) const {
    const Field& f = field(name);  // throws std::invalid_argument("Invalid field for LidarFrame")
    const auto& shp = f.shape();
    const FieldClass fc = field_class_.count(name) ? field_class_.at(name) : FieldClass::PIXEL_FIELD;
    const size_t base = fc == FieldClass::PIXEL_FIELD ? 2 : (fc == FieldClass::FRAME_FIELD ? 0 : 1);
    std::vector<size_t> extra(shp.begin() + std::min(base, shp.size()), shp.end());
    return FieldType(name, f.tag(), extra, fc);
}

LidarFrameFieldTypes LidarFrame::field_types() const {
    LidarFrameFieldTypes out;
    for (const auto& kv : fields_) out.push_back(field_type(kv.first));
    return out;
}

bool LidarFrame::complete(ColumnWindow window) const {
    const size_t a = window.first, b = window.second;
    auto valid = [&](size_t i) { return (status_[i] & 0x01) != 0; };
    if (a <= b) {
        for (size_t i = a; i <= b && i < w; ++i)
            if (!valid(i)) return false;
        return true;
    }
    for (size_t i = 0; i <= b && i < w; ++i)
        if (!valid(i)) return false;
    for (size_t i = a; i < w; ++i)
        if (!valid(i)) return false;
    return true;
}

bool LidarFrame::complete() const {
    if (sensor_info) return complete(sensor_info->format.column_window);
    return complete({0, static_cast<uint16_t>(w - 1)});
}

bool LidarFrame::equals(const LidarFrame& o) const {
    return w == o.w && h == o.h && frame_id == o.frame_id && frame_status == o.frame_status &&
           shutdown_countdown == o.shutdown_countdown &&
           shot_limiting_countdown == o.shot_limiting_countdown && fields_ == o.fields_ &&
           timestamp_ == o.timestamp_ && measurement_id_ == o.measurement_id_ &&
           status_ == o.status_ && packet_timestamp_ == o.packet_timestamp_ &&
           alert_flags_ == o.alert_flags_ && body_to_world_ == o.body_to_world_;  // lidar_frame.cpp:1016
}

bool operator==(const LidarFrame& a, const LidarFrame& b) { return a.equals(b); }

}  // namespace core
}  // namespace sdk
}  // namespace ouster
