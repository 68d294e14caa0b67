// lidar_frame.h -- LidarFrame (a.k.a. LidarScan), Field, FrameBatcher (a.k.a. ScanBatcher) and the
// destagger entry points (mirrors ouster_core/include/ouster/core/lidar_frame.h:36-1160 and
// field.h for the hot path).
//
// LidarFrame stays a host container with the reference's layout (one dense row-major buffer per
// named field).  FrameBatcher keeps the reference's per-packet state machine on the host
// (ordering by frame id, cache, init-id changes, header bookkeeping) and moves the per-pixel
// work -- field decode, zero fill, and optionally destagger + XYZ -- to one fused GPU launch
// when a frame completes.  Pixel fields are therefore materialised when batch() returns true
// (the moment the reference declares the frame "ready to use"), or on flush().
#pragma once
#include <cstddef>
#include <cstdint>
#include <deque>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "ouster/core/chanfield.h"
#include "ouster/core/packet.h"
#include "ouster/core/sensor_info.h"
#include "ouster/core/typedefs.h"
#include "ouster/core/types.h"
#include "ouster/core/visibility.h"

struct ob_decoder;
struct ob_decode_job;
struct ob_stream;
struct ob_lut;

namespace ouster {
namespace sdk {
namespace core {

enum class FieldClass { NONE = 0, PIXEL_FIELD = 1, COLUMN_FIELD = 2, PACKET_FIELD = 3, FRAME_FIELD = 4 };

/// Describes one field of a LidarFrame (lidar_frame.h:76-122).
struct OUSTER_API_CLASS FieldType {
    std::string name;
    ChanFieldType element_type{ChanFieldType::VOID};
    std::vector<size_t> extra_dims;
    FieldClass field_class{FieldClass::PIXEL_FIELD};
    FieldType() = default;
    FieldType(const std::string& name_, ChanFieldType element_type_,
              const std::vector<size_t>& extra_dims_ = {}, FieldClass field_class_ = FieldClass::PIXEL_FIELD)
        : name(name_), element_type(element_type_), extra_dims(extra_dims_), field_class(field_class_) {}
    bool operator==(const FieldType& o) const {
        return name == o.name && element_type == o.element_type && extra_dims == o.extra_dims &&
               field_class == o.field_class;
    }
};
using LidarFrameFieldTypes = std::vector<FieldType>;

/// Zero-initialised host byte buffer.  When a CUDA device is present the bytes come from a
/// process-wide pool of page-locked blocks so that H2D/D2H of fields run at full PCIe rate;
/// without a device it is plain calloc memory (the reference uses calloc, field.cpp:252-254).
class OUSTER_API_CLASS HostBuffer {
   public:
    HostBuffer() = default;
    OUSTER_API_FUNCTION explicit HostBuffer(size_t bytes);
    OUSTER_API_FUNCTION HostBuffer(const HostBuffer& o);
    OUSTER_API_FUNCTION HostBuffer(HostBuffer&& o) noexcept;
    OUSTER_API_FUNCTION HostBuffer& operator=(const HostBuffer& o);
    OUSTER_API_FUNCTION HostBuffer& operator=(HostBuffer&& o) noexcept;
    OUSTER_API_FUNCTION ~HostBuffer();
    uint8_t* data() { return p_; }
    const uint8_t* data() const { return p_; }
    size_t size() const { return n_; }
    OUSTER_API_FUNCTION void resize(size_t bytes);  ///< contents are zeroed
    /// Shares ownership of the memory block behind this buffer: an asynchronous writer (a decode job in
    /// flight) holds it so that the block cannot return to the pool -- and be handed to another frame --
    /// before the writer is done, even if the buffer's owner is destroyed, moved or resized meanwhile.
    std::shared_ptr<void> keepalive() const { return arena_; }
    OUSTER_API_FUNCTION bool operator==(const HostBuffer& o) const;
    /// Zero-initialised buffers of the given sizes laid out back to back (256-byte aligned) in ONE
    /// block, so that a device->host transfer of all of them is a single copy (ob_decode_job_submit
    /// merges outputs that are adjacent in memory).  Each buffer keeps the block alive.
    OUSTER_API_FUNCTION static std::vector<HostBuffer> carve(const std::vector<size_t>& sizes);

   private:
    void release();
    uint8_t* p_{nullptr};
    size_t n_{0};
    size_t cap_{0};
    bool pinned_{false};
    std::shared_ptr<void> arena_;  ///< owner of the block p_ lies in (shared with other buffers after carve())
};

/// Typed dense buffer (field.h:828+).  Zero-initialised like the reference's calloc (field.cpp:254).
class OUSTER_API_CLASS Field {
   public:
    Field() = default;
    Field(ChanFieldType tag, const std::vector<size_t>& shape);
    /// adopts `storage` (resized to the field's byte size when it does not match)
    Field(ChanFieldType tag, const std::vector<size_t>& shape, HostBuffer&& storage);
    ChanFieldType tag() const { return tag_; }
    const std::vector<size_t>& shape() const { return shape_; }
    size_t element_size() const { return field_type_size(tag_); }
    size_t bytes() const { return buf_.size(); }
    size_t size() const { return element_size() ? buf_.size() / element_size() : 0; }
    void* get() { return buf_.data(); }
    const void* get() const { return buf_.data(); }
    std::shared_ptr<void> keepalive() const { return buf_.keepalive(); }  ///< see HostBuffer::keepalive
    template <typename T>
    T* get() { return reinterpret_cast<T*>(buf_.data()); }
    template <typename T>
    const T* get() const { return reinterpret_cast<const T*>(buf_.data()); }
    void set_zero();
    bool operator==(const Field& o) const { return tag_ == o.tag_ && shape_ == o.shape_ && buf_ == o.buf_; }
    bool operator!=(const Field& o) const { return !(*this == o); }

   private:
    ChanFieldType tag_{ChanFieldType::VOID};
    std::vector<size_t> shape_;
    HostBuffer buf_;
};

/// 1-D header view with Eigen-like element access.
template <typename T>
class HeaderRef {
   public:
    HeaderRef(T* d, size_t n) : d_(d), n_(n) {}
    /// view-of-mutable converts to view-of-const
    template <typename U, typename = typename std::enable_if<std::is_same<const U, T>::value>::type>
    HeaderRef(const HeaderRef<U>& o) : d_(o.data()), n_(o.size()) {}
    T* data() const { return d_; }
    size_t rows() const { return n_; }
    size_t size() const { return n_; }
    T& operator[](size_t i) const { return d_[i]; }
    T& operator()(size_t i) const { return d_[i]; }
    void setZero() const { for (size_t i = 0; i < n_; ++i) d_[i] = T{}; }
    size_t count() const {  ///< number of non-zero entries (Eigen's .count())
        size_t c = 0;
        for (size_t i = 0; i < n_; ++i) c += (d_[i] != T{});
        return c;
    }

   private:
    T* d_;
    size_t n_;
};

class OUSTER_API_CLASS LidarFrame {
   public:
    template <typename T>
    using Header = HeaderRef<T>;

    size_t w{0};
    size_t h{0};
    uint64_t frame_status{0};
    uint8_t shutdown_countdown{0};
    uint8_t shot_limiting_countdown{0};
    int64_t frame_id{-1};
    std::shared_ptr<SensorInfo> sensor_info;

    OUSTER_API_FUNCTION LidarFrame();
    OUSTER_API_FUNCTION explicit LidarFrame(std::shared_ptr<SensorInfo> sensor_info);
    OUSTER_API_FUNCTION explicit LidarFrame(const SensorInfo& sensor_info);
    OUSTER_API_FUNCTION LidarFrame(std::shared_ptr<SensorInfo> sensor_info,
                                   const std::vector<FieldType>& field_types);
    OUSTER_API_FUNCTION LidarFrame(size_t h, size_t w, UDPProfileLidar profile,
                                   size_t columns_per_packet = DEFAULT_COLUMNS_PER_PACKET);
    OUSTER_API_FUNCTION LidarFrame(size_t h, size_t w, const LidarFrameFieldTypes& field_types,
                                   size_t columns_per_packet = DEFAULT_COLUMNS_PER_PACKET);

    OUSTER_API_FUNCTION ThermalShutdownStatus thermal_shutdown() const;
    OUSTER_API_FUNCTION ShotLimitingStatus shot_limiting() const;

    /// Typed 2-D view of a pixel field.  Throws std::invalid_argument on unknown field or
    /// mismatched element type.
    template <typename T>
    ArrayRef<T> field(const std::string& name) {
        Field& f = checked(name, FieldTag<T>::tag);
        return ArrayRef<T>(f.get<T>(), h, f.shape().size() > 1 ? f.size() / h : 1);
    }
    template <typename T>
    ArrayRef<const T> field(const std::string& name) const {
        const Field& f = const_cast<LidarFrame*>(this)->checked(name, FieldTag<T>::tag);
        return ArrayRef<const T>(f.get<T>(), h, f.shape().size() > 1 ? f.size() / h : 1);
    }
    OUSTER_API_FUNCTION Field& field(const std::string& name);
    OUSTER_API_FUNCTION const Field& field(const std::string& name) const;
    OUSTER_API_FUNCTION bool has_field(const std::string& name) const;
    OUSTER_API_FUNCTION Field& add_field(const std::string& name, ChanFieldType type,
                                         const std::vector<size_t>& extra_dims = {},
                                         FieldClass field_class = FieldClass::PIXEL_FIELD);
    OUSTER_API_FUNCTION Field& add_field(const FieldType& type);
    OUSTER_API_FUNCTION Field del_field(const std::string& name);
    OUSTER_API_FUNCTION FieldType field_type(const std::string& name) const;
    OUSTER_API_FUNCTION LidarFrameFieldTypes field_types() const;
    const std::map<std::string, Field>& fields() const { return fields_; }
    std::map<std::string, Field>& fields() { return fields_; }

    Header<uint64_t> timestamp() { return {timestamp_.data(), timestamp_.size()}; }
    Header<const uint64_t> timestamp() const { return {timestamp_.data(), timestamp_.size()}; }
    Header<uint16_t> measurement_id() { return {measurement_id_.data(), measurement_id_.size()}; }
    Header<const uint16_t> measurement_id() const { return {measurement_id_.data(), measurement_id_.size()}; }
    Header<uint32_t> status() { return {status_.data(), status_.size()}; }
    Header<const uint32_t> status() const { return {status_.data(), status_.size()}; }
    Header<uint64_t> packet_timestamp() { return {packet_timestamp_.data(), packet_timestamp_.size()}; }
    Header<const uint64_t> packet_timestamp() const { return {packet_timestamp_.data(), packet_timestamp_.size()}; }
    Header<uint8_t> alert_flags() { return {alert_flags_.data(), alert_flags_.size()}; }
    Header<const uint8_t> alert_flags() const { return {alert_flags_.data(), alert_flags_.size()}; }

    /// Per-column poses, w x 4 x 4 doubles, identity on construction (lidar_frame.h:732-736,
    /// lidar_frame.cpp:350-358); pose() is the deprecated spelling.
    Field& body_to_world() { return body_to_world_; }
    const Field& body_to_world() const { return body_to_world_; }
    Field& pose() { return body_to_world_; }
    const Field& pose() const { return body_to_world_; }
    /// throw std::out_of_range("Column index out of range") (lidar_frame.cpp:959-980)
    OUSTER_API_FUNCTION void set_column_pose(int index, const mat4d& pose);
    OUSTER_API_FUNCTION mat4d get_column_pose(int index) const;
    /// first / last column with status & 1; throw std::runtime_error("No valid columns in
    /// LidarFrame") (lidar_frame.cpp:907-925)
    OUSTER_API_FUNCTION int get_first_valid_column() const;
    OUSTER_API_FUNCTION int get_last_valid_column() const;

    /// Host timestamp of the first / last / earliest / latest lidar packet that carries at least one
    /// valid column (lidar_frame.cpp:643-790, lidar packets only -- IMU and zone packets are outside
    /// this path).  The `_lidar_` spellings return 0 when there is none; the others throw
    /// std::runtime_error("No valid packets in LidarFrame").
    OUSTER_API_FUNCTION uint64_t get_first_valid_packet_timestamp() const;
    OUSTER_API_FUNCTION uint64_t get_last_valid_packet_timestamp() const;
    OUSTER_API_FUNCTION uint64_t get_min_valid_packet_timestamp() const;
    OUSTER_API_FUNCTION uint64_t get_max_valid_packet_timestamp() const;
    OUSTER_API_FUNCTION uint64_t get_first_valid_lidar_packet_timestamp() const;
    OUSTER_API_FUNCTION uint64_t get_last_valid_lidar_packet_timestamp() const;
    /// number of lidar packets per frame (lidar_frame.cpp:1008-1010)
    size_t packet_count() const { return packet_timestamp_.size(); }

    /// true when every column inside the window carries status & 1 (lidar_frame.cpp:421-446)
    OUSTER_API_FUNCTION bool complete(ColumnWindow window) const;
    OUSTER_API_FUNCTION bool complete() const;

    OUSTER_API_FUNCTION bool equals(const LidarFrame& other) const;

   private:
    Field& checked(const std::string& name, ChanFieldType tag);
    void init_headers(size_t columns_per_packet);
    std::vector<size_t> field_shape(FieldClass field_class, const std::vector<size_t>& extra_dims) const;
    std::map<std::string, Field> fields_;
    std::map<std::string, FieldClass> field_class_;
    std::vector<uint64_t> timestamp_;
    std::vector<uint16_t> measurement_id_;
    std::vector<uint32_t> status_;
    std::vector<uint64_t> packet_timestamp_;
    std::vector<uint8_t> alert_flags_;
    Field body_to_world_;
};

OUSTER_API_FUNCTION bool operator==(const LidarFrame& a, const LidarFrame& b);
inline bool operator!=(const LidarFrame& a, const LidarFrame& b) { return !(a == b); }

/// Default field set of a profile / sensor (lidar_frame.cpp:228-256, 1038-1117): WINDOW is
/// dropped for firmware older than 3.2.0.
OUSTER_API_FUNCTION LidarFrameFieldTypes get_field_types(UDPProfileLidar profile);
OUSTER_API_FUNCTION LidarFrameFieldTypes get_field_types(const DataFormat& format, const Version& fw);
OUSTER_API_FUNCTION LidarFrameFieldTypes get_field_types(const SensorInfo& info);

/// Field-level destagger (field.h:920, src/field.cpp:329-337): dispatches on the element type;
/// 2-D integer/float fields only -- other types yield a zero-filled field, as in the reference
/// (impl/lidar_frame_impl.h:143-161).  Throws like destagger_into<T>.
OUSTER_API_FUNCTION Field destagger(const SensorInfo& info, const Field& field, bool inverse = false);

/// Timestamp of the staggered column a destaggered pixel came from (lidar_frame.h:955,
/// src/lidar_frame.cpp:893-905).  Throws std::invalid_argument("row or column is out of range").
OUSTER_API_FUNCTION uint64_t column_timestamp_at_destaggered_pixel(
    size_t row, size_t col, const std::vector<int>& pixel_shift_by_row,
    const HeaderRef<const uint64_t>& column_timestamps);

/// Optional fused products of FrameBatcher's GPU pass (an extension over the reference API):
/// when attached, the same launch that decodes the frame also writes the XYZ of every return
/// (lut(frame), xyzlut.h:145-150) and the destaggered range images
/// (destagger<uint32_t>(info, range), impl/lidar_frame_impl.h:875-884).
struct OUSTER_API_CLASS FusedCloud {
    std::shared_ptr<ob_lut> lut;           ///< float or double device LUT (see XYZLutT::device_lut())
    bool lut_is_f64{false};
    std::vector<int> pixel_shift_by_row;   ///< empty: no destaggered range
    HostBuffer xyz[2];                     ///< (h*w) x 3 of float|double per return, staggered order
    HostBuffer range_destaggered[2];       ///< h x w uint32 per return
    /// (Re)allocate the outputs of n_returns returns for an h x w frame in one page-locked block;
    /// no-op when they already have the right sizes.  FrameBatcher calls this before each launch.
    OUSTER_API_FUNCTION void reserve(size_t h, size_t w, int n_returns);
    const float* xyz_f32(int r) const { return reinterpret_cast<const float*>(xyz[r].data()); }
    const double* xyz_f64(int r) const { return reinterpret_cast<const double*>(xyz[r].data()); }
    const uint32_t* rd(int r) const { return reinterpret_cast<const uint32_t*>(range_destaggered[r].data()); }
};

class OUSTER_API_CLASS FrameBatcher {
   public:
    PacketFormat pf;  ///< The packet format object used for decoding

    OUSTER_API_FUNCTION explicit FrameBatcher(const SensorInfo& info);
    OUSTER_API_FUNCTION explicit FrameBatcher(const std::shared_ptr<SensorInfo>& info);
    OUSTER_API_FUNCTION ~FrameBatcher();
    FrameBatcher(const FrameBatcher&) = delete;
    FrameBatcher& operator=(const FrameBatcher&) = delete;

    /// Add a packet to the frame; returns true when the frame is ready to use.
    /// Throws std::invalid_argument("unexpected frame dimensions"),
    /// std::invalid_argument("unexpected frame columns_per_packet: N"),
    /// std::runtime_error("32-bit frame id did not increase since the last frame").
    OUSTER_API_FUNCTION bool batch(const Packet& packet, LidarFrame& lidar_frame);
    OUSTER_API_FUNCTION bool operator()(const Packet& packet, LidarFrame& lidar_frame);
    /// Raw-buffer form of batch() (no Packet copy): buf holds one lidar packet.
    OUSTER_API_FUNCTION bool batch(const uint8_t* buf, size_t size, uint64_t host_timestamp,
                                   LidarFrame& lidar_frame);
    OUSTER_API_FUNCTION void reset();
    OUSTER_API_FUNCTION size_t batched_packets() const;
    OUSTER_API_FUNCTION size_t dropped_packets() const;
    OUSTER_API_FUNCTION void set_max_cache_size(size_t max_cache_size);
    OUSTER_API_FUNCTION size_t get_max_cache_size() const;

    // ---- B200 extensions ----
    /// Decode whatever has been staged for the current (incomplete) frame into lidar_frame.
    OUSTER_API_FUNCTION void flush(LidarFrame& lidar_frame);
    /// Attach fused XYZ / destaggered-range outputs (nullptr detaches).
    OUSTER_API_FUNCTION void set_fused_cloud(FusedCloud* cloud);
    /// Header-only batching: keep the state machine and the per-column / per-packet headers but
    /// skip the pixel decode (no GPU work; pixel fields are left untouched).
    OUSTER_API_FUNCTION void set_headers_only(bool on);
    /// Device-resident pixel outputs (the f-4 row of SURVEY 8f: results that stay in HBM for the next
    /// GPU consumer): fields named here are decoded straight into the given DEVICE buffers
    /// (h*w elements of the field's type each) instead of the LidarFrame's host storage, and, with a
    /// fused cloud attached, XYZ / destaggered range go to `xyz` / `range_destaggered` (device, per
    /// return; null = the FusedCloud's host buffers).  Headers still land in the LidarFrame.  The
    /// buffers must stay valid until batch() has returned true (or wait() with a pipeline).
    struct DeviceOutputs {
        std::vector<std::pair<std::string, void*>> fields;
        void* xyz[2]{nullptr, nullptr};
        uint32_t* range_destaggered[2]{nullptr, nullptr};
    };
    OUSTER_API_FUNCTION void set_device_outputs(const DeviceOutputs* outputs);  ///< nullptr detaches (copied)
    /// Kernel launches issued by this batcher.
    OUSTER_API_FUNCTION size_t gpu_launches() const;
    /// Cumulative host-side time of the calling thread, by phase (nanoseconds).  ns_burst is the
    /// total spent inside batch_burst(); ns_upload_wait (waiting for zero-copy uploads to leave the
    /// caller's memory) and ns_submit (decode table, launch, D2H enqueue) are parts of it; ns_wait
    /// is time blocked on finished GPU passes (wait(), ring reuse, synchronous mode).
    struct Stats {
        uint64_t ns_burst{0}, ns_upload_wait{0}, ns_submit{0}, ns_wait{0};
        size_t frames{0};
    };
    OUSTER_API_FUNCTION const Stats& stats() const;
    /// Feed `n` packets laid out `stride` bytes apart (host_timestamps[i] belongs to packet i) until
    /// a frame completes; returns the number of packets consumed.  When the burst lies in
    /// page-locked (cudaHostAlloc / cudaHostRegister) or managed memory the copy engine reads the
    /// packets where they are (no staging memcpy); the memory may be reused as soon as the call returns.
    /// The host state machine parses the headers with the CPU: plain device memory is rejected
    /// (std::invalid_argument "batch_burst needs host-readable packets ...").
    OUSTER_API_FUNCTION size_t batch_burst(const uint8_t* packets, size_t n, size_t stride, size_t size,
                                           const uint64_t* host_timestamps, LidarFrame& lidar_frame,
                                           bool& complete);
    /// Frames in flight.  1 (default): batch() returns true with the frame materialised, like the
    /// reference.  n >= 2: batch() returns true as soon as the frame's GPU pass is *submitted*; the
    /// caller batches the next frame into another LidarFrame (and FusedCloud) meanwhile and calls
    /// wait(frame) before reading pixel fields / fused outputs.  Column and packet headers are
    /// host-written and valid immediately.  The in-flight job shares ownership of the page-locked blocks
    /// it writes (fields and fused outputs): a LidarFrame (or FusedCloud) that is destroyed, moved or resized
    /// before wait(frame) does not send the device->host copy into memory of another frame -- its results
    /// are simply lost with the old blocks.  See FramePipeline (frame_pipeline.h) for a ready-made ring of
    /// frames.
    OUSTER_API_FUNCTION void set_pipeline_depth(size_t n);
    OUSTER_API_FUNCTION size_t pipeline_depth() const;
    /// Block until the GPU pass that fills lidar_frame has landed (no-op when none is pending).
    OUSTER_API_FUNCTION void wait(const LidarFrame& lidar_frame);
    OUSTER_API_FUNCTION void wait_all();

   private:
    struct CachedPacket {
        std::vector<uint8_t> buf;
        uint64_t host_timestamp;
        uint64_t seq;
    };
    size_t max_cache_size_{4};
    uint16_t next_valid_m_id_{0};
    std::deque<CachedPacket> cache_;
    uint64_t cache_seq_{0};
    int64_t finished_frame_id_{-1};
    int64_t last_frame_id_{-1};
    int64_t last_init_id_;
    std::shared_ptr<SensorInfo> sensor_info_;
    bool reset_frame_{true};
    size_t expected_lidar_packets_;
    size_t batched_lidar_packets_{0};
    size_t dropped_packets_{0};

    // staging of the frame in flight
    struct Staging;
    std::unique_ptr<Staging> stg_;
    FusedCloud* fused_{nullptr};
    bool headers_only_{false};
    bool dev_out_on_{false};
    DeviceOutputs dev_out_;
    size_t launches_{0};
    int n_returns_{0};
    Stats stats_;

    bool batch_impl(const uint8_t* buf, size_t size, uint64_t host_ts, LidarFrame& f);
    void cache_packet(const uint8_t* buf, size_t size, uint64_t host_ts);
    size_t cache_top() const;
    void batch_lidar_packet(const uint8_t* buf, uint64_t host_ts, LidarFrame& f);
    void start_frame(int64_t f_id, const uint8_t* packet_buf, LidarFrame& f);
    void finalize_frame(LidarFrame& f);
    bool handle_init_id_change(const uint8_t* buf, size_t size, uint64_t host_ts, LidarFrame& f);
    bool batch_with_caching(const uint8_t* buf, size_t size, uint64_t host_ts, LidarFrame& f);
    bool check_frame_complete(const LidarFrame& f) const;
    void decode_staged(LidarFrame& f);
    std::vector<std::string> ensure_decoder(LidarFrame& f);
    void upload_runs(LidarFrame& f);
    void settle_user_uploads();
};

/// Deprecated spellings kept by the reference (lidar_frame.h:1157-1160).
using LidarScan = LidarFrame;
using LidarScanFieldTypes = LidarFrameFieldTypes;
using ScanBatcher = FrameBatcher;

}  // namespace core
}  // namespace sdk
}  // namespace ouster

#include "ouster/core/impl/lidar_frame_impl.h"
