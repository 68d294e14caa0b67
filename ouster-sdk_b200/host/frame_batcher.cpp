// frame_batcher.cpp -- FrameBatcher (ScanBatcher): the reference's per-packet state machine on the
// host (ouster_core/src/lidar_frame.cpp:1248-1267, 1530-1576, 1698-1959) with the per-pixel work
// deferred to one fused GPU launch per frame (ob_decode_frames).
//
// Host side, per packet (cheap: 16 column headers): frame-id ordering / cache / init-id handling,
// per-packet and per-column headers written straight into the LidarFrame, and a column map
// col_src[w] that records, with the same sequencing as parse_by_block / parse_by_col /
// zero_fields, which packet column (or "zero") is the last operation applied to every frame
// column.  The packet bytes are appended to a pinned staging buffer.
// Device side, when the frame is finalized: decode of every field of every pixel + zero fill
// (+ optional destagger and XYZ), reading the staged packets once.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include "ouster/core/b200_runtime.h"
#include "ouster/core/lidar_frame.h"

namespace ouster {
namespace sdk {
namespace core {

namespace {
constexpr size_t kUploadChunk = 32;  // packets per early upload of a zero-copy burst (~1 MB)
inline uint64_t now_ns() {
    return static_cast<uint64_t>(std::chrono::duration_cast<std::chrono::nanoseconds>(
                                     std::chrono::steady_clock::now().time_since_epoch())
                                     .count());
}
struct ScopedNs {  // adds the lifetime of the object to a counter
    uint64_t& acc;
    uint64_t t0;
    explicit ScopedNs(uint64_t& a) : acc(a), t0(now_ns()) {}
    ~ScopedNs() { acc += now_ns() - t0; }
};
}  // namespace

// One frame in flight: an ob_decode_job (device packet slots + output slab + events) on its own
// stream, and a page-locked bounce buffer for packets that arrive in pageable memory.
struct FrameBatcher::Staging {
    struct Job {
        ob_decode_job* job{nullptr};
        ob_stream* stream{nullptr};
        uint8_t* bounce{nullptr};
        size_t bounce_cap{0};  // packets
        bool bounce_cuda{false};
        const LidarFrame* owner{nullptr};  // frame whose outputs the job is writing
        std::vector<std::shared_ptr<void>> keep;  // the host blocks it writes, held until the job has landed
        bool user_uploads{false};          // uploads that read caller memory are in flight
    };
    // pending upload: `count` packets, src_stride apart, into slots [first, first+count)
    struct Run {
        const uint8_t* src;
        size_t src_stride, first, count;
        bool user;  // reads caller-owned (page-locked / device) memory
    };
    std::vector<Job> jobs;
    size_t cur{0};
    size_t depth{1};
    size_t stride{0};   // bytes per slot (multiple of 16)
    size_t n_slots{0};  // packets accepted into the frame in flight
    std::vector<Run> runs;
    std::vector<int32_t> col_src;
    // the burst being fed (batch_burst): zero-copy window when its memory is DMA-able
    const uint8_t* burst_begin{nullptr};
    const uint8_t* burst_end{nullptr};
    size_t burst_stride{0};
    ob_decoder* dec{nullptr};
    std::string signature;
    int device{0};

    ~Staging() {
        for (Job& j : jobs) {
            if (j.job) ob_decode_job_destroy(j.job);
            if (j.stream) ob_stream_destroy(j.stream);
            if (j.bounce) {
                if (j.bounce_cuda) ob_host_free(j.bounce);
                else std::free(j.bounce);
            }
        }
        if (dec) ob_decoder_destroy(dec);
    }
    Job& job() { return jobs[cur]; }
    void reserve_bounce(size_t packets) {
        Job& j = job();
        if (packets <= j.bounce_cap) return;
        const size_t cap = std::max<size_t>(packets, j.bounce_cap ? j.bounce_cap * 2 : 16);
        void* p = nullptr;
        // pinned memory when a device exists; plain memory keeps the host state machine usable
        // on machines without one (decode itself still fails loudly there)
        const bool cuda = ob_device_count() > 0;
        if (cuda) b200::check(ob_host_alloc(cap * stride, &p));
        else p = std::malloc(cap * stride);
        if (!p) throw std::runtime_error("out of memory staging lidar packets");
        if (j.bounce) {
            // pending runs may point into the old block
            for (Run& r : runs)
                if (!r.user) r.src = static_cast<uint8_t*>(p) + (r.src - j.bounce);
            std::memcpy(p, j.bounce, j.bounce_cap * stride);
            if (j.bounce_cuda) ob_host_free(j.bounce);
            else std::free(j.bounce);
        }
        j.bounce = static_cast<uint8_t*>(p);
        j.bounce_cuda = cuda;
        j.bounce_cap = cap;
    }
    void add_run(const uint8_t* src, size_t src_stride, size_t slot, bool user) {
        if (!runs.empty()) {
            Run& r = runs.back();
            if (r.user == user && r.first + r.count == slot &&
                (r.count == 1 || r.src_stride == src_stride) && r.src + r.count * src_stride == src) {
                r.src_stride = src_stride;
                r.count++;
                return;
            }
        }
        runs.push_back(Run{src, src_stride, slot, 1, user});
    }
};

FrameBatcher::FrameBatcher(const std::shared_ptr<SensorInfo>& info)
    : pf(get_format(*info)), last_init_id_(info->init_id), sensor_info_(info) {
    if (info->format.columns_per_packet == 0)
        throw std::invalid_argument("unexpected columns_per_packet: 0");
    if (info->format.pixels_per_column == 0)
        throw std::invalid_argument("unexpected pixels_per_column: 0");
    expected_lidar_packets_ = static_cast<size_t>(info->format.lidar_packets_per_frame());
    stg_ = std::make_unique<Staging>();
    stg_->stride = (pf.lidar_packet_size + 15) & ~static_cast<size_t>(15);
    stg_->col_src.assign(info->format.columns_per_frame, -1);
    stg_->device = b200::device();
    stg_->jobs.resize(1);
}

FrameBatcher::FrameBatcher(const SensorInfo& info) : FrameBatcher(std::make_shared<SensorInfo>(info)) {}

FrameBatcher::~FrameBatcher() {
    try {
        wait_all();
    } catch (...) {
    }
}

size_t FrameBatcher::batched_packets() const { return batched_lidar_packets_; }
size_t FrameBatcher::dropped_packets() const { return dropped_packets_; }
size_t FrameBatcher::gpu_launches() const { return launches_; }
const FrameBatcher::Stats& FrameBatcher::stats() const { return stats_; }
void FrameBatcher::set_fused_cloud(FusedCloud* cloud) { fused_ = cloud; }

void FrameBatcher::set_max_cache_size(size_t n) {
    if (n == 0) throw std::invalid_argument("max_cache_size must be > 0");
    max_cache_size_ = n;
}
size_t FrameBatcher::get_max_cache_size() const { return max_cache_size_; }

void FrameBatcher::reset() {
    reset_frame_ = true;
    finished_frame_id_ = -1;
    next_valid_m_id_ = 0;
    batched_lidar_packets_ = 0;
    cache_.clear();
    settle_user_uploads();
    stg_->n_slots = 0;
    stg_->runs.clear();
    std::fill(stg_->col_src.begin(), stg_->col_src.end(), -1);
}

void FrameBatcher::set_pipeline_depth(size_t n) {
    if (n == 0) n = 1;
    wait_all();
    if (batched_lidar_packets_ != 0 && n != stg_->depth)
        throw std::logic_error("set_pipeline_depth: a frame is being batched; call between frames or after reset()");
    stg_->depth = n;
    if (stg_->jobs.size() < n) stg_->jobs.resize(n);
    if (stg_->cur >= n) stg_->cur = 0;
}
size_t FrameBatcher::pipeline_depth() const { return stg_->depth; }

void FrameBatcher::wait(const LidarFrame& f) {
    ScopedNs t(stats_.ns_wait);
    for (Staging::Job& j : stg_->jobs)
        if (j.owner == &f && j.job) {
            j.owner = nullptr;
            b200::check(ob_decode_job_wait(j.job));
            j.keep.clear();
        }
}

void FrameBatcher::wait_all() {
    ScopedNs t(stats_.ns_wait);
    for (Staging::Job& j : stg_->jobs)
        if (j.job) {
            j.owner = nullptr;
            b200::check(ob_decode_job_wait(j.job));
            j.keep.clear();
        }
}

// uploads that read caller memory must have left it before control returns to the caller
void FrameBatcher::settle_user_uploads() {
    ScopedNs t(stats_.ns_upload_wait);
    for (Staging::Job& j : stg_->jobs)
        if (j.user_uploads && j.job) {
            j.user_uploads = false;
            b200::check(ob_decode_job_uploads_done(j.job));
        }
}

void FrameBatcher::cache_packet(const uint8_t* buf, size_t size, uint64_t host_ts) {
    CachedPacket c;
    c.buf.assign(buf, buf + size);
    c.buf.resize(size + 8, 0);  // header getters read 8 bytes
    c.host_timestamp = host_ts;
    c.seq = cache_seq_++;
    cache_.push_back(std::move(c));
}

// the packet a frame-id ordered priority queue would pop: the oldest frame id, FIFO among equals
size_t FrameBatcher::cache_top() const {
    size_t best = 0;
    for (size_t i = 1; i < cache_.size(); ++i) {
        const int d = pf.frame_id_difference(pf.frame_id(cache_[best].buf.data()),
                                             pf.frame_id(cache_[i].buf.data()));
        if (d < 0 || (d == 0 && cache_[i].seq < cache_[best].seq)) best = i;
    }
    return best;
}

void FrameBatcher::start_frame(int64_t f_id, const uint8_t* packet_buf, LidarFrame& f) {
    finished_frame_id_ = -1;
    next_valid_m_id_ = 0;
    batched_lidar_packets_ = 0;
    f.frame_id = f_id;
    f.timestamp().setZero();
    f.measurement_id().setZero();
    f.status().setZero();
    f.packet_timestamp().setZero();
    const uint8_t thermal = static_cast<uint8_t>(pf.thermal_shutdown(packet_buf));
    const uint8_t shot = static_cast<uint8_t>(pf.shot_limiting(packet_buf));
    f.frame_status = static_cast<uint64_t>(thermal & 0x0f) | (static_cast<uint64_t>(shot & 0x0f) << 4);
    f.shutdown_countdown = static_cast<uint8_t>(pf.countdown_thermal_shutdown(packet_buf));
    f.shot_limiting_countdown = static_cast<uint8_t>(pf.countdown_shot_limiting(packet_buf));
    f.sensor_info = sensor_info_;
    Staging& s = *stg_;
    // a frame object that is still the target of an earlier submission must land first
    wait(f);
    // next job of the ring; its previous frame (depth frames ago) must have completed
    if (s.depth > 1) s.cur = (s.cur + 1) % s.depth;
    if (s.job().job) {
        ScopedNs t(stats_.ns_wait);
        s.job().owner = nullptr;
        b200::check(ob_decode_job_wait(s.job().job));
        s.job().keep.clear();
    }
    s.n_slots = 0;
    s.runs.clear();
    std::fill(s.col_src.begin(), s.col_src.end(), -1);
}

void FrameBatcher::batch_lidar_packet(const uint8_t* packet_buf, uint64_t host_ts, LidarFrame& f) {
    const int cpp = pf.columns_per_packet;
    // column header getters, inlined: the decode infos are copied once per call so that the 8-byte
    // load / mask / shift of FieldDecodeInfo::get sits in this loop instead of behind a call
    const FieldDecodeInfo i_ts = pf.col_timestamp_info(), i_mid = pf.col_measurement_id_info(),
                          i_st = pf.col_status_info();
    const size_t col0 = pf.packet_header_size, col_size = pf.col_size;
    auto nth_col = [&](int i) { return packet_buf + col0 + static_cast<size_t>(i) * col_size; };
    auto col_mid = [&](const uint8_t* c) { return i_mid.get<uint16_t>(c); };
    auto col_st = [&](const uint8_t* c) { return i_st.get<uint32_t>(c); };
    auto col_ts = [&](const uint8_t* c) { return i_ts.get<uint64_t>(c); };
    const uint16_t first_m_id = col_mid(nth_col(0));
    const uint16_t packet_id = static_cast<uint16_t>(first_m_id / cpp);
    if (packet_id < f.packet_timestamp().rows()) {
        f.packet_timestamp()[packet_id] = host_ts;
        f.alert_flags()[packet_id] = pf.alert_flags(packet_buf);
    }

    // stage the wire bytes: packets of a DMA-able burst are uploaded from where they are,
    // anything else goes through the page-locked bounce buffer of the job
    Staging& s = *stg_;
    const size_t slot = s.n_slots++;
    if (!headers_only_) {
        if (packet_buf >= s.burst_begin && packet_buf < s.burst_end) {
            s.add_run(packet_buf, s.burst_stride, slot, true);
        } else {
            s.reserve_bounce(slot + 1);
            uint8_t* dst = s.job().bounce + slot * s.stride;
            std::memcpy(dst, packet_buf, pf.lidar_packet_size);
            s.add_run(dst, s.stride, slot, false);
        }
    }
    const int32_t src0 = static_cast<int32_t>(slot) * cpp;

    // block path preconditions (lidar_frame.cpp:1542-1567)
    int block = pf.block_parsable();
    for (int icol = 0; icol < cpp && block; icol++) {
        const uint8_t* col = nth_col(icol);
        if (!(col_st(col) & 0x01) || col_mid(col) >= f.w) block = 0;
    }
    for (int icol = 0; icol < cpp && block; icol += block) {
        if (static_cast<size_t>(col_mid(nth_col(icol))) + block > f.w)
            block = 0;
    }

    auto timestamp = f.timestamp();
    auto measurement_id = f.measurement_id();
    auto status = f.status();
    auto zero_gap = [&](size_t from, size_t to) {  // zero_fields + zero_header_cols
        for (size_t j = from; j < to; ++j) {
            s.col_src[j] = -1;
            timestamp[j] = 0;
            measurement_id[j] = 0;
            status[j] = 0;
        }
    };

    if (block != 0) {  // parse_by_block, lidar_frame.cpp:1492-1528
        if (first_m_id >= next_valid_m_id_) {
            zero_gap(next_valid_m_id_, first_m_id);
            next_valid_m_id_ = static_cast<uint16_t>(first_m_id + cpp);
        }
        for (int icol = 0; icol < cpp; icol++) {
            const uint8_t* col = nth_col(icol);
            const uint16_t m_id = col_mid(col);
            measurement_id[m_id] = m_id;
            timestamp[m_id] = col_ts(col);
            status[m_id] = col_st(col);
        }
        // block_field places a group at the measurement id of its first column (parsing.cpp:647-653)
        for (int icol = 0; icol < cpp; icol += block) {
            const uint16_t m0 = col_mid(nth_col(icol));
            for (int x = 0; x < block; ++x) s.col_src[m0 + x] = src0 + icol + x;
        }
    } else {  // parse_by_col, lidar_frame.cpp:1422-1466
        for (int icol = 0; icol < cpp; icol++) {
            const uint8_t* col = nth_col(icol);
            const uint16_t m_id = col_mid(col);
            const uint32_t st = col_st(col);
            if (m_id >= f.w) continue;
            if (!(st & 0x01)) continue;
            if (m_id >= next_valid_m_id_) {
                zero_gap(next_valid_m_id_, m_id);
                next_valid_m_id_ = static_cast<uint16_t>(m_id + 1);
            }
            timestamp[m_id] = col_ts(col);
            measurement_id[m_id] = m_id;
            status[m_id] = st;
            s.col_src[m_id] = src0 + icol;
        }
    }
    batched_lidar_packets_++;
}

bool FrameBatcher::check_frame_complete(const LidarFrame& f) const {
    return pf.udp_profile_lidar == UDPProfileLidar::OFF ||
           (batched_lidar_packets_ >= expected_lidar_packets_ &&
            f.packet_timestamp().count() == expected_lidar_packets_);
}

void FrameBatcher::finalize_frame(LidarFrame& f) {
    // tail zero fill (lidar_frame.cpp:1906-1908), then the GPU pass materialises the pixel fields
    for (size_t j = next_valid_m_id_; j < f.w; ++j) stg_->col_src[j] = -1;
    decode_staged(f);

    if (f.sensor_info && f.sensor_info->init_id == last_init_id_ && f.frame_id <= last_frame_id_ &&
        pf.header_type == HeaderType::FUSA)
        throw std::runtime_error("32-bit frame id did not increase since the last frame");

    finished_frame_id_ = f.frame_id;
    last_frame_id_ = f.frame_id;
    batched_lidar_packets_ = 0;
}

void FrameBatcher::flush(LidarFrame& f) {
    decode_staged(f);
    wait(f);
}

void FrameBatcher::set_headers_only(bool on) { headers_only_ = on; }

void FrameBatcher::set_device_outputs(const DeviceOutputs* outputs) {
    wait_all();  // nothing in flight may still target the previous buffers
    dev_out_on_ = outputs != nullptr;
    dev_out_ = outputs ? *outputs : DeviceOutputs{};
}

// (Re)build the device decode table for the fields this frame shares with the profile
// (foreach_channel_field: profile order, only fields the frame has; impl/lidar_frame_impl.h:367-375)
// and make sure the current job exists.  Returns the names of the decoded fields, in table order.
std::vector<std::string> FrameBatcher::ensure_decoder(LidarFrame& f) {
    Staging& s = *stg_;
    std::string sig;
    std::vector<std::string> names;
    std::vector<ob_field_desc> descs;
    for (auto it = pf.begin(); it != pf.end(); ++it) {
        const std::string& name = it->first;
        if (!f.has_field(name) || name == ChanField::RAW_HEADERS) continue;
        const Field& fld = f.field(name);
        const FieldDecodeInfo& info = pf.field_decode_info(name);
        size_t elem = fld.element_size();
        if (fld.shape().size() > 2)
            for (size_t d = 2; d < fld.shape().size(); ++d) elem *= fld.shape()[d];
        if (elem < field_type_size(info.ty_tag) * static_cast<size_t>(info.num_elements))
            throw std::invalid_argument("Dest type too small for specified field");
        if (fld.shape().size() < 2 || fld.shape()[0] != f.h || fld.shape()[1] != f.w) continue;
        ob_field_desc d{};
        d.offset = static_cast<uint32_t>(info.offset);
        d.elem_size = static_cast<uint32_t>(elem);
        d.mask = info.mask;
        d.shift = info.shift;
        d.range_return = name == ChanField::RANGE ? 0 : (name == ChanField::RANGE2 ? 1 : -1);
        if (d.range_return >= 0 && elem != 4) d.range_return = -1;
        d.zero_pattern = fld.tag() == ChanFieldType::FLOAT16 ? 0x7e00u : 0u;
        descs.push_back(d);
        names.push_back(name);
        sig += name + ":" + std::to_string(elem) + ":" + std::to_string(info.offset) + ":" +
               std::to_string(info.mask) + ":" + std::to_string(info.shift) + ":" +
               std::to_string(d.zero_pattern) + ":" + std::to_string(d.range_return) + ";";
    }
    if (descs.size() > OB_MAX_FIELDS) throw std::invalid_argument("too many fields to decode");
    n_returns_ = 0;
    for (const auto& d : descs) n_returns_ = std::max(n_returns_, d.range_return + 1);
    if (!s.dec || sig != s.signature) {
        wait_all();
        for (Staging::Job& j : s.jobs) {  // jobs are bound to the decoder
            if (j.job) ob_decode_job_destroy(j.job);
            j.job = nullptr;
        }
        if (s.dec) ob_decoder_destroy(s.dec);
        s.dec = nullptr;
        ob_packet_layout L{};
        L.packet_header_size = static_cast<uint32_t>(pf.packet_header_size);
        L.col_header_size = static_cast<uint32_t>(pf.col_header_size);
        L.channel_data_size = static_cast<uint32_t>(pf.channel_data_size);
        L.col_size = static_cast<uint32_t>(pf.col_size);
        L.packet_size = static_cast<uint32_t>(pf.lidar_packet_size);
        L.columns_per_packet = static_cast<uint32_t>(pf.columns_per_packet);
        L.pixels_per_column = static_cast<uint32_t>(pf.pixels_per_column);
        L.columns_per_frame = static_cast<uint32_t>(f.w);
        // column headers are written on the host; the device copies are not requested here
        auto conv = [](const FieldDecodeInfo& fi) {
            ob_field_desc d{};
            d.offset = static_cast<uint32_t>(fi.offset);
            d.elem_size = 8;
            d.mask = fi.mask;
            d.shift = fi.shift;
            d.range_return = -1;
            return d;
        };
        L.col_timestamp = conv(pf.col_timestamp_info());
        L.col_measurement_id = conv(pf.col_measurement_id_info());
        L.col_status = conv(pf.col_status_info());
        b200::check(ob_decoder_create(&L, descs.data(), descs.size(), s.device, &s.dec));
        s.signature = sig;
    }
    Staging::Job& j = s.job();
    if (!j.stream) b200::check(ob_stream_create(s.device, &j.stream));
    if (!j.job) b200::check(ob_decode_job_create(s.dec, expected_lidar_packets_, j.stream, &j.job));
    return names;
}

// enqueue the H2D of every packet accepted since the last call
void FrameBatcher::upload_runs(LidarFrame& f) {
    Staging& s = *stg_;
    if (s.runs.empty() || headers_only_) {
        s.runs.clear();
        return;
    }
    if (!s.job().job) ensure_decoder(f);
    Staging::Job& j = s.job();
    for (const Staging::Run& r : s.runs) {
        b200::check(ob_decode_job_upload(j.job, r.src, r.src_stride, r.first, r.count));
        if (r.user) j.user_uploads = true;
    }
    s.runs.clear();
}

void FrameBatcher::decode_staged(LidarFrame& f) {
    Staging& s = *stg_;
    if (headers_only_) return;
    ScopedNs t_submit(stats_.ns_submit);
    const std::vector<std::string> names = ensure_decoder(f);
    upload_runs(f);
    Staging::Job& j = s.job();

    ob_decode_io io{};
    io.n_slots = s.n_slots;
    bool identity = s.n_slots * static_cast<size_t>(pf.columns_per_packet) >= f.w;
    for (size_t c = 0; c < f.w && identity; ++c) identity = s.col_src[c] == static_cast<int32_t>(c);
    io.col_src = identity ? nullptr : s.col_src.data();
    j.keep.clear();
    auto hold = [&j](std::shared_ptr<void> k) {
        if (k && (j.keep.empty() || j.keep.back() != k)) j.keep.push_back(std::move(k));
    };
    for (size_t i = 0; i < names.size(); ++i) {
        io.fields[i] = f.field(names[i]).get();
        hold(f.field(names[i]).keepalive());
        if (dev_out_on_)
            for (const auto& df : dev_out_.fields)
                if (df.first == names[i] && df.second) io.fields[i] = df.second;
    }

    const ob_lut* lut = nullptr;
    const int32_t* shifts = nullptr;
    size_t n_shifts = 0;
    if (fused_ && fused_->lut) {
        lut = fused_->lut.get();
        fused_->reserve(f.h, f.w, n_returns_);
        for (int r = 0; r < n_returns_; ++r) {
            io.xyz[r] = fused_->xyz[r].data();
            hold(fused_->xyz[r].keepalive());
            if (dev_out_on_ && dev_out_.xyz[r]) io.xyz[r] = dev_out_.xyz[r];
            if (!fused_->pixel_shift_by_row.empty()) {
                io.range_destaggered[r] = reinterpret_cast<uint32_t*>(fused_->range_destaggered[r].data());
                hold(fused_->range_destaggered[r].keepalive());
                if (dev_out_on_ && dev_out_.range_destaggered[r]) io.range_destaggered[r] = dev_out_.range_destaggered[r];
            }
        }
        if (!fused_->pixel_shift_by_row.empty()) {
            shifts = fused_->pixel_shift_by_row.data();
            n_shifts = fused_->pixel_shift_by_row.size();
        }
    }
    if (names.empty() && !lut) return;
    b200::check(ob_decode_job_submit(j.job, &io, lut, shifts, n_shifts));
    j.owner = &f;
    launches_++;
    stats_.frames++;
    if (s.depth <= 1) {
        ScopedNs t(stats_.ns_wait);  // synchronous: the frame is materialised when batch() returns true
        j.owner = nullptr;
        j.user_uploads = false;
        b200::check(ob_decode_job_wait(j.job));
        j.keep.clear();
    }
}

size_t FrameBatcher::batch_burst(const uint8_t* packets, size_t n, size_t stride, size_t size,
                                 const uint64_t* host_timestamps, LidarFrame& f, bool& complete) {
    complete = false;
    if (n == 0) return 0;
    ScopedNs t_total(stats_.ns_burst);
    if (!packets || !host_timestamps) throw std::invalid_argument("null pointer");
    if (n > 1 && stride < size) throw std::invalid_argument("packet stride smaller than the packet size");
    Staging& s = *stg_;
    // The host state machine reads the packet headers with the CPU, so the burst must be host-readable.
    // Page-locked host memory (and managed memory, which reports as device-kind) can additionally be read
    // by the copy engine where it lies; plain device memory is rejected instead of segfaulting below.
    const int kind0 = ob_pointer_kind(packets), kind1 = ob_pointer_kind(packets + (n - 1) * stride + size - 1);
    if ((kind0 == 2 || kind1 == 2) && !ob_pointer_host_readable(packets))
        throw std::invalid_argument("batch_burst needs host-readable packets (pageable, page-locked or managed memory)");
    const bool dma = !headers_only_ && kind0 != 0 && kind1 != 0;
    if (dma) {
        s.burst_begin = packets;
        s.burst_end = packets + (n - 1) * stride + size;
        s.burst_stride = stride;
    }
    struct Window {  // the zero-copy window never outlives this call
        Staging& s;
        FrameBatcher& b;
        LidarFrame& f;
        bool ok{false};
        ~Window() {
            s.burst_begin = s.burst_end = nullptr;
            if (!ok) {  // unwinding: drop pending reads of caller memory
                s.runs.erase(std::remove_if(s.runs.begin(), s.runs.end(),
                                            [](const Staging::Run& r) { return r.user; }),
                             s.runs.end());
                try {
                    b.settle_user_uploads();
                } catch (...) {
                }
                // the slots of the dropped runs were never uploaded: forget every column that points into
                // the current frame's staging so that a later finalize cannot decode stale device bytes
                try {
                    b.reset();
                } catch (...) {
                }
            }
        }
    } window{s, *this, f};
    size_t i = 0;
    for (; i < n && !complete; ++i) {
        complete = batch_impl(packets + i * stride, size, host_timestamps[i], f);
        // start the DMA of what has been accepted so far: it overlaps the parsing of the rest
        if (dma && !complete && s.runs.size() == 1 && s.runs[0].user && s.runs[0].count >= kUploadChunk)
            upload_runs(f);
    }
    if (dma) {
        // packets of a frame that is still open: upload now, the caller may reuse its memory
        bool user_pending = false;
        for (const Staging::Run& r : s.runs) user_pending |= r.user;
        if (user_pending) upload_runs(f);
        settle_user_uploads();
    }
    window.ok = true;
    return i;
}

bool FrameBatcher::batch_with_caching(const uint8_t* buf, size_t size, uint64_t host_ts, LidarFrame& f) {
    cache_packet(buf, size, host_ts);
    while (!cache_.empty()) {
        const size_t top = cache_top();
        const uint8_t* pbuf = cache_[top].buf.data();
        const int64_t f_id = pf.frame_id(pbuf);

        if (finished_frame_id_ >= 0 &&
            pf.frame_id_difference(static_cast<uint32_t>(finished_frame_id_), static_cast<uint32_t>(f_id)) <= 0) {
            dropped_packets_++;  // belongs to a frame that was already released
            cache_.erase(cache_.begin() + static_cast<std::ptrdiff_t>(top));
            continue;
        }
        if (f.frame_id == -1 || finished_frame_id_ >= 0) start_frame(f_id, pbuf, f);

        const int diff = pf.frame_id_difference(static_cast<uint32_t>(f.frame_id), static_cast<uint32_t>(f_id));
        if (diff < 0) {
            dropped_packets_++;
            cache_.erase(cache_.begin() + static_cast<std::ptrdiff_t>(top));
        } else if (diff > 0) {
            if (cache_.size() >= max_cache_size_) {  // give up waiting for the current frame
                finalize_frame(f);
                return true;
            }
            return false;
        } else {
            batch_lidar_packet(pbuf, cache_[top].host_timestamp, f);
            cache_.erase(cache_.begin() + static_cast<std::ptrdiff_t>(top));
            if (check_frame_complete(f)) {
                finalize_frame(f);
                return true;
            }
        }
    }
    return false;
}

bool FrameBatcher::handle_init_id_change(const uint8_t* buf, size_t size, uint64_t host_ts, LidarFrame& f) {
    last_init_id_ = pf.init_id(buf);
    if (f.frame_id == -1 || finished_frame_id_ >= 0) {  // no frame in flight: just start over
        reset();
        reset_frame_ = false;
        start_frame(pf.frame_id(buf), buf, f);
        batch_lidar_packet(buf, host_ts, f);
        if (check_frame_complete(f)) {
            finalize_frame(f);
            return true;
        }
        return false;
    }
    finalize_frame(f);  // release what we have, keep this packet for the next frame
    reset();
    cache_packet(buf, size, host_ts);
    return true;
}

bool FrameBatcher::batch_impl(const uint8_t* buf_in, size_t size, uint64_t host_ts, LidarFrame& f) {
    if (reset_frame_) {
        f.frame_id = -1;
        reset_frame_ = false;
    }
    if (f.w != sensor_info_->format.columns_per_frame || f.h != sensor_info_->format.pixels_per_column)
        throw std::invalid_argument("unexpected frame dimensions");
    if (f.packet_timestamp().rows() != f.w / static_cast<size_t>(pf.columns_per_packet))
        throw std::invalid_argument("unexpected frame columns_per_packet: " +
                                    std::to_string(pf.columns_per_packet));
    if (size < pf.lidar_packet_size) throw std::invalid_argument("lidar packet buffer too small");

    // every header getter reads 8 bytes that lie inside the packet (packet header >= 32 bytes or,
    // for LEGACY, column header 16 bytes; the LEGACY status window ends with its column)
    const uint8_t* buf = buf_in;

    if (pf.udp_profile_lidar != UDPProfileLidar::LEGACY && pf.init_id(buf) != last_init_id_)
        return handle_init_id_change(buf, pf.lidar_packet_size, host_ts, f);

    const int64_t f_id = pf.frame_id(buf);
    if (cache_.empty()) {
        if (finished_frame_id_ >= 0 &&
            pf.frame_id_difference(static_cast<uint32_t>(finished_frame_id_), static_cast<uint32_t>(f_id)) <= 0) {
            dropped_packets_++;
            return false;
        }
        if (f.frame_id == -1 || finished_frame_id_ >= 0) {
            start_frame(f_id, buf, f);
            batch_lidar_packet(buf, host_ts, f);
            if (check_frame_complete(f)) {
                finalize_frame(f);
                return true;
            }
            return false;
        }
    }
    if (f.frame_id == f_id && finished_frame_id_ < 0) {
        batch_lidar_packet(buf, host_ts, f);
        if (check_frame_complete(f)) {
            finalize_frame(f);
            return true;
        }
        return false;
    }
    return batch_with_caching(buf, pf.lidar_packet_size, host_ts, f);
}

bool FrameBatcher::batch(const uint8_t* buf, size_t size, uint64_t host_timestamp, LidarFrame& f) {
    return batch_impl(buf, size, host_timestamp, f);
}

bool FrameBatcher::batch(const Packet& packet, LidarFrame& f) {
    // lidar_frame.cpp batch_packet: only Lidar packets are batched (a LEGACY stream has no packet type
    // word, so its untyped buffers are taken as lidar); IMU / zone-monitor packets are outside the
    // accelerated path
    if (packet.type() != PacketType::Lidar &&
        !(packet.type() == PacketType::Unknown && pf.udp_profile_lidar == UDPProfileLidar::LEGACY))
        return false;
    return batch_impl(packet.buf.data(), packet.buf.size(), packet.host_timestamp, f);
}

bool FrameBatcher::operator()(const Packet& packet, LidarFrame& f) { return batch(packet, f); }

}  // namespace core
}  // namespace sdk
}  // namespace ouster
