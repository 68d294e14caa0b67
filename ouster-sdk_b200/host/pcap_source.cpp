// pcap_source.cpp -- PcapLidarSource: classic pcap -> page-locked ring of lidar packets (see the header).
#include "ouster/core/pcap_source.h"

#include <cstring>
#include <stdexcept>

namespace ouster {
namespace sdk {
namespace core {

namespace {
uint32_t rd32(const uint8_t* p, bool swap) {
    uint32_t v;
    std::memcpy(&v, p, 4);
    return swap ? __builtin_bswap32(v) : v;
}
uint16_t be16(const uint8_t* p) { return static_cast<uint16_t>((p[0] << 8) | p[1]); }
}  // namespace

PcapLidarSource::PcapLidarSource(const std::string& path, size_t lidar_packet_size, uint16_t dst_port,
                                 size_t ring_packets)
    : packet_size_(lidar_packet_size),
      stride_((lidar_packet_size + 15) & ~static_cast<size_t>(15)),  // 16-byte slots: TMA-copyable on the device side too
      ring_packets_(ring_packets ? ring_packets : 1),
      dst_port_(dst_port) {
    if (lidar_packet_size == 0) throw std::invalid_argument("lidar_packet_size must be positive");
    f_ = std::fopen(path.c_str(), "rb");
    if (!f_) throw std::runtime_error("Failed to open pcap file");
    uint8_t gh[24];
    if (std::fread(gh, 1, 24, f_) != 24) {
        std::fclose(f_);
        f_ = nullptr;
        throw std::runtime_error("Unsupported pcap format");
    }
    uint32_t magic;
    std::memcpy(&magic, gh, 4);
    if (magic == 0xa1b2c3d4u) { swap_ = false; nanos_ = false; }
    else if (magic == 0xd4c3b2a1u) { swap_ = true; nanos_ = false; }
    else if (magic == 0xa1b23c4du) { swap_ = false; nanos_ = true; }
    else if (magic == 0x4d3cb2a1u) { swap_ = true; nanos_ = true; }
    else {
        std::fclose(f_);
        f_ = nullptr;
        throw std::runtime_error("Unsupported pcap format");  // pcapng et al.
    }
    linktype_ = rd32(gh + 20, swap_);
    ring_.resize(stride_ * ring_packets_ + 16);
    ts_.resize(ring_packets_);
}

PcapLidarSource::~PcapLidarSource() {
    if (f_) std::fclose(f_);
}

// one capture record; true when it was a whole lidar datagram (copied to dst)
bool PcapLidarSource::read_record(uint8_t* dst, uint64_t* ts_ns) {
    uint8_t rh[16];
    if (std::fread(rh, 1, 16, f_) != 16) {
        eof_ = true;
        return false;
    }
    const uint32_t sec = rd32(rh, swap_), frac = rd32(rh + 4, swap_), incl = rd32(rh + 8, swap_);
    if (incl > (1u << 24)) {  // corrupt record header
        eof_ = true;
        return false;
    }
    rec_.resize(incl);
    if (incl && std::fread(rec_.data(), 1, incl, f_) != incl) {
        eof_ = true;
        return false;
    }
    const uint8_t* p = rec_.data();
    size_t n = incl;
    if (linktype_ == 1) {  // Ethernet II, optional 802.1Q tag
        if (n < 14) return false;
        uint16_t et = be16(p + 12);
        p += 14;
        n -= 14;
        if (et == 0x8100) {
            if (n < 4) return false;
            et = be16(p + 2);
            p += 4;
            n -= 4;
        }
        if (et != 0x0800) return false;
    } else if (linktype_ == 113) {  // Linux cooked capture
        if (n < 16 || be16(p + 14) != 0x0800) return false;
        p += 16;
        n -= 16;
    } else if (linktype_ != 101 && linktype_ != 228) {  // raw IP otherwise
        return false;
    }
    if (n < 20 || (p[0] >> 4) != 4) return false;
    const size_t ihl = static_cast<size_t>(p[0] & 0x0f) * 4;
    if (ihl < 20 || n < ihl + 8 || p[9] != 17) return false;       // not UDP
    const uint16_t frag = be16(p + 6);
    if ((frag & 0x2000) || (frag & 0x1fff)) return false;            // fragmented: not handled here
    const uint8_t* udp = p + ihl;
    const uint16_t dport = be16(udp + 2), ulen = be16(udp + 4);
    if (dst_port_ && dport != dst_port_) return false;
    if (ulen < 8 || static_cast<size_t>(ulen) - 8 != packet_size_ || n < ihl + ulen) return false;
    std::memcpy(dst, udp + 8, packet_size_);
    *ts_ns = static_cast<uint64_t>(sec) * 1000000000ull + static_cast<uint64_t>(frac) * (nanos_ ? 1ull : 1000ull);
    return true;
}

size_t PcapLidarSource::next_burst(size_t max_packets, const uint8_t** packets, const uint64_t** timestamps_ns) {
    if (max_packets > ring_packets_) max_packets = ring_packets_;
    size_t n = 0;
    while (n < max_packets && !eof_) {
        if (read_record(ring_.data() + n * stride_, &ts_[n])) {
            ++n;
            ++packets_read_;
        } else if (!eof_) {
            ++skipped_;
        }
    }
    if (packets) *packets = ring_.data();
    if (timestamps_ns) *timestamps_ns = ts_.data();
    return n;
}

}  // namespace core
}  // namespace sdk
}  // namespace ouster
