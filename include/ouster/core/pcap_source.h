// pcap_source.h -- lidar-packet ingest for the scan->pointcloud path ("host ingest ->
// pinned staging"): reads lidar UDP payloads out of a capture file into a page-locked ring, in bursts
// that FrameBatcher::batch_burst / FramePipeline::push_burst consume in place (the copy engine reads the
// packets where the reader put them: no staging memcpy between the file and the GPU).
//
// Replaces, for this path only, the read loop of ouster_pcap/src/pcap_packet_source.cpp and the
// libtins-based reader under it (ouster_pcap/src/os_pcap.cpp): classic pcap files (micro- or nanosecond
// timestamps, either byte order), Ethernet II (one optional 802.1Q tag) -> IPv4 -> UDP, unfragmented
// datagrams (every fixture of the reference's test-suite is of this kind).  Fragmented
// datagrams, IPv6 and pcapng are counted in `skipped()` and ignored; indexing, seeking and the IMU /
// zone streams of the reference's source are outside SURVEY 8's scope.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "ouster/core/lidar_frame.h"
#include "ouster/core/visibility.h"

namespace ouster {
namespace sdk {
namespace core {

class OUSTER_API_CLASS PcapLidarSource {
   public:
    /// lidar_packet_size: UDP payloads of exactly this size are lidar packets (PacketFormat::lidar_packet_size);
    /// dst_port: 0 = any, else only datagrams to this UDP port (7502 by default on a sensor);
    /// ring_packets: capacity of the page-locked ring = the largest burst.
    /// Throws std::runtime_error("Failed to open pcap file") / ("Unsupported pcap format").
    OUSTER_API_FUNCTION PcapLidarSource(const std::string& path, size_t lidar_packet_size, uint16_t dst_port = 0,
                                        size_t ring_packets = 256);
    OUSTER_API_FUNCTION ~PcapLidarSource();
    PcapLidarSource(const PcapLidarSource&) = delete;
    PcapLidarSource& operator=(const PcapLidarSource&) = delete;

    /// Read up to max_packets lidar packets; they lie `stride()` bytes apart starting at `*packets`
    /// (page-locked when a GPU is present), capture timestamps in nanoseconds at `*timestamps_ns`.
    /// The burst stays valid until the next call.  Returns the number of packets (0 at end of file).
    OUSTER_API_FUNCTION size_t next_burst(size_t max_packets, const uint8_t** packets, const uint64_t** timestamps_ns);
    size_t stride() const { return stride_; }
    size_t packet_size() const { return packet_size_; }
    size_t packets_read() const { return packets_read_; }
    size_t skipped() const { return skipped_; }  ///< records that were not a whole lidar datagram
    bool eof() const { return eof_; }

   private:
    bool read_record(uint8_t* dst, uint64_t* ts_ns);
    std::FILE* f_{nullptr};
    bool swap_{false}, nanos_{false}, eof_{false};
    uint32_t linktype_{1};
    size_t packet_size_, stride_, ring_packets_;
    uint16_t dst_port_;
    size_t packets_read_{0}, skipped_{0};
    HostBuffer ring_;
    std::vector<uint64_t> ts_;
    std::vector<uint8_t> rec_;
};

}  // namespace core
}  // namespace sdk
}  // namespace ouster
