// packet.h -- packet buffers handed to FrameBatcher
// (mirrors ouster_core/include/ouster/core/packet.h:24-130 for lidar packets).
#pragma once
#include <cstdint>
#include <memory>
#include <vector>

#include "ouster/core/types.h"
#include "ouster/core/visibility.h"

namespace ouster {
namespace sdk {
namespace core {

enum class PacketType { Unknown, Lidar, Imu, Zone };

struct OUSTER_API_CLASS Packet {
   protected:
    PacketType type_{PacketType::Unknown};

   public:
    uint64_t host_timestamp{0};
    std::vector<uint8_t> buf;
    std::shared_ptr<PacketFormat> format;

    Packet() = default;
    explicit Packet(PacketType t, size_t size = 0) : type_(t), buf(size, 0) {}
    PacketType type() const { return type_; }
};

struct OUSTER_API_CLASS LidarPacket : public Packet {
    LidarPacket() : Packet(PacketType::Lidar) {}
    explicit LidarPacket(size_t size) : Packet(PacketType::Lidar, size) {}
    explicit LidarPacket(const std::shared_ptr<PacketFormat>& pf)
        : Packet(PacketType::Lidar, pf->lidar_packet_size) {
        format = pf;
    }
};

}  // namespace core
}  // namespace sdk
}  // namespace ouster
