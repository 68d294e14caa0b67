// host_only_example.cpp -- the parts of the replacement headers that need no GPU: SensorInfo,
// PacketFormat, LidarFrame (fields, headers, column poses), frame_to_packets with CRC, the
// FrameBatcher state machine in header-only mode, column_timestamp_at_destaggered_pixel.
// Built with plain g++ and run on the CPU by tests/test_host_layer.py.  Prints "HOST OK".
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "ouster/core/frame_pipeline.h"
#include "ouster/core/lidar_scan.h"

using namespace ouster::sdk::core;

#define CHECK(cond)                                                                       \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #cond);  \
            std::exit(1);                                                                 \
        }                                                                                 \
    } while (0)

template <typename Ex, typename F>
static void expect_throw(F&& fn, const std::string& text) {
    try {
        fn();
    } catch (const Ex& e) {
        if (std::string(e.what()).find(text) == std::string::npos) {
            std::fprintf(stderr, "wrong message: '%s' (wanted '%s')\n", e.what(), text.c_str());
            std::exit(1);
        }
        return;
    }
    std::fprintf(stderr, "expected exception '%s'\n", text.c_str());
    std::exit(1);
}

int main() {
    auto info = SensorInfo::from_default(LidarMode{512, 10});
    info->format.udp_profile_lidar = UDPProfileLidar::RNG19_RFL8_SIG16_NIR16_DUAL;
    info->fw_rev = "v3.2.1";
    const size_t h = info->format.pixels_per_column, w = info->format.columns_per_frame;
    PacketFormat pf(*info);
    // geometry of the dual-return profile (parsing.cpp:459-469): 16-byte pixels, 12-byte column header
    CHECK(pf.channel_data_size == 16 && pf.col_header_size == 12 && pf.packet_header_size == 32);
    CHECK(pf.col_size == 12 + h * 16 && pf.lidar_packet_size == 32 + 16 * pf.col_size + 32);
    CHECK(pf.field_bitness(ChanField::RANGE) == 19 && pf.field_value_mask(ChanField::REFLECTIVITY2) == 0xff);
    CHECK(pf.block_parsable() == 16);

    LidarScan scan(info);
    CHECK(scan.has_field(ChanField::RANGE2) && scan.has_field(ChanField::WINDOW) && scan.w == w && scan.h == h);
    auto range = scan.field<uint32_t>(ChanField::RANGE);
    for (size_t i = 0; i < range.size(); ++i) range(i) = static_cast<uint32_t>((i * 2654435761u) & 0x7ffff);
    for (size_t c = 0; c < w; ++c) {
        scan.status()[c] = 1;
        scan.measurement_id()[c] = static_cast<uint16_t>(c);
        scan.timestamp()[c] = 5000 + 3 * c;
    }
    for (size_t p = 0; p < w / 16; ++p) scan.packet_timestamp()[p] = 10 + p;
    scan.frame_id = 41;
    expect_throw<std::invalid_argument>([&] { scan.field<uint16_t>(ChanField::RANGE); }, "Accessed field at wrong type");
    CHECK(scan.field_type(ChanField::RANGE) == FieldType(ChanField::RANGE, ChanFieldType::UINT32));
    CHECK(scan.field_types().size() == scan.fields().size());
    expect_throw<std::invalid_argument>([&] { scan.field_type("NOPE"); }, "Invalid field for LidarFrame");
    expect_throw<std::invalid_argument>([&] { scan.add_field(ChanField::RANGE, ChanFieldType::UINT32); }, "Duplicated field");

    // per-column poses: identity on construction, round trip, bounds (lidar_frame.cpp:350-358, 959-980)
    CHECK(scan.get_column_pose(7) == mat4d::Identity());
    mat4d m = mat4d::Identity();
    m(1, 3) = -2.5;
    scan.set_column_pose(7, m);
    CHECK(scan.get_column_pose(7) == m && scan.body_to_world().get<double>()[7 * 16 + 7] == -2.5);
    expect_throw<std::out_of_range>([&] { scan.get_column_pose(static_cast<int>(w)); }, "Column index out of range");
    CHECK(scan.get_first_valid_column() == 0 && scan.get_last_valid_column() == static_cast<int>(w) - 1);
    scan.status()[0] = 0;
    scan.status()[w - 1] = 2;
    CHECK(scan.get_first_valid_column() == 1 && scan.get_last_valid_column() == static_cast<int>(w) - 2);
    scan.status()[0] = 1;
    scan.status()[w - 1] = 1;
    // packet timestamps of the packets that carry valid columns (lidar_frame.cpp:643-790)
    CHECK(scan.packet_count() == w / 16 && scan.get_first_valid_packet_timestamp() == 10);
    CHECK(scan.get_last_valid_packet_timestamp() == 10 + w / 16 - 1 && scan.get_min_valid_packet_timestamp() == 10);
    for (size_t c = 0; c < 16; ++c) scan.status()[c] = 0;  // first packet: no valid column left
    scan.packet_timestamp()[5] = 3;                        // out-of-order host time
    CHECK(scan.get_first_valid_lidar_packet_timestamp() == 11 && scan.get_min_valid_packet_timestamp() == 3);
    CHECK(scan.get_max_valid_packet_timestamp() == 10 + w / 16 - 1);
    for (size_t c = 0; c < 16; ++c) scan.status()[c] = 1;
    scan.packet_timestamp()[5] = 15;
    {
        LidarScan empty(info);
        expect_throw<std::runtime_error>([&] { empty.get_first_valid_packet_timestamp(); }, "No valid packets in LidarFrame");
        CHECK(empty.get_last_valid_lidar_packet_timestamp() == 0);
        expect_throw<std::runtime_error>([&] { empty.get_first_valid_column(); }, "No valid columns in LidarFrame");
        LidarScan copy(scan);
        CHECK(copy == scan);
        copy.set_column_pose(3, m);
        CHECK(copy != scan);
    }

    // frame_to_packets: every packet carries a valid CRC64 and the right frame id / measurement ids
    std::vector<Packet> packets = impl::frame_to_packets(scan, pf, 9, 1234);
    CHECK(packets.size() == w / 16);
    for (size_t i = 0; i < packets.size(); ++i) {
        const uint8_t* b = packets[i].buf.data();
        CHECK(pf.frame_id(b) == 41 && pf.init_id(b) == 9 && pf.prod_sn(b) == 1234);
        CHECK(pf.col_measurement_id(pf.nth_col(0, b)) == 16 * i && (pf.col_status(pf.nth_col(15, b)) & 1));
        CHECK(pf.crc(b, packets[i].buf.size()).value() == pf.calculate_crc(b, packets[i].buf.size()));
        std::vector<uint32_t> px(h);
        pf.col_field<uint32_t>(pf.nth_col(2, b), ChanField::RANGE, px.data(), 1);  // column 16i+2, all rows
        CHECK(px[5] == range(5, 16 * i + 2) && px[h - 1] == range(h - 1, 16 * i + 2));
    }

    // FrameBatcher state machine without a GPU: header-only mode keeps ordering, completion and
    // the per-column / per-packet headers (lidar_frame.cpp:1698-1959)
    info->init_id = 9;
    ScanBatcher batcher(*info);
    batcher.set_headers_only(true);
    LidarScan out(info);
    size_t completed = 0;
    for (size_t i = 0; i < packets.size(); ++i) completed += batcher(packets[i], out) ? 1 : 0;
    CHECK(completed == 1 && out.frame_id == 41 && batcher.dropped_packets() == 0);
    for (size_t c = 0; c < w; ++c) CHECK(out.timestamp()[c] == 5000 + 3 * c && out.measurement_id()[c] == c);
    CHECK(out.packet_timestamp()[3] == 13 && out.complete());
    CHECK(batcher(packets[0], out) == false && batcher.dropped_packets() == 1);  // packet of a released frame

    // timestamp of the staggered column behind a destaggered pixel (lidar_frame.cpp:893-905)
    std::vector<int> shifts(h, 0);
    shifts[2] = 5;
    CHECK(column_timestamp_at_destaggered_pixel(2, 9, shifts, out.timestamp()) == out.timestamp()[4]);
    expect_throw<std::invalid_argument>(
        [&] { column_timestamp_at_destaggered_pixel(h, 0, shifts, out.timestamp()); }, "row or column is out of range");

    // FramePipeline ring logic without a GPU (header-only batcher): frames come out in order,
    // `depth` frames late, slots are recycled, drain() empties the ring
    {
        FramePipeline pipe(info, 2);
        pipe.batcher().set_headers_only(true);
        std::vector<int64_t> seen;
        for (int k = 0; k < 7; ++k) {
            scan.frame_id = 100 + k;
            for (const Packet& p : impl::frame_to_packets(scan, pf, 9, 1234))
                if (const FramePipeline::Slot* sl = pipe.push(p)) seen.push_back(sl->frame.frame_id);
            CHECK(pipe.in_flight() == static_cast<size_t>(k < 2 ? k + 1 : 2));
        }
        CHECK(seen.size() == 5);
        while (const FramePipeline::Slot* sl = pipe.drain()) seen.push_back(sl->frame.frame_id);
        CHECK(seen.size() == 7 && pipe.in_flight() == 0);
        for (int k = 0; k < 7; ++k) CHECK(seen[static_cast<size_t>(k)] == 100 + k);
    }

    std::printf("HOST OK\n");
    return 0;
}
