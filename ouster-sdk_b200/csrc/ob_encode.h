// ob_encode.h -- K4 (fields -> packets + CRC64) internal declarations, see ob_encode.cu.
#pragma once
#include "ob_internal.h"

namespace ob {

struct EncodeFrame {  // one per frame of a launch, lives in device memory; all pointers device memory
    const void* fields[OB_MAX_FIELDS];  // row-major h x w images in the decoder's field order (null: field left 0)
    const uint64_t* timestamp;          // w column timestamps (null: 0)
    const uint32_t* status;             // w column status words (null: every column invalid)
    const uint8_t* packet_headers;      // n_packets x header_bytes leading bytes of every packet (host-built)
    uint32_t header_bytes;
    uint32_t pad;
    uint8_t* packets;                   // out: n_packets x packet_stride
    unsigned long long packet_stride;
};

cudaError_t launch_encode(const DecodeLayout& L, const EncodeFrame* frames_dev, uint32_t n_frames, bool with_crc,
                          int device, cudaStream_t st);

}  // namespace ob
