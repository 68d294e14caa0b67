// field_decode_info.h -- bit-field descriptor of one packet field
// (mirrors ouster_core/include/ouster/core/field_decode_info.h:24-79).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "ouster/core/chanfield.h"
#include "ouster/core/visibility.h"

namespace ouster {
namespace sdk {
namespace core {

struct OUSTER_API_CLASS FieldDecodeInfo {
    ChanFieldType ty_tag{ChanFieldType::VOID};
    size_t offset{0};
    uint64_t mask{0};
    int shift{0};
    int num_elements{1};

    /// NOTE: reads 8 bytes at buffer + offset (host-side use: headers and tests; the per-pixel
    /// decode of whole frames runs on the GPU through the same (offset, mask, shift) triple).
    template <typename T>
    T get(const uint8_t* buffer) const {
        uint64_t word;
        std::memcpy(&word, buffer + offset, sizeof(word));
        word &= mask;
        if (shift > 0) word >>= shift;
        else if (shift < 0) word <<= -shift;
        T out{};
        std::memcpy(&out, &word, sizeof(out));
        return out;
    }

    template <typename T>
    void set(uint8_t* buffer, T value) const {
        uint64_t word = 0;
        std::memcpy(&word, &value, sizeof(value));
        if (shift > 0) word <<= shift;
        if (shift < 0) word >>= -shift;
        word &= mask;
        uint64_t cur;
        std::memcpy(&cur, buffer + offset, sizeof(cur));
        cur = (cur & ~mask) | word;
        std::memcpy(buffer + offset, &cur, sizeof(cur));
    }
};

/// Factory (ouster_core/src/parsing.cpp:57-122): bit_start/bit_size in bits, optional up-shift,
/// optional buffer length limit. Throws std::invalid_argument on impossible requests.
OUSTER_API_FUNCTION FieldDecodeInfo field_info(size_t bit_start, size_t bit_size, size_t upshift = 0,
                                               size_t max_length = 0, size_t num_elements = 1);

namespace impl {
OUSTER_API_FUNCTION uint64_t get_value_mask(const FieldDecodeInfo& f);
OUSTER_API_FUNCTION int get_bitness(const FieldDecodeInfo& f);
}  // namespace impl

}  // namespace core
}  // namespace sdk
}  // namespace ouster
