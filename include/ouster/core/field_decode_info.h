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

    /// Raw 64-bit little-endian window at `offset` <-> field value.  The window is always 8 bytes
    /// wide, whatever the field width (host-side use: headers and tests; the per-pixel decode of
    /// whole frames runs on the GPU from the same (offset, mask, shift) triple).
    uint64_t value_of(uint64_t window) const {
        const uint64_t bits = window & mask;
        return shift >= 0 ? bits >> shift : bits << -shift;
    }
    uint64_t window_of(uint64_t value) const {
        return (shift >= 0 ? value << shift : value >> -shift) & mask;
    }

    template <typename T>
    T get(const uint8_t* buffer) const {
        const uint64_t v = value_of(load64(buffer + offset));
        T out{};
        std::memcpy(&out, &v, sizeof(T));  // truncation to T, as the reference does
        return out;
    }

    template <typename T>
    void set(uint8_t* buffer, T value) const {
        uint64_t v = 0;
        std::memcpy(&v, &value, sizeof(T));
        uint8_t* at = buffer + offset;
        store64(at, (load64(at) & ~mask) | window_of(v));
    }

   private:
    static uint64_t load64(const uint8_t* p) {
        uint64_t w;
        std::memcpy(&w, p, sizeof(w));
        return w;
    }
    static void store64(uint8_t* p, uint64_t w) { std::memcpy(p, &w, sizeof(w)); }
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
