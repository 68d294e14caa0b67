// b200_runtime.h -- glue between the C++ replacement headers and the C ABI (ouster_b200.h):
// device selection, the per-thread stream, and status -> exception translation.
#pragma once
#include <stdexcept>
#include <string>

#include "ouster/core/visibility.h"
#include "ouster_b200.h"

namespace ouster {
namespace sdk {
namespace core {
namespace b200 {

/// CUDA device used by the calling thread's implicit stream (default 0, or env OUSTER_B200_DEVICE).
OUSTER_API_FUNCTION void set_device(int device);
OUSTER_API_FUNCTION int device();
/// The calling thread's ob_stream on the selected device (created on first use).
OUSTER_API_FUNCTION ob_stream* thread_stream();
/// Block until everything queued on the calling thread's stream has finished.
OUSTER_API_FUNCTION void synchronize();

/// Rethrow a C-ABI failure as the exception type the reference throws, with the same text.
inline void check(ob_status st) {
    if (st == OB_OK) return;
    const std::string msg = ob_last_error();
    if (st == OB_INVALID_ARGUMENT) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
}

}  // namespace b200
}  // namespace core
}  // namespace sdk
}  // namespace ouster
