// visibility.h -- export macros of the replacement headers (mirrors
// ouster_core/include/ouster/core/visibility.h of the reference).
#pragma once
#if defined(_WIN32)
#define OUSTER_API_FUNCTION
#define OUSTER_API_CLASS
#else
#define OUSTER_API_FUNCTION __attribute__((visibility("default")))
#define OUSTER_API_CLASS __attribute__((visibility("default")))
#endif
