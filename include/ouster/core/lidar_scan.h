// Old spelling of the frame header.  Code written before the LidarScan -> LidarFrame rename keeps
// compiling: the aliases LidarScan / ScanBatcher / LidarScanFieldTypes live in lidar_frame.h.
#ifndef OUSTER_B200_LIDAR_SCAN_FORWARD_H
#define OUSTER_B200_LIDAR_SCAN_FORWARD_H
#include "ouster/core/lidar_frame.h"
#endif
