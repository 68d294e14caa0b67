// lidar_scan.h -- deprecated forwarding header, as in the reference
// (ouster_core/include/ouster/core/lidar_scan.h:6-10).
#pragma once
#include "ouster/core/lidar_frame.h"
