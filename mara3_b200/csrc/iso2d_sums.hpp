/** Index layout of the 16 per-stage running sums shared by kernels.cu and scheme.cpp. */
#pragma once
namespace m3b { namespace sums {
    enum { ACC_MASS = 0, ACC_PX = 2, ACC_PY = 4, ACC_LZ = 6, GRV_FX = 8, GRV_FY = 10, GRV_TQ = 12, BUF_M = 14, BUF_L = 15, NUM_SUMS = 16 };
}}
