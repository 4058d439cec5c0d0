// gps_config.h -- compile-time kernel configurations baked into libgps_b200.so.
// Chosen by measurement on B200 (tools/gps_tune sweeps the alternatives; the numbers
// behind each choice are in DESIGN.md and profiles/).
#pragma once

// CSDL register path: NT threads, V uint4 per thread -> tile = NT * V * 4 keys
#ifndef GPS_REG_NT
#define GPS_REG_NT 256
#endif
#ifndef GPS_REG_V
#define GPS_REG_V 4
#endif

// CSDL TMA path: NT threads x 16 keys per tile, STAGES shared-memory tile buffers
#ifndef GPS_TMA_NT
#define GPS_TMA_NT 512
#endif
#ifndef GPS_TMA_STAGES
#define GPS_TMA_STAGES 3
#endif
// 0 = as many CTAs per SM as fit
#ifndef GPS_TMA_MAX_CTAS_PER_SM
#define GPS_TMA_MAX_CTAS_PER_SM 0
#endif

// Reduce-then-Scan tiles
#ifndef GPS_RTS_NT
#define GPS_RTS_NT 256
#endif
#ifndef GPS_RTS_V
#define GPS_RTS_V 4
#endif

// GPS_ALGO_CSDL resolves to: 1 = register path, 2 = TMA path
#ifndef GPS_CSDL_DEFAULT_VARIANT
#define GPS_CSDL_DEFAULT_VARIANT 1
#endif

// spin budget of GPS_ALGO_CSDL_FALLBACK (no emulated stalls): polls of a not-ready
// window before the CTA recomputes the blocking tile itself
#ifndef GPS_FALLBACK_DEFAULT_SPIN
#define GPS_FALLBACK_DEFAULT_SPIN 64
#endif
