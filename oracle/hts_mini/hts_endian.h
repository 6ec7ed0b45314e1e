/* TEST INFRASTRUCTURE ONLY: the two little-endian readers of htslib's hts_endian.h that bam_handler.cpp:370,407 use. */
#ifndef PV_HTS_MINI_ENDIAN_H
#define PV_HTS_MINI_ENDIAN_H
#include <cstdint>
#include <cstring>
static inline uint32_t le_to_u32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline float le_to_float(const uint8_t* p) { float v; memcpy(&v, p, 4); return v; }
#endif
