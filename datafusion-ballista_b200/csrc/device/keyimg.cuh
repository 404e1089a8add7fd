// Sign/zero-extended 64-bit image of an integer-like key column value (shared by the join and partition kernels).
#pragma once
#include <stdint.h>

#include "kernels.h"

namespace b200 {

__device__ __forceinline__ bool jkey_valid(const KeyCol& c, int64_t i) { return !c.valid || c.valid[i]; }
__device__ __forceinline__ uint64_t jkey_image(const KeyCol& c, int64_t i) {
  switch (c.width) {
    case 8: return ((const uint64_t*)c.data)[i];
    case 4: return (c.phys == PH_U32) ? (uint64_t)((const uint32_t*)c.data)[i] : (uint64_t)(int64_t)((const int32_t*)c.data)[i];
    case 2: return (c.phys == PH_U16) ? (uint64_t)((const uint16_t*)c.data)[i] : (uint64_t)(int64_t)((const int16_t*)c.data)[i];
    default: return (c.phys == PH_U8 || c.phys == PH_BOOL8) ? (uint64_t)((const uint8_t*)c.data)[i] : (uint64_t)(int64_t)((const int8_t*)c.data)[i];
  }
}

}  // namespace b200
