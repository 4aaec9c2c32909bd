// sort.cu — sort-by-cell (and sort-by-voxel) via CUB's LSD radix sort.
//
// This is the "hash map" of the reference's rasterize() (elevation_mapping.cpp:41-92)
// turned into a sort: (cell key, point index) pairs ordered by key.  LSD radix sort is
// STABLE, so inside one cell the points stay in input order — the two order-dependent
// tie-breaks of rasterize() (lowest index wins equal min_z, highest index wins colour)
// depend on that.  Only the key bits that can be set are sorted (end_bit).
// Kept in its own translation unit: CUB's templates dominate compile time.
#include <cub/device/device_radix_sort.cuh>

#include "device_types.h"

namespace fdem {

size_t sort_pairs_u32_temp_bytes(uint32_t n, int end_bit) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, static_cast<const uint32_t*>(nullptr),
                                  static_cast<uint32_t*>(nullptr),
                                  static_cast<const uint32_t*>(nullptr),
                                  static_cast<uint32_t*>(nullptr), static_cast<int>(n), 0, end_bit);
  return bytes;
}

cudaError_t sort_pairs_u32(void* temp, size_t temp_bytes, const uint32_t* kin, uint32_t* kout,
                           const uint32_t* vin, uint32_t* vout, uint32_t n, int end_bit,
                           cudaStream_t s, LaunchCounter& lc) {
  // onesweep: 1 histogram + 1 scan + ceil(end_bit/8) passes
  lc.library += 2 + (end_bit + 7) / 8;
  return cub::DeviceRadixSort::SortPairs(temp, temp_bytes, kin, kout, vin, vout,
                                         static_cast<int>(n), 0, end_bit, s);
}

size_t sort_pairs_u64_temp_bytes(uint32_t n, int end_bit) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, static_cast<const uint64_t*>(nullptr),
                                  static_cast<uint64_t*>(nullptr),
                                  static_cast<const uint32_t*>(nullptr),
                                  static_cast<uint32_t*>(nullptr), static_cast<int>(n), 0, end_bit);
  return bytes;
}

cudaError_t sort_pairs_u64(void* temp, size_t temp_bytes, const uint64_t* kin, uint64_t* kout,
                           const uint32_t* vin, uint32_t* vout, uint32_t n, int end_bit,
                           cudaStream_t s, LaunchCounter& lc) {
  lc.library += 2 + (end_bit + 7) / 8;
  return cub::DeviceRadixSort::SortPairs(temp, temp_bytes, kin, kout, vin, vout,
                                         static_cast<int>(n), 0, end_bit, s);
}

}  // namespace fdem
