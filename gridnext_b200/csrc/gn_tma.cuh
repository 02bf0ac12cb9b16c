// Host-side TMA descriptor (CUtensorMap) construction.  cuTensorMapEncodeTiled is resolved at run time
// through cudaGetDriverEntryPoint so the library links only the CUDA runtime.
#pragma once
#include "gn_common.cuh"
#include <cuda.h>

int gn_tmap_encode(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* gaddr, const uint64_t* dims,
                   const uint64_t* strides_bytes /* rank-1 entries, dims 1.. */, const uint32_t* box, CUtensorMapSwizzle swizzle);

// 2-D bf16 row-major matrix [rows, cols] with row pitch ld (elements); box = {box_cols, box_rows}
static inline int gn_tmap_bf16_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
                                  uint32_t box_rows, CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B) {
    uint64_t dims[2] = {cols, rows};
    uint64_t strides[1] = {ld * 2};
    uint32_t box[2] = {box_cols, box_rows};
    return gn_tmap_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, sw);
}
