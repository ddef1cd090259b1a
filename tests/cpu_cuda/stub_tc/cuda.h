// cuda.h STAND-IN (driver API types for tensor maps) for the CPU emulation of the tcgen05 / TMA kernels.
// TEST INFRASTRUCTURE (tests/cpu_cuda).  The "tensor map" is a plain record of what cuTensorMapEncodeTiled was given.
#pragma once
#include <stdint.h>
typedef uint32_t cuuint32_t;
typedef uint64_t cuuint64_t;
typedef int CUresult;
#define CUDA_SUCCESS 0
typedef enum { CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 = 9 } CUtensorMapDataType;
typedef enum { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 } CUtensorMapInterleave;
typedef enum { CU_TENSOR_MAP_SWIZZLE_NONE = 0, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_SWIZZLE_128B } CUtensorMapSwizzle;
typedef enum { CU_TENSOR_MAP_L2_PROMOTION_NONE = 0, CU_TENSOR_MAP_L2_PROMOTION_L2_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B } CUtensorMapL2promotion;
typedef enum { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 } CUtensorMapFloatOOBfill;
struct CUtensorMap_st {
    uint64_t base;
    uint64_t strides[4];     // bytes, dims 1..4
    uint32_t dims[5];        // elements, innermost first
    uint32_t box[5];         // bounding box (elements)
    uint32_t estr[5];        // traversal strides
    uint32_t rank, swizzle;
};
typedef struct CUtensorMap_st CUtensorMap;
