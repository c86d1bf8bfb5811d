// tests/cusim -- driver-API subset (tensor maps).  TEST INFRASTRUCTURE, see cuda_runtime.h.
#pragma once
#include <cstdint>
typedef int CUresult;
enum : int { CUDA_SUCCESS = 0, CUDA_ERROR_INVALID_VALUE = 1 };
typedef uint64_t cuuint64_t;
typedef uint32_t cuuint32_t;
struct alignas(64) CUtensorMap { uint64_t opaque[16]; };
enum CUtensorMapDataType { CU_TENSOR_MAP_DATA_TYPE_UINT8 = 0, CU_TENSOR_MAP_DATA_TYPE_UINT16 = 1, CU_TENSOR_MAP_DATA_TYPE_UINT32 = 2 };
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 };
enum CUtensorMapSwizzle { CU_TENSOR_MAP_SWIZZLE_NONE = 0 };
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_NONE = 0, CU_TENSOR_MAP_L2_PROMOTION_L2_128B = 2 };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 };
extern "C" CUresult cuTensorMapEncodeTiled(CUtensorMap *, CUtensorMapDataType, cuuint32_t rank, void *global, const cuuint64_t *dims, const cuuint64_t *strides,
                                           const cuuint32_t *box, const cuuint32_t *elem_strides, CUtensorMapInterleave, CUtensorMapSwizzle,
                                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
