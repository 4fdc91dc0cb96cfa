/* Minimal DLPack (v0.8 ABI) tensor structs used at the b3d C-ABI boundary.
 * Layout-compatible with dmlc/dlpack `dlpack.h`; a `DLManagedTensor*` obtained from
 * `torch.utils.dlpack.to_dlpack` may be passed wherever a `const DLTensor*` is expected
 * (the DLTensor is the first member).  All tensors are BORROWED: the callee never calls
 * `deleter` and keeps no reference past return (+ stream order). */
#ifndef B3D_DLPACK_H_
#define B3D_DLPACK_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { kDLCPU = 1, kDLCUDA = 2, kDLCUDAHost = 3 } DLDeviceType;
typedef struct { int32_t device_type; int32_t device_id; } DLDevice;
typedef enum { kDLInt = 0, kDLUInt = 1, kDLFloat = 2, kDLBfloat = 4 } DLDataTypeCode;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } DLDataType;

typedef struct {
  void* data;
  DLDevice device;
  int32_t ndim;
  DLDataType dtype;
  int64_t* shape;
  int64_t* strides; /* in elements; NULL = compact row-major */
  uint64_t byte_offset;
} DLTensor;

typedef struct DLManagedTensor {
  DLTensor dl_tensor;
  void* manager_ctx;
  void (*deleter)(struct DLManagedTensor* self);
} DLManagedTensor;

#ifdef __cplusplus
}
#endif
#endif
