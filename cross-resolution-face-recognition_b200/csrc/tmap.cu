// Tensor-map (TMA descriptor) encoding shared by every TMA-fed kernel of the library: bf16 tiles, SWIZZLE_128B,
// zero fill outside the tensor.  cuTensorMapEncodeTiled is resolved through the runtime's driver entry point, so the
// library has no link-time dependency on libcuda.
#include <cudaTypedefs.h>

#include "common.cuh"
#include "internal.h"

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

}  // namespace

int crfr_tmap_encode_bf16(CUtensorMap* m, const void* ptr, int rank, const unsigned long long* dims,
                          const unsigned long long* strides_bytes, const unsigned int* box, const char* what) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    crfr_set_error("cuTensorMapEncodeTiled entry point not available");
    return CRFR_ECUDA;
  }
  cuuint64_t d[5], s[4];
  cuuint32_t b[5], es[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    if (i + 1 < rank) s[i] = strides_bytes[i];
  }
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), d, s, b, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    crfr_set_error("cuTensorMapEncodeTiled(%s: rank %d, dims %llu x %llu, box %u x %u) failed: %d", what, rank, dims[0],
                   rank > 1 ? dims[1] : 0ull, box[0], rank > 1 ? box[1] : 0u, (int)r);
    return CRFR_ECUDA;
  }
  return CRFR_OK;
}
