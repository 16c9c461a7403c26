// Host side of tma.cuh: cuTensorMapEncodeTiled resolved through cudaGetDriverEntryPoint.
#include "tma.cuh"

namespace fdg {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    cudaGetLastError();
    return (EncodeTiledFn)p;
  }();
  return fn;
}

static bool make_tmap(CUtensorMap* map, const void* base, CUtensorMapDataType dt, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box);

bool make_tmap_f32(CUtensorMap* map, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
  return make_tmap(map, base, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, dims, strides_bytes, box);
}

bool make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
  return make_tmap(map, base, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, dims, strides_bytes, box);
}

static bool make_tmap(CUtensorMap* map, const void* base, CUtensorMapDataType dt, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box) {
  EncodeTiledFn fn = encode_fn();
  if (!fn || rank < 2 || rank > 4) return false;
  if (reinterpret_cast<uintptr_t>(base) & 15) return false;
  cuuint64_t gd[4], gs[3];
  cuuint32_t bx[4], es[4];
  for (int i = 0; i < rank; ++i) {
    if (dims[i] == 0 || box[i] == 0 || box[i] > 256) return false;
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) {
    if (strides_bytes[i] % 16 != 0 || strides_bytes[i] >= (1ull << 40)) return false;
    gs[i] = strides_bytes[i];
  }
  CUresult r = fn(map, dt, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace fdg
