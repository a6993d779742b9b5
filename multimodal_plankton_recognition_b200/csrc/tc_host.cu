// Host side of the tensor-core path: TMA tensor-map encoding through the driver entry point
// (no link-time dependency on libcuda).
#include <mutex>
#include "tc_common.cuh"

namespace plk {
namespace tc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld,
                   int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  PLK_REQUIRE(fn != nullptr, PLK_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  // The driver entry point needs a context bound to the CALLING thread.  The runtime binds the
  // primary context lazily, and e.g. PyTorch's autograd worker threads may not have issued any
  // runtime call yet when they reach us: bind it once per thread.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    PLK_CUDA(cudaFree(nullptr));
    ctx_bound = true;
  }
  PLK_REQUIRE(((uintptr_t)base & 15) == 0, PLK_ERR_INVALID, "bf16 operand must be 16-byte aligned");
  PLK_REQUIRE((ld * 2) % 16 == 0, PLK_ERR_INVALID, "bf16 operand leading dimension must be a multiple of 8");
  PLK_REQUIRE(cols % kChunkK == 0, PLK_ERR_INVALID, "bf16 operand width must be padded to a multiple of 64");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)kChunkK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PLK_REQUIRE(r == CUDA_SUCCESS, PLK_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return PLK_OK;
}

// fp32 [slabs][rows][cols] output tensor, box = 128 rows x 32 columns of one slab (128-byte rows,
// SWIZZLE_128B): the store side of the backward's accumulator drain.
int make_tmap_f32_slabs(CUtensorMap* out, const void* base, int64_t slabs, int64_t rows, int64_t cols) {
  EncodeTiledFn fn = get_encode_fn();
  PLK_REQUIRE(fn != nullptr, PLK_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    PLK_CUDA(cudaFree(nullptr));
    ctx_bound = true;
  }
  PLK_REQUIRE(((uintptr_t)base & 15) == 0 && cols % 4 == 0, PLK_ERR_INVALID,
              "fp32 slab tensor must be 16-byte aligned with a row pitch that is a multiple of 16 bytes");
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)slabs};
  cuuint64_t gstr[2] = {(cuuint64_t)cols * 4, (cuuint64_t)rows * (cuuint64_t)cols * 4};
  cuuint32_t box[3] = {32, (cuuint32_t)kTileRows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PLK_REQUIRE(r == CUDA_SUCCESS, PLK_ERR_CUDA, "cuTensorMapEncodeTiled (fp32 slabs) failed with CUresult %d", (int)r);
  return PLK_OK;
}

}  // namespace tc
}  // namespace plk
