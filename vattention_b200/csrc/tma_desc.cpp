#include "tma_desc.h"

#include <stdexcept>
#include <string>

#include "vmm_driver.h"

namespace vattn {

namespace {
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                              CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                              CUtensorMapFloatOOBfill);
EncodeFn encoder() {
  static EncodeFn fn = reinterpret_cast<EncodeFn>(cuda_driver_symbol("cuTensorMapEncodeTiled"));
  if (!fn) throw std::runtime_error("[vattn] cuTensorMapEncodeTiled not available (libcuda missing?)");
  return fn;
}
}  // namespace

CUtensorMap make_headdim128_map(const void* base, int64_t seq_extent, int64_t heads, int64_t slots,
                                int64_t row_stride_bytes, int64_t head_stride_bytes,
                                int64_t batch_stride_bytes, int box_rows, int box_atoms) {
  CUtensorMap m;
  // a size-1 dimension may carry any stride; TMA wants a non-zero multiple of 16
  auto fix = [](int64_t s) { return static_cast<cuuint64_t>(s > 0 ? s : 16); };
  const cuuint64_t dims[5] = {64, static_cast<cuuint64_t>(seq_extent), 2, static_cast<cuuint64_t>(heads),
                              static_cast<cuuint64_t>(slots)};
  const cuuint64_t strides[4] = {fix(row_stride_bytes), 128, fix(head_stride_bytes), fix(batch_stride_bytes)};
  const cuuint32_t box[5] = {64, static_cast<cuuint32_t>(box_rows), static_cast<cuuint32_t>(box_atoms), 1, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encoder()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 5, const_cast<void*>(base), dims, strides,
                         box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw std::runtime_error("[vattn] cuTensorMapEncodeTiled failed (" + std::to_string((int)r) +
                             "): base/strides must be 16-byte aligned");
  return m;
}

CUtensorMap make_kmajor_map(const void* base, int64_t rows, int64_t cols, int64_t row_stride_bytes, int box_rows) {
  CUtensorMap m;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(row_stride_bytes)};
  const cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = encoder()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), dims, strides, box,
                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw std::runtime_error("[vattn] cuTensorMapEncodeTiled (2-D) failed (" + std::to_string((int)r) +
                             "): base/strides must be 16-byte aligned");
  return m;
}

int safe_tail_rows(int64_t pitch) {
  if (pitch <= 0) return 0;
  if (16384 % pitch == 0) return 128;
  const int64_t page = 2ll << 20;  // minimum VMM granularity on B200; larger pages are multiples of it
  if (page % pitch != 0) return 0;
  const int64_t tpp = page / pitch;
  int r = 128;
  while (r > 1 && tpp % r != 0) r >>= 1;
  return r >= 8 ? r : 0;
}

}  // namespace vattn
