// Host-side TMA tensor-map construction for K/V cache views and Q/O tensors.
// cuTensorMapEncodeTiled is resolved from libcuda at run time (vmm_driver.cpp).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace vattn {

// 5-D map over a [slots, S, H, 128] 16-bit tensor with arbitrary outer strides, arranged
// so that ONE box of (64, rows, 2, 1, 1) lands in shared memory as two consecutive
// [rows x 128 B] SWIZZLE_128B atoms (dims 0..63 then 64..127) -- the canonical UMMA
// layout for a 128-wide head dimension:
//   dim0 = d % 64 (stride 2 B)       dim1 = row   (row stride)     dim2 = d / 64 (128 B)
//   dim3 = head  (head stride)       dim4 = slot  (batch stride)
// Rows past `seq_extent` are zero-filled by the TMA unit without touching memory.
CUtensorMap make_headdim128_map(const void* base, int64_t seq_extent, int64_t heads, int64_t slots,
                                int64_t row_stride_bytes, int64_t head_stride_bytes,
                                int64_t batch_stride_bytes, int box_rows, int box_atoms = 2);

// 2-D map over a [rows, cols] 16-bit K-major operand (GEMM A or B): one box of (64 cols, box_rows)
// lands as a [box_rows x 128 B] SWIZZLE_128B atom column; rows past `rows` are zero-filled.
CUtensorMap make_kmajor_map(const void* base, int64_t rows, int64_t cols, int64_t row_stride_bytes, int box_rows);

// rows per TMA box that can never reach past a request's mapped prefix: 128 when the row pitch
// divides 16 KB (tokens_per_page is a multiple of 128), otherwise the largest power of two <= 128
// dividing tokens_per_page = 2 MiB / pitch (megacache views).  0 = layout not usable with TMA.
int safe_tail_rows(int64_t row_pitch_bytes);

}  // namespace vattn
