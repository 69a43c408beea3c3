// Device self test of the UMMA descriptor encodings used by the attention kernels:
//   test 1  D[128 x 16] = A[128 x 128] . B[16 x 128]^T   A, B K-major SW128 (the QK^T form)
//   test 2  D[128 x 16] = V[128 x 128]^T . P[16 x 128]^T  A read MN-major from the same tile
//           (the PV form), for both candidate (LBO, SBO) assignments
// Operands are written to shared memory by ordinary stores in the swizzled layout (no TMA), so
// a failure here isolates the descriptor, not the copy.  Host reference in double precision.
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "attn_common.cuh"
#include "sm100_ptx.cuh"

namespace vattn {

namespace {
using namespace ptx;

struct __align__(1024) TestSmem {
  uint8_t a[2][128 * 128];  // two 64-col atoms of [128 rows x 128 B]
  uint8_t b[2][16 * 128];   // two 64-col atoms of [16 rows x 128 B]
  uint64_t bar;
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t sw_off(int r, int c) {
  return r * 128 + ((((c >> 3) ^ (r & 7)) << 4) | ((c & 7) << 1));
}

// mode 0: K-major A.  mode 1: MN-major A with (lbo, sbo).
__global__ void __launch_bounds__(128)
umma_test_kernel(const __nv_bfloat16* __restrict__ ag, const __nv_bfloat16* __restrict__ bg,
                 float* __restrict__ dout, int mode, uint32_t lbo, uint32_t sbo) {
  extern __shared__ uint8_t raw[];
  TestSmem& sm = *reinterpret_cast<TestSmem*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int t = threadIdx.x, warp = t >> 5;
  for (int i = t; i < 128 * 128; i += 128) {
    const int r = i / 128, c = i % 128;
    *reinterpret_cast<__nv_bfloat16*>(sm.a[c >> 6] + sw_off(r, c & 63)) = ag[i];
  }
  for (int i = t; i < 16 * 128; i += 128) {
    const int r = i / 128, c = i % 128;
    *reinterpret_cast<__nv_bfloat16*>(sm.b[c >> 6] + sw_off(r, c & 63)) = bg[i];
  }
  if (t == 0) {
    mbar_init(&sm.bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&sm.tmem_base, 32);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  if (t == 0) {
    const uint32_t a0 = smem_u32(sm.a[0]), b0 = smem_u32(sm.b[0]);
    for (int ks = 0; ks < 8; ks++) {
      const uint32_t bb = b0 + (ks >> 2) * (16 * 128) + (ks & 3) * 32;
      if (mode == 0) {
        const uint32_t aa = a0 + (ks >> 2) * (128 * 128) + (ks & 3) * 32;
        umma_ss(tmem, make_smem_desc(aa, 16, 1024, kLayoutSw128), make_smem_desc(bb, 16, 1024, kLayoutSw128),
                make_idesc(kFmtBF16, 128, 16, 0, 0), ks > 0);
      } else {
        const uint32_t aa = a0 + ks * (16 * 128);
        umma_ss(tmem, make_smem_desc(aa, lbo, sbo, kLayoutSw128), make_smem_desc(bb, 16, 1024, kLayoutSw128),
                make_idesc(kFmtBF16, 128, 16, 1, 0), ks > 0);
      }
    }
    umma_commit(&sm.bar);
  }
  mbar_wait(&sm.bar, 0);
  tc_fence_after();
  uint32_t r[16];
  tmem_ld_x16(tmem + ((uint32_t)(warp * 32) << 16), r);
  tmem_wait_ld();
  for (int n = 0; n < 16; n++) dout[t * 16 + n] = __uint_as_float(r[n]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

}  // namespace

int run_umma_selftest(char* buf, size_t len, cudaStream_t stream) {
  std::vector<__nv_bfloat16> a(128 * 128), b(16 * 128);
  std::vector<float> af(128 * 128), bf(16 * 128);
  uint32_t s = 12345;
  auto rnd = [&]() {
    s = s * 1664525u + 1013904223u;
    return ((int)((s >> 9) & 0xff) - 128) / 64.0f;
  };
  for (size_t i = 0; i < a.size(); i++) {
    a[i] = __float2bfloat16(rnd());
    af[i] = __bfloat162float(a[i]);
  }
  for (size_t i = 0; i < b.size(); i++) {
    b[i] = __float2bfloat16(rnd());
    bf[i] = __bfloat162float(b[i]);
  }
  std::vector<double> ref1(128 * 16), ref2(128 * 16);
  for (int m = 0; m < 128; m++)
    for (int n = 0; n < 16; n++) {
      double x = 0, y = 0;
      for (int k = 0; k < 128; k++) {
        x += (double)af[m * 128 + k] * bf[n * 128 + k];  // A[m][k] . B[n][k]
        y += (double)af[k * 128 + m] * bf[n * 128 + k];  // V[key=k][d=m] . P[n][key=k]
      }
      ref1[m * 16 + n] = x;
      ref2[m * 16 + n] = y;
    }
  __nv_bfloat16 *da, *db;
  float* dd;
  VATTN_CUDA(cudaMalloc(&da, a.size() * 2));
  VATTN_CUDA(cudaMalloc(&db, b.size() * 2));
  VATTN_CUDA(cudaMalloc(&dd, 128 * 16 * 4));
  VATTN_CUDA(cudaMemcpyAsync(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice, stream));
  VATTN_CUDA(cudaMemcpyAsync(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice, stream));
  const size_t smem = sizeof(TestSmem) + 1024;
  VATTN_CUDA(cudaFuncSetAttribute(umma_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  struct Case {
    const char* name;
    int mode;
    uint32_t lbo, sbo;
    const std::vector<double>* ref;
  } cases[] = {{"qk  K-major A (sbo 1024)", 0, 16, 1024, &ref1},
               {"pv  MN-major A lbo=16384 sbo=1024", 1, 128 * 128, 1024, &ref2},
               {"pv  MN-major A lbo=1024 sbo=16384", 1, 1024, 128 * 128, &ref2}};
  std::string report;
  int first_bad = 0;
  std::vector<float> out(128 * 16);
  for (int c = 0; c < 3; c++) {
    VATTN_CUDA(cudaMemsetAsync(dd, 0xff, 128 * 16 * 4, stream));
    umma_test_kernel<<<1, 128, smem, stream>>>(da, db, dd, cases[c].mode, cases[c].lbo, cases[c].sbo);
    count_launch();
    VATTN_CUDA(cudaGetLastError());
    VATTN_CUDA(cudaMemcpyAsync(out.data(), dd, out.size() * 4, cudaMemcpyDeviceToHost, stream));
    VATTN_CUDA(cudaStreamSynchronize(stream));
    double maxerr = 0;
    for (size_t i = 0; i < out.size(); i++) {
      double e = std::fabs((double)out[i] - (*cases[c].ref)[i]);
      if (!(e == e)) e = 1e30;
      if (e > maxerr) maxerr = e;
    }
    const bool ok = maxerr < 1e-2;
    char line[160];
    std::snprintf(line, sizeof(line), "%s: max abs err %.3e %s\n", cases[c].name, maxerr, ok ? "OK" : "MISMATCH");
    report += line;
    // case 2 is the alternative encoding: it is expected to mismatch when case 1 passes
    if (!ok && c < 2 && !first_bad) first_bad = c + 1;
  }
  cudaFree(da);
  cudaFree(db);
  cudaFree(dd);
  if (buf && len) {
    std::snprintf(buf, len, "%s", report.c_str());
  }
  return first_bad;
}

}  // namespace vattn
