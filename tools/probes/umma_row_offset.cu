// Probe: does a K-major SWIZZLE_128B UMMA descriptor accept a start address that is offset by
// whole 128-B rows (not a multiple of the 1024-B swizzle atom)?  Variants: base_offset field 0,
// and base_offset = (addr >> 7) & 7.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -I../../realtime-st-gcn_b200/csrc
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "kernels_tc.cuh"
using namespace stgcn::tc;

__device__ __forceinline__ uint64_t desc_var(uint32_t addr, int use_base_offset) {
  uint64_t d = umma_desc_sw128(addr);
  if (use_base_offset) d |= (uint64_t)((addr >> 7) & 7) << 49;
  return d;
}

__global__ void __launch_bounds__(128, 1)
    probe(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, float *out,
          int n_off) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 256 * 128, bar = sB + 64 * 128, bar2 = bar + 8, tptr = bar + 16;
  volatile uint32_t *tp = reinterpret_cast<volatile uint32_t *>(raw + (tptr - smem_u32(raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar2, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tptr, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tp;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 256 * 128 + 64 * 128);
    tma_load_4d(sA, &tm_a, bar, 0, 0, 0, 0);
    tma_load_4d(sB, &tm_b, bar, 0, 0, 0, 0);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  uint32_t ph = 0;
  for (int var = 0; var < 2; ++var)
    for (int off = 0; off < n_off; ++off) {
      if (threadIdx.x == 0) {
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem, desc_var(sA + off * 128 + k * 32, var), umma_desc_sw128(sB + k * 32),
                    umma_idesc_bf16(128, 64), k);
        umma_commit(bar2);
      }
      mbar_wait(bar2, ph);
      ph ^= 1;
      tc_fence_after();
      float v[32];
      for (int cb = 0; cb < 64; cb += 32) {
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + cb, v);
        for (int i = 0; i < 32; ++i)
          out[(((size_t)var * n_off + off) * 128 + warp * 32 + lane) * 64 + cb + i] = v[i];
      }
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
  if (warp == 0) tmem_dealloc(tmem, 64);
}

int main() {
  const int R = 256, K = 64, N = 64, NOFF = 40;
  std::vector<__nv_bfloat16> ha(R * K), hb(N * K);
  std::vector<float> fa(R * K), fb(N * K);
  srand(1);
  for (int i = 0; i < R * K; ++i) { fa[i] = (float)(rand() % 7 - 3); ha[i] = __float2bfloat16(fa[i]); }
  for (int i = 0; i < N * K; ++i) { fb[i] = (float)(rand() % 5 - 2); hb[i] = __float2bfloat16(fb[i]); }
  __nv_bfloat16 *da, *db; float *dout;
  cudaMalloc(&da, R * K * 2); cudaMalloc(&db, N * K * 2);
  cudaMalloc(&dout, sizeof(float) * 2 * NOFF * 128 * 64);
  cudaMemcpy(da, ha.data(), R * K * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), N * K * 2, cudaMemcpyHostToDevice);
  CUtensorMap ta, tb;
  uint64_t ad[4] = {64, 256, 1, 1}, as[3] = {128, 128 * 256, 128 * 256};
  uint32_t ab[4] = {64, 256, 1, 1};
  uint64_t bd[4] = {64, 64, 1, 1}, bs[3] = {128, 128 * 64, 128 * 64};
  uint32_t bb[4] = {64, 64, 1, 1};
  if (make_tmap_bf16(&ta, da, 4, ad, as, ab) || make_tmap_bf16(&tb, db, 4, bd, bs, bb)) { printf("tmap fail %s\n", stgcn::err_buf()); return 1; }
  int smem = 256 * 128 + 64 * 128 + 64 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 128, smem>>>(ta, tb, dout, NOFF);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel error %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> ho((size_t)2 * NOFF * 128 * 64);
  cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
  for (int var = 0; var < 2; ++var) {
    printf("base_offset variant %d: ", var);
    for (int off = 0; off < NOFF; ++off) {
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
          float ref = 0;
          for (int k = 0; k < K; ++k) ref += fa[(off + m) * K + k] * fb[n * K + k];
          if (ho[(((size_t)var * NOFF + off) * 128 + m) * 64 + n] != ref) ++bad;
        }
      printf("%s", bad ? "X" : ".");
    }
    printf("\n");
  }
  return 0;
}
