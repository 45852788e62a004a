// Probe: cp.async.bulk.tensor.2d...tile::gather4 on sm_100a -- four arbitrary rows of a 2-D bf16
// tensor per instruction.  Questions for the round-2 graph-conv design:
//   * which tensor-map box does the instruction want ({64, 1} / {64, 4})?
//   * do the four rows land as four consecutive 128-B rows of shared memory?
//   * is CU_TENSOR_MAP_SWIZZLE_128B honoured (so that the rows can feed a K-major UMMA descriptor)?
//   * how many bytes does the transaction count (4 x 128)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I../../realtime-st-gcn_b200/csrc -lcuda tma_gather4.cu
//
// Measured on B200 (round 1): box {64, 1} works -- rows {5, 200, 17, 3} land as four consecutive 128-B
// rows, 512 transaction bytes, and with CU_TENSOR_MAP_SWIZZLE_128B the 16-B chunks of destination row i
// are XOR-permuted by (i & 7) exactly like a tiled load (so the rows can feed a K-major SWIZZLE_128B UMMA
// descriptor); box {64, 4} raises "illegal instruction".
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "kernels_tc.cuh"
using namespace stgcn::tc;

__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap *m, uint32_t bar, int col, int r0, int r1,
                                            int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tm, uint16_t *out, int tx_bytes) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t *gen = raw + (base - smem_u32(raw));
  const uint32_t bar = base + 4096;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) reinterpret_cast<uint32_t *>(gen)[i] = 0xdeaddeadu;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    fence_proxy_async();
    mbar_expect_tx(bar, (uint32_t)tx_bytes);
    tma_gather4(base, &tm, bar, 0, 5, 200, 17, 3);
  }
  mbar_wait(bar, 0);
  __syncthreads();
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) out[i] = reinterpret_cast<uint16_t *>(gen)[i];
}

int main(int argc, char **argv) {
  const int rows = 256, cols = 64;
  std::vector<uint16_t> h(rows * cols);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) h[r * cols + c] = (uint16_t)(r * 64 + c);   // raw 16-bit payload: row*64 + col
  uint16_t *d, *o;
  cudaMalloc(&d, h.size() * 2);
  cudaMalloc(&o, 4096);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  for (int box_rows = 1; box_rows <= 4; box_rows += 3)
    for (int sw = 0; sw < 2; ++sw) {
      CUtensorMap tm;
      const uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows};
      const uint64_t strides[1] = {(uint64_t)cols * 2};
      const uint32_t box[2] = {(uint32_t)cols, (uint32_t)box_rows};
      if (make_tmap(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, d,
                    2, dims, strides, box)) {
        printf("box rows %d swizzle %d: tensor map rejected: %s\n", box_rows, sw, stgcn::err_buf());
        continue;
      }
      cudaMemset(o, 0, 4096);
      probe<<<1, 128, 8192>>>(tm, o, 512);
      cudaError_t e = cudaDeviceSynchronize();
      printf("box rows %d swizzle %d: %s\n", box_rows, sw, cudaGetErrorString(e));
      if (e != cudaSuccess) return 1;
      std::vector<uint16_t> r(2048);
      cudaMemcpy(r.data(), o, 4096, cudaMemcpyDeviceToHost);
      for (int row = 0; row < 6; ++row) {
        printf("  smem row %d:", row);
        for (int chunk = 0; chunk < 8; ++chunk) {
          const uint16_t v = r[row * 64 + chunk * 8];
          if (v == 0xdead) printf(" [----]");
          else printf(" [r%3d c%2d]", v / 64, v % 64);
        }
        printf("\n");
      }
    }
  return 0;
}
