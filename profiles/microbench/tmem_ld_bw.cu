// Microbenchmark: tcgen05.ld throughput per SM on sm_100a (how fast can an epilogue drain TMEM?).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bw tmem_ld_bw.cu && ./tmem_ld_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int SHAPE>  // 0: 32x32b.x32, 1: 32x32b.x16, 2: 16x256b.x8 (32 regs), 3: 32x32b.x64
__global__ void k(int iters, int nwarps_active, unsigned long long *out) {
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = s_tmem + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  long long t0 = clock64();
  if (warp < nwarps_active) {
    for (int it = 0; it < iters; ++it) {
      const uint32_t col = (uint32_t)((it * 32 + (warp >> 2) * 64) & 255);
      if (SHAPE == 0) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(base + col) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= r[i];
      } else if (SHAPE == 1) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(base + col) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) acc ^= r[i];
      } else if (SHAPE == 2) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(base + col) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= r[i];
      }
    }
  }
  long long t1 = clock64();
  if (acc == 0x12345678u) out[100] = acc;
  if (lane == 0 && warp < nwarps_active) atomicMax(out + blockIdx.x % 1, (unsigned long long)(t1 - t0));
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(512u) : "memory");
  }
}

int main() {
  unsigned long long *d, h;
  cudaMalloc(&d, 1024);
  const int iters = 20000;
  const char *names[] = {"32x32b.x32 (4 KB/warp-load)", "32x32b.x16 (2 KB)", "16x256b.x8 (4 KB)"};
  for (int shape = 0; shape < 3; ++shape)
    for (int nw : {1, 4, 8, 16, 24}) {
      cudaMemset(d, 0, 1024);
      const int threads = 32 * (nw < 4 ? 4 : nw);
      if (shape == 0) k<0><<<1, threads>>>(iters, nw, d);
      if (shape == 1) k<1><<<1, threads>>>(iters, nw, d);
      if (shape == 2) k<2><<<1, threads>>>(iters, nw, d);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      const double bytes = (shape == 1 ? 2048.0 : 4096.0) * iters * nw;
      printf("%-28s warps %2d: %8.1f clk/load/warp, %7.1f B/clk/SM  (%s)\n", names[shape], nw, (double)h / iters,
             bytes / (double)h, cudaGetErrorString(e));
    }
  return 0;
}
