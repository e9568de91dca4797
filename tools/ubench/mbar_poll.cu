// Microbenchmark: cost of probing an mbarrier whose phase has long completed — mbarrier.try_wait (potentially blocking form)
// vs mbarrier.test_wait (non-blocking test), one warp, dependent probes.   nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void poll_kernel(long long* out, int iters) {
  __shared__ __align__(8) unsigned long long bar;
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory");   // phase 0 completes
  }
  __syncthreads();
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(b), "r"(acc & 0u) : "memory");
    acc += done;
  }
  long long t1 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(b), "r"(acc & 0u) : "memory");
    acc += done;
  }
  long long t2 = clock64();
  // a not-yet-complete phase (parity 1): what one failed probe costs
  for (int i = 0; i < iters; ++i) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(b), "r"(1u | (acc & 0u)) : "memory");
    acc += done;
  }
  long long t3 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(b), "r"(1u | (acc & 0u)) : "memory");
    acc += done;
  }
  long long t4 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3; out[4] = acc; }
}

int main() {
  long long* d; cudaMalloc(&d, 5 * sizeof(long long));
  const int iters = 2000;
  for (int threads : {32, 128}) {
    poll_kernel<<<1, threads>>>(d, iters);
    cudaDeviceSynchronize();
    long long h[5]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("threads %3d: try_wait(complete) %.1f clk, test_wait(complete) %.1f clk, try_wait(pending) %.1f clk, test_wait(pending) %.1f clk  [%lld]\n",
           threads, (double)h[0] / iters, (double)h[1] / iters, (double)h[2] / iters, (double)h[3] / iters, h[4]);
  }
  return 0;
}
