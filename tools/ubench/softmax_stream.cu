// Microbenchmark: the instruction stream of ONE softmax warp of attn_tcgen05_kernel over a 64-key tile, in isolation
// (no MMA, no loaders, no mbarrier hand-overs): tcgen05.ld of the S row, row max, FFMA + MUFU.EX2, bf16 pack, tcgen05.st of P.
// Reports clocks per tile for 1 or 2 warps per SM sub-partition and with parts of the stream switched off, i.e. what the
// pipes themselves allow.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_stream softmax_stream.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../mmduet_b200/csrc/ptx.cuh"
using namespace mmd;

__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MODE bits: 1 = tcgen05.ld, 2 = max, 4 = FFMA, 8 = MUFU, 16 = pack, 32 = tcgen05.st, 64 = polynomial for 3 of 8 pairs,
// 128 = polynomial for 2 of 8 pairs, 256 = polynomial for 1 of 8 pairs
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;
  const float f = x - (t - 12582912.f);
  float p = fmaf(f, 0.0551714599f, 0.2426108569f);
  p = fmaf(p, f, 0.6932609677f);
  p = fmaf(p, f, 0.9999281168f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) stream_kernel(int iters, float sl2, long long* clk_out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = slot;
  const uint32_t t_row = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t a[32], b[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) { a[i] = __float_as_uint(-0.01f * (i + threadIdx.x % 7)); b[i] = __float_as_uint(-0.02f * i); }
  // seed TMEM with finite values
  {
    uint32_t z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = __float_as_uint(-0.5f - 0.01f * i);
    for (int c = 0; c < 128; c += 16) st16(t_row + c, z);
    st_wait();
  }
  __syncthreads();
  float m_ref = 0.f, acc = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int sb = it & 1;
    if (MODE & 1) {
      tmem_ld_32x32b_x32(t_row + sb * 64, a);
      tmem_ld_32x32b_x32(t_row + sb * 64 + 32, b);
      tmem_ld_wait();
    }
    auto val = [&](int i) { return __uint_as_float(i < 32 ? a[i] : b[i - 32]); };
    if (MODE & 2) {
      float m0 = fmaxf(val(0), val(1)), m1 = fmaxf(val(2), val(3)), m2 = fmaxf(val(4), val(5)), m3 = fmaxf(val(6), val(7));
#pragma unroll
      for (int i = 8; i < 64; i += 8) {
        m0 = fmaxf(m0, fmaxf(val(i), val(i + 1)));
        m1 = fmaxf(m1, fmaxf(val(i + 2), val(i + 3)));
        m2 = fmaxf(m2, fmaxf(val(i + 4), val(i + 5)));
        m3 = fmaxf(m3, fmaxf(val(i + 6), val(i + 7)));
      }
      m_ref = fmaxf(m_ref, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
    }
    const float msc = m_ref * sl2;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float x[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = (MODE & 4) ? fmaf(val(32 * c + i), sl2, -msc) : val(32 * c + i);
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const bool poly = ((MODE & 64) && ((i & 7) == 1 || (i & 7) == 4 || (i & 7) == 6)) || ((MODE & 128) && ((i & 7) == 2 || (i & 7) == 6)) ||
                          ((MODE & 256) && (i & 7) == 3);
        float p0, p1;
        if (MODE & 8) {
          p0 = poly ? exp2_poly(x[2 * i]) : exp2f(x[2 * i]);
          p1 = poly ? exp2_poly(x[2 * i + 1]) : exp2f(x[2 * i + 1]);
        } else { p0 = x[2 * i]; p1 = x[2 * i + 1]; }
        if (MODE & 16) {
          __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
          pk[i] = *reinterpret_cast<uint32_t*>(&h);
        } else {
          pk[i] = __float_as_uint(p0) ^ __float_as_uint(p1);
        }
      }
      if (MODE & 32) st16(t_row + sb * 64 + 16 * c, pk);
      else {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc += __uint_as_float(pk[i]);
      }
    }
    if (MODE & 32) st_wait();
    if (!(MODE & 1)) {   // keep the inputs changing so that nothing is hoisted
#pragma unroll
      for (int i = 0; i < 32; ++i) { a[i] ^= (it & 1); b[i] ^= (it & 1); }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) clk_out[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + m_ref + __uint_as_float(a[3]);
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

template <int MODE>
void run(const char* name, int warps_per_quadrant) {
  const int iters = 2000, blocks = 148, threads = 128 * warps_per_quadrant;
  long long* clk; float* sink;
  cudaMalloc(&clk, blocks * sizeof(long long));
  cudaMalloc(&sink, blocks * 512 * sizeof(float));
  for (int rep = 0; rep < 2; ++rep) stream_kernel<MODE><<<blocks, threads>>>(iters, 0.18f, clk, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  long long h[148];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < blocks; ++i) s += (double)h[i];
  printf("%-46s warps/quadrant %d : %7.1f clk per 64-key tile per warp slot\n", name, warps_per_quadrant, s / blocks / iters);
  cudaFree(clk); cudaFree(sink);
}

int main() {
  for (int w = 1; w <= 2; w *= 2) {
    run<63>("full stream (ld max ffma mufu pack st)", w);
    run<63 + 64>("full stream, 3/8 of the exps on the FMA pipe", w);
    run<63 + 128>("full stream, 2/8 of the exps on the FMA pipe", w);
    run<63 + 256>("full stream, 1/8 of the exps on the FMA pipe", w);
    run<63 - 8>("no MUFU", w);
    run<63 - 16>("no pack (F2FP)", w);
    run<63 - 8 - 16>("no MUFU, no pack", w);
    run<4 + 8>("FFMA + MUFU only", w);
    run<8>("MUFU only", w);
    run<4 + 8 + 16>("FFMA + MUFU + pack", w);
    run<1 + 32>("ld + st only", w);
    run<1>("ld only", w);
    run<1 + 2>("ld + max", w);
  }
  return 0;
}
