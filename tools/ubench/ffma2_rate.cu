// Micro-benchmark: issue rate of packed fp32 FMA (fma.rn.f32x2 -> FFMA2) vs scalar FFMA per SM sub-partition, as a
// function of resident warps per scheduler.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(float2& d, float a, const float2& b) {
  asm volatile("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2, %2}; mov.b64 rb, {%3, %4}; mov.b64 rc, {%0, %1};\n"
               "fma.rn.f32x2 rc, ra, rb, rc; mov.b64 {%0, %1}, rc;}" : "+f"(d.x), "+f"(d.y) : "f"(a), "f"(b.x), "f"(b.y));
}

template <int MODE>
__global__ void k(float* out, long long* clk, int iters, float s) {
  float2 acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = s + i * 0.25f + threadIdx.x;
  float2 w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) w[i] = make_float2(s * (i + 1), s * (i + 2));
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (MODE == 0) ffma2(acc[r * 4 + c], a[r], w[c]);                     // FFMA2, weight pair reused
        if (MODE == 1) { acc[r * 4 + c].x = fmaf(a[r], w[c].x, acc[r * 4 + c].x); acc[r * 4 + c].y = fmaf(a[r], w[c].y, acc[r * 4 + c].y); }
        if (MODE == 2) ffma2(acc[r * 4 + c], a[(r + c) & 7], w[(r + c) & 3]);  // FFMA2, no operand reuse between neighbours
      }
  }
  long long t1 = clock64();
  float sum = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) sum += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

int main() {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int warps = 4; warps <= 32; warps *= 2) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(out, clk, iters, 1.0001f);
        if (mode == 1) k<1><<<148, warps * 32>>>(out, clk, iters, 1.0001f);
        if (mode == 2) k<2><<<148, warps * 32>>>(out, clk, iters, 1.0001f);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
      const double pairs = double(iters) * 32 * (warps / 4);   // packed-FMA pairs issued per scheduler
      printf("mode %d (%s) warps/SM %2d: %.2f clk per 2-FMA pair per scheduler -> %.1f FMA/clk/SM\n", mode,
             mode == 0 ? "FFMA2 reuse" : mode == 1 ? "2x FFMA" : "FFMA2 no-reuse", warps, h / pairs, 4 * 64.0 / (h / pairs));
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
