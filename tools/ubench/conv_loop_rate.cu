// Micro-benchmark of the encoder's conv inner loop (lane = agent image, warp = output-channel pair, FFMA2) to separate
// the cost of the shared-memory operand loads from the FFMA2 issue rate.  MODE 0: as in the kernel; 1: activations kept
// in registers (no v loads); 2: filter taps kept in registers (no w loads); 3: neither (pure FFMA2 + loop).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o conv_loop_rate conv_loop_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(float2& d, float a, const float2& b) {
  asm volatile("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2, %2}; mov.b64 rb, {%3, %4}; mov.b64 rc, {%0, %1};\n"
               "fma.rn.f32x2 rc, ra, rb, rc; mov.b64 {%0, %1}, rc;}" : "+f"(d.x), "+f"(d.y) : "f"(a), "f"(b.x), "f"(b.y));
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(float* out, long long* clk, int cin, int reps) {
  extern __shared__ float sm[];
  float* act = sm;                 // [16*25][32]
  float* w = sm + 16 * 25 * 32;    // [4][16][9][4]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 16 * 25 * 32; i += 256) act[i] = (i % 13) * 0.01f;
  for (int i = tid; i < 4 * 16 * 9 * 4; i += 256) w[i] = (i % 7) * 0.02f - 0.05f;
  __syncthreads();
  float2 acc[25];
#pragma unroll
  for (int p = 0; p < 25; ++p) acc[p] = make_float2(0.f, 0.f);
  const int pr = warp;
  const float2* wp = reinterpret_cast<const float2*>(w) + (pr >> 1) * cin * 9 * 2 + (pr & 1);
  float v[25];
  float2 wv[9];
#pragma unroll
  for (int p = 0; p < 25; ++p) v[p] = act[p * 32 + lane];
#pragma unroll
  for (int t = 0; t < 9; ++t) wv[t] = wp[t * 2];
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll 2
    for (int ci = 0; ci < cin; ++ci) {
      if (MODE == 0 || MODE == 2) {
        const float* src = act + ci * 25 * 32 + lane;
#pragma unroll
        for (int p = 0; p < 25; ++p) v[p] = src[p * 32];
      }
      if (MODE == 0 || MODE == 1) {
#pragma unroll
        for (int t = 0; t < 9; ++t) wv[t] = wp[(ci * 9 + t) * 2];
      }
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int y = 0; y < 5; ++y)
#pragma unroll
            for (int x = 0; x < 5; ++x) {
              const int yy = y + ky - 1, xx = x + kx - 1;
              if (yy >= 0 && yy < 5 && xx >= 0 && xx < 5) ffma2(acc[y * 5 + x], v[yy * 5 + xx], wv[ky * 3 + kx]);
            }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int p = 0; p < 25; ++p) s += acc[p].x + acc[p].y;
  out[blockIdx.x * 256 + tid] = s;
  if (tid == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

int main() {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&clk, 8);
  const int smem = (16 * 25 * 32 + 4 * 16 * 9 * 4) * 4, cin = 16, reps = 50;
  for (int mode = 0; mode < 4; ++mode) {
    long long h = 0;
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) { cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<0><<<148, 256, smem>>>(out, clk, cin, reps); }
      if (mode == 1) { cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<1><<<148, 256, smem>>>(out, clk, cin, reps); }
      if (mode == 2) { cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<2><<<148, 256, smem>>>(out, clk, cin, reps); }
      if (mode == 3) { cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<3><<<148, 256, smem>>>(out, clk, cin, reps); }
      cudaDeviceSynchronize();
    }
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    const double ffma2s = double(reps) * cin * 169 * 2;   // per scheduler (2 warps)
    printf("mode %d: %.3f clk per FFMA2 per scheduler (%lld clk)\n", mode, h / ffma2s, h);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
