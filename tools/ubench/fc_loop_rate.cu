// Micro-benchmark of the encoder's FC inner loop (thread tile 4 agents x 8 outputs, K split 4 ways over 256 threads).
// MODE 0: as in the kernel; 1: weights hoisted (no w loads); 2: activations hoisted; 3: both hoisted.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fc_loop_rate fc_loop_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(float2& d, float a, const float2& b) {
  asm volatile("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2, %2}; mov.b64 rb, {%3, %4}; mov.b64 rc, {%0, %1};\n"
               "fma.rn.f32x2 rc, ra, rb, rc; mov.b64 {%0, %1}, rc;}" : "+f"(d.x), "+f"(d.y) : "f"(a), "f"(b.x), "f"(b.y));
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(float* out, long long* clk, int fc_in, int fc_out, int reps) {
  extern __shared__ __align__(16) float sm[];
  float* act = sm;                  // [fc_in][32]
  float* wt = sm + fc_in * 32;      // [fc_in][fc_out]
  const int tid = threadIdx.x;
  for (int i = tid; i < fc_in * 32; i += 256) act[i] = (i % 13) * 0.01f;
  for (int i = tid; i < fc_in * fc_out; i += 256) wt[i] = (i % 7) * 0.02f - 0.05f;
  __syncthreads();
  const int kg = tid >> 6, t6 = tid & 63, aq = t6 & 7, jq = t6 >> 3;
  const int kq = fc_in >> 2, k0 = kg * kq, k1 = k0 + kq, wstride = fc_out >> 2, j = jq * 8;
  float2 acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int h = 0; h < 4; ++h) acc[r][h] = make_float2(0.f, 0.f);
  const float4* wps = reinterpret_cast<const float4*>(wt + j);
  const float4* ap = reinterpret_cast<const float4*>(act + aq * 4);
  long long t0 = clock64();
  for (int rp = 0; rp < reps; ++rp) {
#pragma unroll 4
    for (int i = k0; i < k1; ++i) {
      const int ia = (MODE & 2) ? k0 + ((i - k0) & 3) : i, iw = (MODE & 1) ? k0 + ((i - k0) & 3) : i;
      const float4 a4 = ap[ia * 8];
      const float4 w0 = wps[iw * wstride], w1 = wps[iw * wstride + 1];
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      const float2 wp0 = make_float2(w0.x, w0.y), wp1 = make_float2(w0.z, w0.w);
      const float2 wp2 = make_float2(w1.x, w1.y), wp3 = make_float2(w1.z, w1.w);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        ffma2(acc[r][0], av[r], wp0); ffma2(acc[r][1], av[r], wp1);
        ffma2(acc[r][2], av[r], wp2); ffma2(acc[r][3], av[r], wp3);
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int h = 0; h < 4; ++h) s += acc[r][h].x + acc[r][h].y;
  out[blockIdx.x * 256 + tid] = s;
  if (tid == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

int main() {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&clk, 8);
  const int fc_in = 400, fc_out = 64, reps = 20, smem = (fc_in * 32 + fc_in * fc_out) * 4;
  for (int mode = 0; mode < 4; ++mode) {
    long long h = 0;
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) { cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<0><<<148, 256, smem>>>(out, clk, fc_in, fc_out, reps); }
      if (mode == 1) { cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<1><<<148, 256, smem>>>(out, clk, fc_in, fc_out, reps); }
      if (mode == 2) { cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<2><<<148, 256, smem>>>(out, clk, fc_in, fc_out, reps); }
      if (mode == 3) { cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<3><<<148, 256, smem>>>(out, clk, fc_in, fc_out, reps); }
      cudaDeviceSynchronize();
    }
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    const double ffma2s = double(reps) * (fc_in / 4) * 16 * 2;   // per scheduler (2 warps)
    printf("mode %d: %.3f clk per FFMA2 per scheduler (%lld clk, %.0f per pass)\n", mode, h / ffma2s, h, double(h) / reps);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
