// fp32-accurate tensor-core GEMM (3xTF32 on tcgen05, accumulators in TMEM):
//   C[M,N] = act(A[M,K] . B[N,K]^T + bias[N])       A, B row-major with K contiguous ("K-major")
// One 128-thread CTA per 128-row tile; N <= 256 columns live in TMEM (one fp32 column each).  K is consumed in
// chunks of 32 floats: all threads stage the chunk of A and B into shared memory in the canonical no-swizzle
// K-major UMMA layout, split into TF32 hi / lo parts, through a 2-stage ring; one elected thread issues
// 3 x 4 tcgen05.mma (128 x N x 8) per chunk and commits them to the stage's mbarrier, which frees the stage.
// Epilogue: each warp reads its 32 TMEM lanes (= 32 output rows) with tcgen05.ld and writes rows of C.
#include "common.cuh"
#include "tc_gemm.cuh"

namespace {

constexpr int kTcThreads = 128;
constexpr int kTcM = 128;
constexpr int kTcKC = 32;        // floats of K per stage
constexpr int kTcStages = 2;

struct TcGemmArgs {
  int M, N, K;
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  float* C; int64_t ldc;
  const float* bias;
  int relu;
};

// stage a [rows x 32] chunk (row-major source, leading dimension ld) as hi / lo canonical blocks
__device__ __forceinline__ void stage_chunk(const float* __restrict__ src, int64_t ld, int row0, int rows_total,
                                            int rows_tile, int k0, int K, uint8_t* hi, uint8_t* lo, int tid) {
  // one float4 (= one core-matrix row) per thread and iteration; lanes walk rows so that the 16-byte shared
  // stores of a quarter-warp fill one 128-byte core matrix (conflict-free)
  const int n4 = rows_tile * (kTcKC / 4);
  for (int idx = tid; idx < n4; idx += kTcThreads) {
    const int r8 = idx & 7, k4 = (idx >> 3) & 7, rg = idx >> 6;
    const int row = rg * 8 + r8;
    const int grow = row0 + row, gk = k0 + k4 * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (grow < rows_total) {
      if (gk + 3 < K) {
        v = __ldg(reinterpret_cast<const float4*>(src + int64_t(grow) * ld + gk));
      } else {
        const float* p = src + int64_t(grow) * ld;
        if (gk + 0 < K) v.x = __ldg(p + gk + 0);
        if (gk + 1 < K) v.y = __ldg(p + gk + 1);
        if (gk + 2 < K) v.z = __ldg(p + gk + 2);
      }
    }
    float4 h, l;
    tc::split4(v, h, l);
    const uint32_t off = tc::canon_off(row, k4 * 4, kTcKC);
    *reinterpret_cast<float4*>(hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = l;
  }
}

template <int NT>  // columns of the tile (multiple of 16, <= 256); TMEM columns = next power of two >= 32
__global__ void __launch_bounds__(kTcThreads) tc_gemm_kernel(TcGemmArgs g) {
  constexpr uint32_t kCols = NT <= 32 ? 32 : NT <= 64 ? 64 : NT <= 128 ? 128 : 256;
  constexpr int kABytes = kTcM * kTcKC * 4, kBBytes = NT * kTcKC * 4;
  constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
  extern __shared__ __align__(128) uint8_t s_tc[];
  __shared__ __align__(8) uint64_t s_bar[kTcStages];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * kTcM, n0 = blockIdx.y * NT;

  if (warp == 0) tc::tmem_alloc(&s_tmem, kCols);
  if (tid == 0) {
    for (int s = 0; s < kTcStages; ++s) tc::mbar_init(&s_bar[s], 1);
    tc::fence_mbar_init();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem;
  constexpr uint32_t idesc = tc::idesc_tf32(kTcM, NT);

  const int nchunks = (g.K + kTcKC - 1) / kTcKC;
  uint32_t phase[kTcStages] = {0, 0};
  for (int c = 0; c < nchunks; ++c) {
    const int s = c % kTcStages;
    uint8_t* a_hi = s_tc + s * kStageBytes;
    uint8_t* a_lo = a_hi + kABytes;
    uint8_t* b_hi = a_lo + kABytes;
    uint8_t* b_lo = b_hi + kBBytes;
    if (c >= kTcStages) {  // the MMAs that read this stage two chunks ago must have completed
      tc::mbar_wait(&s_bar[s], phase[s]);
      phase[s] ^= 1;
    }
    stage_chunk(g.A, g.lda, m0, g.M, kTcM, c * kTcKC, g.K, a_hi, a_lo, tid);
    stage_chunk(g.B, g.ldb, n0, g.N, NT, c * kTcKC, g.K, b_hi, b_lo, tid);
    tc::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      const uint32_t sbo = (kTcKC / 4) * 128, lbo = 128;
#pragma unroll
      for (int ks = 0; ks < kTcKC / 8; ++ks) {
        const uint64_t dah = tc::smem_desc(tc::smem_u32(a_hi) + ks * 256, lbo, sbo);
        const uint64_t dal = tc::smem_desc(tc::smem_u32(a_lo) + ks * 256, lbo, sbo);
        const uint64_t dbh = tc::smem_desc(tc::smem_u32(b_hi) + ks * 256, lbo, sbo);
        const uint64_t dbl = tc::smem_desc(tc::smem_u32(b_lo) + ks * 256, lbo, sbo);
        tc::mma_tf32(tmem, dal, dbh, idesc, (c | ks) != 0);   // small terms first
        tc::mma_tf32(tmem, dah, dbl, idesc, true);
        tc::mma_tf32(tmem, dah, dbh, idesc, true);
      }
      tc::mma_commit(&s_bar[s]);
    }
  }
  // drain: wait for the last commit of every stage that was used
  for (int s = 0; s < kTcStages && s < nchunks; ++s) tc::mbar_wait(&s_bar[s], phase[s]);
  tc::fence_after_sync();

  // epilogue: warp w owns TMEM lanes [32w, 32w+32) = rows m0 + 32w + lane
  const int row = m0 + warp * 32 + lane;
#pragma unroll
  for (int cb = 0; cb < NT; cb += 32) {
    float v[32];
    tc::tmem_ld32(tmem + (uint32_t(warp * 32) << 16) + cb, v);
    if (row < g.M) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int n = n0 + cb + j;
        if (cb + j < NT && n < g.N) {
          float o = v[j];
          if (g.bias) o += __ldg(g.bias + n);
          if (g.relu) o = fmaxf(o, 0.0f);
          g.C[int64_t(row) * g.ldc + n] = o;
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, kCols);
}

template <int NT>
int launch_tc(const TcGemmArgs& g, cudaStream_t s) {
  constexpr size_t smem = size_t(kTcStages) * (2 * kTcM * kTcKC * 4 + 2 * NT * kTcKC * 4) + 128;
  auto kernel = tc_gemm_kernel<NT>;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess)
    return MAPF_ERR_CUDA;
  dim3 grid((g.M + kTcM - 1) / kTcM, (g.N + NT - 1) / NT);
  kernel<<<grid, kTcThreads, smem, s>>>(g);
  return mapf_launch_status();
}

}  // namespace

// C[M,N] = act(A[M,K] B[N,K]^T + bias): fp32 in / out, 3xTF32 on the tensor cores.
extern "C" int mapf_tc_gemm(int32_t M, int32_t N, int32_t K, const float* A, int64_t lda, const float* B, int64_t ldb,
                            float* C, int64_t ldc, const float* bias, int32_t relu, mapf_stream_t stream) {
  MAPF_REQUIRE(A && B && C, MAPF_ERR_NULL);
  MAPF_REQUIRE(M > 0 && N > 0 && K > 0, MAPF_ERR_SHAPE);
  MAPF_REQUIRE((lda & 3) == 0 && (ldb & 3) == 0 && mapf_aligned(A, 16) && mapf_aligned(B, 16), MAPF_ERR_ALIGN);
  TcGemmArgs g{M, N, K, A, lda, B, ldb, C, ldc, bias, relu};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (N <= 64) return launch_tc<64>(g, s);
  return launch_tc<128>(g, s);
}
