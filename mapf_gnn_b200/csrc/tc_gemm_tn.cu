// fp32-accurate tensor-core GEMM for the WEIGHT GRADIENTS of the training step (train.py:96 backward): both operands are
// activation matrices whose contraction index is the (huge) row index,
//   C[m, n] += sum_r A[r, m] * B[r, n]            A [K, M] row-major (m contiguous), B [K, N] row-major, K = B*N rows
// e.g. dW_eff = dY^T Z (gnn.py:114-118 backward), dW_fc = dX^T a_L, dW_a = dlogits^T relu(Y).  M and N are at most a few
// hundred, K is 10^4 .. 10^6, so the grid is (M tiles, N tiles, K splits) and the partial products of the K splits meet in
// C through red.global.add (C is zeroed first unless `accumulate`).
//
// 3xTF32 on tcgen05 like tc_gemm.cu (a = hi + lo; lo*hi + hi*lo + hi*hi; fp32 accumulators in TMEM, consecutive chunks
// alternate between two accumulators that are summed in the epilogue).  What is new is the operand path: tcgen05 wants
// K-major operands (contraction index contiguous) and here it is the slow index, so the six producer warps TRANSPOSE on
// the way into shared memory -- a lane reads float4s along m (coalesced 512-byte rows), keeps KC = 16 rows in registers
// and writes, for each of its 4 columns, four k-contiguous float4s into the canonical no-swizzle K-major layout
// [(k/4)][(row/8)][row%8][k%4].  The 8-row groups are 144 bytes apart (SBO = 144, not 128): with 128 the eight lanes of
// a quarter warp would hit the same banks four at a time, with 144 every 128-bit store is conflict free.
//   warps 0-11   producers: warps w and w + 6 own stage w (chunk c = w, w+6, ...), one the A chunk, the other the B
//                chunk, hi and lo (12 converting warps: the kernel is paced by its producers, not by the tensor pipe)
//   warps 12,13  MMA issuers (even / odd chunks, one accumulator each; elect.sync inside a warp-uniform branch)
//   all 16 warps epilogue: TMEM -> registers -> red.global.add.f32 (.v4 along n, or scalar along m for a transposed C)
#include "common.cuh"
#include "tc_gemm.cuh"

namespace {

constexpr int kTnThreads = 512;
constexpr int kTnProd = 6;                     // stages; two producer warps each
constexpr int kTnT = 128;                      // tile edge in M and in N
constexpr int kTnKC = 16;                      // contraction rows per stage
constexpr int kTnSbo = 144;                    // bytes between 8-row groups
constexpr int kTnLbo = (kTnT / 8) * kTnSbo;    // bytes between K-adjacent core matrices (2304)
constexpr int kTnBlock = (kTnKC / 4) * kTnLbo; // bytes of one (operand, hi|lo) block of a stage (9216)
constexpr int kTnStage = 4 * kTnBlock;         // A hi | A lo | B hi | B lo

struct TnArgs {
  int M, N;
  int64_t K;
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  float* C; int64_t ldc;
  int transpose_c;        // 0: C[m * ldc + n]   1: C[n * ldc + m]
  int chunks_per_split;   // KC-row chunks per blockIdx.z
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// one operand chunk: rows k0 .. k0+KC of src ([K, W] row-major, ld) restricted to columns c0 .. c0+128 -> hi / lo blocks
__device__ __forceinline__ void tn_load(const float* __restrict__ src, int64_t ld, int W, int64_t K, int64_t k0, int c0,
                                        int lane, float4 (&v)[kTnKC]) {
  const int c = c0 + 4 * lane;
  const bool vec = (ld & 3) == 0 && c + 3 < W && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
#pragma unroll
  for (int k = 0; k < kTnKC; ++k) {
    v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k0 + k < K) {
      const float* p = src + (k0 + k) * ld + c;
      if (vec) {
        v[k] = mapf_ld_dep(reinterpret_cast<const float4*>(p));
      } else {
        if (c + 0 < W) v[k].x = mapf_ld_dep(p + 0);
        if (c + 1 < W) v[k].y = mapf_ld_dep(p + 1);
        if (c + 2 < W) v[k].z = mapf_ld_dep(p + 2);
        if (c + 3 < W) v[k].w = mapf_ld_dep(p + 3);
      }
    }
  }
}

__device__ __forceinline__ void tn_store(uint8_t* hi_blk, uint8_t* lo_blk, int lane, const float4 (&v)[kTnKC]) {
  // lane holds columns 4*lane + j (j = 0..3) of KC rows; column = MMA row index r
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = 4 * lane + j;
    const uint32_t roff = uint32_t(r >> 3) * kTnSbo + uint32_t(r & 7) * 16;
#pragma unroll
    for (int q = 0; q < kTnKC / 4; ++q) {
      float4 h, l;
      const float a0 = j == 0 ? v[4 * q].x : j == 1 ? v[4 * q].y : j == 2 ? v[4 * q].z : v[4 * q].w;
      const float a1 = j == 0 ? v[4 * q + 1].x : j == 1 ? v[4 * q + 1].y : j == 2 ? v[4 * q + 1].z : v[4 * q + 1].w;
      const float a2 = j == 0 ? v[4 * q + 2].x : j == 1 ? v[4 * q + 2].y : j == 2 ? v[4 * q + 2].z : v[4 * q + 2].w;
      const float a3 = j == 0 ? v[4 * q + 3].x : j == 1 ? v[4 * q + 3].y : j == 2 ? v[4 * q + 3].z : v[4 * q + 3].w;
      tc::split4(make_float4(a0, a1, a2, a3), h, l);
      *reinterpret_cast<float4*>(hi_blk + q * kTnLbo + roff) = h;
      *reinterpret_cast<float4*>(lo_blk + q * kTnLbo + roff) = l;
    }
  }
}

__global__ void __launch_bounds__(kTnThreads, 1) tc_gemm_tn_kernel(TnArgs g) {
  extern __shared__ __align__(128) uint8_t s_tn[];
  __shared__ __align__(8) uint64_t s_full[kTnProd], s_empty[kTnProd], s_done;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * kTnT, n0 = blockIdx.y * kTnT;
  const int64_t nchunks_all = (g.K + kTnKC - 1) / kTnKC;
  const int64_t c_begin = int64_t(blockIdx.z) * g.chunks_per_split;
  const int64_t c_end = c_begin + g.chunks_per_split < nchunks_all ? c_begin + g.chunks_per_split : nchunks_all;
  const int nchunks = c_end > c_begin ? int(c_end - c_begin) : 0;

  if (warp == 0) tc::tmem_alloc(&s_tmem, 256);
  if (tid == 0) {
    for (int s = 0; s < kTnProd; ++s) { tc::mbar_init(&s_full[s], 64); tc::mbar_init(&s_empty[s], 1); }
    tc::mbar_init(&s_done, 2);
    tc::fence_mbar_init();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
  mapf_pdl_wait();      // both operands come from kernels in front
  mapf_pdl_launch();

  if (warp_u < 2 * kTnProd) {
    // ===== producers: warps w and w + 6 own stage w; the first converts the A chunk, the second the B chunk =====
    const int stage = warp_u % kTnProd, opnd = warp_u / kTnProd;
    uint8_t* st = s_tn + stage * kTnStage + opnd * (2 * kTnBlock);
    const float* src = opnd == 0 ? g.A : g.B;
    const int64_t ld = opnd == 0 ? g.lda : g.ldb;
    const int W = opnd == 0 ? g.M : g.N, c0 = opnd == 0 ? m0 : n0;
    float4 v[kTnKC];
    uint32_t ph = 0;
    for (int c = stage; c < nchunks; c += kTnProd, ph ^= 1) {
      tn_load(src, ld, W, g.K, (c_begin + c) * kTnKC, c0, lane, v);
      tc::mbar_wait(&s_empty[stage], ph ^ 1);
      tn_store(st, st + kTnBlock, lane, v);
      tc::fence_proxy_async();
      tc::mbar_arrive(&s_full[stage]);
    }
  } else if (warp_u < 2 * kTnProd + 2) {
    // ===== MMA issuers (warps 12 and 13: even / odd chunks, one accumulator each) =====
    if (tc::elect_one()) {
      constexpr uint32_t idesc = tc::idesc_tf32(kTnT, kTnT);
      const uint32_t base = tc::smem_u32(s_tn);
      const uint64_t dah0 = tc::smem_desc(base, kTnLbo, kTnSbo);
      const uint64_t dal0 = tc::smem_desc(base + kTnBlock, kTnLbo, kTnSbo);
      const uint64_t dbh0 = tc::smem_desc(base + 2 * kTnBlock, kTnLbo, kTnSbo);
      const uint64_t dbl0 = tc::smem_desc(base + 3 * kTnBlock, kTnLbo, kTnSbo);
      const int which = warp_u - 2 * kTnProd;
      const uint32_t acc = tmem_u + uint32_t(which) * kTnT;
      int s = which;
      uint32_t ph = 0;
      for (int c = which; c < nchunks; c += 2) {
        tc::mbar_wait(&s_full[s], ph);
        tc::fence_after_sync();
        const uint64_t soff = uint64_t(uint32_t(s) * (kTnStage >> 4));
#pragma unroll
        for (int ks = 0; ks < kTnKC / 8; ++ks) {
          const uint64_t o = soff + uint64_t(ks * ((2 * kTnLbo) >> 4));   // one 8-wide K slice = 2 core-matrix columns
          tc::mma_tf32(acc, dal0 + o, dbh0 + o, idesc, !(c < 2 && ks == 0));   // small terms first
          tc::mma_tf32(acc, dah0 + o, dbl0 + o, idesc, true);
          tc::mma_tf32(acc, dah0 + o, dbh0 + o, idesc, true);
        }
        tc::mma_commit(&s_empty[s]);
        s += 2;
        if (s >= kTnProd) { s -= kTnProd; ph ^= 1; }
      }
      tc::mma_commit(&s_done);
    }
    __syncwarp();
  }

  if (nchunks > 0) {
    // ===== epilogue: warp w reads TMEM lanes 32(w%4).. (rows m) and the column quarter w/4 (32 columns n) =====
    tc::mbar_wait(&s_done, 0);
    tc::fence_after_sync();
    const int quad = warp & 3;
    const int m = m0 + quad * 32 + lane;
    const int nacc = nchunks < 2 ? 1 : 2;
    {
      const int cb = (warp >> 2) * 32;
      uint32_t raw[2][32];
      tc::tmem_ld32_nowait(tmem + (uint32_t(quad * 32) << 16) + cb, raw[0]);
      if (nacc == 2) tc::tmem_ld32_nowait(tmem + (uint32_t(quad * 32) << 16) + kTnT + cb, raw[1]);
      tc::tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[0][j]) + (nacc == 2 ? __uint_as_float(raw[1][j]) : 0.0f);
      if (m < g.M) {
        if (!g.transpose_c) {
          float* row = g.C + int64_t(m) * g.ldc + n0 + cb;
          const bool vec = (g.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(g.C) & 15) == 0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const int n = n0 + cb + j;
            if (vec && n + 3 < g.N) {
              red_add_v4(row + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int t = 0; t < 4; ++t)
                if (n + t < g.N) atomicAdd(row + j + t, v[j + t]);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = n0 + cb + j;
            if (n < g.N) atomicAdd(g.C + int64_t(n) * g.ldc + m, v[j]);   // lanes = consecutive m: coalesced
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

}  // namespace

// C (+)= A^T B with A [K, M] (lda), B [K, N] (ldb) row-major fp32; C [M, N] (ldc) or, transpose_c != 0, C^T [N, M].
// accumulate == 0: C is zeroed first (the K splits meet in it through atomic adds either way).
extern "C" int mapf_tc_gemm_tn(int32_t M, int32_t N, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb,
                               float* C, int64_t ldc, int32_t transpose_c, int32_t accumulate, mapf_stream_t stream) {
  MAPF_REQUIRE(A && B && C, MAPF_ERR_NULL);
  MAPF_REQUIRE(M > 0 && N > 0 && K > 0 && lda >= M && ldb >= N, MAPF_ERR_SHAPE);
  MAPF_REQUIRE(ldc >= (transpose_c ? M : N), MAPF_ERR_SHAPE);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!accumulate) {
    const size_t bytes = sizeof(float) * size_t(transpose_c ? N : M) * size_t(ldc);
    if (cudaMemsetAsync(C, 0, bytes, s) != cudaSuccess) return MAPF_ERR_CUDA;
  }
  TnArgs g{};
  g.M = M; g.N = N; g.K = K;
  g.A = A; g.lda = lda; g.B = B; g.ldb = ldb;
  g.C = C; g.ldc = ldc; g.transpose_c = transpose_c ? 1 : 0;
  const int mt = (M + kTnT - 1) / kTnT, nt = (N + kTnT - 1) / kTnT;
  const int64_t nchunks = (K + kTnKC - 1) / kTnKC;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // one CTA per SM (216 KB of stages + 256 TMEM columns): split K so that the grid covers the machine about once, but keep
  // at least 8 chunks per split so that the pipeline fills
  int64_t splits = sms / (mt * nt);               // one wave: a CTA's fixed cost (~8 us) is paid once
  if (splits < 1) splits = 1;
  if (splits > (nchunks + 7) / 8) splits = (nchunks + 7) / 8;
  g.chunks_per_split = int((nchunks + splits - 1) / splits);
  splits = (nchunks + g.chunks_per_split - 1) / g.chunks_per_split;
  const size_t smem = size_t(kTnProd) * kTnStage;
  static int cache[16] = {0};
  if (!mapf_ensure_smem(tc_gemm_tn_kernel, smem, cache)) return MAPF_ERR_CUDA;
  mapf_launch_pdl(tc_gemm_tn_kernel, dim3(mt, nt, unsigned(splits)), dim3(kTnThreads), smem, s, g);
  return mapf_launch_status();
}
