// fp32-accurate tensor-core GEMM (3xTF32 on tcgen05, accumulators in TMEM):
//   D[M,N] = A[M,K] . B[N,K]^T        A, B row-major with K contiguous ("K-major"); B may come in two K segments
// followed by a fused epilogue:
//   EPI_STORE : C = act(D + bias)                                   (encoder FC, compressMLP framework_gnn.py:93)
//   EPI_HEAD  : y = relu(D); logits = y Wa^T + ba; first-max argmax (graph-filter mix gnn.py:117-118 + ReLU +
//               actionMLP framework_gnn.py:171-181 + np.argmax evaluate.py:238)
//
// One 128-thread CTA per 128-row tile; N <= 128 columns x NACC accumulators live in TMEM (one fp32 column each).
// K is consumed in chunks of KC floats through a 2-stage shared-memory ring: all threads stage the chunk of A and B
// in the canonical no-swizzle K-major UMMA layout, split into TF32 hi / lo parts (round-to-nearest); one elected
// thread issues 3 tcgen05.mma (128 x N x 8) per 8-wide K slice and commits them to the stage's mbarrier, which
// frees the stage.  The tensor core's fp32 accumulation loses a little precision per accumulate, so consecutive
// chunks rotate over NACC independent TMEM accumulators that are summed in registers in the epilogue.
// Each thread then owns one output row (its TMEM lane), so the per-row head needs no cross-thread reduction.
#include "common.cuh"
#include "tc_gemm.cuh"

namespace {

constexpr int kTcThreads = 128;
constexpr int kTcM = 128;
constexpr int kTcStages = 2;

enum { EPI_STORE = 0, EPI_HEAD = 1 };

struct TcGemmArgs {
  int M, N;
  const float* A; int64_t lda;
  // B segments: columns [0,K0) of A meet B0[N,K0], columns [K0,K0+K1) meet B1[N,K1]
  const float* B0; int64_t ldb0; int K0;
  const float* B1; int64_t ldb1; int K1;
  // EPI_STORE
  float* C; int64_t ldc;
  const float* bias;
  int relu;
  // EPI_HEAD
  const float* act_w;  // [A, N]
  const float* act_b;  // [A]
  int num_actions;
  float* logits;       // [M, A]
  int32_t* actions;    // [M] or NULL
  float* y;            // optional [M, N] post-ReLU features (training) or NULL
};

// stage a [rows x KC] chunk (row-major source, leading dimension ld) as hi / lo canonical blocks
template <int KC>
__device__ __forceinline__ void stage_chunk(const float* __restrict__ src, int64_t ld, int row0, int rows_total,
                                            int rows_tile, int k0, int kend, uint8_t* hi, uint8_t* lo, int tid) {
  // one float4 (= one core-matrix row) per thread and iteration; consecutive lanes walk the 8 rows of a core
  // matrix so that the 16-byte shared stores of a quarter-warp fill 128 contiguous bytes (conflict-free)
  constexpr int KQ = KC / 4;
  const int n4 = rows_tile * KQ;
  for (int idx = tid; idx < n4; idx += kTcThreads) {
    const int r8 = idx & 7, k4 = (idx >> 3) % KQ, rg = idx / (8 * KQ);
    const int row = rg * 8 + r8;
    const int grow = row0 + row, gk = k0 + k4 * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (grow < rows_total) {
      const float* p = src + int64_t(grow) * ld;
      if (gk + 3 < kend) {
        v = __ldg(reinterpret_cast<const float4*>(p + gk));
      } else {
        if (gk + 0 < kend) v.x = __ldg(p + gk + 0);
        if (gk + 1 < kend) v.y = __ldg(p + gk + 1);
        if (gk + 2 < kend) v.z = __ldg(p + gk + 2);
      }
    }
    float4 h, l;
    tc::split4(v, h, l);
    const uint32_t off = tc::canon_off(row, k4 * 4, KC);
    *reinterpret_cast<float4*>(hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = l;
  }
}

template <int NT, int KC, int NACC, int EPI>
__global__ void __launch_bounds__(kTcThreads) tc_gemm_kernel(TcGemmArgs g) {
  constexpr uint32_t kColsNeeded = NT * NACC;
  constexpr uint32_t kCols = kColsNeeded <= 32 ? 32 : kColsNeeded <= 64 ? 64 : kColsNeeded <= 128 ? 128
                             : kColsNeeded <= 256 ? 256 : 512;
  constexpr int kABytes = kTcM * KC * 4, kBBytes = NT * KC * 4;
  constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
  extern __shared__ __align__(128) uint8_t s_tc[];
  __shared__ __align__(8) uint64_t s_bar[kTcStages];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(16) float s_head[EPI == EPI_HEAD ? kMaxActions * NT + kMaxActions : 4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * kTcM, n0 = blockIdx.y * NT;

  if (warp == 0) tc::tmem_alloc(&s_tmem, kCols);
  if (tid == 0) {
    for (int s = 0; s < kTcStages; ++s) tc::mbar_init(&s_bar[s], 1);
    tc::fence_mbar_init();
  }
  if (EPI == EPI_HEAD) {
    for (int i = tid; i < g.num_actions * NT; i += kTcThreads) {
      const int q = i / NT, n = i - q * NT;
      s_head[i] = n < g.N ? __ldg(g.act_w + int64_t(q) * g.N + n) : 0.0f;
    }
    if (tid < g.num_actions) s_head[kMaxActions * NT + tid] = __ldg(g.act_b + tid);
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem;
  constexpr uint32_t idesc = tc::idesc_tf32(kTcM, NT);

  const int nch0 = (g.K0 + KC - 1) / KC, nch1 = (g.K1 + KC - 1) / KC;
  const int nchunks = nch0 + nch1;
  uint32_t phase0 = 0, phase1 = 0;
  for (int c = 0; c < nchunks; ++c) {
    const int s = c & 1;
    uint8_t* a_hi = s_tc + s * kStageBytes;
    uint8_t* a_lo = a_hi + kABytes;
    uint8_t* b_hi = a_lo + kABytes;
    uint8_t* b_lo = b_hi + kBBytes;
    if (c >= kTcStages) {  // the MMAs that read this stage two chunks ago must have completed
      if (s == 0) { tc::mbar_wait(&s_bar[0], phase0); phase0 ^= 1; }
      else        { tc::mbar_wait(&s_bar[1], phase1); phase1 ^= 1; }
    }
    if (c < nch0) {
      stage_chunk<KC>(g.A, g.lda, m0, g.M, kTcM, c * KC, g.K0, a_hi, a_lo, tid);
      stage_chunk<KC>(g.B0, g.ldb0, n0, g.N, NT, c * KC, g.K0, b_hi, b_lo, tid);
    } else {
      const int c1 = c - nch0;
      stage_chunk<KC>(g.A + g.K0, g.lda, m0, g.M, kTcM, c1 * KC, g.K1, a_hi, a_lo, tid);
      stage_chunk<KC>(g.B1, g.ldb1, n0, g.N, NT, c1 * KC, g.K1, b_hi, b_lo, tid);
    }
    tc::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      constexpr uint32_t sbo = (KC / 4) * 128, lbo = 128;
      const uint32_t acc = tmem + uint32_t(c % NACC) * NT;
      const bool first = c < NACC;   // first touch of this accumulator overwrites
#pragma unroll
      for (int ks = 0; ks < KC / 8; ++ks) {
        const uint64_t dah = tc::smem_desc(tc::smem_u32(a_hi) + ks * 256, lbo, sbo);
        const uint64_t dal = tc::smem_desc(tc::smem_u32(a_lo) + ks * 256, lbo, sbo);
        const uint64_t dbh = tc::smem_desc(tc::smem_u32(b_hi) + ks * 256, lbo, sbo);
        const uint64_t dbl = tc::smem_desc(tc::smem_u32(b_lo) + ks * 256, lbo, sbo);
        tc::mma_tf32(acc, dal, dbh, idesc, !(first && ks == 0));   // small terms first
        tc::mma_tf32(acc, dah, dbl, idesc, true);
        tc::mma_tf32(acc, dah, dbh, idesc, true);
      }
      tc::mma_commit(&s_bar[s]);
    }
  }
  // drain: the last commit of every stage that was used
  if (nchunks > 0) tc::mbar_wait(&s_bar[0], phase0);
  if (nchunks > 1) tc::mbar_wait(&s_bar[1], phase1);
  tc::fence_after_sync();

  // epilogue: warp w owns TMEM lanes [32w, 32w+32) = rows m0 + 32w + lane
  const int row = m0 + warp * 32 + lane;
  const int nacc_used = nchunks < NACC ? nchunks : NACC;
  float part[kMaxActions];
#pragma unroll
  for (int q = 0; q < kMaxActions; ++q) part[q] = 0.0f;
#pragma unroll
  for (int cb = 0; cb < NT; cb += 32) {
    float v[32];
    tc::tmem_ld32(tmem + (uint32_t(warp * 32) << 16) + cb, v);
#pragma unroll
    for (int a = 1; a < NACC; ++a) {
      if (a < nacc_used) {   // warp-uniform
        float w[32];
        tc::tmem_ld32(tmem + (uint32_t(warp * 32) << 16) + a * NT + cb, w);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += w[j];
      }
    }
    if (EPI == EPI_STORE) {
      if (row < g.M) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = n0 + cb + j;
          if (n < g.N) {
            float o = v[j];
            if (g.bias) o += __ldg(g.bias + n);
            if (g.relu) o = fmaxf(o, 0.0f);
            v[j] = o;
          }
        }
        float* dst = g.C + int64_t(row) * g.ldc + n0 + cb;
        if (n0 + cb + 32 <= g.N && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + cb + j < g.N) dst[j] = v[j];
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
      if (g.y != nullptr && row < g.M) {
        float* dst = g.y + int64_t(row) * g.N + cb;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (cb + j < g.N) dst[j] = v[j];
      }
#pragma unroll
      for (int q = 0; q < kMaxActions; ++q) {
        if (q < g.num_actions) {
          const float4* wq = reinterpret_cast<const float4*>(s_head + q * NT + cb);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 w4 = wq[j >> 2];
            part[q] = fmaf(v[j], w4.x, part[q]);
            part[q] = fmaf(v[j + 1], w4.y, part[q]);
            part[q] = fmaf(v[j + 2], w4.z, part[q]);
            part[q] = fmaf(v[j + 3], w4.w, part[q]);
          }
        }
      }
    }
  }
  if (EPI == EPI_HEAD && row < g.M) {
    float best = 0.0f;
    int arg = 0;
#pragma unroll
    for (int q = 0; q < kMaxActions; ++q) {
      if (q < g.num_actions) {
        const float o = part[q] + s_head[kMaxActions * NT + q];
        g.logits[int64_t(row) * g.num_actions + q] = o;
        if (q == 0 || o > best) { best = o; arg = q; }   // first max wins (np.argmax)
      }
    }
    if (g.actions != nullptr) g.actions[row] = arg;
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, kCols);
}

template <int NT, int KC, int NACC, int EPI>
int launch_tc(const TcGemmArgs& g, cudaStream_t s) {
  constexpr size_t smem = size_t(kTcStages) * (2 * kTcM * KC * 4 + 2 * NT * KC * 4);
  auto kernel = tc_gemm_kernel<NT, KC, NACC, EPI>;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess)
    return MAPF_ERR_CUDA;
  dim3 grid((g.M + kTcM - 1) / kTcM, EPI == EPI_HEAD ? 1 : (g.N + NT - 1) / NT);
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
  TcGemmArgs g{};
  g.M = M; g.N = N;
  g.A = A; g.lda = lda;
  g.B0 = B; g.ldb0 = ldb; g.K0 = K;
  g.B1 = nullptr; g.ldb1 = 0; g.K1 = 0;
  g.C = C; g.ldc = ldc; g.bias = bias; g.relu = relu;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (N <= 64) return launch_tc<64, 32, 4, EPI_STORE>(g, s);
  return launch_tc<128, 16, 2, EPI_STORE>(g, s);
}

// logits / greedy actions = head(relu(Z [W | W2]^T)): Z [M, K0+K1] row-major, W [N,K0], W2 [N,K1] (K1 may be 0).
extern "C" int mapf_tc_mix_head(int32_t M, int32_t N, const float* Z, int64_t ldz, const float* W, int32_t K0,
                                const float* W2, int32_t K1, const float* act_w, const float* act_b, int32_t num_actions,
                                float* logits, int32_t* actions, float* y, mapf_stream_t stream) {
  MAPF_REQUIRE(Z && W && act_w && act_b && logits, MAPF_ERR_NULL);
  MAPF_REQUIRE(M > 0 && N > 0 && K0 > 0 && K1 >= 0 && (K1 == 0 || W2), MAPF_ERR_SHAPE);
  MAPF_REQUIRE(N <= 128 && (N & 15) == 0 && num_actions > 0 && num_actions <= kMaxActions, MAPF_ERR_UNSUPPORTED);
  MAPF_REQUIRE((ldz & 3) == 0 && (K0 & 3) == 0 && (K1 & 3) == 0 && mapf_aligned(Z, 16) && mapf_aligned(W, 16) &&
                   (K1 == 0 || mapf_aligned(W2, 16)),
               MAPF_ERR_ALIGN);
  TcGemmArgs g{};
  g.M = M; g.N = N;
  g.A = Z; g.lda = ldz;
  g.B0 = W; g.ldb0 = K0; g.K0 = K0;
  g.B1 = W2; g.ldb1 = K1; g.K1 = K1;
  g.act_w = act_w; g.act_b = act_b; g.num_actions = num_actions;
  g.logits = logits; g.actions = actions; g.y = y;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (N <= 64) return launch_tc<64, 32, 4, EPI_HEAD>(g, s);
  return launch_tc<128, 16, 2, EPI_HEAD>(g, s);
}
