// Filter taps of LARGE teams (N = 64, the scaling-sweep shape C5) on the tensor cores (A6: gnn.py:87-109):
//   S = D^-1/2 (A + I) D^-1/2,   x_k = x_{k-1} S,   Z[row] = [x_0 | x_1 | ... | x_{K-1}]          (row = b*N + n)
// At N = 64 the hops are 37 % of the policy's MACs (3 x 64^3 per episode) and the FFMA kernel (graph_taps_kernel) runs
// them at a quarter of the CUDA-core peak (issue-bound, 673 us for 8192 episodes).  Here a hop is ONE GEMM per episode,
//   D[j, f] = sum_i  A[j, i] * B[f, i],      A = S^T (M = 64 agents j, K = 64 agents i),   B = x_{k-1}^T (N = 64 features)
// 3xTF32 on tcgen05 (a = hi + lo; lo*hi + hi*lo + hi*hi), M = 64 MMAs, fp32 accumulator in tensor memory.  Both operands
// are RESIDENT in shared memory in the canonical no-swizzle K-major layout: A is built once per episode from the
// adjacency, B is written by the epilogue of the previous hop straight from the accumulator rows -- a transposing store
// (thread = agent i writes element (f, i) of every feature f) that is bank-conflict free because K-adjacent core matrices
// are 144 bytes apart (LBO = 144: lane i lands in bank i).  So a hop is 24 back-to-back MMAs behind one commit, one
// mbarrier round trip, and an epilogue (tcgen05.ld -> tap k of Z -> next B operand); no operand pipeline at all.
// One CTA (128 threads) per episode, 72 KB of shared memory and 64 TMEM columns: three CTAs per SM.
// With M = 64 the accumulator row r lives in TMEM lane (r % 16) + 32 * (r / 16): warp w, lanes 0-15, owns rows 16w..16w+15.
#include "common.cuh"
#include "tc_gemm.cuh"

namespace {

constexpr int kThThreads = 128;
constexpr int kThN = 64;                        // agents per episode (MMA M and K)
constexpr int kThF = 64;                        // features (MMA N)
constexpr int kThLbo = 144;                     // bytes between K-adjacent core matrices (8 rows x 16 bytes + 16 pad)
constexpr int kThSbo = (kThN / 4) * kThLbo;     // bytes between 8-row groups: 16 k-quads x 144 = 2304
constexpr int kThOpBytes = (kThN / 8) * kThSbo; // one operand, hi or lo: 8 row groups = 18432 bytes
#ifndef TH_LANE_MAP
#define TH_LANE_MAP 0                           // 0: row r in lane (r % 16) + 32 (r / 16);  1: row r in lane r (probe)
#endif

struct HopsArgs {
  int batch, taps, add_self_loop;
  const float* x;      // [B*64, 64] agent-major
  const float* gso;    // [B, 64, 64]
  float* z;            // [B*64, taps*64]
};

// byte offset of element (row, k) of a 64 x 64 K-major operand block: [(row/8)][(k/4)][row%8][k%4] with padded strides
__device__ __forceinline__ uint32_t th_off(int row, int k) {
  return uint32_t(row >> 3) * kThSbo + uint32_t(k >> 2) * kThLbo + uint32_t(row & 7) * 16 + uint32_t(k & 3) * 4;
}

__global__ void __launch_bounds__(kThThreads, 3) graph_hops_tc_kernel(HopsArgs g) {
  extern __shared__ __align__(128) uint8_t s_th[];
  __shared__ __align__(8) uint64_t s_hop;
  __shared__ uint32_t s_tmem;
  __shared__ float s_dinv[kThN];
  uint8_t* a_hi = s_th;                          // S^T, hi | lo
  uint8_t* a_lo = a_hi + kThOpBytes;
  uint8_t* b_hi = a_lo + kThOpBytes;             // x_{k-1}^T, hi | lo; before that: fp32 scratch for the adjacency
  uint8_t* b_lo = b_hi + kThOpBytes;
  float* s_S = reinterpret_cast<float*>(b_hi);   // [64][64] scratch (16 KB <= 2 x 18 KB)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = g.taps, KT = K * kThF;
  const int e = blockIdx.x;
  const int64_t row0 = int64_t(e) * kThN;

#ifndef TH_NACC
#define TH_NACC 1      // independent accumulators (K slices alternate); 2 and 4 measured no faster (340 / 377 us): the hop is not bound by its MMA chain
#endif
  constexpr uint32_t kCols = 64 * TH_NACC;
  if (warp == 0) tc::tmem_alloc(&s_tmem, kCols);
  if (tid == 0) {
    tc::mbar_init(&s_hop, 1);
    tc::fence_mbar_init();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
  mapf_pdl_wait();      // x comes from the encoder FC in front, gso from the env kernel
  mapf_pdl_launch();

  // ---- adjacency -> fp32 scratch (eight 128-bit loads in flight per thread) ----
  {
    const float4* src = reinterpret_cast<const float4*>(g.gso + int64_t(e) * kThN * kThN);
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = mapf_ld_dep(src + u * kThThreads + tid);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = u * kThThreads + tid;
      if (g.add_self_loop) {
        const int i = idx >> 4, j0 = (idx & 15) * 4;     // row i, columns j0 .. j0+3
        if (i == j0) v[u].x += 1.0f;
        if (i == j0 + 1) v[u].y += 1.0f;
        if (i == j0 + 2) v[u].z += 1.0f;
        if (i == j0 + 3) v[u].w += 1.0f;
      }
      *reinterpret_cast<float4*>(s_S + idx * 4) = v[u];
    }
  }
  // x_0 row of this thread's agent (threads 0-63: features 0-31, threads 64-127: features 32-63), in flight meanwhile
  const int ag = tid & 63, fh = tid >> 6;
  float xr[32];
  {
    const float4* src = reinterpret_cast<const float4*>(g.x + (row0 + ag) * kThF + fh * 32);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float4 v = mapf_ld_dep(src + u);
      xr[4 * u] = v.x; xr[4 * u + 1] = v.y; xr[4 * u + 2] = v.z; xr[4 * u + 3] = v.w;
    }
  }
  __syncthreads();
  if (tid < kThN) {      // row degrees (gnn.py:92-96: sums of 0 / 1 / 2, exact in any order); inf -> 0
    float deg = 0.0f;
    const float* srow = s_S + tid * kThN;
#pragma unroll 8
    for (int q = 0; q < kThN; ++q) deg += srow[(q + tid) & 63];      // (skewed start: conflict free)
    s_dinv[tid] = deg > 0.0f ? __fdiv_rn(1.0f, __fsqrt_rn(deg)) : 0.0f;   // IEEE rn like ATen's pow(-0.5)
  }
  __syncthreads();
  // ---- A = S^T as the resident K-major operand: element (row j, k = i) = (dinv_i * a_ij) * dinv_j ----
  // thread (j = tid % 64, half = tid / 64) writes the k-quads of its half: lane-contiguous reads of the scratch
  {
    const int j = tid & 63;
    const float dj = s_dinv[j];
#pragma unroll 2
    for (int q = (tid >> 6) * 8; q < (tid >> 6) * 8 + 8; ++q) {
      const int i0 = 4 * q;
      float4 v;
      v.x = (s_dinv[i0] * s_S[i0 * kThN + j]) * dj;
      v.y = (s_dinv[i0 + 1] * s_S[(i0 + 1) * kThN + j]) * dj;
      v.z = (s_dinv[i0 + 2] * s_S[(i0 + 2) * kThN + j]) * dj;
      v.w = (s_dinv[i0 + 3] * s_S[(i0 + 3) * kThN + j]) * dj;
      float4 h, l;
      tc::split4(v, h, l);
      const uint32_t off = th_off(j, i0);
      *reinterpret_cast<float4*>(a_hi + off) = h;
      *reinterpret_cast<float4*>(a_lo + off) = l;
    }
  }
  __syncthreads();      // the scratch is free: the B operand may overwrite it

  // one half row (32 features of one agent): tap k of Z and, unless it is the last tap, elements (f, agent) of the next
  // hop's B operand
  auto emit_half = [&](const float (&v)[32], int agent, int f0, int k) {
    float4* zp = reinterpret_cast<float4*>(g.z + (row0 + agent) * KT + k * kThF + f0);
#pragma unroll
    for (int u = 0; u < 8; ++u) zp[u] = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
    if (k + 1 < K) {
#pragma unroll
      for (int t = 0; t < 32; ++t) {
        const float hi = tc::tf32_rn(v[t]);
        const float lo = tc::tf32_rn(v[t] - hi);
        const uint32_t off = th_off(f0 + t, agent);
        *reinterpret_cast<float*>(b_hi + off) = hi;
        *reinterpret_cast<float*>(b_lo + off) = lo;
      }
    }
  };
  emit_half(xr, ag, fh * 32, 0);

  constexpr uint32_t idesc = tc::idesc_tf32(kThN, kThF);
  const uint32_t base = tc::smem_u32(s_th);
  // accumulator row of this thread: with M = 64 the rows sit in 16 lanes of every 32-lane TMEM quadrant
#if TH_LANE_MAP == 0
  const bool row_owner = lane < 16;
  const int my_row = warp * 16 + lane;
#else
  const bool row_owner = warp < 2;
  const int my_row = tid & 63;
#endif
  for (int k = 1; k < K; ++k) {
    tc::fence_proxy_async();      // operand stores (generic proxy) -> visible to the tensor core
    tc::fence_before_sync();
    __syncthreads();
    if (warp_u == 0) {
      tc::fence_after_sync();
      if (tc::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < kThN / 8; ++ks) {
          const uint32_t o = uint32_t(ks) * 2 * kThLbo;     // one 8-wide K slice = 2 core-matrix columns
          const uint64_t dah = tc::smem_desc(base + o, kThLbo, kThSbo);
          const uint64_t dal = tc::smem_desc(base + kThOpBytes + o, kThLbo, kThSbo);
          const uint64_t dbh = tc::smem_desc(base + 2 * kThOpBytes + o, kThLbo, kThSbo);
          const uint64_t dbl = tc::smem_desc(base + 3 * kThOpBytes + o, kThLbo, kThSbo);
          const uint32_t acc = tmem_u + uint32_t(ks % TH_NACC) * 64;
          tc::mma_tf32(acc, dal, dbh, idesc, ks >= TH_NACC);     // small terms first; an accumulator's first MMA overwrites
          tc::mma_tf32(acc, dah, dbl, idesc, true);
          tc::mma_tf32(acc, dah, dbh, idesc, true);
        }
        tc::mma_commit(&s_hop);
      }
      __syncwarp();
    }
    tc::mbar_wait(&s_hop, uint32_t(k - 1) & 1u);
    tc::fence_after_sync();
    // ---- epilogue: every warp loads its TMEM quadrant (warp-collective), the row owners emit their rows ----
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float v[32];
      tc::tmem_ld32(tmem + (uint32_t(warp * 32) << 16) + half * 32, v);
#pragma unroll
      for (int a2 = 1; a2 < TH_NACC; ++a2) {
        float w[32];
        tc::tmem_ld32(tmem + (uint32_t(warp * 32) << 16) + a2 * 64 + half * 32, w);
#pragma unroll
        for (int t = 0; t < 32; ++t) v[t] += w[t];
      }
      if (row_owner) emit_half(v, my_row, half * 32, k);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, kCols);
}

}  // namespace

// 1 if the tensor-core hops kernel takes these dimensions (else graph_taps_kernel)
int mapf_hops_tc_supported(const mapf_graph_filter_fwd_args* a) {
  return a->agents == kThN && a->in_dim == kThF && a->taps >= 1 && !a->has_skip;
}

int mapf_graph_hops_tc(const mapf_graph_filter_fwd_args* a, mapf_stream_t stream) {
  MAPF_REQUIRE(a && a->x && a->gso && a->z, MAPF_ERR_NULL);
  MAPF_REQUIRE(mapf_hops_tc_supported(a), MAPF_ERR_UNSUPPORTED);
  MAPF_REQUIRE(mapf_aligned(a->x, 16) && mapf_aligned(a->z, 16) && mapf_aligned(a->gso, 16), MAPF_ERR_ALIGN);
  HopsArgs g{a->batch, a->taps, a->add_self_loop ? 1 : 0, a->x, a->gso, a->z};
  const size_t smem = size_t(4) * kThOpBytes;
  static int cache[16] = {0};
  if (!mapf_ensure_smem(graph_hops_tc_kernel, smem, cache)) return MAPF_ERR_CUDA;
  mapf_launch_pdl(graph_hops_tc_kernel, dim3(a->batch), dim3(kThThreads), smem, static_cast<cudaStream_t>(stream), g);
  return mapf_launch_status();
}
