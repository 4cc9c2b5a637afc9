// Filter taps of LARGE teams (N = 64, the scaling-sweep shape C5) on the tensor cores (A6: gnn.py:87-109):
//   S = D^-1/2 (A + I) D^-1/2,   x_k = x_{k-1} S,   Z[row] = [x_0 | x_1 | ... | x_{K-1}]          (row = b*N + n)
// At N = 64 the hops are 37 % of the policy's MACs (3 x 64^3 per episode) and the FFMA kernel (graph_taps_kernel) runs
// them at a quarter of the CUDA-core peak (issue-bound, 673 us for 8192 episodes).  Here a hop is a GEMM per tile of two
// episodes (128 rows):
//   D[(e, j), f] = sum_i  A[(e, j), (e', i)] * B[f, (e', i)],   A = blockdiag(S_0^T, S_1^T),  B = x_{k-1}^T
// 3xTF32 on tcgen05 (a = hi + lo; lo*hi + hi*lo + hi*hi), fp32 accumulator in tensor memory.  K runs over the 128
// agents of the tile in chunks of 8: the rows of the OTHER episode are zero in a chunk (half of the MMA is wasted; the
// tensor pipe is not the limiter).  Both operands are built in shared memory by the tile's 128 row threads from fp32
// copies of S and x_{k-1}: the A chunk is a lane-contiguous read of S, the B chunk a transposing gather of x (four agents
// of one feature -> one k-contiguous float4 of the canonical no-swizzle K-major layout).  After the 16 chunks of a hop
// the row threads read D (tcgen05.ld, thread = row), store it as the next x and write the tap to Z.
//   warps 0-3  row threads: operand chunks (3 stages), hop epilogue, Z write-out
//   warp  4    MMA issuer (elect.sync inside a warp-uniform branch): 3 MMAs 128 x 64 x 8 per chunk
// Persistent CTAs (two per SM), tiles strided over the grid.
#include "common.cuh"
#include "tc_gemm.cuh"

namespace {

constexpr int kThThreads = 160;
constexpr int kThM = 128;                       // rows per tile = 2 episodes of 64 agents
constexpr int kThN = 64;                        // agents per episode
constexpr int kThF = 64;                        // features (MMA N)
#ifndef TH_STAGES
#define TH_STAGES 3
#endif
constexpr int kThStages = TH_STAGES;
constexpr int kThABytes = kThM * 8 * 4;         // one (hi | lo) A block: 128 rows x 8 K
constexpr int kThBBytes = kThF * 8 * 4;         // one (hi | lo) B block: 64 rows x 8 K
constexpr int kThStageBytes = 2 * kThABytes + 2 * kThBBytes;   // 12 KB
constexpr int kThXPitch = kThF + 4;             // floats per row of the x copy (16-byte aligned, conflict-free gathers)

// mbarrier wait of this kernel: polling (TH_POLL) or the suspended try_wait of tc::mbar_wait
__device__ __forceinline__ void th_wait(uint64_t* bar, uint32_t parity) {
#ifdef TH_POLL
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(tc::smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return;
  }
  asm volatile("trap;");
#else
  tc::mbar_wait(bar, parity);
#endif
}

struct HopsArgs {
  int batch, taps, add_self_loop;
  const float* x;      // [B*64, 64] agent-major
  const float* gso;    // [B, 64, 64]
  float* z;            // [B*64, taps*64]
};

__global__ void __launch_bounds__(kThThreads, TH_STAGES <= 3 ? 2 : 1) graph_hops_tc_kernel(HopsArgs g) {
  extern __shared__ __align__(128) uint8_t s_th[];
  __shared__ __align__(8) uint64_t s_full[kThStages], s_empty[kThStages], s_hop;
  __shared__ uint32_t s_tmem;
  uint8_t* stages = s_th;                                                   // [3][12 KB]
  float* s_S = reinterpret_cast<float*>(s_th + kThStages * kThStageBytes);   // [2][64][64]  S_e[i][j]
  float* s_x = s_S + 2 * kThN * kThN;                                       // [128][68]    x_{k-1}
  float* s_dinv = s_x + kThM * kThXPitch;                                   // [128]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = g.taps, KT = K * kThF;
  const int ntiles = (g.batch + 1) >> 1;

  if (warp == 0) tc::tmem_alloc(&s_tmem, 64);
  if (tid == 0) {
    for (int s = 0; s < kThStages; ++s) { tc::mbar_init(&s_full[s], kThM); tc::mbar_init(&s_empty[s], 1); }
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
  auto rows_sync = [&]() { asm volatile("bar.sync 1, 128;" ::: "memory"); };   // the 128 row threads only

  uint32_t chunk_no = 0;     // chunks built / issued so far by this CTA (same sequence in both roles)
  uint32_t hop_no = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int e0 = tile * 2;
    const int nep = min(2, g.batch - e0);
    const int rows_used = nep * kThN;
    const int64_t row0 = int64_t(e0) * kThN;
    if (warp_u < 4) {
      // ===================== row threads =====================
      const int r = tid;                         // row of the tile: episode r / 64, agent r % 64
      const int er = r >> 6, j = r & 63;
      const bool live = r < rows_used;
      // ---- S of the tile's episodes (fp32, gnn.py:87-101: row degrees, IEEE rn div / sqrt like ATen's pow(-0.5)) ----
      // (eight 128-bit loads in flight per thread: the tile's 64 KB arrive in four round trips instead of 64)
      {
        const float4* src = reinterpret_cast<const float4*>(g.gso + int64_t(e0) * kThN * kThN);
        const int nq = nep * kThN * kThN / 4;
#pragma unroll 1
        for (int base = 0; base < nq; base += 8 * kThM) {
          float4 v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int idx = base + u * kThM + tid;
            v[u] = idx < nq ? mapf_ld_dep(src + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int idx = base + u * kThM + tid;
            if (idx < nq) {
              if (g.add_self_loop) {
                const int i = (idx >> 4) & 63, j0 = (idx & 15) * 4;     // row i of its episode, columns j0 .. j0+3
                if (i == j0) v[u].x += 1.0f;
                if (i == j0 + 1) v[u].y += 1.0f;
                if (i == j0 + 2) v[u].z += 1.0f;
                if (i == j0 + 3) v[u].w += 1.0f;
              }
              *reinterpret_cast<float4*>(s_S + idx * 4) = v[u];
            }
          }
        }
      }
      // ---- x_0: copy to shared memory and to tap 0 of Z ----
      {
        const float4* src = reinterpret_cast<const float4*>(g.x + row0 * kThF);
        const int nq = rows_used * (kThF / 4);
#pragma unroll 1
        for (int base = 0; base < nq; base += 8 * kThM) {
          float4 v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int idx = base + u * kThM + tid;
            v[u] = idx < nq ? mapf_ld_dep(src + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int idx = base + u * kThM + tid;
            if (idx < nq) {
              const int rr = idx >> 4, c4 = idx & 15;
              *reinterpret_cast<float4*>(s_x + rr * kThXPitch + c4 * 4) = v[u];
              *reinterpret_cast<float4*>(g.z + (row0 + rr) * KT + c4 * 4) = v[u];
            }
          }
        }
      }
      rows_sync();
      {
        float deg = 0.0f;
        if (live) {
          const float* srow = s_S + (er * kThN + j) * kThN;
          for (int q = 0; q < kThN; ++q) deg += srow[q];
        }
        s_dinv[r] = deg > 0.0f ? __fdiv_rn(1.0f, __fsqrt_rn(deg)) : 0.0f;
      }
      rows_sync();
      for (int idx = tid; idx < nep * kThN * kThN; idx += kThM) {
        const int e = idx >> 12, i = (idx >> 6) & 63, jj = idx & 63;
        s_S[idx] = (s_dinv[e * kThN + i] * s_S[idx]) * s_dinv[e * kThN + jj];
      }
      rows_sync();
      for (int k = 1; k < K; ++k) {
        // ---- 16 operand chunks: chunk c covers agents i0 .. i0+7 of episode ec ----
        for (int c = 0; c < 2 * (kThN / 8); ++c, ++chunk_no) {
          const int s = chunk_no % kThStages;
          const uint32_t use = chunk_no / kThStages;
          uint8_t* st = stages + s * kThStageBytes;
          const int ec = c >> 3, i0 = (c & 7) * 8;
          // A row r: S_ec[i0 + kk][j] for the rows of episode ec, zero for the other episode's rows
          float a[8];
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) a[kk] = 0.0f;
          if (live && er == ec && ec < nep) {
            const float* sp = s_S + (ec * kThN + i0) * kThN + j;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) a[kk] = sp[kk * kThN];
          }
          // B: this thread's (feature f, k-quad q): x_{k-1}[ec*64 + i0 + 4q + 0..3][f]
          const int f = r & 63, q = r >> 6;
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ec < nep) {
            const float* xp = s_x + (ec * kThN + i0 + 4 * q) * kThXPitch + f;
            b = make_float4(xp[0], xp[kThXPitch], xp[2 * kThXPitch], xp[3 * kThXPitch]);
          }
          float4 ah0, al0, ah1, al1, bh, bl;
          tc::split4(make_float4(a[0], a[1], a[2], a[3]), ah0, al0);
          tc::split4(make_float4(a[4], a[5], a[6], a[7]), ah1, al1);
          tc::split4(b, bh, bl);
          th_wait(&s_empty[s], (use & 1u) ^ 1u);
          const uint32_t aoff = tc::canon_off(r, 0, 8);
          *reinterpret_cast<float4*>(st + aoff) = ah0;
          *reinterpret_cast<float4*>(st + aoff + 128) = ah1;
          *reinterpret_cast<float4*>(st + kThABytes + aoff) = al0;
          *reinterpret_cast<float4*>(st + kThABytes + aoff + 128) = al1;
          const uint32_t boff = tc::canon_off(f, 4 * q, 8);
          *reinterpret_cast<float4*>(st + 2 * kThABytes + boff) = bh;
          *reinterpret_cast<float4*>(st + 2 * kThABytes + kThBBytes + boff) = bl;
          tc::fence_proxy_async();
          tc::mbar_arrive(&s_full[s]);
        }
        // ---- hop epilogue: D -> x_k (shared) and tap k of Z ----
        th_wait(&s_hop, hop_no & 1u);
        ++hop_no;
        tc::fence_after_sync();
        float* xrow = s_x + r * kThXPitch;
        float* zrow = g.z + (row0 + r) * KT + k * kThF;
#pragma unroll
        for (int cb = 0; cb < kThF; cb += 32) {
          float v[32];
          tc::tmem_ld32(tmem + (uint32_t(warp * 32) << 16) + cb, v);
#pragma unroll
          for (int t = 0; t < 32; t += 4) {
            const float4 o = make_float4(v[t], v[t + 1], v[t + 2], v[t + 3]);
            *reinterpret_cast<float4*>(xrow + cb + t) = o;
            if (live) *reinterpret_cast<float4*>(zrow + cb + t) = o;
          }
        }
        tc::fence_before_sync();
        rows_sync();      // x_k complete (and every read of D done) before the next hop's chunks / MMAs
      }
    } else {
      // ===================== MMA issuer =====================
      if (tc::elect_one()) {
        constexpr uint32_t idesc = tc::idesc_tf32(kThM, kThF);
        const uint32_t base = tc::smem_u32(stages);
        const uint64_t dah0 = tc::smem_desc(base, 128, 256);
        const uint64_t dal0 = tc::smem_desc(base + kThABytes, 128, 256);
        const uint64_t dbh0 = tc::smem_desc(base + 2 * kThABytes, 128, 256);
        const uint64_t dbl0 = tc::smem_desc(base + 2 * kThABytes + kThBBytes, 128, 256);
        for (int k = 1; k < K; ++k) {
          for (int c = 0; c < 2 * (kThN / 8); ++c, ++chunk_no) {
            const int s = chunk_no % kThStages;
            const uint32_t use = chunk_no / kThStages;
            th_wait(&s_full[s], use & 1u);
            tc::fence_after_sync();
            const uint64_t o = uint64_t(uint32_t(s) * (kThStageBytes >> 4));
            tc::mma_tf32(tmem_u, dal0 + o, dbh0 + o, idesc, c > 0);     // small terms first; the first chunk overwrites D
            tc::mma_tf32(tmem_u, dah0 + o, dbl0 + o, idesc, true);
            tc::mma_tf32(tmem_u, dah0 + o, dbh0 + o, idesc, true);
            tc::mma_commit(&s_empty[s]);
          }
          tc::mma_commit(&s_hop);
        }
      }
      __syncwarp();
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 64);
}

}  // namespace

// 1 if the tensor-core hops kernel takes these dimensions (else graph_taps_kernel)
int mapf_hops_tc_supported(const mapf_graph_filter_fwd_args* a) {
  return a->agents == kThN && a->in_dim == kThF && a->taps >= 1 && !a->has_skip;
}

int mapf_graph_hops_tc(const mapf_graph_filter_fwd_args* a, mapf_stream_t stream) {
  MAPF_REQUIRE(a && a->x && a->gso && a->z, MAPF_ERR_NULL);
  MAPF_REQUIRE(mapf_hops_tc_supported(a), MAPF_ERR_UNSUPPORTED);
  MAPF_REQUIRE(mapf_aligned(a->x, 16) && mapf_aligned(a->z, 16), MAPF_ERR_ALIGN);
  HopsArgs g{a->batch, a->taps, a->add_self_loop ? 1 : 0, a->x, a->gso, a->z};
  const size_t smem = size_t(kThStages) * kThStageBytes +
                      sizeof(float) * (size_t(2) * kThN * kThN + size_t(kThM) * kThXPitch + kThM);
  static int cache[16] = {0};
  if (!mapf_ensure_smem(graph_hops_tc_kernel, smem, cache)) return MAPF_ERR_CUDA;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ntiles = (a->batch + 1) / 2;
  const int per_sm = kThStages <= 3 ? 2 : 1;
  const int grid = ntiles < per_sm * sms ? ntiles : per_sm * sms;
  mapf_launch_pdl(graph_hops_tc_kernel, dim3(grid), dim3(kThThreads), smem, static_cast<cudaStream_t>(stream), g);
  return mapf_launch_status();
}
