// Fused graph filter on the tensor cores (A6/A7/A8): per 128-row tile of whole episodes, ONE kernel builds the graph
// shift operator, runs the hops, feeds every filter tap straight into tcgen05 MMAs and finishes with ReLU + action head +
// argmax.  The taps never touch HBM (the unfused path writes and re-reads Z [rows, K*F]).
//
//   S = D^-1/2 (A [+ I]) D^-1/2, x_k = x_{k-1} S                         gnn.py:87-109 / :185-199
//   logits = actionMLP(relu([x_0 .. x_{K-1} (, Xr)] [W | W2]^T))          gnn.py:111-118,201-227; framework_gnn.py:171-181
//
// The hops act on the node axis independently per feature, so the K taps of an 8-feature slice can be produced from
// that slice alone.  Warp roles (256 threads):
//   warps 0-3  producers : warp p owns shared-memory stage p and the feature slices j = p, p+4, ...  For each slice it
//                          loads x[:, 8j..8j+8) (128 rows, 4 per lane), and for k = 0..K-1 emits the chunk (tap k,
//                          slice j) -- TF32 hi/lo split, canonical UMMA layout, B chunk by TMA bulk copy -- then
//                          advances the slice one hop through a warp-private shared-memory exchange (a lane needs the
//                          rows of its episode that other lanes hold).  The message-passing skip chunk is gathered from
//                          global memory (raw [B,F,N]->[B,N,F] reshape).
//   warps 4,5  MMA issuers (stages {0,2} / {1,3}, one TMEM accumulator each): wait `full`, 3 x tcgen05.mma 128xGx8
//                          (lo*hi, hi*lo, hi*hi), commit to `empty`.
//   all warps  epilogue  : tcgen05.ld, ReLU, 5-way head, exchange of the two column halves, first-max argmax.
#include "common.cuh"
#include "env_step.cuh"
#include "tc_gemm.cuh"

namespace {

constexpr int kTgThreads = 256;
constexpr int kTgProd = 4;
constexpr int kTgM = 128;
constexpr int kTgNT = 128;
constexpr int kTgABytes = kTgM * 8 * 4, kTgBBytes = kTgNT * 8 * 4;
constexpr int kTgStageBytes = 2 * kTgABytes + 2 * kTgBBytes;   // 16 KB

// x / n for 0 <= x < 2^20 and small n, with rcp = 1.0f / n: no integer-division sequence (dozens of instructions each; this
// kernel runs its code once per CTA from a cold instruction cache, so code size is latency).  Exact: x + 0.5 is at least
// 0.5 away from every multiple of n, far more than the rounding error of the product.
__device__ __forceinline__ int div_small(int x, float rcp) { return int((float(x) + 0.5f) * rcp); }

__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}

struct TcGraphArgs {
  int batch, agents, F, K, G;
  int has_skip, add_self_loop, num_actions;
  const float* x;       // [B*N, F] agent-major
  const float* gso;     // [B, N, N]
  const float* Bp;      // packed [W | W2] (mapf_tc_pack_b, NT = 128, KC = 8)
  const float* act_w;   // [A, G]
  const float* act_b;   // [A]
  float* logits;        // [B*N, A]
  int32_t* actions;     // [B*N] or NULL
  int gso_is_old;       // 1: gso was complete before the kernel IN FRONT of this one started (read it before the grid wait)
  mapf_env_step_args env;   // ENVL > 0: the environment step of the tile's episodes runs in the tail of this kernel
};

// ENVL = 0: graph filter + head only.  ENVL = 8 / 16 / 32 (mapf_rollout): LANES of the fused environment step -- the CTA
// that chose the actions of its episodes also moves them and writes their next observation (env_step.cuh), so a rollout
// step is three kernels; the boards / positions are staged in the prologue, ahead of the grid wait (they were last written
// two kernels back).
#ifndef TG_MINB
#define TG_MINB 2
#endif
template <int ENVL>
__global__ void __launch_bounds__(kTgThreads, TG_MINB) tc_graph_kernel(TcGraphArgs g, int epc) {
  extern __shared__ __align__(128) uint8_t s_tg[];
  __shared__ __align__(8) uint64_t s_full[kTgProd], s_empty[kTgProd], s_done;
  __shared__ uint32_t s_tmem;
  __shared__ __align__(16) float s_head[kMaxActions * kTgNT + kMaxActions];
  const int N = g.agents, F = g.F, K = g.K;
  uint8_t* stages = s_tg;                                                     // [4][16 KB]
  float* scratch = reinterpret_cast<float*>(s_tg + kTgProd * kTgStageBytes);  // [4 warps][128][8]
  float* sm = scratch + kTgProd * kTgM * 8;                                   // [epc][N][N]
  float* dinv = sm + ((epc * N * N + 3) & ~3);                                // [128]
  // fused env step: [128] actions | per-thread arrays | boards of the tile's episodes
  constexpr int kEL = ENVL > 0 ? ENVL : 8;
  int32_t* s_act = reinterpret_cast<int32_t*>(dinv + kTgM);
  EnvShared<kEL, kTgThreads>& s_env = *reinterpret_cast<EnvShared<kEL, kTgThreads>*>(s_act + kTgM);
  uint8_t* s_boards = reinterpret_cast<uint8_t*>(&s_env) + ((sizeof(EnvShared<kEL, kTgThreads>) + 15) & ~size_t(15));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int e0 = blockIdx.x * epc;
  const int nep = min(epc, g.batch - e0);
  const int rows_used = nep * N;
  const int64_t row0 = int64_t(e0) * N;

  if (warp == 0) tc::tmem_alloc(&s_tmem, 256);
  if (tid == 0) {
    for (int s = 0; s < kTgProd; ++s) { tc::mbar_init(&s_full[s], 32); tc::mbar_init(&s_empty[s], 1); }
    tc::mbar_init(&s_done, 2);
    tc::fence_mbar_init();
  }
  for (int i = tid; i < g.num_actions * kTgNT; i += kTgThreads) {
    const int q = i / kTgNT, n = i - q * kTgNT;
    s_head[i] = n < g.G ? __ldg(g.act_w + int64_t(q) * g.G + n) : 0.0f;
  }
  if (tid < g.num_actions) s_head[kMaxActions * kTgNT + tid] = __ldg(g.act_b + tid);
  __syncthreads();      // the mbarriers are initialised
  if (warp < kTgProd && lane == 0) {
    // the B chunk of every producer's first emit (tap 0, slice = warp) is packed weights: fetch it now, under the shift
    // operator build (and, inside mapf_policy_fwd, under the encoder FC that runs in front of this kernel)
    tc::mbar_expect_tx(&s_full[warp], 2 * kTgBBytes);
    tc::bulk_g2s(s_tg + warp * kTgStageBytes + 2 * kTgABytes, g.Bp + int64_t(warp) * (2 * kTgNT * 8), 2 * kTgBBytes,
                 &s_full[warp]);
  }
  // Inside mapf_policy_fwd the adjacency is an input of the whole launch sequence -- complete before the encoder kernels in
  // front of this one started -- so the shift operator is built while the encoder FC still runs; the node features come
  // from that FC and are read behind the grid wait.
  EnvAgent env_ag{};
  if (ENVL > 0) env_ag = env_step_stage<kEL, kTgThreads>(g.env, e0, nep, tid, s_boards, s_env);
  if (!g.gso_is_old) { mapf_pdl_wait(); mapf_pdl_launch(); }
  // ---- graph shift operator of the tile's episodes ----
  const float rcp_n = 1.0f / float(N), rcp_nn = 1.0f / float(N * N);
  for (int idx = tid; idx < nep * N * N; idx += kTgThreads) {
    const int el = div_small(idx, rcp_nn), rem = idx - el * N * N;
    const int i = div_small(rem, rcp_n), j = rem - i * N;
    float v = mapf_ld_dep(g.gso + int64_t(e0 + el) * N * N + rem);
    if (g.add_self_loop && i == j) v += 1.0f;
    sm[idx] = v;
  }
  __syncthreads();
  if (tid < rows_used) {
    const int el = div_small(tid, rcp_n), i = tid - el * N;
    float deg = 0.0f;
    for (int j = 0; j < N; ++j) deg += sm[(el * N + i) * N + j];
    dinv[tid] = deg > 0.0f ? __fdiv_rn(1.0f, __fsqrt_rn(deg)) : 0.0f;
  }
  __syncthreads();
  for (int idx = tid; idx < nep * N * N; idx += kTgThreads) {
    const int el = div_small(idx, rcp_nn), rem = idx - el * N * N;
    const int i = div_small(rem, rcp_n), j = rem - i * N;
    sm[idx] = (dinv[el * N + i] * sm[idx]) * dinv[el * N + j];
  }
  if (g.gso_is_old) { mapf_pdl_wait(); mapf_pdl_launch(); }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem;
  // warp-uniform copies for the MMA issuers (elect.sync inside a warp-uniform branch: plain UTCHMMAs, see tc_gemm.cu)
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
  const int slices = F >> 3;                       // 8-feature slices (host guarantees slices % 4 == 0)
  const int per_warp = slices / kTgProd;
  const int chunks_per_slice = K + g.has_skip;

  if (warp_u < kTgProd) {
    // ===================== producers =====================
    const int p = warp;
    uint8_t* a_hi = stages + p * kTgStageBytes;
    uint8_t* a_lo = a_hi + kTgABytes;
    uint8_t* b_hi = a_lo + kTgABytes;
    float* xs = scratch + p * kTgM * 8;
    uint32_t ne = 0;
    // per own row: episode-local bookkeeping
    int rr[4], rep[4], rjl[4];
    bool rlive[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      rr[i] = lane + 32 * i;
      rlive[i] = rr[i] < rows_used;
      rep[i] = rlive[i] ? div_small(rr[i], rcp_n) : 0;
      rjl[i] = rr[i] - rep[i] * N;
    }
    auto emit = [&](int c, const float (*v)[8]) {
      tc::mbar_wait(&s_empty[p], (ne & 1) ^ 1);
      ++ne;
      if (lane == 0 && ne > 1) {          // (the first emit's B chunk was fetched in the prologue)
        tc::mbar_expect_tx(&s_full[p], 2 * kTgBBytes);
        tc::bulk_g2s(b_hi, g.Bp + int64_t(c) * (2 * kTgNT * 8), 2 * kTgBBytes, &s_full[p]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 h0, l0, h1, l1;
        tc::split4(make_float4(v[i][0], v[i][1], v[i][2], v[i][3]), h0, l0);
        tc::split4(make_float4(v[i][4], v[i][5], v[i][6], v[i][7]), h1, l1);
        const uint32_t off = tc::canon_off(rr[i], 0, 8);
        *reinterpret_cast<float4*>(a_hi + off) = h0;
        *reinterpret_cast<float4*>(a_hi + off + 128) = h1;
        *reinterpret_cast<float4*>(a_lo + off) = l0;
        *reinterpret_cast<float4*>(a_lo + off + 128) = l1;
      }
      tc::fence_proxy_async();
      tc::mbar_arrive(&s_full[p]);
    };
    // x[:, 8j..8j+8) of this lane's four rows; the next slice is loaded while the current one is emitted
    auto load_slice = [&](int j, float4 (&lo4)[4], float4 (&hi4)[4]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        lo4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        hi4[i] = lo4[i];
        if (rlive[i] && j < slices) {
          const float4* src = reinterpret_cast<const float4*>(g.x + (row0 + rr[i]) * F + 8 * j);
          lo4[i] = mapf_ld_dep(src);
          hi4[i] = mapf_ld_dep(src + 1);
        }
      }
    };
    float4 nlo[4], nhi[4];
    load_slice(p, nlo, nhi);
    // The chunk loops are deliberately NOT unrolled: a CTA runs this code once, from a cold instruction cache, and the
    // fully unrolled kernel (111 KB of SASS) spent most of its time in instruction fetch (ncu: top stall no_instruction).
#pragma unroll 1
    for (int si = 0; si < per_warp; ++si) {
      const int j = p + kTgProd * si;
      float xv[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        xv[i][0] = nlo[i].x; xv[i][1] = nlo[i].y; xv[i][2] = nlo[i].z; xv[i][3] = nlo[i].w;
        xv[i][4] = nhi[i].x; xv[i][5] = nhi[i].y; xv[i][6] = nhi[i].z; xv[i][7] = nhi[i].w;
      }
      load_slice(si + 1 < per_warp ? j + kTgProd : slices, nlo, nhi);
#ifndef TG_CHUNK_UNROLL
#pragma unroll 1
#endif
      for (int c = 0; c < chunks_per_slice; ++c) {
        if (c > 0 && c < K) {
          // x_k[row j] = sum_i S[i][j] x_{k-1}[row i] over the rows of the same episode (held by other lanes)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            *reinterpret_cast<float4*>(xs + rr[i] * 8) = make_float4(xv[i][0], xv[i][1], xv[i][2], xv[i][3]);
            *reinterpret_cast<float4*>(xs + rr[i] * 8 + 4) = make_float4(xv[i][4], xv[i][5], xv[i][6], xv[i][7]);
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float acc[8];
#pragma unroll
            for (int f = 0; f < 8; ++f) acc[f] = 0.0f;
            if (rlive[i]) {
              const float* scol = sm + rep[i] * N * N + rjl[i];
              const float* prev = xs + rep[i] * N * 8;
#ifndef TG_HOP_UNROLL
#pragma unroll 1
#endif
              for (int n = 0; n < N; ++n) {
                const float s = scol[n * N];
                const float4 p0 = *reinterpret_cast<const float4*>(prev + n * 8);
                const float4 p1 = *reinterpret_cast<const float4*>(prev + n * 8 + 4);
                acc[0] = fmaf(s, p0.x, acc[0]); acc[1] = fmaf(s, p0.y, acc[1]);
                acc[2] = fmaf(s, p0.z, acc[2]); acc[3] = fmaf(s, p0.w, acc[3]);
                acc[4] = fmaf(s, p1.x, acc[4]); acc[5] = fmaf(s, p1.y, acc[5]);
                acc[6] = fmaf(s, p1.z, acc[6]); acc[7] = fmaf(s, p1.w, acc[7]);
              }
            }
#pragma unroll
            for (int f = 0; f < 8; ++f) xv[i][f] = acc[f];
          }
          __syncwarp();   // everyone has read the exchange buffer before it is overwritten
        } else if (c == K) {
          // message-passing skip chunk: Xr[n'][f'] = X_flat[n'F + f'] with X_flat index of X[b,f,n] = f*N + n
          // (gnn.py:201-204); q = n'F + f' walks 8 consecutive flat indices: (q % N, q / N) advance incrementally
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int q0 = rjl[i] * F + 8 * j;
            int qd = div_small(q0, rcp_n);
            int qn = q0 - qd * N;
            const float* xe = g.x + (row0 + rep[i] * N) * F;
#pragma unroll
            for (int f = 0; f < 8; ++f) {
              xv[i][f] = rlive[i] ? mapf_ld_dep(xe + qn * F + qd) : 0.0f;
              if (++qn == N) { qn = 0; ++qd; }
            }
          }
        }
        emit(c * slices + j, xv);
      }
    }
  } else if (warp_u < kTgProd + 2) {
    // ===================== MMA issuers =====================
    if (tc::elect_one()) {
      const int which = warp_u - kTgProd;
      constexpr uint32_t idesc = tc::idesc_tf32(kTgM, kTgNT);
      constexpr uint32_t sbo = 256, lbo = 128;
      const uint32_t base = tc::smem_u32(stages);
      const uint64_t dah0 = tc::smem_desc(base, lbo, sbo);
      const uint64_t dal0 = tc::smem_desc(base + kTgABytes, lbo, sbo);
      const uint64_t dbh0 = tc::smem_desc(base + 2 * kTgABytes, lbo, sbo);
      const uint64_t dbl0 = tc::smem_desc(base + 2 * kTgABytes + kTgBBytes, lbo, sbo);
      const uint32_t acc = tmem_u + uint32_t(which) * kTgNT;
      const int T = per_warp * chunks_per_slice;
      bool first = true;
      for (int t = 0; t < T; ++t) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int s = which + 2 * h;
          tc::mbar_wait(&s_full[s], uint32_t(t & 1));
          tc::fence_after_sync();
          const uint64_t o = uint64_t(uint32_t(s) * (kTgStageBytes >> 4));
          tc::mma_tf32(acc, dal0 + o, dbh0 + o, idesc, !first);
          tc::mma_tf32(acc, dah0 + o, dbl0 + o, idesc, true);
          tc::mma_tf32(acc, dah0 + o, dbh0 + o, idesc, true);
          first = false;
          tc::mma_commit(&s_empty[s]);
        }
      }
      tc::mma_commit(&s_done);
    }
    __syncwarp();
  }

  {
    // ===================== epilogue (all 8 warps) =====================
    tc::mbar_wait(&s_done, 0);
    tc::fence_after_sync();
    const int quad = warp & 3, half = warp >> 2;
    const int row = quad * 32 + lane;
    float part[kMaxActions];
#pragma unroll
    for (int q = 0; q < kMaxActions; ++q) part[q] = 0.0f;
#ifndef TG_EPI_COLS
#define TG_EPI_COLS 32
#endif
#if TG_EPI_COLS == 8
    // eight columns per (rolled) iteration: small code, same accumulation order as ever (column ascending per action)
#pragma unroll 1
    for (int cbl = 0; cbl < kTgNT / 2; cbl += 8) {
      const int cb = half * (kTgNT / 2) + cbl;
      uint32_t r0[8], r1[8];
      tmem_ld8_nowait(tmem + (uint32_t(quad * 32) << 16) + cb, r0);
      tmem_ld8_nowait(tmem + (uint32_t(quad * 32) << 16) + kTgNT + cb, r1);
      tc::tmem_ld_wait();
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(__uint_as_float(r0[j]) + __uint_as_float(r1[j]), 0.0f);
#pragma unroll
      for (int q = 0; q < kMaxActions; ++q) {
        if (q < g.num_actions) {
          const float4* wq = reinterpret_cast<const float4*>(s_head + q * kTgNT + cb);
#pragma unroll
          for (int j = 0; j < 8; j += 4) {
            const float4 w4 = wq[j >> 2];
            part[q] = fmaf(v[j], w4.x, part[q]);
            part[q] = fmaf(v[j + 1], w4.y, part[q]);
            part[q] = fmaf(v[j + 2], w4.z, part[q]);
            part[q] = fmaf(v[j + 3], w4.w, part[q]);
          }
        }
      }
    }
#else
    // 32 columns per iteration (rolled: the two iterations share their code; column ascending per action as ever)
#pragma unroll 1
    for (int cbl = 0; cbl < kTgNT / 2; cbl += 32) {
      const int cb = half * (kTgNT / 2) + cbl;
      uint32_t r0[32], r1[32];
      tc::tmem_ld32_nowait(tmem + (uint32_t(quad * 32) << 16) + cb, r0);
      tc::tmem_ld32_nowait(tmem + (uint32_t(quad * 32) << 16) + kTgNT + cb, r1);
      tc::tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(__uint_as_float(r0[j]) + __uint_as_float(r1[j]), 0.0f);
#pragma unroll
      for (int q = 0; q < kMaxActions; ++q) {
        if (q < g.num_actions) {
          const float4* wq = reinterpret_cast<const float4*>(s_head + q * kTgNT + cb);
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
#endif
    float* ex = reinterpret_cast<float*>(stages);   // [128 rows][kMaxActions], the stages are idle now
    if (half == 1) {
#pragma unroll
      for (int q = 0; q < kMaxActions; ++q) ex[row * kMaxActions + q] = part[q];
    }
    __syncthreads();
    if (half == 0 && row < rows_used) {
      float best = 0.0f;
      int arg = 0;
#pragma unroll
      for (int q = 0; q < kMaxActions; ++q) {
        if (q < g.num_actions) {
          const float o = part[q] + ex[row * kMaxActions + q] + s_head[kMaxActions * kTgNT + q];
          g.logits[(row0 + row) * g.num_actions + q] = o;
          if (q == 0 || o > best) { best = o; arg = q; }   // first max wins (np.argmax)
        }
      }
      if (g.actions != nullptr) g.actions[row0 + row] = arg;
      if (ENVL > 0) s_act[row] = arg;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
  // the operand stages are idle now (every MMA has completed): they stage the observations on their way out
  if (ENVL > 0)
    env_step_finish<kEL, kTgThreads>(g.env, e0, nep, tid, s_boards, s_env, env_ag, s_act, reinterpret_cast<float*>(stages));
}

}  // namespace

// 1 if mapf_tc_graph_head supports these dimensions (else use mapf_graph_taps + mapf_tc_mix_head)
extern "C" int mapf_tc_graph_supported(int32_t agents, int32_t in_dim, int32_t out_dim, int32_t taps) {
  return agents >= 1 && agents <= 64 && in_dim >= 32 && (in_dim & 31) == 0 && out_dim > 64 && out_dim <= 128 &&
         (out_dim & 15) == 0 && taps >= 1;   // out_dim > 64: the packed B uses the 128-row tiling
}

// Whole graph-filter + head tail of Network.forward from agent-major x [B*N, F] and gso [B,N,N]; Wp = packed [W | W2].
int mapf_tc_graph_head_ex(int32_t batch, int32_t agents, int32_t in_dim, int32_t out_dim, int32_t taps,
                          int32_t add_self_loop, int32_t has_skip, const float* x, const float* gso, const float* Wp,
                          const float* act_w, const float* act_b, int32_t num_actions, float* logits, int32_t* actions,
                          int gso_is_old, const mapf_env_step_args* env, mapf_stream_t stream);

extern "C" int mapf_tc_graph_head(int32_t batch, int32_t agents, int32_t in_dim, int32_t out_dim, int32_t taps,
                                  int32_t add_self_loop, int32_t has_skip, const float* x, const float* gso,
                                  const float* Wp, const float* act_w, const float* act_b, int32_t num_actions,
                                  float* logits, int32_t* actions, mapf_stream_t stream) {
  return mapf_tc_graph_head_ex(batch, agents, in_dim, out_dim, taps, add_self_loop, has_skip, x, gso, Wp, act_w, act_b,
                               num_actions, logits, actions, 0, nullptr, stream);
}

// 1 if the environment step of `agents`-agent episodes on boards of `cell_stride` bytes can run in the tail of the graph
// kernel: an episode's LANES times the episodes of a 128-row tile must fit the CTA's 256 threads
int mapf_tc_graph_env_fusable(int32_t agents, int32_t cell_stride) {
  // measured: small teams gain (one launch less), at N = 20 the separate env kernel with its many small CTAs is faster
  // than 6 episodes per 256-thread CTA in the graph kernel's tail (336 -> 345 us per C3 step)
  if (agents < 1 || agents > 8) return 0;
  const int lanes = agents <= 8 ? 8 : agents <= 16 ? 16 : 32;
  const int epc = kTgM / agents;
  return epc * lanes <= kTgThreads && size_t(epc) * cell_stride <= 64 * 1024 &&
         size_t(env_stage_floats(epc, agents)) * sizeof(float) <= size_t(kTgProd) * kTgStageBytes;
}

// gso_is_old = 1 (mapf_policy_fwd behind its encoder kernels): see TcGraphArgs.  env != NULL (mapf_rollout): the env step
// of the same episodes is fused into the kernel (caller checked mapf_tc_graph_env_fusable; needs actions != NULL).
int mapf_tc_graph_head_ex(int32_t batch, int32_t agents, int32_t in_dim, int32_t out_dim, int32_t taps,
                          int32_t add_self_loop, int32_t has_skip, const float* x, const float* gso, const float* Wp,
                          const float* act_w, const float* act_b, int32_t num_actions, float* logits, int32_t* actions,
                          int gso_is_old, const mapf_env_step_args* env, mapf_stream_t stream) {
  MAPF_REQUIRE(x && gso && Wp && act_w && act_b && logits, MAPF_ERR_NULL);
  MAPF_REQUIRE(batch > 0 && num_actions > 0 && num_actions <= kMaxActions, MAPF_ERR_SHAPE);
  MAPF_REQUIRE(mapf_tc_graph_supported(agents, in_dim, out_dim, taps), MAPF_ERR_UNSUPPORTED);
  MAPF_REQUIRE(mapf_aligned(x, 16) && mapf_aligned(Wp, 16), MAPF_ERR_ALIGN);
  TcGraphArgs g{batch, agents, in_dim, taps, out_dim, has_skip ? 1 : 0, add_self_loop ? 1 : 0, num_actions,
                x, gso, Wp, act_w, act_b, logits, actions, gso_is_old};
  const int epc = kTgM / agents;
  const int ntiles = (batch + epc - 1) / epc;
  size_t smem = size_t(kTgProd) * kTgStageBytes + sizeof(float) * (size_t(kTgProd) * kTgM * 8 +
                ((size_t(epc) * agents * agents + 3) & ~size_t(3)) + kTgM);
  int envl = 0;
  if (env != nullptr) {
    MAPF_REQUIRE(actions != nullptr && env->episodes == batch && env->agents == agents, MAPF_ERR_SHAPE);
    MAPF_REQUIRE(mapf_tc_graph_env_fusable(agents, env->cell_stride) && gso_is_old, MAPF_ERR_UNSUPPORTED);
    envl = 8;   // mapf_tc_graph_env_fusable: agents <= 8
    g.env = *env;
    smem += sizeof(int32_t) * kTgM + ((sizeof(EnvShared<8, kTgThreads>) + 15) & ~size_t(15)) + size_t(epc) * env->cell_stride;
  }
  MAPF_REQUIRE(smem <= 227 * 1024, MAPF_ERR_UNSUPPORTED);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto launch = [&](auto kernel, int* cache) -> int {
    if (!mapf_ensure_smem(kernel, smem, cache)) return MAPF_ERR_CUDA;
    mapf_launch_pdl(kernel, dim3(ntiles), dim3(kTgThreads), smem, s, g, epc);
    return MAPF_OK;
  };
  static int c0[16] = {0}, c8[16] = {0};
  const int rc = envl == 0 ? launch(tc_graph_kernel<0>, c0) : launch(tc_graph_kernel<8>, c8);
  if (rc != MAPF_OK) return rc;
  return mapf_launch_status();
}
