// K-tap polynomial graph filter sum_k S^k X H_k (GCNLayer, models/networks/gnn.py:72-126) and its
// message-passing variant with the W2 skip term (A7: gnn.py:178-235), optionally fused with the ReLU + actionMLP +
// first-max argmax tail of Network.forward (A8: framework_gnn.py:171-181, evaluate.py:238).
//
// Design (B200): latency-oriented.  One small CTA (4*TM threads) per tile of TM agent rows holding whole
// episodes (TM = 32 when N <= 16, else 64), several CTAs resident per SM so that the short serial phases
// of different tiles overlap.  Per tile, entirely in shared memory:
//   S = D^-1/2 (A [+ I]) D^-1/2 (row-sum degrees, 1/sqrt(0) -> 0)      gnn.py:87-101 / :185-199
//   Zt[k*F+f][row] : tap 0 = X, tap k = tap k-1 right-multiplied by S   gnn.py:104-109
//   skip rows      : raw reshape of X[B,F,N] to [B,N,F]                gnn.py:201-204
//   Y = Zt^T W_eff^T as a register-tiled fp32 GEMM (4 rows x 8 columns per thread) whose B operand, the packed
//   64-128 KB filter matrix, is read through the read-only path and stays L1/L2 resident; ReLU; 5-way head
//   partials reduced by warp shuffles + shared memory; first-max argmax.
// The hops are a register-tiled [F x N][N x N] product when N is a multiple of 4 (they cost N MACs per output and
// would otherwise dominate at N = 64).  Unfused, the recurrence alone would be an HBM-bound kernel (1 044 B/agent
// at C1); fused, a tile reads 4F+4N bytes and writes 24 bytes per agent and the kernel is FFMA-bound.
#include "common.cuh"

namespace {

constexpr int kGfColTile = 128;
constexpr int CN = 8;  // columns per thread

template <int TM>
__global__ void __launch_bounds__(TM * 4) graph_filter_kernel(mapf_graph_filter_fwd_args a, int epc) {
  constexpr int NT = TM * 4;          // threads: (TM/4 row groups) x (16 column groups of 8) = one 128-column tile
  constexpr int ZS = TM + 4;          // row stride of Zt (keeps 16-byte alignment of 4-row groups)
  constexpr int RG = TM / 4;
  constexpr int NW = NT / 32;
  extern __shared__ __align__(16) float s_gf[];
  const int N = a.agents, F = a.in_dim, G = a.out_dim, K = a.taps;
  const int KT = (K + a.has_skip) * F;
  float* zt = s_gf;                                    // [KT][ZS]
  float* sm = zt + KT * ZS;                            // [epc][N][N]
  float* dinv = sm + ((epc * N * N + 3) & ~3);         // [TM]
  float* red = dinv + TM;                              // [NW][TM][kMaxActions]

  const int tid = threadIdx.x;
  const int rq = tid % RG, cq = tid / RG;
  const int A = a.num_actions;
  const int e0 = blockIdx.x * epc;
  const int nep = min(epc, a.batch - e0);
  const int rows_used = nep * N;

  for (int idx = tid; idx < nep * N * N; idx += NT) {
    const int el = idx / (N * N), rem = idx - el * N * N;
    const int i = rem / N, j = rem - i * N;
    float v = mapf_ld_dep(a.gso + int64_t(e0 + el) * N * N + rem);
    if (a.add_self_loop && i == j) v += 1.0f;
    sm[idx] = v;
  }
  // tap 0: Zt[f][row] = x[b, n, f]
  if (a.x_sf == 1 && a.x_sn == F && a.x_sb == int64_t(N) * F && (F & 3) == 0 &&
      (reinterpret_cast<uintptr_t>(a.x) & 15) == 0) {
    // agent-major rows: the tile is one contiguous [rows_used][F] block -> 128-bit loads
    const float4* src = reinterpret_cast<const float4*>(a.x + int64_t(e0) * N * F);
    const int f4 = F >> 2;
    for (int idx = tid; idx < TM * f4; idx += NT) {
      const int row = idx / f4, c = (idx - row * f4) << 2;
      const float4 v = row < rows_used ? mapf_ld_dep(src + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
      zt[(c + 0) * ZS + row] = v.x;
      zt[(c + 1) * ZS + row] = v.y;
      zt[(c + 2) * ZS + row] = v.z;
      zt[(c + 3) * ZS + row] = v.w;
    }
  } else {
    for (int idx = tid; idx < TM * F; idx += NT) {
      const int row = idx / F, f = idx - row * F;
      float v = 0.0f;
      if (row < rows_used) {
        const int el = row / N, n = row - el * N;
        v = mapf_ld_dep(a.x + int64_t(e0 + el) * a.x_sb + int64_t(n) * a.x_sn + int64_t(f) * a.x_sf);
      }
      zt[f * ZS + row] = v;
    }
  }
  __syncthreads();
  if (tid < rows_used) {
    const int el = tid / N, i = tid - el * N;
    float deg = 0.0f;
    for (int j = 0; j < N; ++j) deg += sm[(el * N + i) * N + j];
    dinv[tid] = deg > 0.0f ? __fdiv_rn(1.0f, __fsqrt_rn(deg)) : 0.0f;  // torch.pow(D, -0.5), inf -> 0
  }
  __syncthreads();
  for (int idx = tid; idx < nep * N * N; idx += NT) {
    const int el = idx / (N * N), rem = idx - el * N * N;
    const int i = rem / N, j = rem - i * N;
    sm[idx] = (dinv[el * N + i] * sm[idx]) * dinv[el * N + j];
  }
  if (a.has_skip) {
    // Xr[n'][f'] = X_flat[q], q = n'F + f', X_flat index of X[b,f,n] is f*N + n
    for (int idx = tid; idx < TM * F; idx += NT) {
      const int row = idx % TM, fp = idx / TM;
      float v = 0.0f;
      if (row < rows_used) {
        const int el = row / N, np = row - el * N;
        const int q = np * F + fp;
        v = zt[(q / N) * ZS + el * N + (q % N)];
      }
      zt[(K * F + fp) * ZS + row] = v;
    }
  }
  __syncthreads();
  // x_k = x_{k-1} S : Zt[kF+f][row j] = sum_i S[i][j] Zt[(k-1)F+f][row i]
  if ((N & 3) == 0 && (F & 3) == 0) {
    // register-tiled: one thread = 4 features x 4 columns j of one episode, i in steps of 4
    const int jt = N >> 2, ft = F >> 2;
    const int per_ep = jt * ft;
    for (int k = 1; k < K; ++k) {
      for (int t = tid; t < nep * per_ep; t += NT) {
        const int el = t / per_ep, rem = t - el * per_ep;
        const int f0 = (rem / jt) * 4, j0 = (rem % jt) * 4;
        const float* sp = sm + el * N * N + j0;
        const float* zp = zt + ((k - 1) * F + f0) * ZS + el * N;
        float acc[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) acc[u][v] = 0.0f;
        for (int i0 = 0; i0 < N; i0 += 4) {
          float4 zr[4], sr[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) zr[u] = *reinterpret_cast<const float4*>(zp + u * ZS + i0);
#pragma unroll
          for (int u = 0; u < 4; ++u) sr[u] = *reinterpret_cast<const float4*>(sp + (i0 + u) * N);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float zi[4] = {zr[u].x, zr[u].y, zr[u].z, zr[u].w};
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
              acc[u][0] = fmaf(sr[ii].x, zi[ii], acc[u][0]);
              acc[u][1] = fmaf(sr[ii].y, zi[ii], acc[u][1]);
              acc[u][2] = fmaf(sr[ii].z, zi[ii], acc[u][2]);
              acc[u][3] = fmaf(sr[ii].w, zi[ii], acc[u][3]);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          *reinterpret_cast<float4*>(zt + (k * F + f0 + u) * ZS + el * N + j0) =
              make_float4(acc[u][0], acc[u][1], acc[u][2], acc[u][3]);
      }
      __syncthreads();
    }
  } else {
    const int row = tid % TM;
    const bool live = row < rows_used;
    const int el = live ? row / N : 0, jl = row - el * N;
    const float* scol = sm + el * N * N + jl;
    for (int k = 1; k < K; ++k) {
      for (int f = tid / TM; f < F; f += NT / TM) {
        float s = 0.0f;
        if (live) {
          const float* prev = zt + ((k - 1) * F + f) * ZS + el * N;
          for (int i = 0; i < N; ++i) s = fmaf(scol[i * N], prev[i], s);
        }
        zt[(k * F + f) * ZS + row] = s;
      }
      __syncthreads();
    }
  }
  if (a.z != nullptr) {
    for (int idx = tid; idx < rows_used * KT; idx += NT) {
      const int row = idx / KT, k = idx - row * KT;
      a.z[(int64_t(e0) * N + row) * KT + k] = zt[k * ZS + row];
    }
  }

  float part[4][kMaxActions];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int q = 0; q < kMaxActions; ++q) part[r][q] = 0.0f;

  for (int g0 = 0; g0 < G; g0 += kGfColTile) {
    float acc[4][CN];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < CN; ++c) acc[r][c] = 0.0f;
    const int gb = g0 + cq * CN;
    const bool clive = gb < G;   // G is a multiple of 8: column groups never straddle G
    if (clive) {
      const float* zb = zt + rq * 4;
      const float4* wb = reinterpret_cast<const float4*>(a.wt + gb);
      const int wstride = G >> 2;
#pragma unroll 4
      for (int k = 0; k < KT; ++k) {
        const float4 z4 = *reinterpret_cast<const float4*>(zb + k * ZS);
        const float4 w0 = __ldg(wb + int64_t(k) * wstride);
        const float4 w1 = __ldg(wb + int64_t(k) * wstride + 1);
        const float zv[4] = {z4.x, z4.y, z4.z, z4.w};
        const float wv[CN] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < CN; ++c) acc[r][c] = fmaf(zv[r], wv[c], acc[r][c]);
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = rq * 4 + r;
      const bool live = row < rows_used;
      const int el = live ? row / N : 0, n = row - el * N;
#pragma unroll
      for (int c = 0; c < CN; ++c) {
        float v = acc[r][c];
        if (a.relu) v = fmaxf(v, 0.0f);
        acc[r][c] = v;
        if (a.y != nullptr && live && clive)
          a.y[int64_t(e0 + el) * a.y_sb + int64_t(n) * a.y_sn + int64_t(gb + c) * a.y_sg] = v;
      }
    }
    if (A > 0 && clive) {
#pragma unroll
      for (int q = 0; q < kMaxActions; ++q) {
        if (q < A) {
          const float4 q0 = __ldg(reinterpret_cast<const float4*>(a.act_w + int64_t(q) * G + gb));
          const float4 q1 = __ldg(reinterpret_cast<const float4*>(a.act_w + int64_t(q) * G + gb) + 1);
          const float wq[CN] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < CN; ++c) part[r][q] = fmaf(acc[r][c], wq[c], part[r][q]);
        }
      }
    }
  }

  if (A > 0) {
    // lanes of a warp that share rq hold different column groups: fold them with shuffles, then the warps
    // meet in shared memory
    const int warp = tid >> 5;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int q = 0; q < kMaxActions; ++q) {
        if (q < A) {
          float v = part[r][q];
          if (RG == 8) v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if ((tid & 31) < RG) red[(warp * TM + rq * 4 + r) * kMaxActions + q] = v;
        }
      }
    __syncthreads();
    if (tid < rows_used) {
      const int64_t grow = int64_t(e0) * N + tid;
      float best = 0.0f;
      int arg = 0;
      for (int q = 0; q < A; ++q) {
        float v = __ldg(a.act_b + q);
        for (int w = 0; w < NW; ++w) v += red[(w * TM + tid) * kMaxActions + q];
        a.logits[grow * A + q] = v;
        if (q == 0 || v > best) { best = v; arg = q; }  // first max wins (np.argmax, evaluate.py:238)
      }
      if (a.actions != nullptr) a.actions[grow] = arg;
    }
  }
}

template <int TM>
int launch_gf(const mapf_graph_filter_fwd_args& a, int epc, int ntiles, size_t smem, cudaStream_t s) {
  auto kernel = graph_filter_kernel<TM>;
  if ([&]() { static int cache__[16] = {0}; return !mapf_ensure_smem(kernel, size_t(smem), cache__); }())
    return MAPF_ERR_CUDA;
  kernel<<<ntiles, TM * 4, smem, s>>>(a, epc);
  return mapf_launch_status();
}

}  // namespace

extern "C" int mapf_graph_filter_fwd(const mapf_graph_filter_fwd_args* a, mapf_stream_t stream) {
  MAPF_REQUIRE(a && a->x && a->gso && a->wt, MAPF_ERR_NULL);
  MAPF_REQUIRE(a->batch > 0 && a->agents > 0 && a->in_dim > 0 && a->out_dim > 0 && a->taps > 0, MAPF_ERR_SHAPE);
  MAPF_REQUIRE(a->agents <= 64, MAPF_ERR_UNSUPPORTED);
  MAPF_REQUIRE((a->in_dim & 3) == 0 && (a->out_dim & 7) == 0, MAPF_ERR_UNSUPPORTED);
  MAPF_REQUIRE(a->num_actions >= 0 && a->num_actions <= kMaxActions, MAPF_ERR_UNSUPPORTED);
  MAPF_REQUIRE(a->num_actions == 0 || (a->act_w && a->act_b && a->logits), MAPF_ERR_NULL);
  MAPF_REQUIRE(a->num_actions > 0 || a->y || a->z, MAPF_ERR_NULL);
  MAPF_REQUIRE(mapf_aligned(a->wt, 16), MAPF_ERR_ALIGN);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int N = a->agents;
  const int TM = N <= 16 ? 32 : 64;
  const int epc = TM / N;
  const int ntiles = (a->batch + epc - 1) / epc;
  const int KT = (a->taps + (a->has_skip ? 1 : 0)) * a->in_dim;
  MAPF_REQUIRE(mapf_aligned(a->act_w, 16) || a->num_actions == 0, MAPF_ERR_ALIGN);
  const size_t smem = sizeof(float) * (size_t(KT) * (TM + 4) + ((size_t(epc) * N * N + 3) & ~size_t(3)) + TM +
                                       size_t(TM / 8) * TM * kMaxActions);
  MAPF_REQUIRE(smem <= 227 * 1024, MAPF_ERR_UNSUPPORTED);
  return TM == 32 ? launch_gf<32>(*a, epc, ntiles, smem, s) : launch_gf<64>(*a, epc, ntiles, smem, s);
}

// ---------------------------------------------------------------------------------------------------------
// Filter taps only: Z[row][k*F+f] = x_k[f, n] (k < K) and the message-passing skip columns, row-major [B*N, KT].
// Feeds the tensor-core mix (mapf_tc_mix_head).  Memory-light (reads 4F+4N, writes 4*KT bytes per agent), one
// CTA of TZ*4 threads per tile of TZ rows of whole episodes; taps are kept row-major [row][KT] in shared memory so
// that the hop reads are conflict-free (lanes over features) and the tile leaves as one contiguous block.
// ---------------------------------------------------------------------------------------------------------
namespace {

template <int TZ>
__global__ void __launch_bounds__(TZ * 4) graph_taps_kernel(mapf_graph_filter_fwd_args a, int epc) {
  constexpr int NT = TZ * 4;
  extern __shared__ __align__(16) float s_gt[];
  const int N = a.agents, F = a.in_dim, K = a.taps;
  const int KT = (K + a.has_skip) * F;
  const int ZR = KT + 4;                               // row stride (floats), keeps 16-byte alignment
  float* z = s_gt;                                     // [TZ][ZR]
  float* sm = z + TZ * ZR;                             // [epc][N][N]
  float* dinv = sm + ((epc * N * N + 3) & ~3);         // [TZ]
  const int tid = threadIdx.x;
  const int e0 = blockIdx.x * epc;
  const int nep = min(epc, a.batch - e0);
  const int rows_used = nep * N;

  for (int idx = tid; idx < nep * N * N; idx += NT) {
    const int el = idx / (N * N), rem = idx - el * N * N;
    const int i = rem / N, j = rem - i * N;
    float v = mapf_ld_dep(a.gso + int64_t(e0 + el) * N * N + rem);
    if (a.add_self_loop && i == j) v += 1.0f;
    sm[idx] = v;
  }
  {  // tap 0 (agent-major x rows are contiguous)
    const float4* src = reinterpret_cast<const float4*>(a.x + int64_t(e0) * N * F);
    const int f4 = F >> 2;
    for (int idx = tid; idx < rows_used * f4; idx += NT) {
      const int row = idx / f4, c = (idx - row * f4) << 2;
      *reinterpret_cast<float4*>(z + row * ZR + c) = mapf_ld_dep(src + idx);
    }
  }
  __syncthreads();
  if (tid < rows_used) {
    const int el = tid / N, i = tid - el * N;
    float deg = 0.0f;
    for (int j = 0; j < N; ++j) deg += sm[(el * N + i) * N + j];
    dinv[tid] = deg > 0.0f ? __fdiv_rn(1.0f, __fsqrt_rn(deg)) : 0.0f;
  }
  __syncthreads();
  for (int idx = tid; idx < nep * N * N; idx += NT) {
    const int el = idx / (N * N), rem = idx - el * N * N;
    const int i = rem / N, j = rem - i * N;
    sm[idx] = (dinv[el * N + i] * sm[idx]) * dinv[el * N + j];
  }
  if (a.has_skip) {
    for (int idx = tid; idx < rows_used * F; idx += NT) {
      const int row = idx / F, fp = idx - row * F;
      const int el = row / N, np = row - el * N;
      const int q = np * F + fp;
      z[row * ZR + K * F + fp] = z[(el * N + (q % N)) * ZR + (q / N)];
    }
  }
  __syncthreads();
  const int f4n = F >> 2;
  for (int k = 1; k < K; ++k) {
    if ((N & 3) == 0) {
      // 4 rows j x 4 features per thread, i in steps of 1: S[i][j0..j0+3] and x[i][f0..f0+3] are both 128-bit loads
      const int jt = N >> 2;
      for (int t = tid; t < nep * jt * f4n; t += NT) {
        const int fq = t % f4n, r = t / f4n;
        const int el = r / jt, j0 = (r - el * jt) << 2;
        const float* sp = sm + el * N * N + j0;
        const float* xp = z + (el * N) * ZR + (k - 1) * F + (fq << 2);
        float acc[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) acc[u][v] = 0.0f;
        for (int i = 0; i < N; ++i) {
          const float4 s4 = *reinterpret_cast<const float4*>(sp + i * N);
          const float4 x4 = *reinterpret_cast<const float4*>(xp + i * ZR);
          const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            acc[u][0] = fmaf(sv[u], x4.x, acc[u][0]);
            acc[u][1] = fmaf(sv[u], x4.y, acc[u][1]);
            acc[u][2] = fmaf(sv[u], x4.z, acc[u][2]);
            acc[u][3] = fmaf(sv[u], x4.w, acc[u][3]);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          *reinterpret_cast<float4*>(z + (el * N + j0 + u) * ZR + k * F + (fq << 2)) =
              make_float4(acc[u][0], acc[u][1], acc[u][2], acc[u][3]);
      }
    } else {
      for (int t = tid; t < rows_used * f4n; t += NT) {
        const int fq = t % f4n, row = t / f4n;
        const int el = row / N, jl = row - el * N;
        const float* sp = sm + el * N * N + jl;
        const float* xp = z + (el * N) * ZR + (k - 1) * F + (fq << 2);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = 0; i < N; ++i) {
          const float s = sp[i * N];
          const float4 x4 = *reinterpret_cast<const float4*>(xp + i * ZR);
          acc.x = fmaf(s, x4.x, acc.x); acc.y = fmaf(s, x4.y, acc.y);
          acc.z = fmaf(s, x4.z, acc.z); acc.w = fmaf(s, x4.w, acc.w);
        }
        *reinterpret_cast<float4*>(z + row * ZR + k * F + (fq << 2)) = acc;
      }
    }
    __syncthreads();
  }
  const int k4 = KT >> 2;
  float4* dst = reinterpret_cast<float4*>(a.z + int64_t(e0) * N * KT);
  for (int idx = tid; idx < rows_used * k4; idx += NT) {
    const int row = idx / k4, c = (idx - row * k4) << 2;
    dst[idx] = *reinterpret_cast<const float4*>(z + row * ZR + c);
  }
}

template <int TZ>
int launch_taps(const mapf_graph_filter_fwd_args& a, int epc, int ntiles, size_t smem, cudaStream_t s) {
  auto kernel = graph_taps_kernel<TZ>;
  if ([&]() { static int cache__[16] = {0}; return !mapf_ensure_smem(kernel, size_t(smem), cache__); }())
    return MAPF_ERR_CUDA;
  kernel<<<ntiles, TZ * 4, smem, s>>>(a, epc);
  return mapf_launch_status();
}

}  // namespace

// Z [B*N, (K+has_skip)*F] from agent-major x [B*N, F] and gso [B,N,N]; only x, gso, z and the size / flag fields of
// the argument struct are used.
int mapf_hops_tc_supported(const mapf_graph_filter_fwd_args* a);                       // tc_hops.cu
int mapf_graph_hops_tc(const mapf_graph_filter_fwd_args* a, mapf_stream_t stream);     // tc_hops.cu

extern "C" int mapf_graph_taps(const mapf_graph_filter_fwd_args* a, mapf_stream_t stream) {
  MAPF_REQUIRE(a && a->x && a->gso && a->z, MAPF_ERR_NULL);
  // 64-agent teams: hops on the tensor cores (tc_hops.cu); everything else: the FFMA kernel below
  if (a->x_sf == 1 && a->x_sn == a->in_dim && a->x_sb == int64_t(a->agents) * a->in_dim && mapf_hops_tc_supported(a))
    return mapf_graph_hops_tc(a, stream);
  return mapf_graph_taps_ffma(a, stream);
}

extern "C" int mapf_graph_taps_ffma(const mapf_graph_filter_fwd_args* a, mapf_stream_t stream) {
  MAPF_REQUIRE(a && a->x && a->gso && a->z, MAPF_ERR_NULL);
  MAPF_REQUIRE(a->batch > 0 && a->agents > 0 && a->in_dim > 0 && a->taps > 0, MAPF_ERR_SHAPE);
  MAPF_REQUIRE(a->agents <= 64 && (a->in_dim & 3) == 0, MAPF_ERR_UNSUPPORTED);
  MAPF_REQUIRE(a->x_sf == 1 && a->x_sn == a->in_dim && a->x_sb == int64_t(a->agents) * a->in_dim, MAPF_ERR_UNSUPPORTED);
  MAPF_REQUIRE(mapf_aligned(a->x, 16) && mapf_aligned(a->z, 16), MAPF_ERR_ALIGN);
  const int N = a->agents;
  const int TZ = N <= 16 ? 32 : 64;
  const int epc = TZ / N;
  const int ntiles = (a->batch + epc - 1) / epc;
  const int KT = (a->taps + (a->has_skip ? 1 : 0)) * a->in_dim;
  const size_t smem = sizeof(float) * (size_t(TZ) * (KT + 4) + ((size_t(epc) * N * N + 3) & ~size_t(3)) + TZ);
  MAPF_REQUIRE(smem <= 227 * 1024, MAPF_ERR_UNSUPPORTED);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return TZ == 32 ? launch_taps<32>(*a, epc, ntiles, smem, s) : launch_taps<64>(*a, epc, ntiles, smem, s);
}
