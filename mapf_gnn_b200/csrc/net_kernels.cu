// Eval-mode policy network: weight packing, CNN+MLP observation encoder, K-tap polynomial graph
// filter (+ message-passing skip) fused with ReLU + action head + argmax, baseline head, and the
// host-side launch sequences for Network.forward and the on-device rollout loop.
//
// Semantics follow the reference's models/framework_gnn.py:143-181, framework_gnn_message.py,
// framework_baseline.py:109-129 and models/networks/gnn.py:72-126,178-235 (restated in SURVEY.md
// 8a A5-A9); the decomposition below is new.  All arithmetic is fp32 FFMA: the action argmax must be
// bit-exact against fp32 logits whose top-2 margin goes down to 3e-5, which rules out
// TF32/BF16 tensor-core inputs without error-compensating splits.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// weight packing
// ---------------------------------------------------------------------------------------------
constexpr float kBnEps = 1e-5f;  // nn.BatchNorm2d default, framework_gnn.py:70

__global__ void pack_weights_kernel(mapf_net_params p, PackedLayout L, float* __restrict__ out) {
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < L.total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    float v = 0.0f;
    bool hit = false;
    for (int l = 0; l < p.num_conv && !hit; ++l) {
      const int cin = p.channels[l], cout = p.channels[l + 1];
      const int64_t nw = int64_t(cout) * cin * 9;
      if (idx >= L.conv_w[l] && idx < L.conv_w[l] + nw) {
        // packed [Cout/4][Cin][9][4]
        int64_t r = idx - L.conv_w[l];
        const int c4 = int(r & 3); r >>= 2;
        const int tap = int(r % 9); r /= 9;
        const int ci = int(r % cin);
        const int co = int(r / cin) * 4 + c4;
        const float s = p.bn_gamma[l][co] / sqrtf(p.bn_var[l][co] + kBnEps);
        v = p.conv_w[l][(int64_t(co) * cin + ci) * 9 + tap] * s;
        hit = true;
      } else if (idx >= L.conv_b[l] && idx < L.conv_b[l] + cout) {
        const int co = int(idx - L.conv_b[l]);
        const float s = p.bn_gamma[l][co] / sqrtf(p.bn_var[l][co] + kBnEps);
        v = (p.conv_b[l][co] - p.bn_mean[l][co]) * s + p.bn_beta[l][co];
        hit = true;
      }
    }
    if (!hit) {
      const int F = p.fc_out;
      const int head_in = p.taps > 0 ? p.node_dim : F;
      if (idx >= L.fc_wt && idx < L.fc_wt + int64_t(p.fc_in) * F) {
        const int64_t r = idx - L.fc_wt;
        const int i = int(r / F), j = int(r % F);
        v = p.fc_w[int64_t(j) * p.fc_in + i];
      } else if (idx >= L.fc_b && idx < L.fc_b + F) {
        v = p.fc_b[idx - L.fc_b];
      } else if (p.taps > 0 && idx >= L.gf_wt && idx < L.gf_wt + int64_t(p.taps + 1) * F * p.node_dim) {
        // W_eff^T[j][g] = W.reshape(G, K*F)[g][j]  (raw reinterpretation of the [F,K,G] buffer, gnn.py:114-116);
        // rows K*F.. hold W2.reshape(G, F)^T (gnn.py:222)
        const int G = p.node_dim, KF = p.taps * F;
        const int64_t r = idx - L.gf_wt;
        const int j = int(r / G), g = int(r % G);
        if (j < KF) v = p.gf_w[int64_t(g) * KF + j];
        else if (p.gf_w2 != nullptr) v = p.gf_w2[int64_t(g) * F + (j - KF)];
      } else if (idx >= L.act_w && idx < L.act_w + int64_t(p.num_actions) * head_in) {
        v = p.act_w[idx - L.act_w];
      } else if (idx >= L.act_b && idx < L.act_b + p.num_actions) {
        v = p.act_b[idx - L.act_b];
      }
    }
    out[idx] = v;
  }
}

__global__ void pack_graph_weights_kernel(const float* __restrict__ w, const float* __restrict__ w2, int F, int K,
                                          int G, float* __restrict__ wt) {
  const int KF = K * F;
  const int64_t total = int64_t(KF + (w2 ? F : 0)) * G;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int j = int(idx / G), g = int(idx % G);
    wt[idx] = j < KF ? w[int64_t(g) * KF + j] : w2[int64_t(g) * F + (j - KF)];
  }
}

// ---------------------------------------------------------------------------------------------
// graph filter (+ head): 64 agent rows (whole episodes) per CTA
// ---------------------------------------------------------------------------------------------
constexpr int kGfRows = 64;
constexpr int kGfThreads = 256;
constexpr int kGfZs = 68;      // row stride of the k-major tap matrix Zt[KT][kGfZs]
constexpr int kGfColTile = 128;
constexpr int kGfChunk = 16;   // k rows of W^T staged per cp.async stage

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__global__ void __launch_bounds__(kGfThreads) graph_filter_kernel(mapf_graph_filter_fwd_args a, int epc) {
  extern __shared__ __align__(16) float s_gf[];
  const int N = a.agents, F = a.in_dim, G = a.out_dim, K = a.taps;
  const int KT = (K + a.has_skip) * F;
  float* zt = s_gf;                                   // [KT][kGfZs]
  float* wbuf = zt + KT * kGfZs;                      // [2][kGfChunk][kGfColTile]
  float* sm = wbuf + 2 * kGfChunk * kGfColTile;       // [epc][N][N]
  float* dinv = sm + epc * N * N;                     // [kGfRows]

  const int tid = threadIdx.x;
  const int e0 = blockIdx.x * epc;
  const int nep = min(epc, a.batch - e0);
  const int rows_used = nep * N;

  // ---- S = D^-1/2 (A [+ I]) D^-1/2 with row-sum degrees, 1/sqrt(0) -> 0 (gnn.py:87-101 / :185-199)
  for (int idx = tid; idx < nep * N * N; idx += kGfThreads) {
    const int el = idx / (N * N), rem = idx - el * N * N;
    const int i = rem / N, j = rem - i * N;
    float v = __ldg(a.gso + int64_t(e0 + el) * N * N + rem);
    if (a.add_self_loop && i == j) v += 1.0f;
    sm[idx] = v;
  }
  // ---- taps k = 0: Zt[f][row] = x[b, n, f]
  for (int idx = tid; idx < kGfRows * F; idx += kGfThreads) {
    const int row = idx / F, f = idx - row * F;
    float v = 0.0f;
    if (row < rows_used) {
      const int el = row / N, n = row - el * N;
      v = __ldg(a.x + int64_t(e0 + el) * a.x_sb + int64_t(n) * a.x_sn + int64_t(f) * a.x_sf);
    }
    zt[f * kGfZs + row] = v;
  }
  __syncthreads();
  if (tid < rows_used) {
    const int el = tid / N, i = tid - el * N;
    float deg = 0.0f;
    for (int j = 0; j < N; ++j) deg += sm[(el * N + i) * N + j];
    dinv[tid] = deg > 0.0f ? __fdiv_rn(1.0f, __fsqrt_rn(deg)) : 0.0f;  // torch.pow(D,-0.5) = 1/sqrt, inf -> 0
  }
  __syncthreads();
  for (int idx = tid; idx < nep * N * N; idx += kGfThreads) {
    const int el = idx / (N * N), rem = idx - el * N * N;
    const int i = rem / N, j = rem - i * N;
    sm[idx] = (dinv[el * N + i] * sm[idx]) * dinv[el * N + j];
  }
  if (a.has_skip) {
    // skip input = raw reshape of X[B,F,N] to [B,N,F] (gnn.py:201-204): Xr[n'][f'] = X[f = q / N][n = q % N], q = n'F + f'
    for (int idx = tid; idx < kGfRows * F; idx += kGfThreads) {
      const int row = idx % kGfRows, fp = idx / kGfRows;
      float v = 0.0f;
      if (row < rows_used) {
        const int el = row / N, np = row - el * N;
        const int q = np * F + fp;
        v = zt[(q / N) * kGfZs + el * N + (q % N)];
      }
      zt[(K * F + fp) * kGfZs + row] = v;
    }
  }
  __syncthreads();
  // ---- x_k = x_{k-1} S  (right-multiply, gnn.py:104-109): Zt[kF+f][row j] = sum_i S[i][j] Zt[(k-1)F+f][row i]
  {
    const int row = tid & (kGfRows - 1);
    const bool live = row < rows_used;
    const int el = live ? row / N : 0, jl = row - el * N;
    const float* scol = sm + el * N * N + jl;
    for (int k = 1; k < K; ++k) {
      for (int f = tid >> 6; f < F; f += kGfThreads / kGfRows) {
        float s = 0.0f;
        if (live) {
          const float* prev = zt + ((k - 1) * F + f) * kGfZs + el * N;
          for (int i = 0; i < N; ++i) s = fmaf(scol[i * N], prev[i], s);
        }
        zt[(k * F + f) * kGfZs + row] = s;
      }
      __syncthreads();
    }
  }
  if (a.z != nullptr) {
    for (int idx = tid; idx < rows_used * KT; idx += kGfThreads) {
      const int row = idx / KT, k = idx - row * KT;
      a.z[(int64_t(e0) * N + row) * KT + k] = zt[k * kGfZs + row];
    }
  }

  // ---- Y[64, G] = Zt^T W^T, column tiles of 128, W^T streamed through a 2-stage cp.async ring
  const int rq = tid & 15, cq = tid >> 4;
  const int A = a.num_actions;
  float part[4][kMaxActions];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int q = 0; q < kMaxActions; ++q) part[r][q] = 0.0f;

  const int nchunks = KT / kGfChunk;
  for (int g0 = 0; g0 < G; g0 += kGfColTile) {
    float acc[4][8];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[r][c] = 0.0f;

    auto stage = [&](int c, int b) {
      // 16 x 128 floats = 512 float4, two per thread
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int q = tid + h * kGfThreads;
        const int kk = q >> 5, c4 = (q & 31) << 2;
        float* dst = wbuf + (b * kGfChunk + kk) * kGfColTile + c4;
        if (g0 + c4 < G) cp_async16(dst, a.wt + int64_t(c * kGfChunk + kk) * G + g0 + c4);
        else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      cp_async_commit();
    };
    stage(0, 0);
    for (int c = 0; c < nchunks; ++c) {
      if (c + 1 < nchunks) { stage(c + 1, (c + 1) & 1); cp_async_wait<1>(); }
      else cp_async_wait<0>();
      __syncthreads();
      const float* wb = wbuf + (c & 1) * kGfChunk * kGfColTile + cq * 8;
      const float* zb = zt + (c * kGfChunk) * kGfZs + rq * 4;
#pragma unroll
      for (int kk = 0; kk < kGfChunk; ++kk) {
        const float4 z4 = *reinterpret_cast<const float4*>(zb + kk * kGfZs);
        const float4 w0 = *reinterpret_cast<const float4*>(wb + kk * kGfColTile);
        const float4 w1 = *reinterpret_cast<const float4*>(wb + kk * kGfColTile + 4);
        const float zv[4] = {z4.x, z4.y, z4.z, z4.w};
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) acc[r][cc] = fmaf(zv[r], wv[cc], acc[r][cc]);
      }
      __syncthreads();
    }

    // epilogue of this column tile
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = rq * 4 + r;
      const bool live = row < rows_used;
      const int el = live ? row / N : 0, n = row - el * N;
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        const int g = g0 + cq * 8 + cc;
        float v = acc[r][cc];
        if (a.relu) v = fmaxf(v, 0.0f);
        acc[r][cc] = v;
        if (a.y != nullptr && live && g < G)
          a.y[int64_t(e0 + el) * a.y_sb + int64_t(n) * a.y_sn + int64_t(g) * a.y_sg] = v;
      }
    }
    if (A > 0) {
#pragma unroll
      for (int q = 0; q < kMaxActions; ++q) {
        if (q < A) {
          const int gb = g0 + cq * 8;
          float wq[8];
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) wq[cc] = (gb + cc < G) ? __ldg(a.act_w + int64_t(q) * G + gb + cc) : 0.0f;
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) part[r][q] = fmaf(acc[r][cc], wq[cc], part[r][q]);
        }
      }
    }
  }

  if (A > 0) {
    // reduce the head partials over the 16 column groups: lanes l / l^16 hold the two groups of a warp,
    // the 8 warps meet in shared memory (wbuf is free again)
    float* red = wbuf;  // [8 warps][64 rows][kMaxActions]
    const int warp = tid >> 5;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int q = 0; q < kMaxActions; ++q) {
        float v = part[r][q];
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if ((tid & 16) == 0 && q < A) red[(warp * kGfRows + rq * 4 + r) * kMaxActions + q] = v;
      }
    __syncthreads();
    if (tid < rows_used) {
      const int64_t grow = int64_t(e0) * N + tid;
      float best = 0.0f;
      int arg = 0;
      for (int q = 0; q < A; ++q) {
        float v = __ldg(a.act_b + q);
        for (int w = 0; w < kGfThreads / 32; ++w) v += red[(w * kGfRows + tid) * kMaxActions + q];
        a.logits[grow * A + q] = v;
        if (q == 0 || v > best) { best = v; arg = q; }  // first max wins (np.argmax, evaluate.py:238)
      }
      if (a.actions != nullptr) a.actions[grow] = arg;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// baseline head: logits = x @ Wa^T + b, argmax (framework_baseline.py:96,126)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) head_fwd_kernel(mapf_head_fwd_args a) {
  extern __shared__ __align__(16) float s_head[];  // [32 rows][D + 1]
  const int D = a.in_dim, A = a.num_actions;
  const int64_t row0 = int64_t(blockIdx.x) * 32;
  const int nrows = (a.rows - row0) < 32 ? int(a.rows - row0) : 32;
  for (int idx = threadIdx.x; idx < nrows * D; idx += blockDim.x) {
    const int r = idx / D, c = idx - r * D;
    s_head[r * (D + 1) + c] = __ldg(a.x + row0 * D + idx);
  }
  __shared__ float s_logit[32][kMaxActions];
  __syncthreads();
  const int r = threadIdx.x & 31, q = threadIdx.x >> 5;
  if (r < nrows && q < A) {
    float v = __ldg(a.act_b + q);
    const float* w = a.act_w + int64_t(q) * D;
    for (int c = 0; c < D; ++c) v = fmaf(s_head[r * (D + 1) + c], __ldg(w + c), v);
    s_logit[r][q] = v;
    a.logits[(row0 + r) * A + q] = v;
  }
  __syncthreads();
  if (threadIdx.x < nrows && a.actions != nullptr) {
    float best = s_logit[threadIdx.x][0];
    int arg = 0;
    for (int k = 1; k < A; ++k)
      if (s_logit[threadIdx.x][k] > best) { best = s_logit[threadIdx.x][k]; arg = k; }
    a.actions[row0 + threadIdx.x] = arg;
  }
}


}  // namespace

extern "C" int64_t mapf_packed_weights_size(const mapf_net_params* p) {
  if (mapf_check_params(p) != MAPF_OK) return -1;
  return mapf_packed_layout(*p).total;
}

extern "C" int mapf_pack_weights(const mapf_net_params* p, float* packed, mapf_stream_t stream) {
  const int st = mapf_check_params(p);
  if (st != MAPF_OK) return st;
  MAPF_REQUIRE(packed != nullptr && p->fc_w && p->fc_b && p->act_w && p->act_b, MAPF_ERR_NULL);
  for (int l = 0; l < p->num_conv; ++l)
    MAPF_REQUIRE(p->conv_w[l] && p->conv_b[l] && p->bn_gamma[l] && p->bn_beta[l] && p->bn_mean[l] && p->bn_var[l],
                 MAPF_ERR_NULL);
  MAPF_REQUIRE(p->taps == 0 || p->gf_w, MAPF_ERR_NULL);
  MAPF_REQUIRE(mapf_aligned(packed, 16), MAPF_ERR_ALIGN);
  const PackedLayout L = mapf_packed_layout(*p);
  pack_weights_kernel<<<int((L.total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(*p, L, packed);
  return mapf_launch_status();
}

extern "C" int mapf_pack_graph_weights(const float* w, const float* w2, int32_t in_dim, int32_t taps, int32_t out_dim,
                                       float* wt, mapf_stream_t stream) {
  MAPF_REQUIRE(w && wt, MAPF_ERR_NULL);
  MAPF_REQUIRE(in_dim > 0 && taps > 0 && out_dim > 0, MAPF_ERR_SHAPE);
  const int64_t total = int64_t(taps + (w2 ? 1 : 0)) * in_dim * out_dim;
  pack_graph_weights_kernel<<<int((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(w, w2, in_dim, taps,
                                                                                                   out_dim, wt);
  return mapf_launch_status();
}

extern "C" int mapf_graph_filter_fwd(const mapf_graph_filter_fwd_args* a, mapf_stream_t stream) {
  MAPF_REQUIRE(a && a->x && a->gso && a->wt, MAPF_ERR_NULL);
  MAPF_REQUIRE(a->batch > 0 && a->agents > 0 && a->in_dim > 0 && a->out_dim > 0 && a->taps > 0, MAPF_ERR_SHAPE);
  MAPF_REQUIRE(a->agents <= kGfRows, MAPF_ERR_UNSUPPORTED);
  MAPF_REQUIRE((a->in_dim & 15) == 0 && (a->out_dim & 3) == 0, MAPF_ERR_UNSUPPORTED);
  MAPF_REQUIRE(a->num_actions >= 0 && a->num_actions <= kMaxActions, MAPF_ERR_UNSUPPORTED);
  MAPF_REQUIRE(a->num_actions == 0 || (a->act_w && a->act_b && a->logits), MAPF_ERR_NULL);
  MAPF_REQUIRE(a->num_actions > 0 || a->y || a->z, MAPF_ERR_NULL);
  MAPF_REQUIRE(mapf_aligned(a->wt, 16), MAPF_ERR_ALIGN);
  const int epc = max(1, kGfRows / a->agents);
  const int KT = (a->taps + (a->has_skip ? 1 : 0)) * a->in_dim;
  const size_t smem = sizeof(float) * (size_t(KT) * kGfZs + 2 * kGfChunk * kGfColTile +
                                       size_t(epc) * a->agents * a->agents + kGfRows);
  MAPF_REQUIRE(smem <= 220 * 1024, MAPF_ERR_UNSUPPORTED);
  if (cudaFuncSetAttribute(graph_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess)
    return MAPF_ERR_CUDA;
  const int grid = (a->batch + epc - 1) / epc;
  graph_filter_kernel<<<grid, kGfThreads, smem, static_cast<cudaStream_t>(stream)>>>(*a, epc);
  return mapf_launch_status();
}

extern "C" int mapf_head_fwd(const mapf_head_fwd_args* a, mapf_stream_t stream) {
  MAPF_REQUIRE(a && a->x && a->act_w && a->act_b && a->logits, MAPF_ERR_NULL);
  MAPF_REQUIRE(a->rows > 0 && a->in_dim > 0, MAPF_ERR_SHAPE);
  MAPF_REQUIRE(a->num_actions > 0 && a->num_actions <= kMaxActions && a->in_dim <= 1024, MAPF_ERR_UNSUPPORTED);
  const size_t smem = sizeof(float) * 32 * (a->in_dim + 1);
  if (smem > 40 * 1024 &&
      cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess)
    return MAPF_ERR_CUDA;
  head_fwd_kernel<<<unsigned((a->rows + 31) / 32), 256, smem, static_cast<cudaStream_t>(stream)>>>(*a);
  return mapf_launch_status();
}

extern "C" int mapf_policy_fwd(const mapf_net_params* p, const mapf_policy_fwd_args* a, mapf_stream_t stream) {
  const int st = mapf_check_params(p);
  if (st != MAPF_OK) return st;
  MAPF_REQUIRE(a && a->states && a->packed && a->workspace_x && a->logits, MAPF_ERR_NULL);
  MAPF_REQUIRE(a->batch > 0 && a->agents > 0, MAPF_ERR_SHAPE);
  const PackedLayout L = mapf_packed_layout(*p);
  const int64_t rows = int64_t(a->batch) * a->agents;
  mapf_encoder_fwd_args e{rows, a->states, a->packed, a->workspace_x};
  int rc = mapf_encoder_fwd(p, &e, stream);
  if (rc != MAPF_OK) return rc;
  if (p->taps == 0) {
    mapf_head_fwd_args h{rows, p->fc_out, p->num_actions, a->workspace_x, a->packed + L.act_w,
                         a->packed + L.act_b, a->logits, a->actions};
    return mapf_head_fwd(&h, stream);
  }
  MAPF_REQUIRE(a->gso != nullptr, MAPF_ERR_NULL);
  mapf_graph_filter_fwd_args g{};
  g.batch = a->batch; g.agents = a->agents; g.in_dim = p->fc_out; g.out_dim = p->node_dim; g.taps = p->taps;
  g.add_self_loop = a->message ? 0 : 1;
  g.has_skip = a->message ? 1 : 0;
  g.relu = 1;
  g.num_actions = p->num_actions;
  g.x = a->workspace_x; g.x_sb = int64_t(a->agents) * p->fc_out; g.x_sn = p->fc_out; g.x_sf = 1;
  g.gso = a->gso;
  g.wt = a->packed + L.gf_wt;
  g.y = nullptr; g.z = nullptr;
  g.act_w = a->packed + L.act_w; g.act_b = a->packed + L.act_b;
  g.logits = a->logits; g.actions = a->actions;
  return mapf_graph_filter_fwd(&g, stream);
}

extern "C" int mapf_rollout(const mapf_net_params* p, const mapf_rollout_args* a, mapf_stream_t stream) {
  MAPF_REQUIRE(a != nullptr, MAPF_ERR_NULL);
  MAPF_REQUIRE(a->steps >= 0, MAPF_ERR_SHAPE);
  MAPF_REQUIRE(a->policy.actions != nullptr && a->env.actions == a->policy.actions, MAPF_ERR_NULL);
  MAPF_REQUIRE(a->env.fov == a->policy.states && (p->taps == 0 || a->env.adj == a->policy.gso), MAPF_ERR_SHAPE);
  MAPF_REQUIRE(a->env.episodes == a->policy.batch && a->env.agents == a->policy.agents, MAPF_ERR_SHAPE);
  for (int t = 0; t < a->steps; ++t) {
    int rc = mapf_policy_fwd(p, &a->policy, stream);
    if (rc != MAPF_OK) return rc;
    mapf_env_step_args e = a->env;
    e.do_move = 1;
    rc = mapf_env_step(&e, stream);
    if (rc != MAPF_OK) return rc;
  }
  return MAPF_OK;
}

extern "C" const char* mapf_last_error_string(int code) {
  switch (code) {
    case MAPF_OK: return "ok";
    case MAPF_ERR_NULL: return "a required pointer is NULL";
    case MAPF_ERR_SHAPE: return "a size is non-positive or inconsistent";
    case MAPF_ERR_UNSUPPORTED: return "size outside what the kernels are built for";
    case MAPF_ERR_ALIGN: return "pointer or stride alignment requirement violated";
    case MAPF_ERR_CUDA: return "the CUDA runtime reported an error at launch";
    default: return "unknown mapf status";
  }
}

extern "C" const char* mapf_version(void) { return "mapf_b200 0.1 sm_100a"; }
