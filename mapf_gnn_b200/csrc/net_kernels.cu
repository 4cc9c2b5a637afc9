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

int mapf_encoder_tc_pack(const mapf_net_params* p, float* packed, mapf_stream_t stream);   // enc_tc.cu
int mapf_tc_graph_head_ex(int32_t batch, int32_t agents, int32_t in_dim, int32_t out_dim, int32_t taps,
                          int32_t add_self_loop, int32_t has_skip, const float* x, const float* gso, const float* Wp,
                          const float* act_w, const float* act_b, int32_t num_actions, float* logits, int32_t* actions,
                          int gso_is_old, const mapf_env_step_args* env, mapf_stream_t stream);   // tc_graph.cu
int mapf_tc_graph_env_fusable(int32_t agents, int32_t cell_stride);                          // tc_graph.cu

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

// Ends every weight packing.  The step kernels read the packed weights BEFORE their griddepcontrol.wait (weights into
// shared memory under the previous kernel's tail); that is safe because a programmatically launched kernel can only start
// once the kernel right in front of it has started everywhere -- and this normally launched no-op starts only after the
// packing kernels have completed and flushed.
__global__ void pack_fence_kernel() {}

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
// baseline head: logits = x @ Wa^T + b, argmax (framework_baseline.py:96,126)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) head_fwd_kernel(mapf_head_fwd_args a) {
  extern __shared__ __align__(16) float s_head[];  // [32 rows][D + 1]
  mapf_pdl_wait();
  mapf_pdl_launch();
  const int D = a.in_dim, A = a.num_actions;
  const int64_t row0 = int64_t(blockIdx.x) * 32;
  const int nrows = (a.rows - row0) < 32 ? int(a.rows - row0) : 32;
  for (int idx = threadIdx.x; idx < nrows * D; idx += blockDim.x) {
    const int r = idx / D, c = idx - r * D;
    s_head[r * (D + 1) + c] = mapf_ld_dep(a.x + row0 * D + idx);
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
  if (p->fc_out <= 128) {
    const int rc = mapf_tc_pack_b(p->fc_w, p->fc_in, nullptr, 0, p->fc_out, packed + L.tc_fc, stream);
    if (rc != MAPF_OK) return rc;
  }
  if (p->taps > 0 && p->node_dim <= 128) {
    // B operand of the tensor-core mix: GFL.0.W viewed as [G, K*F] (+ W2 viewed as [G, F]) -- the raw reshapes of
    // gnn.py:114-116,222 are exactly the K-major operand the MMA wants
    const int rc = mapf_tc_pack_b(p->gf_w, p->taps * p->fc_out, p->gf_w2, p->gf_w2 ? p->fc_out : 0, p->node_dim,
                                  packed + L.tc_gf, stream);
    if (rc != MAPF_OK) return rc;
  }
  if (mapf_enc_tc_dims_ok(*p)) {
    const int rc = mapf_encoder_tc_pack(p, packed, stream);
    if (rc != MAPF_OK) return rc;
  }
  pack_fence_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>();
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

extern "C" int mapf_head_fwd(const mapf_head_fwd_args* a, mapf_stream_t stream) {
  MAPF_REQUIRE(a && a->x && a->act_w && a->act_b && a->logits, MAPF_ERR_NULL);
  MAPF_REQUIRE(a->rows > 0 && a->in_dim > 0, MAPF_ERR_SHAPE);
  MAPF_REQUIRE(a->num_actions > 0 && a->num_actions <= kMaxActions && a->in_dim <= 1024, MAPF_ERR_UNSUPPORTED);
  const size_t smem = sizeof(float) * 32 * (a->in_dim + 1);
  if (smem > 40 * 1024 &&
      [&]() { static int cache__[16] = {0}; return !mapf_ensure_smem(head_fwd_kernel, size_t(smem), cache__); }())
    return MAPF_ERR_CUDA;
  mapf_launch_pdl(head_fwd_kernel, dim3(unsigned((a->rows + 31) / 32)), dim3(256), smem, static_cast<cudaStream_t>(stream), *a);
  return mapf_launch_status();
}

// env != NULL (mapf_rollout): when the fused tensor-core graph kernel runs and the shapes allow it, the environment step of
// the same episodes is executed in that kernel's tail and *env_done is set; the caller launches mapf_env_step otherwise.
static int policy_fwd_impl(const mapf_net_params* p, const mapf_policy_fwd_args* a, const mapf_env_step_args* env,
                           int* env_done, mapf_stream_t stream);

extern "C" int mapf_policy_fwd(const mapf_net_params* p, const mapf_policy_fwd_args* a, mapf_stream_t stream) {
  return policy_fwd_impl(p, a, nullptr, nullptr, stream);
}

static int policy_fwd_impl(const mapf_net_params* p, const mapf_policy_fwd_args* a, const mapf_env_step_args* env,
                           int* env_done, mapf_stream_t stream) {
  const int st = mapf_check_params(p);
  if (st != MAPF_OK) return st;
  MAPF_REQUIRE(a && a->states && a->packed && a->workspace_x && a->logits, MAPF_ERR_NULL);
  MAPF_REQUIRE(a->batch > 0 && a->agents > 0, MAPF_ERR_SHAPE);
  const MapfPdlScope pdl_scope((a->flags & MAPF_POLICY_NO_PDL) != 0);
  const PackedLayout L = mapf_packed_layout(*p);
  const int64_t rows = int64_t(a->batch) * a->agents;
  int rc;
  if (a->workspace_act != nullptr && p->fc_out <= 128) {
    // conv stack on FFMA2 -> flattened activations; Linear + ReLU as a 3xTF32 tcgen05 GEMM.  Default: the conv kernel
    // emits the GEMM's A operand pre-split and pre-tiled (all-TMA GEMM); MAPF_B200_FC_UNSPLIT=1 keeps the plain
    // [rows, fc_in] hand-over with conversion warps in the GEMM (development switch).
    mapf_encoder_fwd_args e{rows, a->states, a->packed, a->workspace_act};
    const bool unsplit = (a->flags & MAPF_POLICY_FC_UNSPLIT) != 0;      // development switches of the caller
    const bool ffma_conv = (a->flags & MAPF_POLICY_FFMA_CONV) != 0;
    if (!ffma_conv && !unsplit && mapf_enc_tc_dims_ok(*p)) {
      // conv stack on tcgen05 (activations as the A operand in tensor memory, enc_tc.cu) -> FC operand in pixel-major
      // K order -> all-TMA tcgen05 FC against the matching B operand
      rc = mapf_encoder_tc_conv_fwd(p, &e, stream);
      if (rc != MAPF_OK) return rc;
      rc = mapf_tc_fc_presplit_f16(int32_t(rows), p->fc_out, p->fc_in, a->workspace_act, a->packed + L.tc_fc_pm,
                                   a->workspace_x, p->fc_out, a->packed + L.fc_b, 1, stream);
    } else if (unsplit) {
      rc = mapf_encoder_conv_fwd(p, &e, stream);
      if (rc != MAPF_OK) return rc;
      rc = mapf_tc_gemm(int32_t(rows), p->fc_out, p->fc_in, a->workspace_act, p->fc_in, a->packed + L.tc_fc,
                        a->workspace_x, p->fc_out, a->packed + L.fc_b, 1, stream);
    } else {
      rc = mapf_encoder_conv_split_fwd(p, &e, stream);
      if (rc != MAPF_OK) return rc;
      rc = mapf_tc_fc_presplit(int32_t(rows), p->fc_out, p->fc_in, a->workspace_act, a->packed + L.tc_fc, a->workspace_x,
                               p->fc_out, a->packed + L.fc_b, 1, stream);
    }
  } else {
    mapf_encoder_fwd_args e{rows, a->states, a->packed, a->workspace_x};
    rc = mapf_encoder_fwd(p, &e, stream);
  }
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
  const bool tensor_path = a->workspace_z != nullptr && p->node_dim <= 128 && (p->node_dim & 15) == 0 &&
                           (p->fc_out & 3) == 0 && a->agents <= 64;
  if (!tensor_path) return mapf_graph_filter_fwd(&g, stream);   // fused fp32-FFMA kernel
  // The fused kernel runs the hops inside its 4 producer warps: right for small teams, too serial for N = 64 where the
  // hops are 37 % of the MACs (measured: C5 987 us unfused vs 1415 us fused; C2a 24.7 vs 18.5 us).
  if (a->agents <= 32 && mapf_tc_graph_supported(a->agents, p->fc_out, p->node_dim, p->taps) &&
      !(a->flags & MAPF_POLICY_UNFUSED_GRAPH)) {
    // fused tensor-core path: S, hops, 3xTF32 tcgen05 mix, ReLU, head, argmax in one kernel
    // gso is an input of this launch sequence: complete before the encoder kernels above were launched
    // Fusing the env step into the graph kernel's tail saves a launch and its prologue: worth it for small launches
    // (a host-stepped episode group: end to end 1.78e8 -> 1.87e8 agent-steps/s at C2a), not for a full resident batch, where
    // the separate env kernel's many small CTAs finish sooner (2.32e8 against 2.28e8 fused).  MAPF_POLICY_FUSED_ENV_ON / _OFF
    // force either.
    const bool want_fused = (a->flags & MAPF_POLICY_FUSED_ENV_ON) ? true
                            : (a->flags & MAPF_POLICY_FUSED_ENV_OFF) ? false : rows <= 12288;
    const bool fuse_env = want_fused && env != nullptr && a->actions != nullptr && env->actions == a->actions &&
                          env->episodes == a->batch && env->agents == a->agents &&
                          mapf_tc_graph_env_fusable(a->agents, env->cell_stride);
    rc = mapf_tc_graph_head_ex(a->batch, a->agents, p->fc_out, p->node_dim, p->taps, a->message ? 0 : 1,
                               a->message ? 1 : 0, a->workspace_x, a->gso, a->packed + L.tc_gf, a->packed + L.act_w,
                               a->packed + L.act_b, p->num_actions, a->logits, a->actions, 1, fuse_env ? env : nullptr,
                               stream);
    if (rc == MAPF_OK && fuse_env) *env_done = 1;
    return rc;
  }
  // two-kernel tensor-core path: filter taps -> Z (tiny kernel), then 3xTF32 tcgen05 mix + ReLU + head + argmax
  g.z = a->workspace_z;
  rc = (a->flags & MAPF_POLICY_FFMA_HOPS) ? mapf_graph_taps_ffma(&g, stream) : mapf_graph_taps(&g, stream);
  if (rc != MAPF_OK) return rc;
  const int KT = (p->taps + (a->message ? 1 : 0)) * p->fc_out;
  return mapf_tc_mix_head(int32_t(rows), p->node_dim, KT, a->workspace_z, KT, a->packed + L.tc_gf, a->packed + L.act_w,
                          a->packed + L.act_b, p->num_actions, a->logits, a->actions, nullptr, stream);
}

extern "C" int mapf_rollout(const mapf_net_params* p, const mapf_rollout_args* a, mapf_stream_t stream) {
  MAPF_REQUIRE(a != nullptr, MAPF_ERR_NULL);
  MAPF_REQUIRE(a->steps >= 0, MAPF_ERR_SHAPE);
  MAPF_REQUIRE(a->policy.actions != nullptr && a->env.actions == a->policy.actions, MAPF_ERR_NULL);
  MAPF_REQUIRE(a->env.fov == a->policy.states && (p->taps == 0 || a->env.adj == a->policy.gso), MAPF_ERR_SHAPE);
  MAPF_REQUIRE(a->env.episodes == a->policy.batch && a->env.agents == a->policy.agents, MAPF_ERR_SHAPE);
  const MapfPdlScope pdl_scope((a->policy.flags & MAPF_POLICY_NO_PDL) != 0);   // covers the env step launches too
  for (int t = 0; t < a->steps; ++t) {
    mapf_env_step_args e = a->env;
    e.do_move = 1 | 4;   // 4: boards / positions are older than the policy kernels in front (see env_step_kernel)
    int env_done = 0;    // set when the env step ran in the tail of the fused graph kernel
    int rc = policy_fwd_impl(p, &a->policy, &e, &env_done, stream);
    if (rc != MAPF_OK) return rc;
    if (!env_done) {
      rc = mapf_env_step(&e, stream);
      if (rc != MAPF_OK) return rc;
    }
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

extern "C" int mapf_graph_launch(void* graph_exec, mapf_stream_t stream, void* event) {
  MAPF_REQUIRE(graph_exec, MAPF_ERR_NULL);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (cudaGraphLaunch(static_cast<cudaGraphExec_t>(graph_exec), s) != cudaSuccess) return MAPF_ERR_CUDA;
  if (event && cudaEventRecord(static_cast<cudaEvent_t>(event), s) != cudaSuccess) return MAPF_ERR_CUDA;
  return MAPF_OK;
}

extern "C" int mapf_event_wait(void* event) {
  MAPF_REQUIRE(event, MAPF_ERR_NULL);
  return cudaEventSynchronize(static_cast<cudaEvent_t>(event)) == cudaSuccess ? MAPF_OK : MAPF_ERR_CUDA;
}
