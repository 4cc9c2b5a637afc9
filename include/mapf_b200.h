/* mapf_b200.h -- C ABI of the B200-native MAPF-GNN batched decentralized-policy step.
 *
 * One shared library (libmapf_b200.so, hand-written sm_100a CUDA) sits under the Python
 * facade classes (GraphEnv / BatchedGraphEnv, GCNLayer, MessagePassingLayer, Network) and the
 * `mapf::` torch custom ops.  Every entry point
 *   - takes one plain-C argument struct of DEVICE pointers + int sizes and a cudaStream_t,
 *   - is asynchronous on that stream, never allocates, never synchronises, never throws,
 *   - keeps no global mutable state apart from an idempotent per-device cache of "dynamic shared memory already
 *     granted" flags (re-entrant; thread-safe per stream) and never reads the process environment: development
 *     switches are explicit `flags` fields of the argument structs,
 *   - returns MAPF_OK (0) or a negative mapf_status code; mapf_last_error_string(code)
 *     gives the message.  All buffers are caller-owned.
 *
 * The reference (RetamalVictor/MAPF-GNN) is pure Python and has no FFI seam; each entry point
 * cites the reference Python interface it replaces (file:line under the reference root).
 *
 * Layout conventions (all row-major, contiguous):
 *   pos / goal     int32 [E,N,2]   (x, y) per agent                     env_graph_gridv1.py:75-90
 *   cells          uint8 [E,cell_stride]  board[y*H+x]; low 2 bits = board value 0 empty /
 *                  1 agent / 2 obstacle (env_graph_gridv1.py:68-70,273-275); bit 7 = cell is in
 *                  the (immutable) obstacle set used by check_collision_obstacle (:303-313).
 *                  cell_stride >= H*H and a multiple of 16.
 *   fov            f32 [E,N,2,5,5]                                       env_graph_gridv1.py:248-265
 *   adj            f32 [E,N,N]    0/1, row-pruned closest-4              env_graph_gridv1.py:117-136,174-181
 *   states         f32 [B,N,2,5,5] = fov; gso f32 [B,N,N] = adj (+I from the data loader)
 *   node features  f32 [B*N, F] "agent-major" rows (row = b*N+n) inside the fused net; the
 *                  standalone layer entry point also accepts the reference's [B,F,N] via strides.
 */
#ifndef MAPF_B200_H
#define MAPF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* mapf_stream_t; /* cudaStream_t */

typedef enum {
  MAPF_OK = 0,
  MAPF_ERR_NULL = -1,        /* a required pointer is NULL */
  MAPF_ERR_SHAPE = -2,       /* a size is non-positive or inconsistent */
  MAPF_ERR_UNSUPPORTED = -3, /* size outside what the kernels are built for */
  MAPF_ERR_ALIGN = -4,       /* pointer / stride alignment requirement violated */
  MAPF_ERR_CUDA = -5         /* the CUDA runtime reported a launch error */
} mapf_status;

const char* mapf_last_error_string(int code);
/* library / build identification: "mapf_b200 <version> sm_100a" */
const char* mapf_version(void);

/* Host-stepping helpers for callers that drive the evaluate.py:223-241 loop from the host with a captured launch
 * sequence (HostSteppedRollout): launch an instantiated cudaGraphExec_t on `stream` and record `event` (may be NULL)
 * behind it; block the calling thread until `event` has completed.  Two plain driver calls each -- they exist so that
 * the per-step host cost is not dominated by an interpreter's stream / graph wrappers. */
int mapf_graph_launch(void* graph_exec, mapf_stream_t stream, void* event);
int mapf_event_wait(void* event);

/* ------------------------------------------------------------------------------------------
 * Environment (reference: grid/env_graph_gridv1.py GraphEnv)
 * ---------------------------------------------------------------------------------------- */

/* GraphEnv.reset (env_graph_gridv1.py:65-95) for E episodes with explicit starts:
 * cells <- obstacles only (value 2 | obstacle-set flag), pos <- starts, time <- 0, done <- 0.
 * obstacles may be NULL when num_obstacles == 0.  Cells outside the board are ignored. */
typedef struct {
  int32_t episodes, agents, board, num_obstacles;
  int32_t cell_stride;
  const int32_t* starts;    /* [E,N,2] */
  const int32_t* obstacles; /* [E,M,2] */
  int32_t* pos;             /* [E,N,2] out */
  uint8_t* cells;           /* [E,cell_stride] out */
  int32_t* time;            /* [E] out */
  uint8_t* done;            /* [E] out */
} mapf_env_reset_args;
int mapf_env_reset(const mapf_env_reset_args* a, mapf_stream_t stream);

/* On-device scenario generation for E episodes: M unique obstacles, then N unique obstacle-free goals
 * (evaluate.py:207-208: create_obstacles env_graph_gridv1.py:504-529, create_goals :465-501), then N unique
 * obstacle-free starts.  Deterministic in (seed, episode index); distributional (not stream-exact) parity with the
 * reference's numpy generator.  MAPF_ERR_SHAPE when M + N exceeds the number of cells (create_goals' ValueError). */
int mapf_make_scenarios(int32_t episodes, int32_t agents, int32_t board, int32_t num_obstacles, uint64_t seed,
                        int32_t* starts, int32_t* goals, int32_t* obstacles, mapf_stream_t stream);

/* GraphEnv.step (env_graph_gridv1.py:183-198) = _updatePositions (:200-213) -> _computeDistance +
 * _computeClosest (:117-136,174-181) -> preprocessObs (:248-265) -> time += 1 -> checkAllInGoal
 * (:156-159), for E episodes in one fused kernel.  do_move == 0 gives getObservations() of the
 * current state without moving (the observation reset() returns; actions may then be NULL).
 * Episodes whose done flag is already set are frozen (callers of the reference `break`,
 * evaluate.py:251): no move, no time increment, observation recomputed.  do_move == 3 steps finished episodes too
 * (the reference env itself never freezes; data_generation/record.py:78-88 keeps stepping).
 * do_move bit 2 (value 4) is a scheduling hint set by mapf_rollout: boards / positions / goals / done flags were last
 * written BEFORE the kernels launched in front of this call started (only `actions` comes from the kernel right in
 * front), so the kernel may stage them before its grid-dependency wait.  Never set it when that is not guaranteed.
 * d2_limit: smallest integer d2 with sqrt((double)d2) >= sensing_range. */
typedef struct {
  int32_t episodes, agents, board;
  int32_t cell_stride;
  int32_t d2_limit;
  int32_t do_move;
  int32_t* pos;           /* [E,N,2] in/out */
  const int32_t* goal;    /* [E,N,2] */
  const int32_t* actions; /* [E,N] 0 idle,1 +x,2 +y,3 -x,4 -y (:42-48) */
  uint8_t* cells;         /* [E,cell_stride] in/out */
  int32_t* time;          /* [E] in/out */
  uint8_t* done;          /* [E] in/out */
  float* fov;             /* [E,N,2,5,5] out (may be NULL) */
  float* adj;             /* [E,N,N] out (may be NULL) */
  int32_t episodes_per_cta; /* work decomposition: 0 = chosen by the library; else 1 .. 128 / next_pow2(N).  Results are
                             * identical for every value (tests sweep it). */
  int32_t* pos_out;       /* NULL: positions are updated in place.  Else [E,N,2]: the new positions of EVERY agent go here and
                           * `pos` is only read -- for callers that keep the state in two alternating (host-mapped) blocks */
  uint8_t* done_out;      /* NULL: in place.  Else [E]: the done flags after this step (frozen episodes keep theirs) */
} mapf_env_step_args;
int mapf_env_step(const mapf_env_step_args* a, mapf_stream_t stream);

/* The single rules of GraphEnv._updatePositions on caller-supplied candidate positions, for callers (and the
 * reference's tests/test_env.py:97-121,164-194) that poke them one at a time.  stages, applied in the reference's order:
 * bit 0 check_collision_obstacle (env_graph_gridv1.py:303-313: a candidate inside the obstacle set reverts to pos_temp),
 * bit 1 check_boundary (:267-271: clamp to the board), bit 2 check_collisions (:282-301: every member of a shared cell
 * reverts to pos_temp, single pass).  The board is not touched (updateBoard :273-275 is part of mapf_env_step). */
typedef struct {
  int32_t episodes, agents, board, cell_stride;
  int32_t stages;
  int32_t* pos;            /* [E,N,2] in/out: candidate positions (positionX / positionY) */
  const int32_t* pos_temp; /* [E,N,2] positions to revert to (positionX_temp / positionY_temp) */
  const uint8_t* cells;    /* [E,cell_stride] (obstacle-set flag); may be NULL when bit 0 is clear */
} mapf_env_rules_args;
int mapf_env_apply_rules(const mapf_env_rules_args* a, mapf_stream_t stream);

/* GraphEnv.computeMetrics (env_graph_gridv1.py:138-154,168-172): per episode
 * success = (#agents on own goal)/N, flow_time = time if all on goal else N*max_time. */
typedef struct {
  int32_t episodes, agents, max_time;
  const int32_t* pos;
  const int32_t* goal;
  const int32_t* time;
  float* success;     /* [E] out */
  int32_t* flow_time; /* [E] out */
} mapf_env_metrics_args;
int mapf_env_metrics(const mapf_env_metrics_args* a, mapf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Policy network (reference: models/framework_gnn.py, framework_gnn_message.py,
 * framework_baseline.py, models/networks/gnn.py)
 * ---------------------------------------------------------------------------------------- */

#define MAPF_MAX_CONV 4

/* Raw parameters in the reference state_dict layout.  conv_w[l] is
 * convLayers.{3l}.weight [Cout,Cin,3,3]; bn_* are convLayers.{3l+1}.*; fc_* is compressMLP.0;
 * gf_w is GFL.0.W [F,K,G]; gf_w2 is GFL.0.W2 [F,G] (message variant) or NULL; act_* is
 * actionMLP.0 [A, G or F].  gf_w == NULL selects the no-communication baseline. */
typedef struct {
  int32_t num_conv;                   /* 1..MAPF_MAX_CONV */
  int32_t channels[MAPF_MAX_CONV + 1];/* channels[0] = 2 input planes */
  int32_t fc_in, fc_out;              /* 25*channels[num_conv], F */
  int32_t taps, node_dim;             /* K, G (0 when baseline) */
  int32_t num_actions;                /* 5 */
  const float* conv_w[MAPF_MAX_CONV];
  const float* conv_b[MAPF_MAX_CONV];
  const float* bn_gamma[MAPF_MAX_CONV];
  const float* bn_beta[MAPF_MAX_CONV];
  const float* bn_mean[MAPF_MAX_CONV];
  const float* bn_var[MAPF_MAX_CONV];
  const float* fc_w;  /* [F, fc_in] */
  const float* fc_b;  /* [F] */
  const float* gf_w;  /* [F,K,G] or NULL */
  const float* gf_w2; /* [F,G] or NULL */
  const float* act_w; /* [A, G or F] */
  const float* act_b; /* [A] */
} mapf_net_params;

/* Number of floats mapf_pack_weights writes for these dimensions (pointers are ignored). */
int64_t mapf_packed_weights_size(const mapf_net_params* p);

/* Inference-time weight preparation (one tiny kernel): folds eval-mode BatchNorm
 * (framework_gnn.py:70, eps 1e-5) into the conv weights, regroups them per 4-output-channel
 * group, transposes the encoder FC, and materialises W_eff^T = GFL.0.W.reshape(G,K*F)^T with the
 * W2.reshape(G,F)^T rows appended (gnn.py:114-116,222 raw reinterpretations). */
int mapf_pack_weights(const mapf_net_params* p, float* packed, mapf_stream_t stream);

/* A5 eval: encoder of Network.forward (framework_gnn.py:156-166): per agent image
 * 3x{conv3x3+BN+ReLU} -> flatten (C,H,W) -> Linear+ReLU.  x row r = b*N+n. */
typedef struct {
  int64_t rows;          /* B*N agent images */
  const float* states;   /* [rows,2,5,5] */
  const float* packed;   /* from mapf_pack_weights */
  float* x;              /* [rows,F] out */
} mapf_encoder_fwd_args;
int mapf_encoder_fwd(const mapf_net_params* p, const mapf_encoder_fwd_args* a, mapf_stream_t stream);
/* Conv stack only: a->x receives the flattened (C,H,W) activations [rows, fc_in] (framework_gnn.py:160-163); the
 * Linear + ReLU that follows is mapf_tc_gemm on the tensor cores. */
int mapf_encoder_conv_fwd(const mapf_net_params* p, const mapf_encoder_fwd_args* a, mapf_stream_t stream);

/* A6/A7 (+A8): GCNLayer.forward (gnn.py:72-126) / MessagePassingLayer.forward (gnn.py:178-235),
 * optionally fused with the ReLU + actionMLP + argmax tail of Network.forward
 * (framework_gnn.py:171-181; np.argmax first-max tie-break evaluate.py:238).
 *   S = D^-1/2 (A [+ I]) D^-1/2 (row-sum degrees, inf -> 0); x_k = x_{k-1} S;
 *   Y[n,g] = sum_{k,f} x_k[f,n] * W.reshape(G,K*F)[g, k*F+f]  (+ skip: X.reshape(B,N,F) W2.reshape(G,F)^T).
 * Element (b,n,f) of x is at x[b*x_sb + n*x_sn + f*x_sf] (agent-major: N*F, F, 1; reference
 * [B,F,N]: F*N, 1, N); likewise y with (y_sb, y_sn, y_sg).  wt is the packed
 * [(K(+1))*F, G] matrix (mapf_pack_weights / mapf_pack_graph_weights). */
typedef struct {
  int32_t batch, agents, in_dim, out_dim, taps;
  int32_t add_self_loop;  /* 1: GCNLayer (gnn.py:87); 0: MessagePassingLayer */
  int32_t has_skip;       /* 1: W2 rows appended to wt (message variant) */
  int32_t relu;           /* apply ReLU to Y before storing / before the head */
  int32_t num_actions;    /* >0: fuse the action head */
  const float* x; int64_t x_sb, x_sn, x_sf;
  const float* gso;       /* [B,N,N] */
  const float* wt;        /* [(K+has_skip)*F, G] */
  float* y; int64_t y_sb, y_sn, y_sg; /* may be NULL when the head is fused */
  float* z;               /* optional [B*N, (K+has_skip)*F] saved filter taps (training) or NULL */
  const float* act_w;     /* [A,G] */
  const float* act_b;     /* [A] */
  float* logits;          /* [B*N, A] out */
  int32_t* actions;       /* [B*N] out, may be NULL */
} mapf_graph_filter_fwd_args;
int mapf_graph_filter_fwd(const mapf_graph_filter_fwd_args* a, mapf_stream_t stream);

/* Packs a standalone layer's W [F,K,G] (+ W2 [F,G]) into wt [(K(+1))*F, G]. */
int mapf_pack_graph_weights(const float* w, const float* w2, int32_t in_dim, int32_t taps, int32_t out_dim,
                            float* wt, mapf_stream_t stream);

/* A9 head of the baseline: Linear(F->A) on the encoder output + argmax
 * (framework_baseline.py:96,126). */
typedef struct {
  int64_t rows; int32_t in_dim, num_actions;
  const float* x; const float* act_w; const float* act_b;
  float* logits; int32_t* actions;
} mapf_head_fwd_args;
int mapf_head_fwd(const mapf_head_fwd_args* a, mapf_stream_t stream);

/* Whole eval-mode Network.forward(states, gso) -> logits (+ greedy actions):
 * encoder -> graph filter + ReLU -> head, or encoder -> head for the baseline.
 * workspace_x: [B*N, F] floats. */
typedef struct {
  int32_t batch, agents;
  int32_t message;        /* 1: framework_gnn_message.Network, 0: framework_gnn.Network */
  const float* states;    /* [B,N,2,5,5] */
  const float* gso;       /* [B,N,N] (ignored for the baseline) */
  const float* packed;
  float* workspace_x;
  float* workspace_z;     /* [B*N, (K+1)*F] floats: enables the tensor-core graph-filter path; NULL = fused FFMA kernel */
  float* workspace_act;   /* mapf_tc_presplit_a_size(B*N, fc_in) floats, 128-byte aligned: enables the tensor-core encoder
                           * FC; NULL = FC fused in the conv kernel */
  float* logits;          /* [B,N,A] */
  int32_t* actions;       /* [B,N] or NULL */
  int32_t flags;          /* 0 = the library's choices; development switches MAPF_POLICY_* below (every combination gives
                           * the same logits within the stated tolerance and the same actions; tests sweep them) */
} mapf_policy_fwd_args;
#define MAPF_POLICY_FFMA_CONV 1      /* conv stack on the FFMA2 kernel even where the tensor-core stack applies */
#define MAPF_POLICY_FC_UNSPLIT 2     /* plain [rows, fc_in] hand-over to the encoder FC (conversion warps in the GEMM) */
#define MAPF_POLICY_UNFUSED_GRAPH 4  /* graph filter as taps kernel + tensor-core mix instead of the fused kernel */
#define MAPF_POLICY_FUSED_ENV_ON 8   /* mapf_rollout: always run the env step in the tail of the graph kernel */
#define MAPF_POLICY_FUSED_ENV_OFF 16 /* mapf_rollout: never */
#define MAPF_POLICY_NO_PDL 32        /* launch the call's kernels without programmatic stream serialization */
#define MAPF_POLICY_FFMA_HOPS 64     /* 64-agent teams: filter taps on the FFMA kernel instead of the tensor-core hops */
int mapf_policy_fwd(const mapf_net_params* p, const mapf_policy_fwd_args* a, mapf_stream_t stream);

/* `steps` iterations of the rollout protocol of evaluate.py:223-241 for E resident episodes,
 * entirely on the device: policy(fov, adj) -> argmax -> env step -> (fov, adj).  fov / adj must
 * hold the current observation on entry (mapf_env_step with do_move = 0) and hold the last one on
 * exit.  No host round trip, no collective. */
typedef struct {
  mapf_env_step_args env;     /* env.actions must point at `actions` below */
  mapf_policy_fwd_args policy;/* policy.states == env.fov, policy.gso == env.adj, policy.actions == env.actions */
  int32_t steps;
} mapf_rollout_args;
int mapf_rollout(const mapf_net_params* p, const mapf_rollout_args* a, mapf_stream_t stream);


/* ------------------------------------------------------------------------------------------
 * Imitation-training step (reference: train.py:82-100)
 * ---------------------------------------------------------------------------------------- */

/* Gradients, one buffer per parameter tensor, shapes as in mapf_net_params. */
typedef struct {
  float* conv_w[MAPF_MAX_CONV];
  float* conv_b[MAPF_MAX_CONV];
  float* bn_gamma[MAPF_MAX_CONV];
  float* bn_beta[MAPF_MAX_CONV];
  float* fc_w;
  float* fc_b;
  float* gf_w;  /* NULL for the baseline */
  float* gf_w2; /* NULL unless message variant */
  float* act_w;
  float* act_b;
} mapf_net_grads;

/* Floats of caller-owned scratch the train-mode forward/backward pair needs for a [B,N] batch
 * (saved activations, BatchNorm statistics, packed weights, gradient temporaries). -1 if unsupported. */
int64_t mapf_train_workspace_size(const mapf_net_params* p, int32_t batch, int32_t agents);

/* Train-mode Network.forward (train.py:87): BatchNorm uses batch statistics computed PER AGENT SLOT
 * over B*25 values (the reference calls the conv stack once per agent, framework_gnn.py:160-163) and
 * the running statistics receive N sequential momentum updates (p->bn_mean / p->bn_var are written
 * through; bn_batches[l] += N when non-NULL).  gso is used as given (the loader's A + I,
 * data_loader.py:74); GCNLayer adds its own I. */
typedef struct {
  int32_t batch, agents;
  int32_t message;
  float momentum;            /* 0.1 */
  const float* states;       /* [B,N,2,5,5] */
  const float* gso;          /* [B,N,N] (ignored for the baseline) */
  float* workspace;          /* mapf_train_workspace_size floats, 16-byte aligned */
  float* logits;             /* [B,N,A] out */
  int64_t* bn_batches[MAPF_MAX_CONV];
  int32_t flags;             /* development switches MAPF_TRAIN_* below; 0 = tensor-core FC and graph mix */
} mapf_train_fwd_args;
#define MAPF_TRAIN_FFMA_FC 1     /* compressMLP forward on the FFMA GEMM */
#define MAPF_TRAIN_FFMA_GRAPH 2  /* graph filter forward on the fused FFMA kernel */
#define MAPF_TRAIN_FFMA_BWD 4    /* backward GEMMs on the FFMA2 kernels instead of mapf_tc_gemm_tn / mapf_tc_gemm */
#define MAPF_TRAIN_FFMA_CONV 16   /* forward / data-gradient convolutions on the FFMA2 kernel instead of the tensor cores */
#define MAPF_TRAIN_TC_BWD_ALWAYS 8 /* accepted, no effect: the tensor-core paths are the default at every batch size */
#define MAPF_TRAIN_SINGLE_LANE 32 /* every launch on the caller's stream: no auxiliary lanes for weight packs / weight gradients */
#define MAPF_TRAIN_PREPARED 128   /* mapf_train_forward: also prepare what mapf_train_backward needs from the parameters alone
                                     (transposed-weight packs, mirrored filter operands, zeroed BatchNorm sums) on the pack lane;
                                     mapf_train_backward: the forward call on this workspace did so (same parameters, same other
                                     flags).  Without it each call is self-contained. */
#define MAPF_TRAIN_BN_PASS 64     /* BatchNorm + ReLU of every conv layer as its own pass (default: applied on load by the next
                                     layer's tensor-core kernel, which also writes the activations the backward pass reads) */
int mapf_train_forward(const mapf_net_params* p, const mapf_train_fwd_args* a, mapf_stream_t stream);

/* Backward of the above from d(loss)/d(logits); must follow mapf_train_forward on the same
 * workspace.  Every gradient buffer is overwritten. */
typedef struct {
  int32_t batch, agents;
  int32_t message;
  const float* states;
  const float* gso;
  float* workspace;
  const float* dlogits;      /* [B,N,A] */
  mapf_net_grads grads;
  int32_t flags;             /* development switches MAPF_TRAIN_* ; 0 = tensor-core GEMMs */
} mapf_train_bwd_args;
int mapf_train_backward(const mapf_net_params* p, const mapf_train_bwd_args* a, mapf_stream_t stream);

/* nn.CrossEntropyLoss (mean) over rows = B*N (train.py:66,90-93): loss scalar and d loss / d logits. */
typedef struct {
  int64_t rows;
  int32_t num_actions;
  const float* logits;    /* [rows,A] */
  const int32_t* targets; /* [rows] */
  float* loss;            /* [1] out */
  float* dlogits;         /* [rows,A] out */
} mapf_ce_loss_args;
int mapf_ce_loss(const mapf_ce_loss_args* a, mapf_stream_t stream);

/* clip_grad_norm_(max_norm) followed by Adam with L2 weight decay (train.py:64,97,100; torch.optim.Adam
 * defaults) over flat fp32 buffers; `step` is the 1-based step count.  grad_norm receives the
 * pre-clip global L2 norm; scratch is one double owned by the caller. */
typedef struct {
  int64_t n;
  float* params;
  float* grads; /* scaled in place by the clip coefficient */
  float* exp_avg;
  float* exp_avg_sq;
  float lr, beta1, beta2, eps, weight_decay, max_norm;
  int32_t step;
  float* grad_norm; /* [1] out */
  double* scratch;  /* [1] */
  /* optional device-resident hyper-state, so that a captured CUDA graph can be replayed unchanged every step:
   * step_dev (int32, incremented by the call and used instead of `step`), lr_dev (float, used instead of `lr`) */
  int32_t* step_dev;
  const float* lr_dev;
} mapf_clip_adam_args;
int mapf_clip_adam(const mapf_clip_adam_args* a, mapf_stream_t stream);

/* Multi-GPU training step (the reference is single-process, train.py:96-100 is the step body this extends):
 * one-shot all-reduce (mean) of the flat gradient over NVLink PEER MEMORY fused with the gradient-norm reduction, followed
 * by the same clip + Adam kernel as mapf_clip_adam -- two launches, no NCCL call, capturable in a CUDA graph.
 * Every rank passes peer-mapped pointers to all ranks' published gradient buffers of this step and to all ranks' flag
 * blocks (int32[16], zero-initialised: [0..7] arrival epochs written by the peers, [8] the owner's epoch counter).  The
 * caller alternates between two gradient buffers per rank from step to step (see train_kernels.cu).  a->grads receives
 * the averaged gradient (then scaled by the clip coefficient, as in mapf_clip_adam).  A peer that never arrives traps the
 * kernel after ~10 s instead of hanging the GPU. */
typedef struct {
  int32_t world, rank;          /* world <= 8 */
  const float* peer_grads[8];   /* [world] a->n floats each, 16-byte aligned; [rank] = this rank's own buffer */
  int32_t* peer_flags[8];       /* [world] flag blocks; [rank] = this rank's own */
} mapf_peer_reduce_args;
int mapf_peer_reduce_clip_adam(const mapf_peer_reduce_args* r, const mapf_clip_adam_args* a, mapf_stream_t stream);
/* Peer-mapped device memory by CUDA IPC -- the only entry points that allocate: mapf_peer_alloc returns zeroed device
 * memory and its 64-byte IPC handle (to be sent to the other ranks, e.g. with an all-gather); mapf_peer_open maps another
 * process' handle into this process (peer access enabled lazily); close / free undo them. */
int mapf_peer_alloc(int64_t bytes, void** ptr, uint8_t* handle);
int mapf_peer_open(const uint8_t* handle, void** ptr);
int mapf_peer_close(void* ptr);
int mapf_peer_free(void* ptr);

/* Backward of the standalone GCNLayer / MessagePassingLayer (gnn.py:72-126,178-235) in the reference
 * layouts: x [B,F,N], dy [B,G,N] -> dx [B,F,N], dw [F,K,G], dw2 [F,G] (w2/dw2 NULL for GCNLayer).
 * workspace: mapf_graph_filter_bwd_workspace_size(B, N, F, G, K, has_skip) floats. */
typedef struct {
  int32_t batch, agents, in_dim, out_dim, taps;
  int32_t add_self_loop;
  const float* x;
  const float* gso;
  const float* w;
  const float* w2;
  const float* dy;
  float* workspace;
  float* dx;
  float* dw;
  float* dw2;
} mapf_graph_filter_bwd_args;
int64_t mapf_graph_filter_bwd_workspace_size(int32_t batch, int32_t agents, int32_t in_dim, int32_t out_dim,
                                             int32_t taps, int32_t has_skip);
int mapf_graph_filter_bwd(const mapf_graph_filter_bwd_args* a, mapf_stream_t stream);

/* fp32-accurate tensor-core GEMMs (3xTF32 on tcgen05, accumulators in TMEM).
 * mapf_tc_pack_b: B [N,K0] (optionally [N,K1] appended along K), row-major fp32 with K contiguous, N <= 128 ->
 *   TF32 hi / lo parts tiled in the canonical UMMA shared-memory layout, one contiguous block per 16-wide K chunk
 *   (what the kernel's TMA bulk copies fetch); out holds mapf_tc_packed_b_size(N, K0+K1) floats.
 * mapf_tc_gemm: C[M,N] = act(A[M,K] B^T + bias[N]); A row-major fp32, lda multiple of 4, 16-byte aligned.  The
 *   building block behind the encoder FC (compressMLP, framework_gnn.py:93).
 * mapf_tc_mix_head: logits / greedy actions = actionMLP(relu(Z [W | W2]^T)) with Z [M,K] the row-major filter taps
 *   of mapf_graph_taps and the packed B built from GFL.0.W viewed as [G, K*F] and GFL.0.W2 viewed as [G, F] (the raw
 *   reshapes of gnn.py:114-116,222); y (optional) receives relu(...) [M,G]. */
int64_t mapf_tc_packed_b_size(int32_t N, int32_t K);
int mapf_tc_pack_b(const float* b0, int32_t K0, const float* b1, int32_t K1, int32_t N, float* out,
                   mapf_stream_t stream);
int mapf_tc_gemm(int32_t M, int32_t N, int32_t K, const float* A, int64_t lda, const float* Bp, float* C, int64_t ldc,
                 const float* bias, int32_t relu, mapf_stream_t stream);
/* same contract, work decomposition for the training step (4 stages, one accumulator, three CTAs per SM) */
int mapf_tc_gemm_w3(int32_t M, int32_t N, int32_t K, const float* A, int64_t lda, const float* Bp, float* C, int64_t ldc,
                    const float* bias, int32_t relu, mapf_stream_t stream);
/* wide outputs (any N): B packed as ceil(N / 128) tiles of 128 rows, mapf_tc_packed_b_size(128, K) floats apart, and all
 * tiles computed by one launch (the training step's activation gradients dz = dy W_eff, da = dx W_fc) */
int mapf_tc_pack_b_tiles(const float* b, int32_t K, int32_t N, float* out, mapf_stream_t stream);
int mapf_tc_gemm_tiles(int32_t M, int32_t N, int32_t K, const float* A, int64_t lda, const float* Bp, float* C, int64_t ldc,
                       const float* bias, int32_t relu, mapf_stream_t stream);
/* Weight-gradient GEMM of the training step (train.py:96): C (+)= A^T B with A [K, M] (lda) and B [K, N] (ldb) row-major
 * fp32 activation matrices whose contraction index is the row index (K = B*N rows), 3xTF32 on the tensor cores with the
 * transposition done on the way into shared memory and split-K partials meeting in C through red.global.add.
 * transpose_c != 0 writes C^T [N, M] (ldc >= M).  accumulate == 0 zeroes C first. */
int mapf_tc_gemm_tn(int32_t M, int32_t N, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
                    int64_t ldc, int32_t transpose_c, int32_t accumulate, mapf_stream_t stream);
/* Encoder FC with the A operand pre-split by the conv kernel: mapf_encoder_conv_split_fwd writes the flattened conv
 * output (framework_gnn.py:160-163) as TF32 hi / lo halves in the tensor-core tile order (mapf_tc_presplit_a_size
 * floats), mapf_tc_fc_presplit computes relu(A W^T + b) (compressMLP, framework_gnn.py:93) from it with an all-TMA
 * operand pipeline. */
int64_t mapf_tc_presplit_a_size(int64_t M, int32_t K);
int mapf_encoder_conv_split_fwd(const mapf_net_params* p, const mapf_encoder_fwd_args* a, mapf_stream_t stream);
int mapf_tc_fc_presplit(int32_t M, int32_t N, int32_t K, const float* Ap, const float* Bp, float* C, int64_t ldc,
                        const float* bias, int32_t relu, mapf_stream_t stream);
/* Conv stack of the encoder on the tensor cores (framework_gnn.py:55-71,160-163; first layer 16 or 32 channels, 16
 * behind it): each layer is one 128 x 48 x 48 GEMM per 4 agent images on fp16-split operands (3 kind::f16 MMAs per
 * 16-wide K step; horizontal taps in K, filter rows in N, activations as the A operand in tensor memory) followed by
 * the vertical shifted adds in the epilogue (csrc/enc_tc.cu).  a->x (mapf_tc_presplit_a_size floats) receives the FC operand
 * fp16-split (x * 2^-e = hi + lo, one scale per image) with its K index ordered pixel-major, followed by the rows' inverse
 * scales; mapf_tc_fc_presplit_f16 computes relu(A W^T + b) (compressMLP, framework_gnn.py:93) from it and the B operand
 * mapf_tc_pack_b_f16_pixel_major builds from compressMLP.0.weight (mapf_pack_weights does this).  mapf_policy_fwd
 * takes this path when workspace_act is given and the dimensions fit. */
int mapf_encoder_tc_supported(const mapf_net_params* p);
int mapf_encoder_tc_conv_fwd(const mapf_net_params* p, const mapf_encoder_fwd_args* a, mapf_stream_t stream);
int64_t mapf_tc_packed_b_f16_size(int32_t N, int32_t K);
int mapf_tc_pack_b_f16_pixel_major(const float* b, int32_t C, int32_t N, float* out, mapf_stream_t stream);
int mapf_tc_fc_presplit_f16(int32_t M, int32_t N, int32_t K, const float* Ap, const float* Bp, float* C, int64_t ldc,
                            const float* bias, int32_t relu, mapf_stream_t stream);
int mapf_tc_mix_head(int32_t M, int32_t N, int32_t K, const float* Z, int64_t ldz, const float* Wp, const float* act_w,
                     const float* act_b, int32_t num_actions, float* logits, int32_t* actions, float* y,
                     mapf_stream_t stream);
/* Fused graph filter + head on the tensor cores: S, hops, mix, ReLU, actionMLP, argmax in ONE kernel per 128-row tile
 * of whole episodes (the taps never touch HBM).  x agent-major [B*N, F]; Wp as for mapf_tc_mix_head.
 * mapf_tc_graph_supported tells whether the dimensions fit (N <= 64, F % 32 == 0, 64 < G <= 128). */
int mapf_tc_graph_supported(int32_t agents, int32_t in_dim, int32_t out_dim, int32_t taps);
int mapf_tc_graph_head(int32_t batch, int32_t agents, int32_t in_dim, int32_t out_dim, int32_t taps,
                       int32_t add_self_loop, int32_t has_skip, const float* x, const float* gso, const float* Wp,
                       const float* act_w, const float* act_b, int32_t num_actions, float* logits, int32_t* actions,
                       mapf_stream_t stream);
/* Filter taps Z [B*N, (K+has_skip)*F] (x_k = x_{k-1} S, gnn.py:104-109, + skip columns gnn.py:201-204) from
 * agent-major x; uses the x / gso / z / size / flag fields of the argument struct. */
int mapf_graph_taps(const mapf_graph_filter_fwd_args* a, mapf_stream_t stream);
/* same result from the FFMA kernel for every team size (mapf_graph_taps runs the hops of 64-agent teams on the tensor
 * cores: 3xTF32, equal to ~1e-6 relative) */
int mapf_graph_taps_ffma(const mapf_graph_filter_fwd_args* a, mapf_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Conflict-Based Search expert (host code; reference cbs/cbs.py `CBS(env).search()` :326-410 over
 * `Environment(dimension, agents, obstacles)` :148-157 and cbs/a_star.py `AStar.search` :29-102): the offline label
 * generator of the imitation data set.  The plans are identical to the reference's (same neighbour order, same
 * (f, push counter) / (cost, push counter) orderings, same first-conflict rule, same iteration caps).
 * Coordinates are (x, y) pairs as in the reference's YAML; a path is one cell per time step starting at t = 0.
 * mapf_cbs_solve fills path_len[agent] with the TRUE lengths and paths[agent][t < min(len, max_path)]; `truncated` = 1
 * when some path did not fit (call again with a larger max_path).  solved = 0: no plan within the caps (the reference
 * returns {}), path_len = 0.  mapf_cbs_solve_batch solves independent instances on `threads` host threads (0 = all). */
typedef struct {
  int32_t width, height;              /* dimension[0], dimension[1]: 0 <= x < width, 0 <= y < height */
  int32_t num_agents;
  const int32_t* starts;              /* [num_agents][2] */
  const int32_t* goals;               /* [num_agents][2] */
  int32_t num_obstacles;
  const int32_t* obstacles;           /* [num_obstacles][2] */
  int32_t max_high_level_iterations;  /* <= 0: 5000 (cbs.py:353) */
  int32_t max_low_level_iterations;   /* <= 0: 10000 (a_star.py:59) */
  int32_t max_path;                   /* capacity of `paths` per agent, in cells */
  int32_t* path_len;                  /* out [num_agents] */
  int32_t* paths;                     /* out [num_agents][max_path][2] */
  int32_t solved, truncated;          /* out */
  int32_t expanded, generated;        /* out: high-level nodes closed / pushed */
  int64_t cost;                       /* out: sum of path lengths (compute_solution_cost, cbs.py:302-303) */
} mapf_cbs_args;
int mapf_cbs_solve(mapf_cbs_args* a);
int mapf_cbs_solve_batch(mapf_cbs_args* items, int32_t count, int32_t threads);
/* one low-level search for `agent` under explicit constraints (Environment.a_star.search with env.constraints set):
 * vertex constraints [nv][3] = (t, x, y), edge constraints [ne][5] = (t, x1, y1, x2, y2); fills path_len[agent] and
 * paths[agent]; *found = 0 when there is no path within the cap (the reference returns False) */
int mapf_cbs_a_star(mapf_cbs_args* a, int32_t agent, const int32_t* vertex_constraints, int32_t nv,
                    const int32_t* edge_constraints, int32_t ne, int32_t* found);

#ifdef __cplusplus
}
#endif
#endif /* MAPF_B200_H */
