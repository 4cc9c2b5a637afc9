// Shared helpers for the sm_100a kernels behind include/mapf_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "mapf_b200.h"

#define MAPF_REQUIRE(cond, code) \
  do {                           \
    if (!(cond)) return (code);  \
  } while (0)

static inline int mapf_launch_status() {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return MAPF_ERR_CUDA;
  }
  return MAPF_OK;
}

// BatchNorm + ReLU of a training conv layer's INPUT applied on load (enc_tc.cu mapf_conv_tc_layer; train.py:82-100 forward):
// the layer reads the pre-BatchNorm output u of the layer in front, every CTA derives the (slot, channel) scale / shift pairs
// from that layer's fp64 batch sums, CTA 0 also writes the tables the backward pass reads and advances the running
// statistics, and act_out receives relu(bn(u)) -- the activation tensor the weight-gradient kernels read.
struct MapfConvBnIn {
  const double* stats;        // [N][stat_c][2] fp64 (sum, sum of squares) of the layer in front
  const float* gamma;         // [16]
  const float* beta;          // [16]
  double count;               // values per (slot, channel): B * 25
  float momentum;
  int stat_c;
  float* meaninv;             // out [N][16][2] (mean, invstd)
  float* scale_shift;         // out [N][16][2]
  float* running_mean;        // in / out [16] or NULL
  float* running_var;
  int64_t* batches;           // num_batches_tracked or NULL
  float* act_out;             // out [rows][16*25]
};

static inline bool mapf_aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

constexpr int kFov = 5;          // FOV side, env_graph_gridv1.py:252 with pad = 3
constexpr int kFovHalf = 2;
constexpr int kFovCells = 25;
constexpr int kFovFloats = 50;   // 2 channels x 5 x 5
constexpr uint8_t kObstacleFlag = 0x80;
constexpr uint8_t kBoardValueMask = 0x03;

// Packed-weight layout shared by pack_weights and the consumers (offsets in floats).
struct PackedLayout {
  int64_t conv_w[MAPF_MAX_CONV];  // [Cout/4][Cin][9][4]
  int64_t conv_b[MAPF_MAX_CONV];  // [Cout]
  int64_t fc_wt;                  // [fc_in][F]   (transposed compressMLP.0.weight)
  int64_t fc_b;                   // [F]
  int64_t gf_wt;                  // [(K(+1))*F][G]
  int64_t tc_fc;                  // tensor-core B operand of the encoder FC (compressMLP.0.weight [F, fc_in])
  int64_t tc_gf;                  // tensor-core B operand of the mix: [W | W2] as chunked canonical TF32 hi/lo blocks
  int64_t tc_conv;                // tensor-core B operands of the conv layers (enc_tc.cu), 0 floats when unsupported
  int64_t tc_fc_pm;               // tensor-core B operand of the encoder FC with K ordered pixel-major (enc_tc.cu)
  int64_t act_w;                  // [A][G or F]
  int64_t act_b;                  // [A]
  int64_t total;
};

static inline __host__ __device__ int64_t mapf_round4(int64_t v) { return (v + 3) & ~int64_t(3); }

// floats of a pre-tiled tensor-core B operand [N, K] (tc_gemm.cu: KC = 8, NT = 64 or 128, hi | lo per chunk)
static inline __host__ __device__ int64_t mapf_tc_b_floats(int N, int K) {
  return int64_t((K + 7) / 8) * 2 * (N <= 64 ? 64 : 128) * 8;
}

// the tensor-core conv stack (enc_tc.cu) is built for a first layer of 16 or 32 channels and 16-channel layers behind it
static inline __host__ __device__ bool mapf_enc_tc_dims_ok(const mapf_net_params& p) {
  if (p.num_conv < 1 || p.num_conv > MAPF_MAX_CONV || p.fc_out > 128) return false;
  if (p.channels[1] != 16 && p.channels[1] != 32) return false;
  for (int l = 2; l <= p.num_conv; ++l)
    if (p.channels[l] != 16) return false;
  return p.channels[p.num_conv] == 16;
}

static inline __host__ __device__ PackedLayout mapf_packed_layout(const mapf_net_params& p) {
  PackedLayout L;
  int64_t off = 0;
  for (int l = 0; l < MAPF_MAX_CONV; ++l) {
    L.conv_w[l] = off;
    if (l < p.num_conv) off += mapf_round4(int64_t(p.channels[l + 1]) * p.channels[l] * 9);
    L.conv_b[l] = off;
    if (l < p.num_conv) off += mapf_round4(p.channels[l + 1]);
  }
  L.fc_wt = off;
  off += mapf_round4(int64_t(p.fc_in) * p.fc_out);
  L.fc_b = off;
  off += mapf_round4(p.fc_out);
  L.gf_wt = off;
  if (p.taps > 0) off += mapf_round4(int64_t(p.taps + 1) * p.fc_out * p.node_dim);  // room for the W2 rows
  L.tc_fc = off;
  if (p.fc_out <= 128) off += mapf_tc_b_floats(p.fc_out, p.fc_in);
  L.tc_gf = off;
  if (p.taps > 0 && p.node_dim <= 128) off += mapf_tc_b_floats(p.node_dim, (p.taps + 1) * p.fc_out);
  L.tc_conv = off;
  L.tc_fc_pm = off;
  if (mapf_enc_tc_dims_ok(p)) {
    off += 4;   // inverse filter scales
    for (int l = 0; l < p.num_conv; ++l)
      off += int64_t(p.channels[l + 1] / 16) * ((3 * p.channels[l] + 15) >> 4) * 2 * (48 * 8);
    L.tc_fc_pm = off;
    off += 4 + int64_t((p.fc_in + 15) / 16) * (p.fc_out <= 64 ? 64 : 128) * 16;   // fp16 hi|lo blocks + scale header
  }
  L.act_w = off;
  off += mapf_round4(int64_t(p.num_actions) * (p.taps > 0 ? p.node_dim : p.fc_out));
  L.act_b = off;
  off += mapf_round4(p.num_actions);
  L.total = off;
  return L;
}

constexpr int kMaxActions = 8;

static inline int mapf_check_params(const mapf_net_params* p) {
  MAPF_REQUIRE(p != nullptr, MAPF_ERR_NULL);
  MAPF_REQUIRE(p->num_conv >= 1 && p->num_conv <= MAPF_MAX_CONV, MAPF_ERR_UNSUPPORTED);
  MAPF_REQUIRE(p->channels[0] == 2, MAPF_ERR_SHAPE);
  for (int l = 1; l <= p->num_conv; ++l)
    MAPF_REQUIRE(p->channels[l] > 0 && p->channels[l] <= 64 && (p->channels[l] & 3) == 0, MAPF_ERR_UNSUPPORTED);
  MAPF_REQUIRE(p->fc_in == kFovCells * p->channels[p->num_conv], MAPF_ERR_SHAPE);
  MAPF_REQUIRE(p->fc_out > 0 && (p->fc_out & 15) == 0, MAPF_ERR_UNSUPPORTED);
  MAPF_REQUIRE(p->num_actions > 0 && p->num_actions <= kMaxActions, MAPF_ERR_UNSUPPORTED);
  MAPF_REQUIRE(p->taps >= 0 && p->taps <= 8, MAPF_ERR_UNSUPPORTED);
  MAPF_REQUIRE(p->taps == 0 || (p->node_dim > 0 && (p->node_dim & 3) == 0), MAPF_ERR_UNSUPPORTED);
  return MAPF_OK;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) costs microseconds of host time per call; remember, per device and
// per kernel call site, the largest size already granted.  Benign race: the attribute call is idempotent.
template <typename F>
static inline bool mapf_ensure_smem(F func, size_t bytes, int* cache /* [16] zero-initialised, one per call site */) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  const int slot = dev & 15;
  if (int(bytes) <= cache[slot]) return true;
  if (cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)) != cudaSuccess) return false;
  cache[slot] = int(bytes);
  return true;
}

// Programmatic dependent launch: the kernels of a rollout step (encoder conv stack -> encoder FC -> graph filter + head ->
// env step -> next step's encoder ...) are launched with programmatic stream serialization, so a kernel's CTAs may become
// resident -- and run their prologue: barrier init, TMEM allocation, weights into shared memory -- while the previous
// kernel drains.  Every kernel calls mapf_pdl_wait() before its first access to memory another kernel of the chain
// writes or reads (it returns once ALL earlier kernels have completed and flushed), and mapf_pdl_launch() to let its own
// successor start early.  Without the launch attribute both are no-ops (mapf_policy_fwd_args.flags & MAPF_POLICY_NO_PDL).
__device__ __forceinline__ void mapf_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void mapf_pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Data written by an EARLIER KERNEL of a programmatically launched chain must not be read through the non-coherent path
// (__ldg / ld.global.nc): ptxas treats such a load as a read of immutable memory and may schedule it ABOVE
// griddepcontrol.wait, i.e. before the producing kernel has finished (found in round 2: the env kernel read the previous
// step's actions).  mapf_ld_dep is an ordinary coherent load; `asm volatile` keeps it behind the wait, no memory clobber
// so that independent loads can still be batched ahead of shared-memory stores.  Weights (packed long before, behind
// pack_fence_kernel) keep __ldg.
__device__ __forceinline__ float mapf_ld_dep(const float* p) {
  float v;
  asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 mapf_ld_dep(const float4* p) {
  float4 v;
  asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float2 mapf_ld_dep(const float2* p) {
  float2 v;
  asm volatile("ld.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ int mapf_ld_dep(const int* p) {
  int v;
  asm volatile("ld.global.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double mapf_ld_dep(const double* p) {
  double v;
  asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

// Per-thread, per-call launch option: the entry points that take a `flags` field (mapf_policy_fwd, mapf_rollout) switch
// programmatic launches off for the duration of the call when MAPF_POLICY_NO_PDL is set.  Thread-local and restored on
// return: the library stays re-entrant and keeps no process-wide mutable state.
inline thread_local int mapf_tl_no_pdl = 0;
struct MapfPdlScope {
  int saved;
  explicit MapfPdlScope(bool no_pdl) : saved(mapf_tl_no_pdl) { if (no_pdl) mapf_tl_no_pdl = 1; }
  ~MapfPdlScope() { mapf_tl_no_pdl = saved; }
};
static inline bool mapf_pdl_enabled() { return mapf_tl_no_pdl == 0; }

template <typename... KArgs, typename... Args>
static inline void mapf_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                   Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = mapf_pdl_enabled() ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);   // errors surface through mapf_launch_status()
}

// Packed fp32 FMA (Blackwell FFMA2): {d.x, d.y} = a * {w.x, w.y} + {d.x, d.y}, two IEEE fp32 FMAs in one issue slot.
// ptxas folds the duplicated multiplicand into the scalar-broadcast operand form (FFMA2 Rd, Ra.F32, Rb.F32x2, Rc.F32x2).
// Fully packed form: {d.x, d.y} += {a.x * b.x, a.y * b.y}.
__device__ __forceinline__ void mapf_ffma2_pp(float2& d, const float2& a, const float2& b) {
  unsigned long long A, B, C;
  asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(C) : "f"(d.x), "f"(d.y));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(C) : "l"(A), "l"(B));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(C));
}

__device__ __forceinline__ void mapf_ffma2(float2& d, float a, const float2& w) {
  unsigned long long A, B, C;
  asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a), "f"(a));
  asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(w.x), "f"(w.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(C) : "f"(d.x), "f"(d.y));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(C) : "l"(A), "l"(B));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(C));
}
