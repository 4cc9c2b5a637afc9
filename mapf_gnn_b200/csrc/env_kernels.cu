// Batched grid environment: reset, fused move + adjacency + FOV step, metrics.
// New B200-first design of the semantics of the reference's single-episode numpy GraphEnv
// (grid/env_graph_gridv1.py; exact rules restated in SURVEY.md 8a A1-A4).
//
// Work decomposition: one CTA of 128 threads owns up to 128/LANES consecutive episodes, LANES =
// next power of two >= N lanes per episode.  The episodes' board cells are staged into shared
// memory with 128-bit coalesced loads, all integer logic (move, obstacle revert, clamp, single-pass
// collision revert, closest-4 threshold) runs on shared memory, and the two observation tensors of
// the CTA's episodes -- contiguous ranges of fov[E,N,2,5,5] and adj[E,N,N] -- are assembled in a
// shared-memory staging area and leave as one TMA bulk store each (env_step.cuh).  HBM traffic per
// agent-step is the algorithmic minimum: pos + goal + H*H/N board bytes in, 200 B FOV + 4N B
// adjacency + the 2 changed cells out.
#include <limits.h>

#include "common.cuh"
#include "env_step.cuh"

namespace {

constexpr int kEnvThreads = 128;

__global__ void env_reset_kernel(mapf_env_reset_args a) {
  const int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  const int64_t n_agents = int64_t(a.episodes) * a.agents;
  const int64_t n_obst = int64_t(a.episodes) * a.num_obstacles;
  if (t < n_agents) {
    reinterpret_cast<int2*>(a.pos)[t] = reinterpret_cast<const int2*>(a.starts)[t];
  }
  if (t < n_obst) {
    const int e = int(t / a.num_obstacles);
    const int2 o = reinterpret_cast<const int2*>(a.obstacles)[t];
    if (o.x >= 0 && o.x < a.board && o.y >= 0 && o.y < a.board)
      a.cells[int64_t(e) * a.cell_stride + o.y * a.board + o.x] = uint8_t(2) | kObstacleFlag;
  }
}

template <int LANES>
__global__ void __launch_bounds__(kEnvThreads) env_step_kernel(mapf_env_step_args a, int epb) {
  extern __shared__ __align__(16) uint8_t s_env_dyn[];   // [epb boards | staging of the CTA's adj and fov slices]
  __shared__ EnvShared<LANES, kEnvThreads> sh;
  uint8_t* s_cells = s_env_dyn;
  float* s_out = reinterpret_cast<float*>(s_env_dyn + size_t(epb) * a.cell_stride);
  const int tid = threadIdx.x;
  const int e0 = blockIdx.x * epb;
  const int nep = min(epb, a.episodes - e0);
  // do_move & 4 (mapf_rollout): boards, positions, goals and done flags were last written before the policy kernels in
  // front of this launch started, only the actions come from the kernel right in front -- stage the state while it drains
  const bool old_state = (a.do_move & 4) != 0;
  if (!old_state) { mapf_pdl_wait(); mapf_pdl_launch(); }
  const EnvAgent ag = env_step_stage<LANES, kEnvThreads>(a, e0, nep, tid, s_cells, sh);
  if (old_state) { mapf_pdl_wait(); mapf_pdl_launch(); }
  env_step_finish<LANES, kEnvThreads>(a, e0, nep, tid, s_cells, sh, ag, nullptr, s_out);
}

__global__ void env_metrics_kernel(mapf_env_metrics_args a) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.episodes) return;
  int hit = 0;
  for (int n = 0; n < a.agents; ++n) {
    const int2 p = reinterpret_cast<const int2*>(a.pos)[int64_t(e) * a.agents + n];
    const int2 g = reinterpret_cast<const int2*>(a.goal)[int64_t(e) * a.agents + n];
    hit += (p.x == g.x) & (p.y == g.y);
  }
  a.success[e] = float(hit) / float(a.agents);
  a.flow_time[e] = (hit == a.agents) ? a.time[e] : a.agents * a.max_time;
}

// Episodes per CTA: LANES lanes per episode do the per-agent rules, all 128 threads write the observations.  Few episodes
// per CTA keep every SM busy at small E (C2a: 4096 episodes -> 1024 CTAs instead of 256) and the staging area small at
// large N; a multiple of 4 for odd N keeps the CTA's output slices 16-byte aligned (bulk stores).
// The single rules of _updatePositions on caller-supplied candidate positions (the reference's tests and callers poke
// them one at a time): bit 0 check_collision_obstacle (:303-313), bit 1 check_boundary (:267-271), bit 2 check_collisions
// (:282-301), applied in the reference's order.  One CTA per episode, one thread per agent.
__global__ void __launch_bounds__(128) env_rules_kernel(mapf_env_rules_args a) {
  __shared__ int s_x[128], s_y[128];
  const int e = blockIdx.x, i = threadIdx.x, N = a.agents, H = a.board;
  const bool agent = i < N;
  int2 p = make_int2(kEnvFarAway, kEnvFarAway), t = p;
  if (agent) {
    p = reinterpret_cast<const int2*>(a.pos)[int64_t(e) * N + i];
    t = reinterpret_cast<const int2*>(a.pos_temp)[int64_t(e) * N + i];
    if ((a.stages & 1) && p.x >= 0 && p.x < H && p.y >= 0 && p.y < H &&
        (a.cells[int64_t(e) * a.cell_stride + p.y * H + p.x] & kObstacleFlag))
      p = t;
    if (a.stages & 2) { p.x = min(max(p.x, 0), H - 1); p.y = min(max(p.y, 0), H - 1); }
  }
  if (a.stages & 4) {
    s_x[i] = p.x; s_y[i] = p.y;
    __syncthreads();
    int cnt = 0;
    for (int j = 0; j < N; ++j) cnt += (s_x[j] == p.x) & (s_y[j] == p.y);
    if (agent && cnt > 1) p = t;
  }
  if (agent) reinterpret_cast<int2*>(a.pos)[int64_t(e) * N + i] = p;
}

template <int LANES>
int launch_env_step(const mapf_env_step_args& a, cudaStream_t s) {
  constexpr int EPB = kEnvThreads / LANES;
  const auto smem_of = [&](int epb) {
    return size_t(epb) * a.cell_stride + sizeof(float) * size_t(env_stage_floats(epb, a.agents));
  };
  int epb = a.episodes_per_cta;
  if (epb <= 0) {
    const int align = (a.agents & 1) ? 4 : 1;
    epb = EPB;
    while (epb > 1 && epb / 2 >= align && ((a.episodes + epb - 1) / epb < 148 * 8 || smem_of(epb) > 36 * 1024)) epb >>= 1;
  }
  MAPF_REQUIRE(epb >= 1 && epb <= EPB, MAPF_ERR_UNSUPPORTED);
  const size_t smem = smem_of(epb);
  if (smem > 200 * 1024) return MAPF_ERR_UNSUPPORTED;
  if (smem > 40 * 1024) {
    if ([&]() { static int cache__[16] = {0}; return !mapf_ensure_smem(env_step_kernel<LANES>, size_t(smem), cache__); }())
      return MAPF_ERR_CUDA;
  }
  const int grid = (a.episodes + epb - 1) / epb;
  mapf_launch_pdl(env_step_kernel<LANES>, dim3(grid), dim3(kEnvThreads), smem, s, a, epb);
  return mapf_launch_status();
}

}  // namespace

extern "C" int mapf_env_reset(const mapf_env_reset_args* a, mapf_stream_t stream) {
  MAPF_REQUIRE(a != nullptr, MAPF_ERR_NULL);
  MAPF_REQUIRE(a->starts && a->pos && a->cells && a->time && a->done, MAPF_ERR_NULL);
  MAPF_REQUIRE(a->num_obstacles == 0 || a->obstacles, MAPF_ERR_NULL);
  MAPF_REQUIRE(a->episodes > 0 && a->agents > 0 && a->board > 0 && a->num_obstacles >= 0, MAPF_ERR_SHAPE);
  MAPF_REQUIRE(a->cell_stride >= a->board * a->board && (a->cell_stride & 15) == 0, MAPF_ERR_ALIGN);
  MAPF_REQUIRE(mapf_aligned(a->cells, 16) && mapf_aligned(a->pos, 8) && mapf_aligned(a->starts, 8), MAPF_ERR_ALIGN);
  MAPF_REQUIRE(a->num_obstacles == 0 || mapf_aligned(a->obstacles, 8), MAPF_ERR_ALIGN);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(a->cells, 0, size_t(a->episodes) * a->cell_stride, s) != cudaSuccess) return MAPF_ERR_CUDA;
  if (cudaMemsetAsync(a->time, 0, size_t(a->episodes) * sizeof(int32_t), s) != cudaSuccess) return MAPF_ERR_CUDA;
  if (cudaMemsetAsync(a->done, 0, size_t(a->episodes), s) != cudaSuccess) return MAPF_ERR_CUDA;
  const int64_t work = int64_t(a->episodes) * max(a->agents, a->num_obstacles);
  const int threads = 256;
  env_reset_kernel<<<unsigned((work + threads - 1) / threads), threads, 0, s>>>(*a);
  return mapf_launch_status();
}

extern "C" int mapf_env_step(const mapf_env_step_args* a, mapf_stream_t stream) {
  MAPF_REQUIRE(a != nullptr, MAPF_ERR_NULL);
  MAPF_REQUIRE(a->pos && a->goal && a->cells && a->time && a->done, MAPF_ERR_NULL);
  MAPF_REQUIRE(!(a->do_move & 3) || a->actions, MAPF_ERR_NULL);
  MAPF_REQUIRE(a->episodes > 0 && a->agents > 0 && a->board > 0 && a->d2_limit >= 0, MAPF_ERR_SHAPE);
  MAPF_REQUIRE(a->cell_stride >= a->board * a->board && (a->cell_stride & 15) == 0, MAPF_ERR_ALIGN);
  MAPF_REQUIRE(mapf_aligned(a->cells, 16) && mapf_aligned(a->pos, 8) && mapf_aligned(a->goal, 8), MAPF_ERR_ALIGN);
  MAPF_REQUIRE(a->fov == nullptr || mapf_aligned(a->fov, 4), MAPF_ERR_ALIGN);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int n = a->agents;
  if (n <= 8) return launch_env_step<8>(*a, s);
  if (n <= 16) return launch_env_step<16>(*a, s);
  if (n <= 32) return launch_env_step<32>(*a, s);
  if (n <= 64) return launch_env_step<64>(*a, s);
  if (n <= 128) return launch_env_step<128>(*a, s);
  return MAPF_ERR_UNSUPPORTED;
}

extern "C" int mapf_env_apply_rules(const mapf_env_rules_args* a, mapf_stream_t stream) {
  MAPF_REQUIRE(a != nullptr && a->pos && a->pos_temp, MAPF_ERR_NULL);
  MAPF_REQUIRE(!(a->stages & 1) || a->cells, MAPF_ERR_NULL);
  MAPF_REQUIRE(a->episodes > 0 && a->agents > 0 && a->board > 0 && (a->stages & ~7) == 0, MAPF_ERR_SHAPE);
  MAPF_REQUIRE(a->agents <= 128, MAPF_ERR_UNSUPPORTED);
  MAPF_REQUIRE(mapf_aligned(a->pos, 8) && mapf_aligned(a->pos_temp, 8), MAPF_ERR_ALIGN);
  env_rules_kernel<<<a->episodes, 128, 0, static_cast<cudaStream_t>(stream)>>>(*a);
  return mapf_launch_status();
}

extern "C" int mapf_env_metrics(const mapf_env_metrics_args* a, mapf_stream_t stream) {
  MAPF_REQUIRE(a != nullptr, MAPF_ERR_NULL);
  MAPF_REQUIRE(a->pos && a->goal && a->time && a->success && a->flow_time, MAPF_ERR_NULL);
  MAPF_REQUIRE(a->episodes > 0 && a->agents > 0, MAPF_ERR_SHAPE);
  const int threads = 128;
  env_metrics_kernel<<<(a->episodes + threads - 1) / threads, threads, 0, static_cast<cudaStream_t>(stream)>>>(*a);
  return mapf_launch_status();
}

// ---------------------------------------------------------------------------------------------------------
// On-device scenario generation: per episode, M unique obstacle cells, then N unique obstacle-free
// goal cells (evaluate.py:207-208 order: create_obstacles :504-529, create_goals :465-501), then N unique
// obstacle-free start cells (explicit starts instead of reset()'s row/column draw :77-90, which fails on dense maps).
// One CTA per episode: the cell permutation lives in shared memory, a partial Fisher-Yates driven by a counter-based
// hash (no state carried between launches: the stream of episode e depends only on (seed, e)).
// Distributional, not stream-exact, parity with the reference's numpy generator.
// ---------------------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {  // lowbias32 integer hash
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

__global__ void __launch_bounds__(128) scenario_kernel(int episodes, int agents, int board, int num_obstacles,
                                                       uint64_t seed, int32_t* __restrict__ starts,
                                                       int32_t* __restrict__ goals, int32_t* __restrict__ obstacles) {
  extern __shared__ int s_perm[];
  const int e = blockIdx.x, cells = board * board;
  for (int i = threadIdx.x; i < cells; i += blockDim.x) s_perm[i] = i;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t ctr = 0;
    const uint32_t k0 = mix32(uint32_t(seed) ^ mix32(uint32_t(e) * 0x9E3779B9U + 1U));
    const uint32_t k1 = mix32(uint32_t(seed >> 32) + 0x85EBCA6BU + uint32_t(e));
    auto draw = [&](uint32_t n) -> uint32_t {   // uniform in [0, n) (multiply-shift; bias < 2^-20 for n <= 4096)
      const uint32_t r = mix32(k0 + 0x9E3779B9U * (++ctr)) ^ mix32(k1 ^ ctr);
      return uint32_t((uint64_t(r) * n) >> 32);
    };
    auto swap_to = [&](int i, int lo_end) {
      const int j = i + int(draw(uint32_t(lo_end - i)));
      const int t = s_perm[i]; s_perm[i] = s_perm[j]; s_perm[j] = t;
    };
    const int M = num_obstacles, N = agents;
    for (int i = 0; i < M + N; ++i) swap_to(i, cells);            // obstacles, then goals
    for (int i = 0; i < M; ++i) {
      const int c = s_perm[i];
      obstacles[(int64_t(e) * M + i) * 2] = c % board;
      obstacles[(int64_t(e) * M + i) * 2 + 1] = c / board;
    }
    for (int i = 0; i < N; ++i) {
      const int c = s_perm[M + i];
      goals[(int64_t(e) * N + i) * 2] = c % board;
      goals[(int64_t(e) * N + i) * 2 + 1] = c / board;
    }
    for (int i = M; i < M + N; ++i) swap_to(i, cells);            // starts: N fresh draws among the obstacle-free cells
    for (int i = 0; i < N; ++i) {
      const int c = s_perm[M + i];
      starts[(int64_t(e) * N + i) * 2] = c % board;
      starts[(int64_t(e) * N + i) * 2 + 1] = c / board;
    }
  }
}

}  // namespace

extern "C" int mapf_make_scenarios(int32_t episodes, int32_t agents, int32_t board, int32_t num_obstacles, uint64_t seed,
                                   int32_t* starts, int32_t* goals, int32_t* obstacles, mapf_stream_t stream) {
  MAPF_REQUIRE(starts && goals && (num_obstacles == 0 || obstacles), MAPF_ERR_NULL);
  MAPF_REQUIRE(episodes > 0 && agents > 0 && board > 0 && num_obstacles >= 0, MAPF_ERR_SHAPE);
  MAPF_REQUIRE(int64_t(num_obstacles) + agents <= int64_t(board) * board, MAPF_ERR_SHAPE);   // create_goals ValueError
  const size_t smem = sizeof(int) * size_t(board) * board;
  MAPF_REQUIRE(smem <= 200 * 1024, MAPF_ERR_UNSUPPORTED);
  static int cache[16] = {0};
  if (smem > 40 * 1024 && !mapf_ensure_smem(scenario_kernel, smem, cache)) return MAPF_ERR_CUDA;
  scenario_kernel<<<episodes, 128, smem, static_cast<cudaStream_t>(stream)>>>(episodes, agents, board, num_obstacles,
                                                                               seed, starts, goals, obstacles);
  return mapf_launch_status();
}
