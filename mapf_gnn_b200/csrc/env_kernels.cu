// Batched grid environment: reset, fused move + adjacency + FOV step, metrics.
// New B200-first design of the semantics of the reference's single-episode numpy GraphEnv
// (grid/env_graph_gridv1.py; exact rules restated in SURVEY.md 8a A1-A4).
//
// Work decomposition: one CTA of 128 threads owns EPB = 128/LANES consecutive episodes, LANES =
// next power of two >= N lanes per episode.  The episodes' board cells are staged into shared
// memory with 128-bit coalesced loads, all integer logic (move, obstacle revert, clamp, single-pass
// collision revert, closest-4 threshold) runs on shared memory, and the two observation tensors of
// the CTA's episodes -- contiguous ranges of fov[E,N,2,5,5] and adj[E,N,N] -- are written back
// cooperatively with 128-bit coalesced stores.  HBM traffic per agent-step is the algorithmic
// minimum: pos + goal + H*H/N board bytes in, 200 B FOV + 4N B adjacency + the 2 changed cells out.
#include <limits.h>

#include "common.cuh"

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
__global__ void __launch_bounds__(kEnvThreads) env_step_kernel(mapf_env_step_args a) {
  constexpr int EPB = kEnvThreads / LANES;
  extern __shared__ __align__(16) uint8_t s_cells[];
  __shared__ int s_x[kEnvThreads], s_y[kEnvThreads], s_thr[kEnvThreads];
  __shared__ int8_t s_gr[kEnvThreads], s_gc[kEnvThreads];
  __shared__ int s_notgoal[EPB];

  const int tid = threadIdx.x;
  const int el = tid / LANES, lane = tid % LANES;
  const int e0 = blockIdx.x * EPB;
  const int nep = min(EPB, a.episodes - e0);
  const int e = e0 + el;
  const int N = a.agents, H = a.board, stride = a.cell_stride;
  const bool agent = (el < nep) && (lane < N);

  // do_move & 4 (mapf_rollout): boards, positions, goals and done flags were last written before the policy kernels in
  // front of this launch started, only the actions come from the kernel right in front -- stage the state while it drains
  const bool old_state = (a.do_move & 4) != 0;
  if (!old_state) { mapf_pdl_wait(); mapf_pdl_launch(); }
  {  // stage this CTA's boards (contiguous, 16-byte aligned by construction)
    const uint4* src = reinterpret_cast<const uint4*>(a.cells + int64_t(e0) * stride);
    uint4* dst = reinterpret_cast<uint4*>(s_cells);
    const int nvec = nep * (stride >> 4);
    for (int i = tid; i < nvec; i += kEnvThreads) dst[i] = src[i];  // plain load: this kernel writes cells later
  }
  if (tid < EPB) s_notgoal[tid] = 0;

  int ox = 0, oy = 0, gx = 0, gy = 0;
  bool moving = false;
  if (agent) {
    const int2 p = reinterpret_cast<const int2*>(a.pos)[int64_t(e) * N + lane];
    const int2 g = __ldg(reinterpret_cast<const int2*>(a.goal) + int64_t(e) * N + lane);
    ox = p.x; oy = p.y; gx = g.x; gy = g.y;
    moving = (a.do_move & 3) && (!a.done[e] || (a.do_move & 2));   // do_move & 2: keep stepping finished episodes
  }
  if (old_state) { mapf_pdl_wait(); mapf_pdl_launch(); }
  __syncthreads();

  uint8_t* cells = s_cells + el * stride;
  int nx = ox, ny = oy;
  if (moving) {
    // A1: delta -> obstacle-set revert -> clamp   (env_graph_gridv1.py:202-211)
    const int act = __ldg(a.actions + int64_t(e) * N + lane);
    nx = ox + (act == 1) - (act == 3);
    ny = oy + (act == 2) - (act == 4);
    if (nx >= 0 && nx < H && ny >= 0 && ny < H && (cells[ny * H + nx] & kObstacleFlag)) { nx = ox; ny = oy; }
    nx = min(max(nx, 0), H - 1);
    ny = min(max(ny, 0), H - 1);
  }
  s_x[tid] = nx; s_y[tid] = ny;
  __syncthreads();
  const int base = el * LANES;
  if (moving) {
    // single-pass collision revert of every member of a shared cell (:282-301)
    int cnt = 0;
    for (int j = 0; j < N; ++j) cnt += (s_x[base + j] == nx) & (s_y[base + j] == ny);
    if (cnt > 1) { nx = ox; ny = oy; }
  }
  __syncthreads();
  s_x[tid] = nx; s_y[tid] = ny;
  if (moving) cells[oy * H + ox] &= kObstacleFlag;  // updateBoard: old cells <- 0 (:273-275)
  __syncthreads();
  if (moving) cells[ny * H + nx] = (cells[ny * H + nx] & kObstacleFlag) | uint8_t(1);  // new cells <- 1
  __syncthreads();
  if (moving) {
    uint8_t* gcells = a.cells + int64_t(e) * stride;
    gcells[oy * H + ox] = cells[oy * H + ox];
    gcells[ny * H + nx] = cells[ny * H + nx];
    reinterpret_cast<int2*>(a.pos)[int64_t(e) * N + lane] = make_int2(nx, ny);
  }

  // A2: per-row closest-4 threshold on integer squared distances (:117-136,174-181)
  int thr = -1;
  if (agent) {
    int m0 = INT_MAX, m1 = INT_MAX, m2 = INT_MAX, m3 = INT_MAX, cnt = 0, mx = -1;
    for (int j = 0; j < N; ++j) {
      const int dx = s_x[base + j] - nx, dy = s_y[base + j] - ny;
      const int d2 = dx * dx + dy * dy;
      if (d2 > 0 && d2 < a.d2_limit) {
        ++cnt;
        mx = max(mx, d2);
        if (d2 < m3) {
          m3 = d2;
          if (m3 < m2) { const int t = m2; m2 = m3; m3 = t; }
          if (m2 < m1) { const int t = m1; m1 = m2; m2 = t; }
          if (m1 < m0) { const int t = m0; m0 = m1; m1 = t; }
        }
      }
    }
    thr = cnt >= 4 ? m3 : mx;
    // A3 goal channel: single mark, projected onto the FOV border (:218-246)
    const int dgx = min(max(gx - nx, -kFovHalf), kFovHalf), dgy = min(max(gy - ny, -kFovHalf), kFovHalf);
    s_gr[tid] = int8_t(kFovHalf - dgy);
    s_gc[tid] = int8_t(kFovHalf + dgx);
    if (nx != gx || ny != gy) s_notgoal[el] = 1;
  }
  s_thr[tid] = thr;
  __syncthreads();

  if (a.adj != nullptr) {
    const int nn = N * N, total = nep * nn;
    float* out = a.adj + int64_t(e0) * nn;
    for (int idx = tid; idx < total; idx += kEnvThreads) {
      const int el2 = idx / nn, rem = idx - el2 * nn;
      const int i = rem / N, j = rem - i * N;
      const int b = el2 * LANES;
      const int dx = s_x[b + i] - s_x[b + j], dy = s_y[b + i] - s_y[b + j];
      const int d2 = dx * dx + dy * dy;
      out[idx] = (d2 > 0 && d2 < a.d2_limit && d2 <= s_thr[b + i]) ? 1.0f : 0.0f;
    }
  }

  if (a.fov != nullptr) {
    // A3 channel 0: fov[r,c] = board[y+2-r, x-2+c], zero outside (:248-265); channel 1: the goal mark.
    // One work item = one 5-float row of one channel of one agent (10 items per agent): the per-agent state is
    // decoded once per item and consecutive items write consecutive floats, so a warp writes one contiguous span.
    const int items = nep * N * 10;
    float* out = a.fov + int64_t(e0) * N * kFovFloats;
    for (int it = tid; it < items; it += kEnvThreads) {
      const int ag = it / 10, rr = it - ag * 10;
      const int el2 = ag / N, ln = ag - el2 * N;
      const int t = el2 * LANES + ln;
      float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f;
      if (rr < kFov) {
        const int by = s_y[t] + kFovHalf - rr, bx = s_x[t] - kFovHalf;
        if (by >= 0 && by < H) {
          const uint8_t* row = s_cells + el2 * stride + by * H;
          if (bx >= 0 && bx < H) v0 = float(row[bx] & kBoardValueMask);
          if (bx + 1 >= 0 && bx + 1 < H) v1 = float(row[bx + 1] & kBoardValueMask);
          if (bx + 2 >= 0 && bx + 2 < H) v2 = float(row[bx + 2] & kBoardValueMask);
          if (bx + 3 >= 0 && bx + 3 < H) v3 = float(row[bx + 3] & kBoardValueMask);
          if (bx + 4 >= 0 && bx + 4 < H) v4 = float(row[bx + 4] & kBoardValueMask);
        }
      } else if (rr - kFov == s_gr[t]) {
        const int gc = s_gc[t];
        v0 = gc == 0 ? 3.0f : 0.0f; v1 = gc == 1 ? 3.0f : 0.0f; v2 = gc == 2 ? 3.0f : 0.0f;
        v3 = gc == 3 ? 3.0f : 0.0f; v4 = gc == 4 ? 3.0f : 0.0f;
      }
      float* dst = out + it * kFov;
      dst[0] = v0; dst[1] = v1; dst[2] = v2; dst[3] = v3; dst[4] = v4;
    }
  }

  if ((a.do_move & 3) && tid < nep) {  // time += 1; done = checkAllInGoal() (:195-197)
    const int ee = e0 + tid;
    if (!a.done[ee] || (a.do_move & 2)) {
      a.time[ee] += 1;
      a.done[ee] = s_notgoal[tid] ? 0 : 1;
    }
  }
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

template <int LANES>
int launch_env_step(const mapf_env_step_args& a, cudaStream_t s) {
  constexpr int EPB = kEnvThreads / LANES;
  const size_t smem = size_t(EPB) * a.cell_stride;
  if (smem > 200 * 1024) return MAPF_ERR_UNSUPPORTED;
  if (smem > 40 * 1024) {
    if ([&]() { static int cache__[16] = {0}; return !mapf_ensure_smem(env_step_kernel<LANES>, size_t(smem), cache__); }())
      return MAPF_ERR_CUDA;
  }
  const int grid = (a.episodes + EPB - 1) / EPB;
  mapf_launch_pdl(env_step_kernel<LANES>, dim3(grid), dim3(kEnvThreads), smem, s, a);
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
