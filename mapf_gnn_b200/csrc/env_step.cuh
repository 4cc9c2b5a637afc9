// Body of the batched environment step (env_kernels.cu), as device functions over NT threads that own `nep` consecutive
// episodes starting at e0 (LANES = next power of two >= N lanes per episode, nep * LANES <= NT): used by env_step_kernel and
// by the tail of the fused graph-filter kernel (tc_graph.cu), whose CTA steps the episodes whose actions it just chose.
// Semantics: grid/env_graph_gridv1.py, see env_kernels.cu.
#pragma once
#include <limits.h>

#include "common.cuh"

template <int LANES, int NT>
struct EnvShared {
  int x[NT], y[NT], thr[NT];
  int8_t gr[NT], gc[NT];
  int notgoal[NT / LANES];
};

struct EnvAgent {
  int ox, oy, gx, gy;
  bool moving;
};

// part 1: everything that does not need the actions -- boards into shared memory, own position / goal into registers
template <int LANES, int NT>
__device__ __forceinline__ EnvAgent env_step_stage(const mapf_env_step_args& a, int e0, int nep, int tid,
                                                   uint8_t* s_cells, EnvShared<LANES, NT>& sh) {
  constexpr int kEnvThreads = NT;
  constexpr int EPB = NT / LANES;
  int* const s_notgoal = sh.notgoal;
  const int el = tid / LANES, lane = tid % LANES;
  const int e = e0 + el;
  const int N = a.agents, stride = a.cell_stride;
  const bool agent = (el < nep) && (lane < N);
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
  return EnvAgent{ox, oy, gx, gy, moving};
}

// part 2: move, collisions, board update, adjacency, field of view, time / done.  s_actions (optional): the actions of
// this CTA's agents in shared memory, row = episode-local index * N + agent, instead of a.actions
template <int LANES, int NT>
__device__ __forceinline__ void env_step_finish(const mapf_env_step_args& a, int e0, int nep, int tid, uint8_t* s_cells,
                                                EnvShared<LANES, NT>& sh, const EnvAgent& ag,
                                                const int32_t* s_actions) {
  constexpr int kEnvThreads = NT;
  int* const s_x = sh.x; int* const s_y = sh.y; int* const s_thr = sh.thr;
  int8_t* const s_gr = sh.gr; int8_t* const s_gc = sh.gc;
  int* const s_notgoal = sh.notgoal;
  const int el = tid / LANES, lane = tid % LANES;
  const int e = e0 + el;
  const int N = a.agents, H = a.board, stride = a.cell_stride;
  const bool agent = (el < nep) && (lane < N);
  const int ox = ag.ox, oy = ag.oy, gx = ag.gx, gy = ag.gy;
  const bool moving = ag.moving;
  __syncthreads();

  uint8_t* cells = s_cells + el * stride;
  int nx = ox, ny = oy;
  if (moving) {
    // A1: delta -> obstacle-set revert -> clamp   (env_graph_gridv1.py:202-211)
    const int act = s_actions != nullptr ? s_actions[el * N + lane] : __ldg(a.actions + int64_t(e) * N + lane);
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
