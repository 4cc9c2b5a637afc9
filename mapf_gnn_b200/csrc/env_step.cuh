// Body of the batched environment step (env_kernels.cu), as device functions over NT threads that own `nep` consecutive
// episodes starting at e0 (LANES = next power of two >= N lanes per episode, nep * LANES <= NT): used by env_step_kernel and
// by the tail of the fused graph-filter kernel (tc_graph.cu), whose CTA steps the episodes whose actions it just chose.
// Semantics: grid/env_graph_gridv1.py, see env_kernels.cu.
//
// Write-out (round 2): the CTA's slices of adj[E,N,N] and fov[E,N,2,5,5] are contiguous in global memory.  They are
// assembled in a shared-memory staging area in exactly that layout (no runtime divisions: a compact agent row -> thread
// slot table replaces them; stores are bank-conflict free) and leave the SM as ONE TMA bulk copy each
// (cp.async.bulk shared -> global, SASS UBLKCP) when the slice is 16-byte aligned, else as a coalesced scalar copy.
#pragma once
#include <limits.h>

#include "common.cuh"

template <int LANES, int NT>
struct EnvShared {
  int x[NT], y[NT];            // per slot t = episode-local index * LANES + agent; slots without an agent sit far away
  int4 q[NT];                  // after the rules: {x, y, adjacency bound on d2 - 1 (unsigned), goal mark row | column << 8}
  int16_t slot[NT];            // compact agent row (episode-local index * N + agent) -> slot t
  int notgoal[NT / LANES];
};

constexpr int kEnvFarAway = -30000;   // coordinate of the unused slots of an episode: never equal, never in range

struct EnvAgent {
  int ox, oy, gx, gy;
  bool moving;
};

// floats of staging the finish half needs for nep episodes of N agents (adjacency slice | observation slice).  Teams of
// more than 16 agents (LANES >= 32) write their adjacency rows straight to global memory: a warp stores one matrix row
// per instruction (whole 128-byte lines at N = 32 / 64), and the smaller staging area doubles the resident CTAs.
__host__ __device__ inline int64_t env_stage_floats(int nep, int N) {
  return (N > 16 ? 0 : ((int64_t(nep) * N * N + 3) & ~int64_t(3))) + int64_t(nep) * N * kFovFloats;
}

// shared -> global TMA bulk store of `bytes` (multiple of 16, both addresses 16-byte aligned), one thread
__device__ __forceinline__ void env_bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(static_cast<uint32_t>(__cvta_generic_to_shared(ssrc))), "r"(bytes)
               : "memory");
}

// part 1: everything that does not need the actions -- boards into shared memory, own position / goal into registers
template <int LANES, int NT>
__device__ __forceinline__ EnvAgent env_step_stage(const mapf_env_step_args& a, int e0, int nep, int tid,
                                                   uint8_t* s_cells, EnvShared<LANES, NT>& sh) {
  constexpr int EPB = NT / LANES;
  const int el = tid / LANES, lane = tid % LANES;
  const int e = e0 + el;
  const int N = a.agents, stride = a.cell_stride;
  const bool agent = (el < nep) && (lane < N);
  {  // stage this CTA's boards (contiguous, 16-byte aligned by construction)
    const uint4* src = reinterpret_cast<const uint4*>(a.cells + int64_t(e0) * stride);
    uint4* dst = reinterpret_cast<uint4*>(s_cells);
    const int nvec = nep * (stride >> 4);
    for (int i = tid; i < nvec; i += NT) dst[i] = src[i];  // plain load: this kernel writes cells later
  }
  if (tid < EPB) sh.notgoal[tid] = 0;

  int ox = kEnvFarAway, oy = kEnvFarAway, gx = 0, gy = 0;
  bool moving = false;
  if (agent) {
    const int2 p = reinterpret_cast<const int2*>(a.pos)[int64_t(e) * N + lane];
    const int2 g = __ldg(reinterpret_cast<const int2*>(a.goal) + int64_t(e) * N + lane);
    ox = p.x; oy = p.y; gx = g.x; gy = g.y;
    moving = (a.do_move & 3) && (!a.done[e] || (a.do_move & 2));   // do_move & 2: keep stepping finished episodes
    sh.slot[el * N + lane] = int16_t(tid);
  }
  return EnvAgent{ox, oy, gx, gy, moving};
}

// part 2: move, collisions, board update, adjacency, field of view, time / done.  s_actions (optional): the actions of
// this CTA's agents in shared memory, row = episode-local index * N + agent, instead of a.actions.  s_out: staging area
// of env_stage_floats(nep, N) floats, 16-byte aligned.
template <int LANES, int NT>
__device__ __forceinline__ void env_step_finish(const mapf_env_step_args& a, int e0, int nep, int tid, uint8_t* s_cells,
                                                EnvShared<LANES, NT>& sh, const EnvAgent& ag,
                                                const int32_t* s_actions, float* s_out) {
  int* const s_x = sh.x; int* const s_y = sh.y;
  const int el = tid / LANES, lane = tid % LANES;
  const int e = e0 + el;
  const int N = a.agents, H = a.board, stride = a.cell_stride;
  const bool agent = (el < nep) && (lane < N);
  const int ox = ag.ox, oy = ag.oy, gx = ag.gx, gy = ag.gy;
  const bool moving = ag.moving;
  // loads whose latency the barriers below would otherwise serialise: the action, the episode's clock and done flag
  int act = 0;
  if (moving) act = s_actions != nullptr ? 0 : mapf_ld_dep(a.actions + int64_t(e) * N + lane);   // not __ldg (common.cuh)
  int ep_time = 0;
  bool ep_live = false, ep_done_in = false;
  if ((a.do_move & 3) && tid < nep) {
    ep_done_in = a.done[e0 + tid] != 0;
    ep_live = !ep_done_in || (a.do_move & 2);
    if (ep_live) ep_time = a.time[e0 + tid];
  }
  __syncthreads();

  uint8_t* cells = s_cells + el * stride;
  int nx = ox, ny = oy;
  if (moving) {
    // A1: delta -> obstacle-set revert -> clamp   (env_graph_gridv1.py:202-211)
    if (s_actions != nullptr) act = s_actions[el * N + lane];
    nx = ox + (act == 1) - (act == 3);
    ny = oy + (act == 2) - (act == 4);
    if (nx >= 0 && nx < H && ny >= 0 && ny < H && (cells[ny * H + nx] & kObstacleFlag)) { nx = ox; ny = oy; }
    nx = min(max(nx, 0), H - 1);
    ny = min(max(ny, 0), H - 1);
  }
  s_x[tid] = nx; s_y[tid] = ny;
  __syncthreads();
  const int base = el * LANES;
  const int4* const bx4 = reinterpret_cast<const int4*>(s_x + base);   // the episode's LANES slots, four per load
  const int4* const by4 = reinterpret_cast<const int4*>(s_y + base);
  if (moving) {
    // single-pass collision revert of every member of a shared cell (:282-301); unused slots never match
    int cnt = 0;
#pragma unroll 2
    for (int j = 0; j < N; j += 4) {
      const int4 xs = bx4[j >> 2], ys = by4[j >> 2];
      cnt += (xs.x == nx) & (ys.x == ny);
      cnt += (xs.y == nx) & (ys.y == ny);
      cnt += (xs.z == nx) & (ys.z == ny);
      cnt += (xs.w == nx) & (ys.w == ny);
    }
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
    reinterpret_cast<int2*>(a.pos_out != nullptr ? a.pos_out : a.pos)[int64_t(e) * N + lane] = make_int2(nx, ny);
  } else if (agent && a.pos_out != nullptr) {
    reinterpret_cast<int2*>(a.pos_out)[int64_t(e) * N + lane] = make_int2(nx, ny);   // out-of-place: every agent is written
  }

  // A2: per-row closest-4 rule on integer squared distances (:117-136,174-181).  u = d2 - 1 as an unsigned number
  // (co-located agents -> 0xFFFFFFFF, never a neighbour); the four smallest u of the row are kept sorted by a
  // branch-free insertion (min / max network), four candidates per pair of 128-bit loads.  With lim_u = d2_limit - 1:
  // the in-range candidates are those below lim_u; the row's threshold is the 4th smallest in-range u, or the largest
  // one when there are fewer than four -- i.e. the largest kept value below lim_u -- and entry (i, j) of the adjacency
  // is 1 iff u_ij < bound with bound = threshold + 1 (0 when the row has no in-range candidate).
  if (agent) {
    unsigned m0 = 0xFFFFFFFFu, m1 = 0xFFFFFFFFu, m2 = 0xFFFFFFFFu, m3 = 0xFFFFFFFFu;
    auto visit = [&](int xj, int yj) {
      const int dx = xj - nx, dy = yj - ny;
      unsigned t = min(m3, unsigned(dx * dx + dy * dy - 1));
      m3 = max(m2, t); t = min(m2, t);
      m2 = max(m1, t); t = min(m1, t);
      m1 = max(m0, t); m0 = min(m0, t);
    };
#pragma unroll 2
    for (int j = 0; j < N; j += 4) {
      const int4 xs = bx4[j >> 2], ys = by4[j >> 2];
      visit(xs.x, ys.x);
      visit(xs.y, ys.y);
      visit(xs.z, ys.z);
      visit(xs.w, ys.w);
    }
    const unsigned lim_u = a.d2_limit > 0 ? unsigned(a.d2_limit - 1) : 0u;
    const unsigned bound = m3 < lim_u ? m3 + 1 : m2 < lim_u ? m2 + 1 : m1 < lim_u ? m1 + 1 : m0 < lim_u ? m0 + 1 : 0u;
    // A3 goal channel: single mark, projected onto the FOV border (:218-246)
    const int dgx = min(max(gx - nx, -kFovHalf), kFovHalf), dgy = min(max(gy - ny, -kFovHalf), kFovHalf);
    sh.q[tid] = make_int4(nx, ny, int(bound), (kFovHalf - dgy) | ((kFovHalf + dgx) << 8));
    if (nx != gx || ny != gy) sh.notgoal[el] = 1;
  }
  __syncthreads();

  const int rows = nep * N;                       // agents of this CTA, compact row = episode-local index * N + agent
  const int adj_floats = rows * N, fov_floats = rows * kFovFloats;
  constexpr bool stage_adj = LANES <= 16;
  float* const s_adj = s_out;
  float* const s_fov = s_out + (stage_adj ? ((adj_floats + 3) & ~3) : 0);
  if (a.adj != nullptr) {
    const int warp = tid >> 5, l32 = tid & 31;
    if constexpr (stage_adj) {
      // one matrix row per LANES lanes: lane j of the group owns column j; the row's agent is a broadcast read
      constexpr int RPW = 32 / LANES;
      const int sub = l32 / LANES, j = l32 % LANES;
      for (int row = warp * RPW + sub; row < rows; row += (NT / 32) * RPW) {
        const int t = sh.slot[row];
        const int4 qi = sh.q[t];
        const int2 pj = *reinterpret_cast<const int2*>(&sh.q[(t & ~(LANES - 1)) + j]);
        const int dx = qi.x - pj.x, dy = qi.y - pj.y;
        if (j < N) s_adj[row * N + j] = unsigned(dx * dx + dy * dy - 1) < unsigned(qi.z) ? 1.0f : 0.0f;
      }
    } else {
      // a warp stores one matrix row per instruction straight to global memory; its lanes' columns stay in registers
      constexpr int CPL = LANES / 32;
      for (int el2 = 0; el2 < nep; ++el2) {
        const int b = el2 * LANES;
        int xj[CPL], yj[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) { xj[c] = s_x[b + l32 + 32 * c]; yj[c] = s_y[b + l32 + 32 * c]; }
        float* const g_row0 = a.adj + (int64_t(e0 + el2) * N) * N + l32;
#pragma unroll 2
        for (int i = warp; i < N; i += NT / 32) {
          const int4 qi = sh.q[b + i];
          float* const g_row = g_row0 + i * N;
#pragma unroll
          for (int c = 0; c < CPL; ++c) {
            const int dx = qi.x - xj[c], dy = qi.y - yj[c];
            if (l32 + 32 * c < N) g_row[32 * c] = unsigned(dx * dx + dy * dy - 1) < unsigned(qi.z) ? 1.0f : 0.0f;
          }
        }
      }
    }
  }

  if (a.fov != nullptr) {
    // A3 channel 0: fov[r,c] = board[y+2-r, x-2+c], zero outside (:248-265); channel 1: the goal mark.
    // One work item = row r of BOTH channels of one agent (5 items per agent, 10 floats each, no divergence between
    // the lanes of a warp); the staging area has the layout of the global slice.
    const int items = rows * kFov;
    for (int it = tid; it < items; it += NT) {
      const int row = it / kFov, rr = it - row * kFov;
      const int t = sh.slot[row];
      const int4 q = sh.q[t];
      float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f;
      const int by = q.y + kFovHalf - rr, bx = q.x - kFovHalf;
      if (unsigned(by) < unsigned(H)) {
        const uint8_t* brow = s_cells + (t / LANES) * stride + by * H;
        if (unsigned(bx) < unsigned(H)) v0 = float(brow[bx] & kBoardValueMask);
        if (unsigned(bx + 1) < unsigned(H)) v1 = float(brow[bx + 1] & kBoardValueMask);
        if (unsigned(bx + 2) < unsigned(H)) v2 = float(brow[bx + 2] & kBoardValueMask);
        if (unsigned(bx + 3) < unsigned(H)) v3 = float(brow[bx + 3] & kBoardValueMask);
        if (unsigned(bx + 4) < unsigned(H)) v4 = float(brow[bx + 4] & kBoardValueMask);
      }
      float* dst = s_fov + row * kFovFloats + rr * kFov;
      dst[0] = v0; dst[1] = v1; dst[2] = v2; dst[3] = v3; dst[4] = v4;
      const int gc = (q.w & 0xff) == rr ? (q.w >> 8) : -1;      // column of the goal mark if it lies in this row
      dst += kFovCells;
      dst[0] = gc == 0 ? 3.0f : 0.0f; dst[1] = gc == 1 ? 3.0f : 0.0f; dst[2] = gc == 2 ? 3.0f : 0.0f;
      dst[3] = gc == 3 ? 3.0f : 0.0f; dst[4] = gc == 4 ? 3.0f : 0.0f;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staging writes -> visible to the bulk copies
  __syncthreads();

  bool bulk = false;
  if (a.adj != nullptr && stage_adj) {
    float* out = a.adj + int64_t(e0) * N * N;
    if (((reinterpret_cast<uintptr_t>(out) | uintptr_t(adj_floats * 4)) & 15) == 0) {
      if (tid == 0) { env_bulk_s2g(out, s_adj, uint32_t(adj_floats) * 4u); bulk = true; }
    } else {
      for (int i = tid; i < adj_floats; i += NT) out[i] = s_adj[i];
    }
  }
  if (a.fov != nullptr) {
    float* out = a.fov + int64_t(e0) * N * kFovFloats;
    if (((reinterpret_cast<uintptr_t>(out) | uintptr_t(fov_floats * 4)) & 15) == 0) {
      if (tid == 0) { env_bulk_s2g(out, s_fov, uint32_t(fov_floats) * 4u); bulk = true; }
    } else {
      for (int i = tid; i < fov_floats; i += NT) out[i] = s_fov[i];
    }
  }

  if (ep_live) {  // time += 1; done = checkAllInGoal() (:195-197)
    a.time[e0 + tid] = ep_time + 1;
    (a.done_out != nullptr ? a.done_out : a.done)[e0 + tid] = sh.notgoal[tid] ? 0 : 1;
  } else if (a.done_out != nullptr && (a.do_move & 3) && tid < nep) {
    a.done_out[e0 + tid] = ep_done_in ? 1 : 0;      // out-of-place: frozen episodes keep their flag
  }
  if (bulk) {   // the staging area must outlive the copies' reads
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}
