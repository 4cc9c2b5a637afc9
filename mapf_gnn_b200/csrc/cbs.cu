// Conflict-Based Search expert (reference cbs/cbs.py:148-410, cbs/a_star.py:12-102): the offline label
// generator behind the imitation data set.  Host code only -- a branchy best-first search over a handful of agents has no
// data-parallel shape, so the device is not involved; what this file buys is the native search (the reference spends its
// time in Python object hashing) and mapf_cbs_solve_batch, which solves independent instances on all host cores.
//
// The search is the reference's, decision for decision, so that the plans are IDENTICAL to the ones cbs.py returns:
//   low level   space-time A* over (t, x, y); neighbours in the order wait, y+1, y-1, x-1, x+1 (cbs.py:159-182); a state
//               is valid inside the grid, off the obstacles and off the agent's vertex constraints, a move also off its
//               edge constraints (cbs.py:254-270); Manhattan heuristic (cbs.py:275-279); the open list is ordered by
//               (f, push counter) (a_star.py:44-58), every improvement pushes a new entry and stale entries are skipped
//               (a_star.py:66-70,88-100); at most `max_low_level_iterations` pops (a_star.py:59-62); the goal test ignores
//               time (cbs.py:281-283).
//   high level  nodes ordered by (cost, push counter) (cbs.py:345,407); a node whose whole solution was expanded before is
//               skipped (cbs.py:361-365); the first conflict is the earliest time step's first vertex conflict over the
//               agent pairs in order, else that step's first edge swap (cbs.py:184-215, paths padded with their last state
//               cbs.py:247-251); a vertex conflict constrains both agents at (t, cell), an edge conflict constrains agent 1
//               on (t, a->b) and agent 2 on (t, b->a) (cbs.py:217-245); children in the order agent 1, agent 2; at most
//               `max_high_level_iterations` pops (cbs.py:352-357).
// The reference re-plans EVERY agent of a child node (cbs.py:399 -> compute_solution); A* is deterministic and an agent's
// path depends on its own constraints only, so re-planning just the newly constrained agent gives the same node.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <memory>
#include <queue>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "mapf_b200.h"

namespace {

struct Cell { int32_t x, y; };
struct St { int32_t t, x, y; };
inline bool operator==(const St& a, const St& b) { return a.t == b.t && a.x == b.x && a.y == b.y; }
struct Edge { int32_t t, x1, y1, x2, y2; };
inline bool operator==(const Edge& a, const Edge& b) {
  return a.t == b.t && a.x1 == b.x1 && a.y1 == b.y1 && a.x2 == b.x2 && a.y2 == b.y2;
}
inline uint64_t mix(uint64_t h, uint64_t v) {
  h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
  return h;
}
struct StHash {
  size_t operator()(const St& s) const { return size_t(mix(mix(mix(0, uint32_t(s.t)), uint32_t(s.x)), uint32_t(s.y))); }
};
struct EdgeHash {
  size_t operator()(const Edge& e) const {
    return size_t(mix(mix(mix(mix(mix(0, uint32_t(e.t)), uint32_t(e.x1)), uint32_t(e.y1)), uint32_t(e.x2)), uint32_t(e.y2)));
  }
};
struct CellHash {
  size_t operator()(const Cell& c) const { return size_t(mix(mix(0, uint32_t(c.x)), uint32_t(c.y))); }
};
inline bool operator==(const Cell& a, const Cell& b) { return a.x == b.x && a.y == b.y; }

struct Constraints {                       // cbs.py:128-145
  std::unordered_set<St, StHash> vertex;
  std::unordered_set<Edge, EdgeHash> edge;
};

struct Grid {
  int32_t w, h;
  std::vector<uint8_t> blocked;            // [h][w]: 1 = obstacle (obstacles outside the grid can never be stepped on anyway)
  void block(int32_t x, int32_t y) {
    if (x >= 0 && x < w && y >= 0 && y < h) blocked[int64_t(y) * w + x] = 1;
  }
};

using Path = std::vector<Cell>;            // index = time step

// a_star.py:29-102.  Returns false when the open list runs dry or the pop budget is spent.
bool a_star(const Grid& g, const Constraints& c, Cell start, Cell goal, int max_pops, Path* out) {
  struct Node { St s; int32_t gscore; int32_t parent; bool open, closed; };
  std::vector<Node> nodes;
  // state -> node: a dense [time step][W x H] table of node indices in a per-thread scratch buffer that is never
  // cleared -- an entry counts only if its stamp is this call's generation (states outside the grid -- only a start cell
  // can be -- go through the hash map)
  const int64_t cells = int64_t(g.w) * g.h;
  thread_local std::vector<int32_t> tab_idx;
  thread_local std::vector<uint32_t> tab_gen;
  thread_local uint32_t generation = 0;
  if (++generation == 0) { std::fill(tab_gen.begin(), tab_gen.end(), 0u); generation = 1; }
  std::unordered_map<St, int32_t, StHash> outside;
  auto inside = [&](const St& s) { return s.x >= 0 && s.x < g.w && s.y >= 0 && s.y < g.h; };
  auto slot_of = [&](const St& s) { return int64_t(s.t) * cells + int64_t(s.y) * g.w + s.x; };
  auto node_of = [&](const St& s) -> int32_t {          // -1: never seen (no g_score entry)
    if (!inside(s)) {
      auto it = outside.find(s);
      return it == outside.end() ? -1 : it->second;
    }
    const int64_t k = slot_of(s);
    return k < int64_t(tab_gen.size()) && tab_gen[k] == generation ? tab_idx[k] : -1;
  };
  auto remember = [&](const St& s, int32_t idx) {
    if (!inside(s)) { outside.emplace(s, idx); return; }
    const int64_t k = slot_of(s);
    if (k >= int64_t(tab_gen.size())) {
      const size_t want = size_t(k + 1 + 16 * cells);
      tab_gen.resize(want, 0u);
      tab_idx.resize(want, -1);
    }
    tab_gen[k] = generation;
    tab_idx[k] = idx;
  };
  struct Entry { int32_t f; int64_t count; int32_t node; };
  struct Later { bool operator()(const Entry& a, const Entry& b) const { return a.f != b.f ? a.f > b.f : a.count > b.count; } };
  std::priority_queue<Entry, std::vector<Entry>, Later> heap;
  auto heuristic = [&](int32_t x, int32_t y) { return (x > goal.x ? x - goal.x : goal.x - x) + (y > goal.y ? y - goal.y : goal.y - y); };
  const bool any_vertex = !c.vertex.empty(), any_edge = !c.edge.empty();
  auto state_valid = [&](const St& s) {
    return s.x >= 0 && s.x < g.w && s.y >= 0 && s.y < g.h && !(any_vertex && c.vertex.find(s) != c.vertex.end()) &&
           !g.blocked[int64_t(s.y) * g.w + s.x];
  };
  int64_t counter = 0;
  nodes.push_back(Node{St{0, start.x, start.y}, 0, -1, true, false});
  remember(nodes[0].s, 0);
  heap.push(Entry{heuristic(start.x, start.y), counter++, 0});
  static const int dx[5] = {0, 0, 0, -1, 1}, dy[5] = {0, 1, -1, 0, 0};
  int pops = 0;
  while (!heap.empty() && pops < max_pops) {
    ++pops;
    const int32_t cur = heap.top().node;
    heap.pop();
    if (!nodes[cur].open) continue;                      // a stale duplicate entry
    nodes[cur].open = false;
    const St s = nodes[cur].s;
    if (s.x == goal.x && s.y == goal.y) {
      std::vector<Cell> rev;
      for (int32_t n = cur; n >= 0; n = nodes[n].parent) rev.push_back(Cell{nodes[n].s.x, nodes[n].s.y});
      out->assign(rev.rbegin(), rev.rend());
      return true;
    }
    nodes[cur].closed = true;
    for (int k = 0; k < 5; ++k) {
      const St n{s.t + 1, s.x + dx[k], s.y + dy[k]};
      if (!state_valid(n)) continue;
      if (k > 0 && any_edge && c.edge.find(Edge{s.t, s.x, s.y, n.x, n.y}) != c.edge.end()) continue;
      int32_t ni = node_of(n);
      if (ni >= 0 && nodes[ni].closed) continue;
      const int32_t tentative = nodes[cur].gscore + 1;
      if (ni < 0 || tentative < nodes[ni].gscore) {
        if (ni < 0) {
          ni = int32_t(nodes.size());
          nodes.push_back(Node{n, tentative, cur, true, false});
          remember(n, ni);
        } else {
          nodes[ni].gscore = tentative;
          nodes[ni].parent = cur;
          nodes[ni].open = true;
        }
        heap.push(Entry{tentative + heuristic(n.x, n.y), counter++, ni});
      }
    }
  }
  return false;
}

struct Conflict { int type; int32_t t; int a1, a2; Cell l1, l2; };   // type 0: none, 1: vertex, 2: edge

inline Cell at(const Path& p, int32_t t) { return t < int32_t(p.size()) ? p[t] : p.back(); }

Conflict first_conflict(const std::vector<Path>& sol) {              // cbs.py:184-215
  size_t max_t = 0;
  for (const Path& p : sol) max_t = p.size() > max_t ? p.size() : max_t;
  const int A = int(sol.size());
  for (int32_t t = 0; t < int32_t(max_t); ++t) {
    for (int i = 0; i < A; ++i)
      for (int j = i + 1; j < A; ++j)
        if (at(sol[i], t) == at(sol[j], t)) return Conflict{1, t, i, j, at(sol[i], t), Cell{-1, -1}};
    for (int i = 0; i < A; ++i)
      for (int j = i + 1; j < A; ++j) {
        const Cell a0 = at(sol[i], t), a1 = at(sol[i], t + 1), b0 = at(sol[j], t), b1 = at(sol[j], t + 1);
        if (a0 == b1 && a1 == b0) return Conflict{2, t, i, j, a0, a1};
      }
  }
  return Conflict{0, 0, 0, 0, Cell{0, 0}, Cell{0, 0}};
}

struct HighNode {
  std::vector<std::shared_ptr<const Path>> solution;
  std::vector<std::shared_ptr<const Constraints>> constraints;     // unchanged agents share their parent's sets (cbs.py:384-396)
  int64_t cost;
};

struct Result {
  bool solved = false;
  std::vector<Path> paths;
  int64_t cost = 0;
  int32_t expanded = 0;       // nodes that entered the closed set
  int32_t generated = 0;      // nodes pushed
};

Result cbs_search(const Grid& g, const std::vector<Cell>& starts, const std::vector<Cell>& goals, int max_high, int max_low) {
  Result res;
  const int A = int(starts.size());
  auto root = std::make_shared<HighNode>();
  root->cost = 0;
  auto empty = std::make_shared<const Constraints>();
  for (int a = 0; a < A; ++a) {
    Path p;
    if (!a_star(g, *empty, starts[a], goals[a], max_low, &p)) return res;     // cbs.py:339-341
    root->cost += int64_t(p.size());
    root->solution.push_back(std::make_shared<const Path>(std::move(p)));
    root->constraints.push_back(empty);
  }
  struct Entry { int64_t cost; int64_t count; std::shared_ptr<HighNode> node; };
  struct Later { bool operator()(const Entry& a, const Entry& b) const { return a.cost != b.cost ? a.cost > b.cost : a.count > b.count; } };
  std::priority_queue<Entry, std::vector<Entry>, Later> open;
  int64_t counter = 0;
  open.push(Entry{root->cost, counter++, root});
  res.generated = 1;
  // closed set over whole solutions (cbs.py:412-419): the exact serialisation, so hash collisions cannot merge nodes
  struct VecHash {
    size_t operator()(const std::vector<int32_t>& v) const {
      uint64_t h = v.size();
      for (int32_t x : v) h = mix(h, uint32_t(x));
      return size_t(h);
    }
  };
  std::unordered_set<std::vector<int32_t>, VecHash> closed;
  int iterations = 0;
  while (!open.empty() && iterations < max_high) {
    ++iterations;
    std::shared_ptr<HighNode> P = open.top().node;
    open.pop();
    std::vector<int32_t> key;
    for (const auto& p : P->solution) {
      key.push_back(int32_t(p->size()));
      for (const Cell& c : *p) { key.push_back(c.x); key.push_back(c.y); }
    }
    if (!closed.insert(std::move(key)).second) continue;
    ++res.expanded;
    std::vector<Path> sol;
    sol.reserve(A);
    for (const auto& p : P->solution) sol.push_back(*p);
    const Conflict cf = first_conflict(sol);
    if (cf.type == 0) {
      res.solved = true;
      res.paths = std::move(sol);
      res.cost = P->cost;
      return res;
    }
    for (int side = 0; side < 2; ++side) {
      const int agent = side == 0 ? cf.a1 : cf.a2;
      auto cons = std::make_shared<Constraints>(*P->constraints[agent]);
      if (cf.type == 1) {
        cons->vertex.insert(St{cf.t, cf.l1.x, cf.l1.y});
      } else if (side == 0) {
        cons->edge.insert(Edge{cf.t, cf.l1.x, cf.l1.y, cf.l2.x, cf.l2.y});
      } else {
        cons->edge.insert(Edge{cf.t, cf.l2.x, cf.l2.y, cf.l1.x, cf.l1.y});
      }
      Path p;
      if (!a_star(g, *cons, starts[agent], goals[agent], max_low, &p)) continue;     // cbs.py:401-402
      auto child = std::make_shared<HighNode>(*P);
      child->cost = P->cost - int64_t(P->solution[agent]->size()) + int64_t(p.size());
      child->solution[agent] = std::make_shared<const Path>(std::move(p));
      child->constraints[agent] = cons;
      open.push(Entry{child->cost, counter++, child});
      ++res.generated;
    }
  }
  return res;
}

int check_args(const mapf_cbs_args* a) {
  if (a == nullptr) return MAPF_ERR_NULL;
  if (a->width <= 0 || a->height <= 0 || a->num_agents <= 0 || a->num_obstacles < 0 || a->max_path < 0) return MAPF_ERR_SHAPE;
  if (!a->starts || !a->goals || !a->path_len || (a->num_obstacles > 0 && !a->obstacles)) return MAPF_ERR_NULL;
  if (a->max_path > 0 && !a->paths) return MAPF_ERR_NULL;
  return MAPF_OK;
}

void solve_one(mapf_cbs_args* a) {
  Grid g{a->width, a->height, std::vector<uint8_t>(size_t(a->width) * a->height, 0)};
  for (int i = 0; i < a->num_obstacles; ++i) g.block(a->obstacles[2 * i], a->obstacles[2 * i + 1]);
  std::vector<Cell> starts(a->num_agents), goals(a->num_agents);
  for (int i = 0; i < a->num_agents; ++i) {
    starts[i] = Cell{a->starts[2 * i], a->starts[2 * i + 1]};
    goals[i] = Cell{a->goals[2 * i], a->goals[2 * i + 1]};
  }
  const int max_high = a->max_high_level_iterations > 0 ? a->max_high_level_iterations : 5000;
  const int max_low = a->max_low_level_iterations > 0 ? a->max_low_level_iterations : 10000;
  const Result r = cbs_search(g, starts, goals, max_high, max_low);
  a->solved = r.solved ? 1 : 0;
  a->cost = r.solved ? r.cost : 0;
  a->expanded = r.expanded;
  a->generated = r.generated;
  a->truncated = 0;
  for (int i = 0; i < a->num_agents; ++i) {
    const int32_t len = r.solved ? int32_t(r.paths[i].size()) : 0;
    a->path_len[i] = len;
    if (len > a->max_path) a->truncated = 1;
    const int32_t n = len < a->max_path ? len : a->max_path;
    for (int32_t t = 0; t < n; ++t) {
      a->paths[(int64_t(i) * a->max_path + t) * 2] = r.paths[i][t].x;
      a->paths[(int64_t(i) * a->max_path + t) * 2 + 1] = r.paths[i][t].y;
    }
  }
}

}  // namespace

extern "C" int mapf_cbs_solve(mapf_cbs_args* a) {
  const int rc = check_args(a);
  if (rc != MAPF_OK) return rc;
  solve_one(a);
  return MAPF_OK;
}

extern "C" int mapf_cbs_solve_batch(mapf_cbs_args* items, int32_t count, int32_t threads) {
  if (count < 0) return MAPF_ERR_SHAPE;
  if (count == 0) return MAPF_OK;
  if (items == nullptr) return MAPF_ERR_NULL;
  for (int32_t i = 0; i < count; ++i) {
    const int rc = check_args(&items[i]);
    if (rc != MAPF_OK) return rc;
  }
  int hw = int(std::thread::hardware_concurrency());
  if (hw <= 0) hw = 1;
  int nt = threads > 0 ? threads : hw;
  if (nt > count) nt = count;
  std::atomic<int32_t> next{0};
  auto work = [&]() {
    for (int32_t i = next.fetch_add(1); i < count; i = next.fetch_add(1)) solve_one(&items[i]);
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; ++t) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
  return MAPF_OK;
}

// One low-level search with caller-given constraints (the reference's Environment.a_star.search under env.constraints,
// a_star.py:29 + cbs.py:254-270): vertex constraints [nv][3] = (t, x, y), edge constraints [ne][5] = (t, x1, y1, x2, y2).
extern "C" int mapf_cbs_a_star(mapf_cbs_args* a, int32_t agent, const int32_t* vertex_constraints, int32_t nv,
                               const int32_t* edge_constraints, int32_t ne, int32_t* found) {
  const int rc = check_args(a);
  if (rc != MAPF_OK) return rc;
  if (found == nullptr || (nv > 0 && !vertex_constraints) || (ne > 0 && !edge_constraints)) return MAPF_ERR_NULL;
  if (agent < 0 || agent >= a->num_agents || nv < 0 || ne < 0) return MAPF_ERR_SHAPE;
  Grid g{a->width, a->height, std::vector<uint8_t>(size_t(a->width) * a->height, 0)};
  for (int i = 0; i < a->num_obstacles; ++i) g.block(a->obstacles[2 * i], a->obstacles[2 * i + 1]);
  Constraints c;
  for (int i = 0; i < nv; ++i) c.vertex.insert(St{vertex_constraints[3 * i], vertex_constraints[3 * i + 1], vertex_constraints[3 * i + 2]});
  for (int i = 0; i < ne; ++i)
    c.edge.insert(Edge{edge_constraints[5 * i], edge_constraints[5 * i + 1], edge_constraints[5 * i + 2], edge_constraints[5 * i + 3],
                       edge_constraints[5 * i + 4]});
  Path p;
  const int max_low = a->max_low_level_iterations > 0 ? a->max_low_level_iterations : 10000;
  const bool ok = a_star(g, c, Cell{a->starts[2 * agent], a->starts[2 * agent + 1]}, Cell{a->goals[2 * agent], a->goals[2 * agent + 1]},
                         max_low, &p);
  *found = ok ? 1 : 0;
  const int32_t len = ok ? int32_t(p.size()) : 0;
  a->path_len[agent] = len;
  const int32_t n = len < a->max_path ? len : a->max_path;
  for (int32_t t = 0; t < n; ++t) {
    a->paths[(int64_t(agent) * a->max_path + t) * 2] = p[t].x;
    a->paths[(int64_t(agent) * a->max_path + t) * 2 + 1] = p[t].y;
  }
  return MAPF_OK;
}
