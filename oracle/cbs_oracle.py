"""Test infrastructure -- CPU restatement of the reference's CBS expert in plain Python (tuples instead of classes).

Pinned by ``tests/golden/cbs_cases.npz`` (plans of the unmodified reference, ``tests/golden/make_golden_cbs.py``).  Only
``tests/`` may import this; the product is ``mapf_gnn_b200/csrc/cbs.cu`` behind ``mapf_gnn_b200/cbs.py``.

Follows cbs/a_star.py:29-102 (low level) and cbs/cbs.py:184-245,326-410 (conflicts, constraints, high level).  A state
is ``(t, x, y)``, a vertex constraint the same triple, an edge constraint ``(t, x1, y1, x2, y2)``.
"""
import heapq
from itertools import combinations, count

MOVES = ((0, 0), (0, 1), (0, -1), (-1, 0), (1, 0))       # cbs.py:159-182: wait, up, down, left, right


def low_level(dims, obstacles, start, goal, vertex, edge, max_pops=10000):
    """a_star.py:29-102 -> list of (x, y) per time step, or None."""
    first = (0, start[0], start[1])
    tick = count()
    heap = [(abs(start[0] - goal[0]) + abs(start[1] - goal[1]), next(tick), first)]
    open_set, closed, parent, g = {first}, set(), {}, {first: 0}
    pops = 0
    while heap and pops < max_pops:                       # a_star.py:62
        pops += 1
        _, _, cur = heapq.heappop(heap)
        if cur not in open_set:                           # a_star.py:69
            continue
        open_set.remove(cur)
        t, x, y = cur
        if (x, y) == tuple(goal):                         # cbs.py:281-283: the goal test ignores time
            path = [cur]
            while path[-1] in parent:
                path.append(parent[path[-1]])
            return [(s[1], s[2]) for s in reversed(path)]
        closed.add(cur)
        for k, (dx, dy) in enumerate(MOVES):
            n = (t + 1, x + dx, y + dy)
            if not (0 <= n[1] < dims[0] and 0 <= n[2] < dims[1]) or n in vertex or (n[1], n[2]) in obstacles:
                continue                                  # cbs.py:254-263
            if k and (t, x, y, n[1], n[2]) in edge:       # cbs.py:265-270 (the wait is not a transition check)
                continue
            if n in closed:
                continue
            if n not in g or g[cur] + 1 < g[n]:           # a_star.py:88-100
                parent[n], g[n] = cur, g[cur] + 1
                open_set.add(n)
                heapq.heappush(heap, (g[n] + abs(n[1] - goal[0]) + abs(n[2] - goal[1]), next(tick), n))
    return None


def first_conflict(paths):
    """cbs.py:184-215 -> None | ('v', t, i, j, cell) | ('e', t, i, j, cell_a, cell_b)."""
    at = lambda p, t: p[t] if t < len(p) else p[-1]
    for t in range(max(len(p) for p in paths)):
        for i, j in combinations(range(len(paths)), 2):
            if at(paths[i], t) == at(paths[j], t):
                return ("v", t, i, j, at(paths[i], t))
        for i, j in combinations(range(len(paths)), 2):
            if at(paths[i], t) == at(paths[j], t + 1) and at(paths[i], t + 1) == at(paths[j], t):
                return ("e", t, i, j, at(paths[i], t), at(paths[i], t + 1))
    return None


def search(dims, starts, goals, obstacles, max_nodes=5000, max_pops=10000):
    """cbs.py:326-410 -> (paths, cost, expanded) or (None, 0, expanded)."""
    obstacles = {tuple(o) for o in obstacles}
    n = len(starts)
    plan = lambda a, vc, ec: low_level(dims, obstacles, starts[a], goals[a], vc, ec, max_pops)
    root_paths = []
    for a in range(n):
        p = plan(a, frozenset(), frozenset())
        if p is None:
            return None, 0, 0
        root_paths.append(p)
    tick = count()
    cons = [(frozenset(), frozenset())] * n
    heap = [(sum(map(len, root_paths)), next(tick), root_paths, cons)]
    closed = set()
    pops = 0
    while heap and pops < max_nodes:
        pops += 1
        cost, _, paths, cons = heapq.heappop(heap)
        key = tuple(tuple(p) for p in paths)
        if key in closed:
            continue
        closed.add(key)
        cf = first_conflict(paths)
        if cf is None:
            return paths, cost, len(closed)
        for side, agent in enumerate((cf[2], cf[3])):
            vc, ec = cons[agent]
            if cf[0] == "v":
                vc = vc | {(cf[1], cf[4][0], cf[4][1])}
            else:
                a, b = (cf[4], cf[5]) if side == 0 else (cf[5], cf[4])
                ec = ec | {(cf[1], a[0], a[1], b[0], b[1])}
            child_cons = list(cons)
            child_cons[agent] = (vc, ec)
            # the reference re-plans every agent (cbs.py:399); the others' constraints are unchanged and A* is deterministic
            child_paths = []
            for k in range(n):
                p = plan(k, *child_cons[k])
                if p is None:
                    break
                child_paths.append(p)
            else:
                heapq.heappush(heap, (sum(map(len, child_paths)), next(tick), child_paths, child_cons))
    return None, 0, len(closed)
