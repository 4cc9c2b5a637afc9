"""Development: build variants of ONE kernel source (extra -D flags), link each against the other objects of the current
build and run a timing child with it.
    python tools/kernel_variants.py tc_graph.cu graph "" "-DTG_EPI_COLS=8" ...
children: graph = graph-filter stage alone at C2a / C1 / C3 shapes (CUDA events, L2 flushed), enc = conv stack."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHILD_GRAPH = r'''
import os, sys, hashlib
sys.path.insert(0, %r)
import numpy as np, torch
import bench
import mapf_gnn_b200 as m
from mapf_gnn_b200 import _lib
lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = []
for name in ("c2a", "c1", "c3"):
    w = bench.WORKLOADS[name]; E, N = w["E"], w["N"]
    torch.manual_seed(0)
    net = m.build_network(bench.model_config(w))
    net.load_state_dict({k: torch.from_numpy(v) for k, v in bench.load_weights(w["model"]).items()})
    net = net.cuda().eval()
    dims = net.dims(); F, K_, G = dims[7], dims[8], dims[9]
    skip = 1 if net.message else 0; KT = (K_ + skip) * F
    gf = net.GFL[0]
    Wp = torch.empty((lib.mapf_tc_packed_b_size(G, KT),), device="cuda")
    W2 = gf.W2.detach().contiguous() if skip else None
    lib.mapf_tc_pack_b(gf.W.detach().contiguous().data_ptr(), K_ * F, None if W2 is None else W2.data_ptr(), F if skip else 0, G, Wp.data_ptr(), _lib.current_stream())
    aw = net.actionMLP[0].weight.detach().contiguous(); ab = net.actionMLP[0].bias.detach().contiguous()
    x = torch.randn(E * N, F, device="cuda"); adj = (torch.rand(E, N, N, device="cuda") < 0.4).float()
    lg = torch.zeros(E * N, 5, device="cuda"); ac = torch.zeros(E * N, dtype=torch.int32, device="cuda")
    fn = lambda: lib.mapf_tc_graph_head(E, N, F, G, K_, 0 if skip else 1, skip, x.data_ptr(), adj.data_ptr(), Wp.data_ptr(), aw.data_ptr(), ab.data_ptr(), 5, lg.data_ptr(), ac.data_ptr(), _lib.current_stream())
    ts = []
    for i in range(15):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        if i >= 5: ts.append(a.elapsed_time(b))
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): fn()
    b.record(); b.synchronize()
    h = hashlib.sha1(lg.cpu().numpy().tobytes()).hexdigest()[:8]
    out.append(f"{name}: flushed {np.median(ts) * 1e3:.1f} warm {a.elapsed_time(b) / 50 * 1e3:.1f} us [{h}]")
print(" | ".join(out))
'''

CHILD_TAPS = r'''
import os, sys, runpy
sys.argv = ["taps_time.py"]
runpy.run_path(os.path.join(%r, "tools", "taps_time.py"), run_name="__main__")
'''

CHILD_MIX = r'''
import os, sys, runpy
sys.argv = ["mix_time.py"]
runpy.run_path(os.path.join(%r, "tools", "mix_time.py"), run_name="__main__")
'''

def main():
    from mapf_gnn_b200 import _build
    _build.build_library()
    src, child = sys.argv[1], sys.argv[2]
    stem = os.path.basename(src)[:-3]
    if len(sys.argv) > 3 and sys.argv[3].startswith("--replaces="):      # an alternative source standing in for csrc/<name>.cu
        stem = sys.argv.pop(3).split("=", 1)[1][:-3]
    objs = [os.path.join(_build.LIBDIR, f) for f in os.listdir(_build.LIBDIR) if f.endswith(".o") and f != stem + ".o"]
    code = {"graph": CHILD_GRAPH, "taps": CHILD_TAPS, "mix": CHILD_MIX}[child] % (ROOT,)
    for i, flags in enumerate(sys.argv[3:] or [""]):
        obj, so = f"/tmp/{stem}_v{i}.o", f"/tmp/libmapf_{stem}_v{i}.so"
        path = src if os.path.isabs(src) or os.path.exists(src) else os.path.join(_build.CSRC, src)
        cmd = [_build.nvcc_path(), *_build.NVCC_FLAGS, *flags.split(), "-c", path, "-o", obj]
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        subprocess.run([_build.nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", so, obj, *objs],
                       check=True)
        out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, MAPF_B200_LIB_OVERRIDE=so),
                             capture_output=True, text=True)
        print(f"[{flags or 'default'}] {' | '.join(out.stdout.strip().splitlines())} {out.stderr.strip()[-300:]}", flush=True)

main()
