"""Development: build variants of one kernel source (extra -D flags), link each against the other objects of the current
build and time the tensor-core conv stack with it (warm, back to back) -- also checks the emitted FC operand against the
default build bit for bit.     python tools/enc_tc_variants.py "" "-DMAPF_ETC_WAIT=1" "-DMAPF_ETC_L0_NOLO" ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

CHILD = r'''
import os, sys, hashlib
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import torch
from helpers import load_network
from mapf_gnn_b200 import _lib
from mapf_gnn_b200.ops import net_params
lib = _lib.load()
res = []
for model in ("gnn_k2", "baseline"):
    net = load_network(model, num_agents=1).eval()
    dims, packed = net.dims(), net.packed_weights()
    p = net_params(dims); s = _lib.current_stream()
    for rows in ((2560, 5920, 20480, 81920) if model == "gnn_k2" else (20480,)):
        torch.manual_seed(0)
        st = torch.randint(0, 4, (rows, 2, 5, 5), device="cuda").float()
        Ap = torch.zeros((lib.mapf_tc_presplit_a_size(rows, 400),), device="cuda")
        ea = _lib.EncoderFwdArgs(rows, st.data_ptr(), packed.data_ptr(), Ap.data_ptr())
        fn = lambda: lib.mapf_encoder_tc_conv_fwd(p, ea, s)
        for _ in range(5): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(100): fn()
        b.record(); b.synchronize()
        h = hashlib.sha1(Ap.cpu().numpy().tobytes()).hexdigest()[:10]
        res.append(f"{model}/{rows}: {a.elapsed_time(b) / 100 * 1e3:.1f} us [{h}]")
print(" | ".join(res))
'''

def main():
    from mapf_gnn_b200 import _build
    _build.build_library()
    objs = [os.path.join(_build.LIBDIR, f) for f in os.listdir(_build.LIBDIR) if f.endswith(".o") and f != "enc_tc.o"]
    for i, flags in enumerate(sys.argv[1:] or [""]):
        obj, so = f"/tmp/enc_tc_v{i}.o", f"/tmp/libmapf_v{i}.so"
        cmd = [_build.nvcc_path(), *_build.NVCC_FLAGS, *flags.split(), "-c", os.path.join(_build.CSRC, "enc_tc.cu"), "-o", obj]
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        subprocess.run([_build.nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", so, obj, *objs],
                       check=True)
        out = subprocess.run([sys.executable, "-c", CHILD % (ROOT, ROOT)], env=dict(os.environ, MAPF_B200_LIB_OVERRIDE=so),
                             capture_output=True, text=True)
        print(f"[{flags or 'default'}] {out.stdout.strip()} {out.stderr.strip()[-300:]}", flush=True)

main()
