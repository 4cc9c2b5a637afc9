import sys,os
sys.path.insert(0,"/root/repo"); sys.path.insert(0,"/root/repo/tests")
import torch
from helpers import load_network
from mapf_gnn_b200 import _lib
from mapf_gnn_b200.ops import net_params
lib=_lib.load(); net=load_network("gnn_k2",num_agents=1).eval(); p=net_params(net.dims()); packed=net.packed_weights(); s=_lib.current_stream()
def timeit(fn,iters=100):
    for _ in range(5): fn()
    torch.cuda.synchronize(); a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); a.record()
    for _ in range(iters): fn()
    b.record(); b.synchronize(); return a.elapsed_time(b)/iters*1e3
for rows in (2960, 5920, 11840, 20480, 81920):
    st=torch.randint(0,4,(rows,2,5,5),device="cuda").float(); Ap=torch.empty((lib.mapf_tc_presplit_a_size(rows,400),),device="cuda")
    ea=_lib.EncoderFwdArgs(rows,st.data_ptr(),packed.data_ptr(),Ap.data_ptr())
    out=[]
    for skew in ("0","200","400","700","1000","1400","2000"):
        os.environ["MAPF_B200_ETC_SKEW"]=skew
        out.append(f"{skew}:{timeit(lambda: lib.mapf_encoder_tc_conv_fwd(p,ea,s)):.1f}")
    print(rows," ".join(out))
