"""Dev tool: time the bf16 backward entry point alone at cfg2 (CUDA events, L2 flushed between runs)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctcvr_b200 as C
import ctcvr_b200._lib as _L
if os.environ.get('CTCVR_LIB'): _L.LIB_PATH = os.environ['CTCVR_LIB']   # experiment builds (tools/exp_build.sh)
from ctcvr_b200._lib import call, ptr, query, stream
torch.manual_seed(0)
B,T,U1,D,V,blank=32,250,41,512,412,5
dev='cuda'
e=torch.randn(B,T,D,device=dev); p=torch.randn(B,U1,D,device=dev)
w=torch.randn(V,D,device=dev)/D**0.5; b=torch.zeros(V,device=dev)
tgt=torch.randint(6,V,(B,U1-1),dtype=torch.int32,device=dev)
tl=torch.full((B,),T,dtype=torch.int32,device=dev); ul=torch.full((B,),U1-1,dtype=torch.int32,device=dev)
lse=torch.empty(B,T,U1,device=dev); lpb=torch.empty_like(lse); lpl=torch.empty_like(lse)
ws=torch.empty(query("ctcvr_joint_rnnt_fwd_ws_bytes",B,T,U1,D,V,1),dtype=torch.uint8,device=dev)
call("ctcvr_joint_rnnt_fwd",ptr(e),ptr(p),ptr(w),ptr(b),ptr(tgt),ptr(tl),ptr(ul),ptr(lse),ptr(lpb),ptr(lpl),B,T,U1,D,V,blank,1,ptr(ws),ws.numel(),stream())
al=torch.empty(B,T,U1,device=dev); be=torch.empty_like(al); costs=torch.empty(B,device=dev)
call("ctcvr_rnnt_lattice",ptr(lpb),ptr(lpl),ptr(tl),ptr(ul),ptr(al),ptr(be),ptr(costs),B,T,U1,stream())
gc=torch.full((B,),1.0/B,device=dev)
d_e=torch.empty_like(e); d_p=torch.empty_like(p); d_w=torch.empty_like(w); d_b=torch.empty_like(b)
wsb=torch.empty(query("ctcvr_joint_rnnt_bwd_ws_bytes",B,T,U1,D,V,1),dtype=torch.uint8,device=dev)
def runb():
    call("ctcvr_joint_rnnt_bwd",ptr(e),ptr(p),ptr(w),ptr(b),ptr(tgt),ptr(tl),ptr(ul),ptr(lse),ptr(lpb),ptr(lpl),ptr(al),ptr(be),ptr(costs),ptr(gc),-1.0,ptr(d_e),ptr(d_p),ptr(d_w),ptr(d_b),B,T,U1,D,V,blank,1,ptr(wsb),wsb.numel(),stream())
flush=torch.empty(256<<20,dtype=torch.uint8,device=dev)
for _ in range(3): runb()
ts=[]
for _ in range(10):
    flush.zero_(); a=torch.cuda.Event(enable_timing=True); c=torch.cuda.Event(enable_timing=True)
    a.record(); runb(); c.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(c)*1e3)
ts.sort(); print("%s bwd total us: median %.1f min %.1f" % (os.environ.get("CTCVR_LIB","default"), ts[len(ts)//2], ts[0]))
