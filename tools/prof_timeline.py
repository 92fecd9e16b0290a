"""Dev tool: per-role timeline (clock64) of CTA 0 of the tcgen05 forward kernel at cfg2."""
import sys, torch, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctcvr_b200 as C
from ctcvr_b200._lib import call, ptr, query, stream, lib
torch.manual_seed(0)
lib().ctcvr_debug_set_mode(int(os.environ.get('FWD_MODE', '1')))
B,T,U1,D,V,blank=32,250,41,512,412,5
dev='cuda'
e=torch.randn(B,T,D,device=dev); p=torch.randn(B,U1,D,device=dev)
w=torch.randn(V,D,device=dev)/D**0.5; b=torch.zeros(V,device=dev)
tgt=torch.randint(6,V,(B,U1-1),dtype=torch.int32,device=dev)
tl=torch.full((B,),T,dtype=torch.int32,device=dev); ul=torch.full((B,),U1-1,dtype=torch.int32,device=dev)
lse=torch.empty(B,T,U1,device=dev); lpb=torch.empty_like(lse); lpl=torch.empty_like(lse)
ws=torch.empty(query("ctcvr_joint_rnnt_fwd_ws_bytes",B,T,U1,D,V,1),dtype=torch.uint8,device=dev)
prof=torch.zeros(4*2048,dtype=torch.int64,device=dev)
def run():
    call("ctcvr_joint_rnnt_fwd",ptr(e),ptr(p),ptr(w),ptr(b),ptr(tgt),ptr(tl),ptr(ul),ptr(lse),ptr(lpb),ptr(lpl),B,T,U1,D,V,blank,1,ptr(ws),ws.numel(),stream())
for _ in range(3): run()
torch.cuda.synchronize()
lib().ctcvr_debug_set_prof(ptr(prof))
run(); torch.cuda.synchronize()
lib().ctcvr_debug_set_prof(None)
pr=prof.cpu().numpy().reshape(4,2048)
names=['TMA','MMA','EPI','PROD']
t0=min((pr[r][pr[r]!=0] & 0xffffffffffff).min() for r in range(4) if (pr[r]!=0).any())
for r in range(4):
    ev=pr[r][pr[r]!=0]
    tags=(ev>>48); clk=(ev & 0xffffffffffff)-t0
    print(names[r], len(ev))
    # print first 3 tiles worth
    lim={'TMA':48,'MMA':90,'EPI':int(os.environ.get('EPI_LIM', '8')),'PROD':140}[names[r]]
    print(' '.join(f"{int(a)}:{int(c)}" for a,c in zip(tags[:lim],clk[:lim])))
    print(' ... last:', ' '.join(f"{int(a)}:{int(c)}" for a,c in zip(tags[-6:],clk[-6:])))

# ---- backward kernel timeline
al=torch.empty(B,T,U1,device=dev); be=torch.empty_like(al); costs=torch.empty(B,device=dev)
call("ctcvr_rnnt_lattice",ptr(lpb),ptr(lpl),ptr(tl),ptr(ul),ptr(al),ptr(be),ptr(costs),B,T,U1,stream())
gc=torch.full((B,),1.0/B,device=dev)
d_e=torch.empty_like(e); d_p=torch.empty_like(p); d_w=torch.empty_like(w); d_b=torch.empty_like(b)
wsb=torch.empty(query("ctcvr_joint_rnnt_bwd_ws_bytes",B,T,U1,D,V,1),dtype=torch.uint8,device=dev)
def runb():
    call("ctcvr_joint_rnnt_bwd",ptr(e),ptr(p),ptr(w),ptr(b),ptr(tgt),ptr(tl),ptr(ul),ptr(lse),ptr(lpb),ptr(lpl),ptr(al),ptr(be),ptr(costs),ptr(gc),-1.0,ptr(d_e),ptr(d_p),ptr(d_w),ptr(d_b),B,T,U1,D,V,blank,1,ptr(wsb),wsb.numel(),stream())
for _ in range(2): runb()
torch.cuda.synchronize()
prof.zero_()
lib().ctcvr_debug_set_prof(ptr(prof))
runb(); torch.cuda.synchronize()
lib().ctcvr_debug_set_prof(None)
pr=prof.cpu().numpy().reshape(4,2048)
t0=min((pr[r][pr[r]!=0] & 0xffffffffffff).min() for r in range(4) if (pr[r]!=0).any())
print("BWD")
for r in range(4):
    ev=pr[r][pr[r]!=0]
    tags=(ev>>48); clk=(ev & 0xffffffffffff)-t0
    print(names[r], len(ev))
    print(' '.join(f"{int(a)}:{int(c)}" for a,c in zip(tags[:260],clk[:260])))
