import os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from promptcache_b200 import _native as nat
DEV="cuda"
H=Hkv=4; D=128; q=1
for N, max_ctx in [(900, 1024), (686, 1024), (100, 1024)]:
    Q=torch.randn(H,q,D,device=DEV).half(); K=torch.randn(Hkv,max_ctx,D,device=DEV).half(); V=torch.randn(Hkv,max_ctx,D,device=DEV).half()
    out=torch.empty(q,H*D,dtype=torch.half,device=DEV); out2=torch.empty_like(out)
    ws=torch.zeros(nat.attn_workspace_bytes(H,q,D),dtype=torch.uint8,device=DEV)
    n_dev=torch.full((1,),N,dtype=torch.int32,device=DEV)
    print("case",N,max_ctx,"eager...",flush=True)
    nat.attention(Q,K,V,out,N,q,H,Hkv,D,nat.softmax_scale(D),3,workspace=ws,decode=True)
    torch.cuda.synchronize(); print("  eager ok",flush=True)
    nat.attention(Q,K,V,out2,max_ctx-q,q,H,Hkv,D,nat.softmax_scale(D),3,workspace=ws,decode=True,n_cached_dev=n_dev,impl=nat.PC_ATTN_TCGEN05)
    torch.cuda.synchronize(); print("  dyn ok, max diff",float((out.float()-out2.float()).abs().max()),flush=True)
