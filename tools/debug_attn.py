# NOTE: needs a debug build of the library: PC_NVCC_EXTRA=-DPC_DEBUG python tools/debug_attn.py  (product builds compile the env / trace hooks out)
import os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from promptcache_b200 import _native as nat
DEV="cuda"
def ref_fp32(Q,K,V,N,q):
    H,_,D=Q.shape
    k=K[:, :N+q].float().repeat_interleave(H//K.shape[0],0); v=V[:, :N+q].float().repeat_interleave(H//K.shape[0],0)
    s=torch.einsum("hqd,hkd->hqk",Q.float(),k)/math.sqrt(D)
    mask=torch.ones(q,N+q,dtype=torch.bool,device=DEV).tril(N)
    s=s.masked_fill(~mask,float("-inf"))
    return torch.einsum("hqk,hkd->hqd",torch.softmax(s,-1),v)   # [H,q,D]
for (H,Hkv,D,q,N,splits) in [(2,2,128,256,192,1),(2,2,128,256,64,1),(2,2,128,256,0,1),(2,2,128,1024,2048,1)]:
    torch.manual_seed(0)
    mc=((N+q+127)//128)*128
    Q=torch.randn(H,q,D,device=DEV).half(); K=torch.randn(Hkv,mc,D,device=DEV).half(); V=torch.randn(Hkv,mc,D,device=DEV).half()
    out=torch.empty(q,H*D,dtype=torch.half,device=DEV)
    ws=torch.zeros(max(nat.attn_workspace_bytes(H,q,D)*4,1<<20),dtype=torch.uint8,device=DEV)
    os.environ["PC_ATTN_SPLITS"]=str(splits)
    nat.attention(Q,K,V,out,N,q,H,Hkv,D,1/math.sqrt(D),0,workspace=ws)
    torch.cuda.synchronize()
    r=ref_fp32(Q,K,V,N,q)
    o=out.view(q,H,D).permute(1,0,2).float()
    err=(o-r).abs().amax(-1)  # [H,q]
    blocks=err.view(H,-1,min(128,q)).amax(-1)
    print(f"H{H} q{q} N{N} splits{splits}: max err {err.max().item():.4f}; per (head, 128-row block):", [[round(x,3) for x in b] for b in blocks.tolist()])
