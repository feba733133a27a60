import sys, torch
sys.path.insert(0, ".")
from scoreperformer_b200 import kernels as K
import torch.nn.functional as F
BF16=torch.bfloat16
D,H=256,1024
torch.manual_seed(0)
for n in (333, 700):
    dy=torch.randn(n,D,device="cuda").bfloat16(); u=torch.randn(n,2*H,device="cuda").bfloat16()
    w1=(torch.randn(2*H,D,device="cuda")/16).bfloat16(); w2=(torch.randn(D,H,device="cuda")/32).bfloat16()
    w2t=K.transpose_bf16(w2); db=torch.zeros(2*H,device="cuda")
    dxn,du=K.ffn_bwd(dy,w2t,w1,u,db,0.1,3,in_place=False)
    xn=torch.randn(n,D,device="cuda").bfloat16(); b1=torch.randn(2*H,device="cuda")*0.1; resid=torch.randn(n,D,device="cuda")
    out,uu,hh=K.ffn_fwd(xn,w1,b1,w2,resid,0.1,3)
    K.multi_add([db],[torch.ones_like(db)]); K.multi_copy([db],[torch.zeros_like(db)])
    torch.cuda.synchronize()
print("ffn ok")
