import torch, sys
sys.path.insert(0,'.')
from orbit2_b200 import ops
T,D=129600,1024
x=torch.randn(T,D,device='cuda').bfloat16(); dy=torch.randn(T,D,device='cuda').bfloat16(); dres=torch.randn(T,D,device='cuda').bfloat16()
g=torch.randn(D,device='cuda'); b=torch.randn(D,device='cuda'); dg=torch.zeros(D,device='cuda'); db=torch.zeros(D,device='cuda')
y,m,r=ops.layernorm_fwd(x,g,b)
def t(f,n=10):
    for _ in range(3): f()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
f=t(lambda: ops.layernorm_fwd(x,g,b)); bw=t(lambda: ops.layernorm_bwd(dy,x,g,m,r,dg,db,dres=dres))
print(f"ln fwd {f:.3f} ms {2*T*D*2/f/1e6:.0f} GB/s ; ln bwd {bw:.3f} ms {4*T*D*2/bw/1e6:.0f} GB/s (algorithmic 4 tensors)")
