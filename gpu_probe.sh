#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -m gpu -k "frontend or golden or 8m" > gpurun_out/fe_test.log 2>&1; echo "exit $?" >> gpurun_out/fe_test.log; tail -4 gpurun_out/fe_test.log
cat > /tmp/fe_time.py <<'PY'
import torch, sys
sys.path.insert(0,'.')
from orbit2_b200 import ops
B,V,gh,gw,heads,hd=8,23,90,180,16,64
x=torch.randn(B,V,180,360,device='cuda'); ts=torch.randn(V,heads,5,device='cuda'); tv=torch.randn(heads,V*5,hd,device='cuda')*0.1
do=torch.randn(B*gh*gw,heads*hd,device='cuda').bfloat16()
for f,name in ((lambda: ops.frontend_fwd(x,ts,tv,2,gh,gw,hd,torch.bfloat16),'fwd'),(lambda: ops.frontend_bwd(x,ts,tv,do,2,gh,gw,hd),'bwd')):
    for _ in range(2): f()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f()
    e1.record(); torch.cuda.synchronize(); print(name, e0.elapsed_time(e1)/5,'ms')
PY
python /tmp/fe_time.py
