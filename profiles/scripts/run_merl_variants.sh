cat > /tmp/merl_v.py <<'PY'
import torch, numpy as np, sys
import dj_brdf_b200 as djb
dev='cuda'
g=torch.Generator(device=dev).manual_seed(7)
def dirs(n, zmin=0.001):
    z=1.0-(1-zmin)*torch.rand(n,device=dev,generator=g); ph=6.283185307179586*torch.rand(n,device=dev,generator=g)
    r=torch.sqrt(torch.clamp(1-z*z,min=0)); return torch.stack([r*torch.cos(ph),r*torch.sin(ph),z],1).contiguous()
n=100_000_000
wi,wo=dirs(n),dirs(n)
rng=np.random.default_rng(0); tab=rng.uniform(-0.05,3,3*90*90*180); m=djb.merl(tab)
out=m.eval(wi,wo); torch.cuda.synchronize()
a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): out=m.eval(wi,wo)
b.record(); torch.cuda.synchronize()
ms=a.elapsed_time(b)/10
print(sys.argv[1], f"merl eval {n/ms/1e6:.2f} G lookups/s, frac {36*n/ms/1e6/6551.4:.3f}")
PY
for v in 5 101 102 5; do DJB200_MERL_VARIANT=$v PYTHONPATH=$PWD python /tmp/merl_v.py $v; done
