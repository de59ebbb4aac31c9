cat > /tmp/pcie.py <<'PY'
import torch, time
dev='cuda'
for mb in (16, 256, 2048):
    n=mb*1024*1024
    h=torch.empty(n,dtype=torch.uint8).pin_memory(); d=torch.empty(n,dtype=torch.uint8,device=dev)
    for name,(src,dst) in {'H2D':(h,d),'D2H':(d,h)}.items():
        dst.copy_(src,non_blocking=True); torch.cuda.synchronize()
        t=time.perf_counter()
        for _ in range(5): dst.copy_(src,non_blocking=True)
        torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
        print(f"{name} {mb} MB pinned: {n/dt/1e9:.1f} GB/s")
# bidirectional
n=1024*1024*1024
h1=torch.empty(n,dtype=torch.uint8).pin_memory(); h2=torch.empty(n,dtype=torch.uint8).pin_memory()
d1=torch.empty(n,dtype=torch.uint8,device=dev); d2=torch.empty(n,dtype=torch.uint8,device=dev)
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d1.copy_(h1,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
print(f"bidirectional 1 GB each way: {n/dt/1e9:.1f} GB/s per direction")
hp=torch.empty(n,dtype=torch.uint8)
torch.cuda.synchronize(); t=time.perf_counter(); hp.copy_(d1); torch.cuda.synchronize(); print(f"D2H pageable: {n/(time.perf_counter()-t)/1e9:.1f} GB/s")
t=time.perf_counter(); x=torch.empty(4*n,dtype=torch.uint8).pin_memory(); print(f"pin 4 GB: {time.perf_counter()-t:.2f} s")
PY
python /tmp/pcie.py
nvidia-smi -q | grep -A4 "GPU Link Info" | head -12
