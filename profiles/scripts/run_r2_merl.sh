# MERL lookups: where is the ceiling?  Timings of the product path, of the two probes (1: same index math + coalesced gather,
# 2: no index math + random gather) and of the L2-persistence experiment; then ncu counters of the product kernel and of probe 2.
mkdir -p gpurun_out
cat > /tmp/merl_t.py <<'PY'
import sys, os, ctypes as C, numpy as np, torch
sys.path.insert(0, ".")
import dj_brdf_b200 as djb
from dj_brdf_b200 import capi, workloads
lib = capi.load()
n = 100_000_000
g = torch.Generator(device="cuda").manual_seed(1)
def dirs():
    z = 1.0 - 0.999 * torch.rand(n, device="cuda", generator=g); phi = 6.283185307179586 * torch.rand(n, device="cuda", generator=g)
    r = torch.sqrt(torch.clamp(1.0 - z * z, min=0.0)); return torch.stack([r * torch.cos(phi), r * torch.sin(phi), z], dim=1).contiguous()
wi, wo = dirs(), dirs()
out = torch.empty(n, 3, device="cuda")
m = djb.merl(workloads.synthetic_merl_table(0.15))
st = torch.cuda.current_stream(); sp = C.c_void_p(st.cuda_stream)
def run():
    capi.check(lib.djb200_merl_eval(m._h, C.c_void_p(wi.data_ptr()), C.c_void_p(wo.data_ptr()), C.c_int64(n), C.c_void_p(out.data_ptr()), C.c_int(capi.MEM_DEVICE), sp))
for _ in range(3): run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); [run() for _ in range(10)]; b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(f"probe={os.environ.get('DJB200_MERL_PROBE','0')} persist={os.environ.get('DJB200_MERL_PERSIST','0')}: {ms:.4f} ms, {n/ms/1e6:.1f} G lookups/s, {36*n/ms/1e6/6437.9:.3f} of HBM peak")
PY
python /tmp/merl_t.py
DJB200_MERL_PERSIST=1 python /tmp/merl_t.py
DJB200_MERL_PROBE=1 python /tmp/merl_t.py
DJB200_MERL_PROBE=2 python /tmp/merl_t.py
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_sectors.sum,l1tex__m_l1tex2xbar_write_sectors.sum,l1tex__m_xbar2l1tex_read_sectors.sum.pct_of_peak_sustained_elapsed,l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"
ncu --metrics $M --clock-control none -k regex:merl_eval_quad -s 3 -c 1 --csv --log-file gpurun_out/r02_merl_ncu_product.csv python /tmp/merl_t.py > /dev/null 2>&1
DJB200_MERL_PROBE=2 ncu --metrics $M --clock-control none -k regex:merl_eval_quad -s 3 -c 1 --csv --log-file gpurun_out/r02_merl_ncu_probe2.csv python /tmp/merl_t.py > /dev/null 2>&1
DJB200_MERL_PROBE=1 ncu --metrics $M --clock-control none -k regex:merl_eval_quad -s 3 -c 1 --csv --log-file gpurun_out/r02_merl_ncu_probe1.csv python /tmp/merl_t.py > /dev/null 2>&1
python - <<'PY'
import csv
for tag in ("product","probe2","probe1"):
    lines=open(f'gpurun_out/r02_merl_ncu_{tag}.csv').read().split('\n')
    k=[i for i,l in enumerate(lines) if l.startswith('"ID"')][0]
    print(tag)
    for r in csv.DictReader(lines[k:]):
        print('   ', r['Metric Name'], r['Metric Value'], r['Metric Unit'])
PY
