# memcheck + racecheck + synccheck of the section-8f kernels on small inputs (round 1, session e)
mkdir -p gpurun_out
cat > /tmp/san_e.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
import dj_brdf_b200 as djb
from tests import cases
wi, wo, u = cases.pairs(3001)
ewi, ewo, eu = cases.edge_pairs()
wi = np.concatenate([wi, ewi]); wo = np.concatenate([wo, ewo]); u = np.concatenate([u, eu])
n = len(wi)
d = [torch.from_numpy(x).cuda() for x in (wi, wo, u)]
E, alpha = cases.lean_texels(n)
tE, ta = torch.from_numpy(E).cuda(), torch.from_numpy(alpha).cuda()
for fr in (djb.fresnel.ideal(), djb.fresnel.schlick([0.9, 0.5, 0.2]), djb.fresnel.unpolarized([1.5, 1.8, 2.4])):
    for cls in (djb.beckmann, djb.ggx):
        b = cls(fr)
        b.evalp_lean(d[0], d[1], tE, ta); b.pdf_lean(d[0], d[1], tE, ta); b.evalp_is_lean(d[2], d[1], tE, ta)
        b.evalp_lean(wi, wo, E, np.array([0.1, 0.3, 0.4], np.float32), lean_filtering=False)
djb.beckmann.lean_shading_params(tE, ta)
for name in ("gold-metallic-paint", "white-fabric"):
    djb.sgd(name).eval(d[0], d[1]); djb.sgd(name).eval(wi[:5], wo[:5])
djb.abc("aluminium").eval(d[0], d[1])
djb.tabular.fit_batch([djb.sgd("alum-bronze"), djb.abc("alum-bronze")], 24)
t = djb.tabular_anisotropic(djb.utia(cases.random_utia_table(3)), 9, 11)
t.sampling_tables(); t.sample(d[2], d[1]); t.evalp_is(d[2], d[1]); t.eval(d[0], d[1])
t2 = djb.tabular_anisotropic(djb.ggx(), 33, 7); t2.sample(u, wo)
iso = djb.tabular(djb.beckmann(), 24)
x = torch.rand(1000, device="cuda") * 0.98 + 0.01
for b in (iso, djb.ggx(), djb.beckmann()):
    b.p22_radial(x); b.sigma_std_radial(x); b.cdf_radial(x); b.qf_radial(x)
rng = np.random.default_rng(1)
for h, w in ((33, 17), (64, 128), (1, 4), (5, 1)):
    dm = rng.integers(0, 256, (h, w), dtype=np.uint8)
    djb.dmap2nmap(torch.from_numpy(dm).cuda(), 0.1); djb.dmap2nmap(dm, 0.1)
torch.cuda.synchronize()
print("sanitizer workload done")
PY
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  compute-sanitizer --tool $tool python /tmp/san_e.py 2>&1 | grep -E "SUMMARY|sanitizer workload|Error|hazard" | head -8
done > gpurun_out/sanitizer_e.log 2>&1
cat gpurun_out/sanitizer_e.log
cat > /tmp/dm.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
import dj_brdf_b200 as djb
d = torch.randint(0, 256, (8192, 8192), dtype=torch.uint8, device="cuda")
for _ in range(3): djb.dmap2nmap(d, 0.01)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): djb.dmap2nmap(d, 0.01)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
print(f"dmap2nmap 8192^2: {ms:.4f} ms, {4 * 8192 * 8192 / ms / 1e6:.1f} GB/s algorithmic (1 B in + 3 B out per texel)")
PY
python /tmp/dm.py | tee gpurun_out/dmap_e.log
python -m pytest tests/test_gpu_widening.py -m gpu -q -k dmap 2>&1 | tail -2
