# round 2: racecheck / memcheck / synccheck of what the round added -- the 1e-5 tier of sample, the compacting kernel's hand-over of
# declined items to the slow queue (both tiers, off-centre materials), the split-mode isotropic fit (producer / adder slabs in shared
# memory), the iterate history, the member-query kernel, the LEAN half / mip kernels, the in-library anisotropic fit run
mkdir -p gpurun_out
cat > /tmp/san_r2.py <<'PY'
import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, ".")
import dj_brdf_b200 as djb
from dj_brdf_b200 import capi, fit_sharded as fs
from oracle import api
from tests import cases
port = api.PortOracle()
lib = capi.load()
for tier in ("1e-5", "bits"):
    djb.set_precision(tier)
    for n, nm in ((1, 2), (33, 3), (3001, 16), (777, 40)):
        wi, wo, u = cases.pairs(n, stream=50 + nm)
        if n > 1000:
            ewi, ewo, eu = cases.edge_pairs()
            wi, wo, u = np.concatenate([wi, ewi]), np.concatenate([wo, ewo]), np.concatenate([u, eu])
        mats = cases.c2_materials(port, nm, seed=nm)
        mats[0] = djb.params.pdfparams(0.3, 0.2, 0.4, 0.6, -0.7)   # off-centre: ill-conditioned G -> declined items
        mats[-1] = djb.params.pdfparams(0.5, 0.4, -0.3, -0.8, 0.5)
        twi, two, tu = torch.from_numpy(wi).cuda(), torch.from_numpy(wo).cuda(), torch.from_numpy(u).cuda()
        for fr in (djb.fresnel.ideal(), djb.fresnel.schlick([0.9, 0.5, 0.2])):
            for b in (djb.beckmann(fr), djb.ggx(fr)):
                b.eval(twi, two, mats); b.evalp(twi, two, mats); b.pdf(twi, two, mats); b.sample(tu, two, mats)
                b.eval(wi, wo, mats[0]); b.sample(u, wo, mats[1])
djb.set_precision("bits")
# isotropic fit: single launch, every split part count, 4 and 50 iterations
srcs = [djb.merl(cases.smooth_merl_table(21)), djb.ggx()]
for parts in (1, 3, 8, 0):
    capi.check(lib.djb200_debug_fit_parts(C.c_int(parts)))
    for it in (4, 50):
        djb.tabular.fit_packed(srcs, 90, True, it)
capi.check(lib.djb200_debug_fit_parts(C.c_int(0)))
djb.tabular.fit_packed(srcs[:1], 24, True, 4)
# anisotropic fit through the in-library run, members, LEAN half mips
ut = djb.utia(cases.random_utia_table(12))
t = fs.tabular_anisotropic_sharded(ut, 14, 18, True, 4)
phi = np.linspace(-1, 7, 257, dtype=np.float32); th = np.linspace(0, 1.7, 257, dtype=np.float32); uu = np.linspace(0, 1, 257, dtype=np.float32)
t.pdf1(phi); t.cdf1(phi); t.qf1(uu); t.pdf2(th, phi); t.cdf2(th, phi); t.qf2(uu, phi)
g = djb.ggx(); bk = djb.beckmann()
c = np.linspace(1e-3, 1, 257, dtype=np.float32); s = np.sqrt(1 - c * c).astype(np.float32)
for b in (g, bk):
    b.qf1(uu * 0.99998 + 1e-5); q = b.qf2_radial(uu * 0.99998 + 1e-5, c, s); b.qf3_radial(uu * 0.99998 + 1e-5, q)
h = api.directions(300, 3)
for m in (djb.sgd("gold-metallic-paint"), djb.abc("gold-metallic-paint")):
    m.ndf(h); m.gaf(h, h, h); m.fresnel_term(c)
l1, l2 = djb.nmap2leanmap(cases.synthetic_nmap(37, 53, seed=5), 1e-5, 25.0)
djb.leanmap_half_mips(l1); djb.leanmap_half_mips(torch.from_numpy(l2).cuda())
torch.cuda.synchronize()
print("sanitizer workload done")
PY
for tool in racecheck memcheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool python /tmp/san_r2.py 2>&1 | grep -E "SUMMARY|sanitizer workload|Error|hazard|Race|Traceback|rror:" | head -12
done > gpurun_out/sanitizer_r2.log 2>&1
cat gpurun_out/sanitizer_r2.log
