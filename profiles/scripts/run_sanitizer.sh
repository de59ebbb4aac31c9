# memcheck + racecheck of the kernels on small inputs (sanitizers)
cat > /tmp/san.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
import dj_brdf_b200 as djb
from oracle import api
from tests import cases
wi, wo, u = cases.pairs(3001)
ewi, ewo, eu = cases.edge_pairs()
wi = np.concatenate([wi, ewi]); wo = np.concatenate([wo, ewo]); u = np.concatenate([u, eu])
port = api.PortOracle()
mats = cases.c2_materials(port)[:5]
d = [torch.from_numpy(x).cuda() for x in (wi, wo, u)]
for cls in (djb.ggx, djb.beckmann):
    for fr in (djb.fresnel.ideal(), djb.fresnel.schlick([0.9, 0.5, 0.2]), djb.fresnel.unpolarized([1.5, 1.8, 2.4])):
        b = cls(fr)
        b.eval(d[0], d[1], mats); b.pdf(d[0], d[1], mats); b.sample(d[2], d[1], mats); b.evalp_is(d[2], d[1], mats[0])
        b.eval(wi, wo, mats); b.sample(u, wo, mats)
    pp = torch.from_numpy(np.ascontiguousarray(mats[np.arange(len(wi)) % 5])).cuda()
    b.eval(d[0], d[1], pp); b.evalp_is(d[2], d[1], pp)
m = djb.merl(cases.random_merl_table(3)); m.eval(d[0], d[1]); m.eval(wi[:7], wo[:7]); djb.merl.index(d[0], d[1])
spec = torch.from_numpy(wo.copy()).cuda(); spec[:, :2] *= -1
m.eval((spec + 1e-3 * torch.randn_like(spec)).contiguous(), d[1])  # everything through the exact-path queue
t = djb.utia(cases.random_utia_table(3)); t.eval(d[0], d[1])
nm = cases.synthetic_nmap(33, 17); djb.nmap2leanmap(torch.from_numpy(nm).cuda()); djb.nmap2leanmap(nm, 1e-5, 25.0)
djb.tabular.fit_batch([m, djb.ggx(), t], 24)
djb.tabular_anisotropic(t, 8, 10)
torch.cuda.synchronize()
print("sanitizer workload done")
PY
PYTHONPATH=$PWD compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/san.py 2>&1 | tail -4
echo "memcheck rc=$?"
PYTHONPATH=$PWD compute-sanitizer --tool racecheck --error-exitcode 7 python /tmp/san.py 2>&1 | tail -4
PYTHONPATH=$PWD compute-sanitizer --tool synccheck --error-exitcode 7 python /tmp/san.py 2>&1 | tail -3
