# racecheck / memcheck / synccheck of the warp-compacting Beckmann kernel (shared-memory queues, __syncwarp protocol)
mkdir -p gpurun_out
cat > /tmp/san_f.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
import dj_brdf_b200 as djb
from oracle import api
from tests import cases
port = api.PortOracle()
for n, nm in ((1, 2), (33, 3), (3001, 16), (777, 40)):
    wi, wo, _ = cases.pairs(n, stream=50 + nm)
    if n > 1000:
        ewi, ewo, _ = cases.edge_pairs()
        wi, wo = np.concatenate([wi, ewi]), np.concatenate([wo, ewo])
    mats = cases.c2_materials(port, nm, seed=nm)
    twi, two = torch.from_numpy(wi).cuda(), torch.from_numpy(wo).cuda()
    for fr in (djb.fresnel.ideal(), djb.fresnel.schlick([0.9, 0.5, 0.2])):
        b = djb.beckmann(fr)
        b.eval(twi, two, mats); b.evalp(twi, two, mats); b.pdf(twi, two, mats)
        b.eval(wi, wo, mats)
torch.cuda.synchronize()
print("sanitizer workload done")
PY
for tool in racecheck memcheck synccheck; do
  echo "== $tool"
  compute-sanitizer --tool $tool python /tmp/san_f.py 2>&1 | grep -E "SUMMARY|sanitizer workload|Error|hazard|Race" | head -8
done > gpurun_out/sanitizer_f.log 2>&1
cat gpurun_out/sanitizer_f.log
