# round 2 (g): 1e-5 tier of sample -- statistics against the oracle / the exact tier, then the step and the lean-shading leg
mkdir -p gpurun_out
PYTHONPATH=$PWD python profiles/scripts/fast_sample_stats.py > gpurun_out/r02_g_fast_sample_stats.json 2> gpurun_out/r02_g_stats.err; tail -3 gpurun_out/r02_g_stats.err
python -c "
import json
d=json.load(open('gpurun_out/r02_g_fast_sample_stats.json'))
for k,v in d.items(): print(k, {a:(round(b,9) if isinstance(b,float) else b) for a,b in v.items() if a!='frac_le_1e5_per_material'})
"
