# round 2 (p): the pdf kernels decide the shadowing gate of centred lobes without sigma(i): whole GPU suite, then the six step kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -6 | tee gpurun_out/r02_p_tests.log
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r02_p_quick.json 2> gpurun_out/r02_p_quick.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_p_quick.json').read().strip().split('\n')[-1])
print('ms/step',d['ms_per_step'], {k: round(v['ms'],2) for k,v in d['kernels'].items()})
"
