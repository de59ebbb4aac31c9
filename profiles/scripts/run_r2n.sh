# round 2 (n): table-driven double exp / log / atan + polynomial acos in the widened kernels (sgd, abc, utia, tabular): whole GPU
# suite, then the widened legs of the bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -15 | tee gpurun_out/r02_n_tests.log
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_n_quick.json 2> gpurun_out/r02_n_quick.err
tail -2 gpurun_out/r02_n_quick.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_n_quick.json').read().strip().split('\n')[-1])
print('ms/step',d['ms_per_step'])
for k in ('sgd','abc','utia','tabular_eval','tabular_anisotropic_sample','lean_shading','merl'):
    print(k, round(d[k]['ms'],3), round(d[k].get('evals_per_s', d[k].get('samples_per_s', d[k].get('lookups_per_s',0)))/1e9,2), 'G/s', round(d[k]['roofline']['frac'],3))
print('fit', d['fit']['fits_per_s'], d['fit']['device_full']['fits_per_s'], 'aniso', d['aniso_fit']['ms'], d['aniso_fit']['device_ms'])
"
