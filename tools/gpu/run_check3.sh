python bench.py > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2n_bench.json')); print(d['value'], d['e2e'], d.get('parity_ok'), {k:round(v['ms_per_step'],2) for k,v in d['stages'].items()}, d['secondary']['value'], d['secondary'].get('parity_ok'))"
python tests/angle_stats.py --frames 160 > gpurun_out/r2n_angle_stats.json 2> gpurun_out/r2n_angle.err; cat gpurun_out/r2n_angle_stats.json; tail -2 gpurun_out/r2n_angle.err
