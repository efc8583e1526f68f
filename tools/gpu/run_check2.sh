python bench.py > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2n_bench.json')); print(d['value'], d['e2e'], d.get('parity_ok'), {k:round(v['ms_per_step'],2) for k,v in d['stages'].items()}, d['secondary']['value'])"
python tests/angle_stats.py --frames 160 > gpurun_out/r2n_angle_stats.json 2> gpurun_out/r2n_angle.err; cat gpurun_out/r2n_angle_stats.json; tail -2 gpurun_out/r2n_angle.err
ncu --set full --clock-control none --import-source on -k regex:"pyramid_kernel|agast_detect|row_scan|corner_fill|nms_|refine_kernel|compact_kernel|integral_|describe_" -s 54 -c 27 -o gpurun_out/r02_step python tools/profile_step.py 256 2 > gpurun_out/r02_ncu_step.log 2>&1; tail -1 gpurun_out/r02_ncu_step.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "alternative_forms or chunked" 2>&1 | tail -3
