python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1
python bench.py > gpurun_out/r02_bench_C3.json 2> gpurun_out/r02_bench_C3.err
python bench.py --config C2 --steps 5 > gpurun_out/r02_bench_C2.json 2>/dev/null
python bench.py --config C4 --steps 3 > gpurun_out/r02_bench_C4.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 3 --frames 256 --no-knn --parity-frames 0 > gpurun_out/r02_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pyramid_kernel|agast_detect|row_scan|corner_fill|nms_|refine_kernel|compact_kernel|integral_|describe_" -s 54 -c 27 -o gpurun_out/r02_step python tools/profile_step.py 128 2 > gpurun_out/r02_ncu_step.log 2>&1
tail -2 gpurun_out/r02_pytest.log; tail -1 gpurun_out/r02_smoke.log; for f in C3 C2 C4; do python -c "
import json,sys; d=json.load(open('gpurun_out/r02_bench_$f.json')); print('$f', round(d['value'],1), d['unit'], 'e2e', round(d['e2e']['value'],1), 'parity', d.get('parity_ok'), {k:round(v['ms_per_step'],2) for k,v in d['stages'].items()})"; done
