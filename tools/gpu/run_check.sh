python -m pytest tests/test_gpu_parity.py -x -q -k "resident or chunked or full_size or async or fused" 2>&1 | tail -2
python bench.py --steps 5 --no-knn --parity-frames 8 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err || tail -5 gpurun_out/r2q_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2q_bench.json')); print(round(d['value'],1), round(d['e2e']['value'],1), d.get('parity_ok'), {k:round(v['ms_per_step'],2) for k,v in d['stages'].items()})"
