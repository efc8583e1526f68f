python -m pytest tests/test_gpu_parity.py -x -q -k "knn or match or hamming or radius or matcher" 2>&1 | tail -4
python bench.py --steps 3 --parity-frames 2 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err || tail -5 gpurun_out/r2s_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2s_bench.json')); s=d['secondary']; print(round(d['value'],1), s['value'], s.get('parity_ok'), {k:round(v['Gcmp/s'],1) for k,v in s['variants'].items()}, s['roofline']['frac'], s['e2e']['value'])"
python bench.py --config C5 --steps 2 > gpurun_out/r02_bench_C5.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_C5.json')); print('C5', d['value'], d['e2e']['value'], d.get('parity_ok'), d.get('variants'))"
