python -m pytest tests/test_gpu_parity.py -x -q -k "knn or match or hamming or radius or matcher" 2>&1 | tail -3
python bench.py --config C5 --steps 2 > gpurun_out/r02_bench_C5.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_C5.json')); print('C5', d['value'], d['e2e']['value'], d.get('parity_ok'), {k: round(v['Gcmp/s'],1) for k,v in d['variants'].items()}, d['roofline']['frac'])"
