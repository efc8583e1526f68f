python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1
python bench.py > gpurun_out/r02_bench_C3.json 2> gpurun_out/r02_bench_C3.err
python bench.py --impl reference --steps 3 > gpurun_out/r02_bench_C3_reference.json 2>> gpurun_out/r02_bench_C3.err
python bench.py --config C2 --steps 5 > gpurun_out/r02_bench_C2.json 2>/dev/null
python bench.py --config C4 --steps 3 > gpurun_out/r02_bench_C4.json 2>/dev/null
bash tools/gpu/run_sanitize.sh > gpurun_out/r02_sanitize_summary.txt 2>&1
tail -2 gpurun_out/r02_pytest.log; tail -1 gpurun_out/r02_smoke.log; for f in C3 C2 C4; do python -c "
import json,sys; d=json.load(open('gpurun_out/r02_bench_$f.json')); print('$f', round(d['value'],1), d['unit'], 'e2e', round(d['e2e']['value'],1), 'parity', d.get('parity_ok'), {k:round(v['ms_per_step'],2) for k,v in d['stages'].items()}, (d.get('secondary') or {}).get('value'))"; done
grep -E "rc=|SUMMARY" gpurun_out/r02_sanitize_summary.txt
