python -m pytest tests/test_gpu_parity.py -x -q -k "describe or golden or fused or full_size or config4 or harris_batch or cpp_dropin or chunked or device_resident or integral" 2>&1 | tail -5 > gpurun_out/r2k_pytest.log
python bench.py --frames 256 --steps 3 --no-knn --parity-frames 4 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
tail -3 gpurun_out/r2k_pytest.log; python -c "
import json; d=json.load(open('gpurun_out/r2k_bench.json')); print(d['value'], d['e2e']['value'], d.get('parity_ok'), {k:round(v['ms_per_step'],2) for k,v in d['stages'].items()})"
