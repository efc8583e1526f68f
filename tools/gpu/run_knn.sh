timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "knn" 2>&1 | tail -8 > gpurun_out/r2m_pytest.log
timeout 300 python tools/knn_timing.py 100000 1000000 2 > gpurun_out/r2m_knn_ts.json 2> gpurun_out/r2m_knn.err
BRISK_B200_TC5_MODE=ss timeout 300 python tools/knn_timing.py 100000 1000000 2 > gpurun_out/r2m_knn_ss.json 2>> gpurun_out/r2m_knn.err
tail -8 gpurun_out/r2m_pytest.log; cat gpurun_out/r2m_knn_ts.json gpurun_out/r2m_knn_ss.json; tail -3 gpurun_out/r2m_knn.err
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "describe or golden or fused or full_size or config4 or harris_batch or cpp_dropin or chunked or device_resident or capacity" 2>&1 | tail -5 > gpurun_out/r2m_pytest_desc.log
python bench.py --frames 256 --steps 3 --no-knn --parity-frames 4 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
tail -3 gpurun_out/r2m_pytest_desc.log; python -c "
import json; d=json.load(open('gpurun_out/r2m_bench.json')); print(d['value'], d['e2e']['value'], d.get('parity_ok'), {k:round(v['ms_per_step'],2) for k,v in d['stages'].items()})"
