python -m pytest tests/test_gpu_parity.py -x -q -k "async or chunked or fused or device_resident" 2>&1 | tail -8 > gpurun_out/r2m_pytest.log
tail -8 gpurun_out/r2m_pytest.log
python bench.py --steps 5 --no-knn --parity-frames 4 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err || tail -5 gpurun_out/r2m_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2m_bench.json')); print(round(d['value'],1), d['e2e'], d.get('parity_ok'))"
