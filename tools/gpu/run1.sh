python -m pytest tests/test_gpu_parity.py -x -q -k "detect or golden or full_size or config4 or fused or chunked or compute_scale or capacity or dense or raw_corners" 2>&1 | tail -5 > gpurun_out/r2b_pytest.log
BRISK_B200_NMS_TIMING=1 python bench.py --frames 256 --steps 3 --no-knn --parity-frames 4 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -3 gpurun_out/r2b_pytest.log; grep "nms 256" gpurun_out/r2b_bench.err | tail -2; python -c "
import json; d=json.load(open('gpurun_out/r2b_bench.json')); print(d['value'], d['e2e']['value'], d.get('parity_ok'), {k:round(v['ms_per_step'],2) for k,v in d['stages'].items()})"
