python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2h_pytest.log
BRISK_B200_NMS_TIMING=1 python bench.py --frames 256 --steps 3 --no-knn --parity-frames 4 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
tail -6 gpurun_out/r2h_pytest.log; grep "nms 256" gpurun_out/r2h_bench.err | tail -1; python -c "
import json; d=json.load(open('gpurun_out/r2h_bench.json')); print(d['value'], d['e2e']['value'], d.get('parity_ok'), {k:round(v['ms_per_step'],2) for k,v in d['stages'].items()})"
