python -m pytest tests/test_gpu_parity.py -x -q -k "detect or golden or fused or full_size or config4 or compute_scale or nms or thresh or mask" 2>&1 | tail -5 > gpurun_out/r2n_pytest.log
tail -3 gpurun_out/r2n_pytest.log
python bench.py --frames 512 --steps 3 --no-knn --parity-frames 8 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err || tail -5 gpurun_out/r2n_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2n_bench.json')); print(round(d['value'],1), round(d['e2e']['value'],1), d.get('parity_ok'), {k:round(v['ms_per_step'],2) for k,v in d['stages'].items()}, d['stages']['nms'].get('kernels'))"
