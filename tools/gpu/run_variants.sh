# usage: run_variants.sh <tag> ... : check + time the frame bench with each prebuilt library variants/lib_<tag>.so
cp ethzasl_brisk_b200/libbrisk_b200.so /tmp/lib_keep.so
for t in "$@"; do
  cp variants/lib_$t.so ethzasl_brisk_b200/libbrisk_b200.so
  python -m pytest tests/test_gpu_parity.py -x -q -k "golden or fused or describe" 2>&1 | tail -1
  python bench.py --frames 256 --steps 3 --no-knn --parity-frames 2 > gpurun_out/var_$t.json 2> gpurun_out/var_$t.err
  python -c "
import json; d=json.load(open('gpurun_out/var_$t.json')); print('$t', round(d['value'],1), round(d['e2e']['value'],1), d.get('parity_ok'), {k:round(v['ms_per_step'],2) for k,v in d['stages'].items()})"
done
cp /tmp/lib_keep.so ethzasl_brisk_b200/libbrisk_b200.so
