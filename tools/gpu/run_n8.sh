N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n$N.json')); print(d['value'], d['e2e'], d.get('parity_ok'), d['secondary']['value'], d['secondary'].get('parity_ok'), d['secondary']['sharding'])"
grep -c "NCCL INFO" gpurun_out/r02_bench_n$N.err; grep -i "nranks" gpurun_out/r02_bench_n$N.err | head -2; tail -2 gpurun_out/r02_bench_n$N.err | cut -c1-300
