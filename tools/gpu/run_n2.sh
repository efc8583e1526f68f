python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/host_bw_probe.py > gpurun_out/r2i_hostbw_n2.json 2> gpurun_out/r2i_hostbw_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r2i_bench_n2.json')); print(d['value'], d['e2e']['value'], d.get('parity_ok'), d['secondary']['value'], d['secondary'].get('parity_ok'), d['secondary']['sharding'])"
cat gpurun_out/r2i_hostbw_n2.json; grep -c "nranks" gpurun_out/r2i_bench_n2.err; tail -3 gpurun_out/r2i_bench_n2.err
