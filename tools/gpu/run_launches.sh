ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --steps 1 --warmup 3 --frames 256 --no-knn --parity-frames 0 > gpurun_out/r2n_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/r2n_launches.csv 2>/dev/null | head -40
