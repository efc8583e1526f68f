ncu --set full --clock-control none --import-source on -k regex:"tc5mx_kernel" -c 1 -o gpurun_out/r02_tc5mx python tools/knn_timing.py 50000 400000 3 > gpurun_out/r02_ncu_tc5mx.log 2>&1
ls -la gpurun_out/r02_tc5mx.ncu-rep; tail -3 gpurun_out/r02_ncu_tc5mx.log
