timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "knn or matcher or radius" 2>&1 | tail -15 > gpurun_out/r2e_pytest.log
timeout 300 python tools/knn_variants.py 100000 1000000 > gpurun_out/r2e_knn.json 2> gpurun_out/r2e_knn.err
tail -12 gpurun_out/r2e_pytest.log; cat gpurun_out/r2e_knn.json; tail -3 gpurun_out/r2e_knn.err
