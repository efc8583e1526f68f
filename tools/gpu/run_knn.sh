timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "knn" 2>&1 | tail -5 > gpurun_out/r2e_pytest.log
timeout 300 python tools/knn_timing.py 100000 1000000 2 > gpurun_out/r2e_knn256.json 2> gpurun_out/r2e_knn.err
BRISK_B200_TC5_TILE_ROWS=128 timeout 300 python tools/knn_timing.py 100000 1000000 2 > gpurun_out/r2e_knn128.json 2>> gpurun_out/r2e_knn.err
tail -4 gpurun_out/r2e_pytest.log; cat gpurun_out/r2e_knn256.json gpurun_out/r2e_knn128.json; tail -3 gpurun_out/r2e_knn.err
