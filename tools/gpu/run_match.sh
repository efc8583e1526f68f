python -m pytest tests/test_gpu_parity.py -x -q -k "knn or match or hamming or radius or matcher" 2>&1 | tail -60 > gpurun_out/r2p_pytest.log
grep -n "^E " gpurun_out/r2p_pytest.log | head -20; tail -3 gpurun_out/r2p_pytest.log
