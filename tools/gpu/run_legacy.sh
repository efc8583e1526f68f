python -m pytest tests/test_gpu_parity.py -x -q -k "legacy or score_calculator or opencv_mode or samplers_16" 2>&1 | tail -25 > gpurun_out/r2l_pytest.log
tail -25 gpurun_out/r2l_pytest.log
