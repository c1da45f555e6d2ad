python -m pytest tests/test_gpu_parity.py -q -m gpu -k "split_frame_with_crossing or frame_split" 2>&1 | tail -15 > gpurun_out/r3t_tests.log; tail -5 gpurun_out/r3t_tests.log
