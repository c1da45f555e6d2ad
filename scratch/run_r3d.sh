# round 2, call 3D: edge cases of the crossing march
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "crossing_march_edge_cases" 2>&1 | tail -30 > gpurun_out/r3d_tests.log
grep -n "PARITY\|passed\|failed\|Error" gpurun_out/r3d_tests.log | cut -c1-400
