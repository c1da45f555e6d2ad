# round 2, call T: InterpolatingRectilinear generator on the device -- the whole GPU suite
python profiles/source_sha.py > gpurun_out/r2t_sha.txt
python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/r2t_tests.log
tail -5 gpurun_out/r2t_tests.log
