# round 2, call 4H: the host tests on the GPU (the Python mirror of output_image against the executable's picture)
timeout 200 python -m pytest tests/test_host_gen.py tests/test_geotiff.py -q -m gpu 2>&1 | tail -4
