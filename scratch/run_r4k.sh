# round 2, call 4K: gen's sidecar through the parallel gzip writer
timeout 60 python -m pytest tests/test_host_gen.py -q -m gpu -k "end_to_end" 2>&1 | tail -5
