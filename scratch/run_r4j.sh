# round 2, call 4J: gen on a folder with a GeoTIFF tile
timeout 100 python -m pytest tests/test_geotiff.py -q -m gpu 2>&1 | tail -12
