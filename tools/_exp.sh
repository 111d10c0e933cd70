timeout 300 python -m pytest tests/test_stem.py -q -m gpu 2>&1 | tail -3
RECNEXT_STEM_DBG=1 timeout 120 python tools/stem_prof.py 256 64 224 | tail -2
timeout 120 python tools/stem_prof.py 256 80 224
