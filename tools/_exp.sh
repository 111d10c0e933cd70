timeout 300 python -m pytest tests/test_ffn.py -x -q -m gpu 2>&1 | tail -2
for s in "256 64 56 128" "256 128 28 256" "256 256 14 512" "256 512 7 1024" "256 320 14 640"; do timeout 60 python tools/ffn_prof.py $s 7; done
RECNEXT_FFN_PROF=1 timeout 60 python tools/ffn_prof.py 256 256 14 512 3 2>&1 | grep -A14 "MMA warp totals"
