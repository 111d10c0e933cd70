run() { echo -n "$1: "; env $1 timeout 200 python tools/kbench.py --dtype bf16 --shapes m3 2>&1 | grep -v "^  " | sed -n "$2p" | cut -c1-130; }
for cfg in "RECNEXT_MAXW=8" "RECNEXT_MAXW=12" "RECNEXT_MAXW=16" "RECNEXT_MAXW=12 RECNEXT_TW=4 RECNEXT_NT=3" "RECNEXT_MAXW=16 RECNEXT_TW=4 RECNEXT_NT=3" "RECNEXT_MAXW=8 RECNEXT_TW=1 RECNEXT_NT=3"; do run "$cfg" 2; done
for cfg in "RECNEXT_MAXW=8" "RECNEXT_MAXW=12" "RECNEXT_MAXW=16"; do run "$cfg" 3; run "$cfg" 4; run "$cfg" 5; done
