run() { echo -n "$1: "; env $1 timeout 200 python tools/kbench.py --dtype bf16 --shapes m3 2>&1 | grep -v "^  " | sed -n "$2p" | cut -c1-130; }
run "RECNEXT_X=0" 2; run "RECNEXT_X=0" 3; run "RECNEXT_X=0" 4; run "RECNEXT_X=0" 5
run "RECNEXT_TW=2 RECNEXT_NT=3" 2
run "RECNEXT_TW=4 RECNEXT_NT=2" 2
run "RECNEXT_TW=8 RECNEXT_NT=1" 2
