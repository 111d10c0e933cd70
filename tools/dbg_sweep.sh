for d in 0 1 2 4 8 16 31 30; do echo -n "fwd dbg=$d: "; RECNEXT_DBG=$d timeout 200 python tools/kbench.py --dtype bf16 --shapes m3 2>&1 | grep -v "^  " | sed -n 2p | cut -c1-60; done
for d in 0 32 64 128 256 512 992 1023; do echo -n "bwd dbg=$d: "; RECNEXT_DBG=$d timeout 200 python tools/kbench.py --dtype bf16 --shapes m3 2>&1 | grep -v "^  " | sed -n 2p | cut -c60-110; done
