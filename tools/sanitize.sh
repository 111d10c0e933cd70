for t in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $t"; timeout 1500 compute-sanitizer --tool $t --print-limit 20 python tools/sanitize_all.py 2>&1 | grep -v "^recconv\|^ffn\|^dwdown\|^linattn\|^stem" | tail -12
done
