# Launch-shape knob sweep on the whole step (FROST_TUNE presets, include/frost_b200.h): bash tools/sweep_knobs.sh
for t in "" "6=8" "6=12" "6=24" "6=32" "1=4" "1=6" "0=3" "2=4" "10=4" "9=2" ""; do
  echo "== FROST_TUNE=$t"
  FROST_TUNE="$t" timeout 200 python bench.py --quick --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3))"
done
