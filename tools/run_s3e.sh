for cfg in "1024 64" "1536 32" "768 32" "1024 48" "768 64"; do
  set -- $cfg
  for sc in Benchmark_Empty32x32_30robots Congested10x10_7robots Corridor10x10_4robots; do
    echo "== targets $1 controls $2 $sc"
    timeout 300 python tools/profile_solve.py $sc 1 8 0 $1 $2 2>&1 | grep -v "^\[kcbs\]" | sed -e 's/low-level plans/plans/' -e 's/validations [0-9]*, //'
  done
done
