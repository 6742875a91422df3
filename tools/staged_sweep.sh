for smin in ${SWEEP:-1000 0 7 9 13}; do
  echo "== HYDROTM_STAGED_MIN=$smin"
  HYDROTM_STAGED_MIN=$smin python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'], d['gpu_launches'], d['status_counts'])"
done
