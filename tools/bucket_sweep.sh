for cfg in "8:64,5:128,3:128,2:128,1:256" "8:128,5:128,3:128,2:128,1:256" "8:64,5:64,3:64,2:128,1:128" "8:32,5:64,3:128,2:256,1:256" "8:64,6:64,4:128,3:128,2:128,1:256" "8:256:w,5:128,3:128,2:128,1:256"; do
  echo "== $cfg"; HYDROTM_BUCKETS="$cfg" python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'], d['gpu_launches'])"
done
