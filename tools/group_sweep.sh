export HYDROTM_STAGED_MIN=0
for g in 1 2 4 6 8; do for k in 0 2 3; do
  echo "== groups $g spec $k: $(HYDROTM_GROUPS=$g HYDROTM_SPEC=$k python tools/prof_w2.py 4 2>&1 | tail -2 | tr '\n' ' ')"
done; done
