# groups x host threads of the staged pipeline on a core-starved host (taskset emulates 4 cores per rank of an 8-rank box)
# usage: bash tools/thread_sweep.sh   (on a GPU box)
for cfg in "3 3" "6 3" "6 2" "6 1" "8 3" "6 6" "4 4"; do
  set -- $cfg
  echo "== 4 cores: groups $1 threads $2: $(HYDROTM_GROUPS=$1 HYDROTM_THREADS=$2 taskset -c 0-3 python tools/prof_w2.py 6 2>&1 | tail -3 | awk '{print $3}' | tr '\n' ' ')"
done
for cfg in "6 6" "6 3" "8 8" "8 4"; do
  set -- $cfg
  echo "== all cores: groups $1 threads $2: $(HYDROTM_GROUPS=$1 HYDROTM_THREADS=$2 python tools/prof_w2.py 6 2>&1 | tail -3 | awk '{print $3}' | tr '\n' ' ')"
done
