# does CPU placement / the number of hardware connections change the pipeline's wall time?  (on a GPU box)
nvidia-smi topo -m 2>&1 | head -8
lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" | head -8
run() { echo "$1: $(eval "$2 python tools/prof_w2.py 6" 2>&1 | tail -3 | cut -c14-22 | tr '\n' ' ')"; }
run "default" ""
run "taskset 0-3 (4 thr)" "taskset -c 0-3"
run "taskset 0-7" "taskset -c 0-7"
run "taskset 8-15" "taskset -c 8-15"
run "taskset 0-7 threads 3" "HYDROTM_THREADS=3 taskset -c 0-7"
run "max connections 32" "CUDA_DEVICE_MAX_CONNECTIONS=32"
run "max connections 32 + taskset 0-7" "CUDA_DEVICE_MAX_CONNECTIONS=32 taskset -c 0-7"
run "threads 3 groups 6 all cores" "HYDROTM_THREADS=3"
run "default again" ""
