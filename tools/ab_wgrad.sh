mkdir -p gpurun_out
(timeout 200 python -m pytest tests/test_ops_gpu.py -x -q -k "bwd_filter or conv_small") > gpurun_out/s6_ops.log 2>&1; tail -3 gpurun_out/s6_ops.log
for cfg in "1 0" "0 0" "1 32" "2 0" "3 0"; do set -- $cfg; g=$1; px=$2
DPIG_WGRAD_GROUP=$g DPIG_WGRAD_PX=$px timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --detail gpurun_out/s6_detail_g${g}_p${px}.txt > gpurun_out/s6_bench_g${g}_p${px}.log 2>&1
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/s6_bench_g${g}_p${px}.log") if l.startswith("{")][-1])
print("group=$g px=$px", round(d["value"],1), round(d["roofline"]["wgrad_kernel"]["achieved"],1), d["roofline"]["kernel_time_ms_per_iteration"]["conv2d_bwd_filter"])
PY
done
