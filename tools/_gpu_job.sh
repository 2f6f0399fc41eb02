B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --sustain-s 0"
for v in 0 1 3; do STRKIT_B200_LIB=$PWD/build/libstrk_pol$v.so $B > gpurun_out/pol$v.json 2> gpurun_out/pol$v.err; python -c "
import json,sys
d=json.load(open('gpurun_out/pol$v.json')); print('pol$v', d['value']/1e6, d['reads_only']['value']/1e6, d['ref_path']['ms_per_step'], d['parity_sample_bit_exact'])
"; done
N="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --sustain-s 0 --no-ref-path --pool 1"
for v in 0 1 3; do STRKIT_B200_LIB=$PWD/build/libstrk_pol$v.so ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:dp_packed_kernel -c 15 --csv --log-file gpurun_out/r2_traffic_pol$v.csv $N > /dev/null 2>&1; done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dp_|ref_|replay|plan_|hash|dedupe|expand" -c 200 --csv --log-file gpurun_out/r2_ref_launches.csv python tools/bench_ref_path.py 32768 > gpurun_out/ncu_ref.log 2>&1; tail -2 gpurun_out/ncu_ref.log
