python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_4.log 2>&1; tail -8 gpurun_out/r2_gputests_4.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_3.json 2> gpurun_out/r2_bench_3.err; tail -3 gpurun_out/r2_bench_3.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_3.json'))
print('value', d['value']/1e6, 'reads_only', d['reads_only']['value']/1e6, 'ref ms', d['ref_path']['ms_per_step'], 'ref dp ms', d['ref_path']['dp_kernel_ms_per_step'], 'ref frac', d['ref_path']['roofline']['frac'])
print('e2e', d['e2e']['value']/1e6, 'e2e reads only', d['e2e']['reads_only_value']/1e6, 'blocking', d['e2e']['blocking_call_value']/1e6, 'frac', d['roofline']['frac'], 'parity', d['parity_sample'], d['ref_path']['parity_sample'])
PY
python bench.py --config 4 --steps 5 --warmup 2 > gpurun_out/r2_bench_cfg4.json 2> gpurun_out/r2_bench_cfg4.err; tail -c 1800 gpurun_out/r2_bench_cfg4.json; tail -3 gpurun_out/r2_bench_cfg4.err
python bench.py --config 3 --steps 3 --warmup 2 > gpurun_out/r2_bench_cfg3.json 2> gpurun_out/r2_bench_cfg3.err; tail -c 1800 gpurun_out/r2_bench_cfg3.json; tail -3 gpurun_out/r2_bench_cfg3.err
STRK_REF_TIMING=1 python tools/bench_ref_path.py 2>&1 | tail -4
