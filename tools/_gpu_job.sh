python -m pytest tests -m gpu -x -q -k "ref or golden or stream or config4 or block_session" > gpurun_out/r2_gputests_12.log 2>&1; tail -4 gpurun_out/r2_gputests_12.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustain-s 0 > gpurun_out/r2_bench_12.json 2> gpurun_out/r2_bench_12.err;  python -c "
import json; d=json.load(open('gpurun_out/r2_bench_12.json')); print('value', d['value']/1e6, 'reads_only', d['reads_only']['value']/1e6, 'ref', d['ref_path']['ms_per_step'], d['ref_path']['alone_ms_per_block'], d['ref_path']['roofline']['frac'], 'e2e', d['e2e']['value']/1e6, d['e2e']['reads_only_value']/1e6, d['e2e']['blocking_call_value']/1e6, 'frac', d['roofline']['frac'])"; tail -3 gpurun_out/r2_bench_12.err
STRK_REF_TIMING=1 python tools/bench_ref_path.py 2>&1 | tail -3
STRK_REF_WD=8 STRK_REF_TIMING=1 python tools/bench_ref_path.py 2>&1 | tail -3
