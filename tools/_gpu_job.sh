python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_5.log 2>&1; tail -5 gpurun_out/r2_gputests_5.log
python bench.py --config 4 --steps 5 --warmup 2 > gpurun_out/r2_bench_cfg4.json 2> gpurun_out/r2_bench_cfg4.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_cfg4.json')); print('cfg4', d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity_sample_bit_exact'])"; tail -3 gpurun_out/r2_bench_cfg4.err
python bench.py --gpus 1 --steps 32 --warmup 3 --scaling strong > gpurun_out/r2_bench_strong_n1.json 2> gpurun_out/r2_bench_strong_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_strong_n1.json')); print('strong n1', d['value']/1e6, d['partition'], d['parity_sample_bit_exact'])"; tail -3 gpurun_out/r2_bench_strong_n1.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustain-s 0 > gpurun_out/r2_bench_4.json 2> gpurun_out/r2_bench_4.err;  python -c "
import json; d=json.load(open('gpurun_out/r2_bench_4.json')); print('value', d['value']/1e6, 'reads_only', d['reads_only']['value']/1e6, 'ref ms', d['ref_path']['ms_per_step'], 'e2e', d['e2e']['value']/1e6, d['e2e']['reads_only_value']/1e6)"; tail -3 gpurun_out/r2_bench_4.err
