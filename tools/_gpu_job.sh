for v in 0 1; do if [ $v = 1 ]; then export STRK_NO_PRIORITY=1; fi
python bench.py --config 3 --steps 3 --warmup 2 > gpurun_out/r2_bench_cfg3_p$v.json 2> gpurun_out/r2_bench_cfg3_p$v.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_cfg3_p$v.json')); print('cfg3 noprio=$v', d['value']/1e6, d['e2e']['value']/1e6, d['widening_passes'])"
python bench.py --steps 12 --warmup 3 --no-cpu-baseline --sustain-s 0 > gpurun_out/r2_bench_p$v.json 2> gpurun_out/r2_bench_p$v.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_p$v.json')); print('cfg2 noprio=$v value', d['value']/1e6, 'e2e', d['e2e']['value']/1e6, d['e2e']['reads_only_value']/1e6, d['e2e']['blocking_call_value']/1e6)"
done
