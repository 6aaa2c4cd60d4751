#!/bin/bash
# One GPU-box round: parity tests, smoke, bench, ncu launch list.  Usage (under gpurun): bash tools/gpu_round.sh [tag] [steps...]
TAG=${1:-r1}; shift
STEPS=${@:-tc tests smoke bench ncu}
mkdir -p gpurun_out
for S in $STEPS; do case $S in
 tc)    timeout 600 python -m pytest tests/test_conv3d_tc.py -q -m gpu --timeout 120 -x > gpurun_out/${TAG}_tc.log 2>&1; echo "tc rc=$?" >> gpurun_out/${TAG}_tc.log; tail -25 gpurun_out/${TAG}_tc.log;;
 tests) timeout 1200 python -m pytest tests -q -m gpu --timeout 300 --deselect tests/test_conv3d_tc.py > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -25 gpurun_out/${TAG}_pytest.log;;
 smoke) timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log; tail -4 gpurun_out/${TAG}_smoke.log;;
 bench) timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?" >> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err;;
 ncufull) for K in conv3d_tc_kernel warp_var_fwd_fast_kernel softargmin_fwd_kernel; do timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$K -c 2 -f -o gpurun_out/${TAG}_$K python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --ncu-range > gpurun_out/${TAG}_ncufull_$K.log 2>&1; tail -1 gpurun_out/${TAG}_ncufull_$K.log; done;;
 sweepprof) timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3d_tc_kernel -s 3 -c 20 -f -o gpurun_out/${TAG}_sweepprof python tools/conv_sweep.py prof > gpurun_out/${TAG}_sweepprof.log 2>&1; tail -3 gpurun_out/${TAG}_sweepprof.log;;
 batches) for BB in 1 2 4 8; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch $BB > gpurun_out/${TAG}_bench_b$BB.json 2>&1; python -c "import json;d=json.loads(open('gpurun_out/${TAG}_bench_b$BB.json').read().strip().splitlines()[-1]);print('batch $BB', round(d['value']/1e9,3), 'G/s', round(d['ms_per_step'],3), 'ms', {k: round(v,3) for k,v in d['stage_ms'].items()}, 'e2e', round(d['e2e']['value']/1e9,3))"; done;;
 conv0prof) timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv3d_tc_kernel -c 1 -f -o gpurun_out/${TAG}_conv0 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --batch 1 --ncu-range > gpurun_out/${TAG}_conv0prof.log 2>&1; tail -2 gpurun_out/${TAG}_conv0prof.log;;
 sweep) timeout 600 python tools/conv_sweep.py > gpurun_out/${TAG}_sweep.txt 2>&1; cat gpurun_out/${TAG}_sweep.txt;;
 cpt4)  MVS_WARP_CPT=4 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cpt4.json 2>&1; python -c "import json;d=json.loads(open('gpurun_out/${TAG}_bench_cpt4.json').read().strip().splitlines()[-1]);print('cpt4', d['ms_per_step'], d['stage_ms'])";;
 ncu)   timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --batch ${NCU_BATCH:-1} --ncu-range > gpurun_out/${TAG}_ncu_bench.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_bench.log;;
esac; done
