#!/bin/bash
# One GPU-box round.  Usage (under gpurun): bash tools/gpu_round.sh <tag> <step> [step...]
# Every step writes its log under gpurun_out/<tag>_<step>.* (merged back by gpurun) and prints a short tail.
TAG=${1:-r2}; shift
STEPS=${@:-tests smoke bench}
mkdir -p gpurun_out
O=gpurun_out/${TAG}
for S in $STEPS; do echo "=== $S"; case $S in
 tests)   timeout 2400 python -m pytest tests -q -m gpu --timeout 600 > ${O}_pytest.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest.log; tail -30 ${O}_pytest.log;;
 testsx)  timeout 2400 python -m pytest tests -q -x -m gpu --timeout 600 > ${O}_pytest.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest.log; tail -40 ${O}_pytest.log;;
 full)    timeout 1500 python -m pytest tests/test_gpu_fullsize.py -q -m gpu --timeout 600 > ${O}_full.log 2>&1; echo "rc=$?" >> ${O}_full.log; tail -40 ${O}_full.log;;
 smoke)   timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; echo "smoke rc=$?" >> ${O}_smoke.log; tail -6 ${O}_smoke.log;;
 bench)   timeout 900 python bench.py > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$?" >> ${O}_bench.err; cat ${O}_bench.json; tail -3 ${O}_bench.err;;
 benchq)  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity > ${O}_benchq.json 2> ${O}_benchq.err; echo "rc=$?" >> ${O}_benchq.err; cat ${O}_benchq.json; tail -3 ${O}_benchq.err;;
 parity)  timeout 900 python tools/parity_report.py --peaks 0.3,0.5 --out ${O}_parity.json > ${O}_parity.log 2>&1; echo "rc=$?" >> ${O}_parity.log; tail -5 ${O}_parity.log;;
 bars)    timeout 900 python tools/cudnn_bars.py > ${O}_cudnn_bars.json 2> ${O}_cudnn_bars.err; echo "rc=$?" >> ${O}_cudnn_bars.err; cat ${O}_cudnn_bars.json; tail -3 ${O}_cudnn_bars.err;;
 memcheck)  timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_cases.py > ${O}_memcheck.log 2>&1; echo "memcheck rc=$?" >> ${O}_memcheck.log; tail -12 ${O}_memcheck.log;;
 racecheck) timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_cases.py > ${O}_racecheck.log 2>&1; echo "racecheck rc=$?" >> ${O}_racecheck.log; tail -12 ${O}_racecheck.log;;
 synccheck) timeout 1500 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitize_cases.py > ${O}_synccheck.log 2>&1; echo "synccheck rc=$?" >> ${O}_synccheck.log; tail -12 ${O}_synccheck.log;;
 ncu)     timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file ${O}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-graph --batch ${NCU_BATCH:-8} --ncu-range > ${O}_ncu_bench.log 2>&1; tail -2 ${O}_ncu_bench.log;;
 ncufull) for K in ${NCU_KERNELS:-warp_var_fwd_tma_kernel softargmin_fwd_smem_kernel conv3d_tc_kernel featnet_front_kernel}; do timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$K -c ${NCU_COUNT:-1} -f -o ${O}_ncu_$K python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-graph --batch 1 --ncu-range > ${O}_ncufull_$K.log 2>&1; tail -1 ${O}_ncufull_$K.log; done;;
 output)  timeout 600 python bench.py --workload output --steps 10 --warmup 3 > ${O}_output.json 2> ${O}_output.err; echo "rc=$?" >> ${O}_output.err; cat ${O}_output.json; tail -3 ${O}_output.err;;
 train)   timeout 900 python bench.py --workload train --steps 5 --warmup 3 > ${O}_train.json 2> ${O}_train.err; echo "rc=$?" >> ${O}_train.err; cat ${O}_train.json; tail -3 ${O}_train.err;;
 cvp)     timeout 900 python bench.py --workload cvp --steps 10 --warmup 3 > ${O}_cvp.json 2> ${O}_cvp.err; echo "rc=$?" >> ${O}_cvp.err; cat ${O}_cvp.json; tail -3 ${O}_cvp.err;;
 ncutrain) timeout 900 ncu --set full --clock-control none --import-source on -k regex:"warp_var_bwd16s_kernel|conv3d_wgrad_mma_kernel" -c 4 -f -o ${O}_ncu_train python tools/train_kernels_prof.py > ${O}_ncutrain.log 2>&1; tail -1 ${O}_ncutrain.log;;
 *)       echo "unknown step $S";;
esac; done
