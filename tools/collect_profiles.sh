#!/bin/bash
# Everything the profiles/ directory is built from, in one GPU call.  Usage: bash tools/collect_profiles.sh <tag>
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
python bench.py > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err
NCU="ncu --clock-control none --profile-from-start off"
$NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_${TAG}.csv python bench.py --profile --no-extras > /dev/null 2>&1
$NCU --set full --import-source on -k regex:attn_small -c 2 -f -o $O/ncu_attn_small_${TAG} python bench.py --profile --no-extras > /dev/null 2>&1
$NCU --set full --import-source on -k regex:linear_small -c 5 -f -o $O/ncu_linear_small_${TAG} python bench.py --profile --no-extras > /dev/null 2>&1
PC_BENCH_PROFILE_COPIED=1 $NCU --set full --import-source on -k regex:splice_kernel -c 1 -f -o $O/ncu_splice_${TAG} python bench.py --profile --no-extras > /dev/null 2>&1
ncu --clock-control none --set full --import-source on -k regex:attn_pp -c 2 -f -o $O/ncu_attn_pp_${TAG} python tools/attn_point.py 32,32,128,1024,8192 > /dev/null 2>&1
ncu --clock-control none --set full --import-source on -k regex:attn_pp -c 2 -f -o $O/ncu_attn_pp_falcon_${TAG} python tools/attn_point.py 71,1,64,256,2048 > /dev/null 2>&1
python tools/attn_point.py 32,32,128,16,2048 32,32,128,64,2048 32,32,128,256,2048 32,32,128,1024,2048 32,32,128,2048,0 32,32,128,32,8192 32,32,128,1024,8192 71,1,64,21,256 71,1,64,256,2048 2>/dev/null | grep "^{" > $O/attn_points_${TAG}.jsonl
python tools/eval_sys.py critical_point --repeat 2 --out $O/eval_sys_critical_point_${TAG}.json > $O/eval_sys_cp.log 2>&1
python tools/eval_sys.py latency --repeat 2 --out $O/eval_sys_latency_${TAG}.json > $O/eval_sys_lat.log 2>&1
python -m pytest tests/test_reference_gpu.py -m gpu -q -s -k "c2_llama7b" 2>&1 | grep -o "\[C2 full depth\].*" > $O/c2_full_depth_${TAG}.txt
ls -la $O | tail -20
