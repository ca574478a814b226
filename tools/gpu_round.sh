mkdir -p gpurun_out
for pf in "0" "32,-1,0,32,32" "16,-1,0,16,16" "48,-1,0,48,48" "32,-1,32,32,32" "0,-1,0,0,0" "64,-1,0,64,64"; do
  echo "== PC_PREFETCH=$pf"; PC_PREFETCH=$pf timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ttft', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ttft_ms'],3), 'decode', round(d['decode_ms_per_token'],3))"
done
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/t_gpu.log
