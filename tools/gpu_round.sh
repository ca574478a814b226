mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/t_gpu.log
for pdl in 1 0; do
  echo "== PC_PDL=$pdl"; PC_PDL=$pdl timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ttft', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ttft_ms'],3), 'decode', round(d['decode_ms_per_token'],3), 'attn_alone_us', round(d['roofline_attention']['avg_launch_ms']*1e3,2))"
done
echo "== native linears"; PC_LINEAR=native timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ttft', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ttft_ms'],3), 'decode', round(d['decode_ms_per_token'],3))"
