set -x
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02b_launches.csv python tools/run_steps.py --steps 2 > gpurun_out/r02b_run_steps.log 2>&1
tail -2 gpurun_out/r02b_run_steps.log
for spec in "qkv_ln:gemm_bf16_tn_2cta" "proj_ln2:gemm_bf16_tn_2cta" "ffn1_ln:gemm_bf16_tn_2cta" "attn:attention_kernel"; do
  k=${spec%%:*}; r=${spec##*:}
  ncu --set full --clock-control none --import-source on -k regex:$r --launch-skip 1 -c 1 -o gpurun_out/r02b_$k python tools/prof_ops.py $k --iters 1 > /dev/null 2>&1
done
python bench.py --config 3 --steps 10 --warmup 3 > gpurun_out/r02b_config3.json 2> gpurun_out/r02b_config3.err
python bench.py --config 4 --steps 10 --warmup 3 > gpurun_out/r02b_config4.json 2> gpurun_out/r02b_config4.err
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_ops_gpu.py -q -x -k "ln_fold" > gpurun_out/r02b_memcheck.log 2>&1
tail -5 gpurun_out/r02b_memcheck.log
ls -la gpurun_out/r02b_*
