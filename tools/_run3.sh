timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
python tools/show_bench.py gpurun_out/r02c_bench.json 2>/dev/null | head -16
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | head -c 700
