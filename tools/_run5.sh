timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02f_tests.log; tail -3 gpurun_out/r02f_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
python tools/show_bench.py gpurun_out/r02f_bench.json 2>/dev/null | head -5
