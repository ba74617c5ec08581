set -x
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.log 2>&1; tail -1 gpurun_out/smoke_final.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; tail -1 gpurun_out/pytest_gpu_final.log
timeout 900 python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; tail -c 3000 gpurun_out/bench_final_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; tail -c 1200 gpurun_out/bench_final_ref.json
timeout 300 python bench.py --steps 10 --warmup 3 --yield-limit 1e7 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/bench_final_plastic.json; cut -c1-300 gpurun_out/bench_final_plastic.json
timeout 300 python bench.py --size 100 --zones 1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/bench_final_1m.json; cut -c1-300 gpurun_out/bench_final_1m.json
