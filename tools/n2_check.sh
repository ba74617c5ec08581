timeout 600 python -m pytest tests/test_multiprocess.py -m gpu -x -q 2>&1 | tail -5
bash tools/scale_run.sh 2
mv gpurun_out/bench_final_n2.json gpurun_out/bench_n2_native.json
GCM_B200_TORCH_HALO=1 bash tools/scale_run.sh 2
mv gpurun_out/bench_final_n2.json gpurun_out/bench_n2_torch.json
