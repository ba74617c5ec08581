python bench.py --workload contact --steps 3 --warmup 1 > gpurun_out/plain_contact.log 2>&1 && tail -1 gpurun_out/plain_contact.log | cut -c1-400 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"stage_contact_kernel" --launch-skip 4 -c 2 -o gpurun_out/prof_contact -f \
  python bench.py --workload contact --steps 2 --warmup 1 > gpurun_out/ncu_contact.log 2>&1
ls -la gpurun_out/prof_contact.ncu-rep
