set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
timeout 900 python bench.py --steps 30 --warmup 5 --dump-calls gpurun_out/r2_calls_v11.json > gpurun_out/r2_bench_v11.json 2> gpurun_out/r2_bench_v11.err; tail -c 600 gpurun_out/r2_bench_v11.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_v11_reference.json 2> gpurun_out/r2_bench_v11_reference.err; tail -c 800 gpurun_out/r2_bench_v11_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --nvtx --nvtx-include "frost_step" --csv --log-file gpurun_out/r2_step_v11.csv python bench.py --quick --steps 1 --warmup 3 > gpurun_out/r2_step_v11.log 2>&1; tail -2 gpurun_out/r2_step_v11.log
[ "${FULL:-0}" = 1 ] && { timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pw_fused_kernel|pw_chain_bwd" -c 8 -o gpurun_out/r2_fused_big --force-overwrite python tools/microbench_fused.py 3211264 16 96 --iters 1 > gpurun_out/r2_fused_big.log 2>&1; tail -2 gpurun_out/r2_fused_big.log; }
[ "${FULL:-0}" = 1 ] && { timeout 600 ncu --set full --clock-control none -k regex:"pw_fused_kernel" -c 6 -o gpurun_out/r2_fused_k360 --force-overwrite python tools/microbench_fused.py 50176 360 96 --iters 1 > gpurun_out/r2_fused_k360.log 2>&1; tail -2 gpurun_out/r2_fused_k360.log; }
[ "${FULL:-0}" = 1 ] && { timeout 600 ncu --set full --clock-control none -k regex:"pw_fused_kernel" -c 6 -o gpurun_out/r2_fused_k1440 --force-overwrite python tools/microbench_fused.py 12544 1440 192 --iters 1 > gpurun_out/r2_fused_k1440.log 2>&1; tail -2 gpurun_out/r2_fused_k1440.log; }
