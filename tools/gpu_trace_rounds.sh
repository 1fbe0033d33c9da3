mkdir -p gpurun_out
ILQC_TRACE_ROUNDS=1 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/trace_bench.json 2> gpurun_out/trace_rounds.txt
grep -c "ilqc round" gpurun_out/trace_rounds.txt
