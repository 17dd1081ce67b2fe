# N = 1 evidence of a build: bench line, reference arm, per-op breakdown, in-kernel trace, smoke, parity prints, standalone kernel
# timings and the ncu launch list of one eager step (each ncu pass only after the same command has exited 0 without it).
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err || exit 1
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/ref.err
python bench.py --no-sliding-window --no-ranking --no-cpu-baseline --no-dp128 --no-augment --breakdown > /dev/null 2> gpurun_out/r02_breakdown.txt
python tools/trace_step.py > gpurun_out/r02_insitu_trace.txt 2>&1
python __graft_entry__.py smoke > gpurun_out/r02_smoke.txt 2>&1
python -m pytest tests -m gpu -q -s -k 'config1 or config2' 2>&1 | grep '^\[' > gpurun_out/r02_parity_prints.txt
python tools/prof_kernels.py > gpurun_out/r02_prof_kernels.txt 2>&1
python bench.py --no-graph --steps 1 --warmup 3 --no-sliding-window --no-ranking --no-dp128 --no-augment --no-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 1000 --csv --log-file gpurun_out/launches_r02.csv python bench.py --no-graph --steps 1 --warmup 3 --no-sliding-window --no-ranking --no-dp128 --no-augment --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
wc -l gpurun_out/launches_r02.csv
cat gpurun_out/r02_parity_prints.txt
