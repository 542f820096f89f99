# tools/measure_round.sh — the single-GPU measurement pass behind profiles/r2 (run on the GPU box: gpurun -- bash tools/measure_round.sh)
mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -q > gpurun_out/r2_gputests.log 2>&1; tail -3 gpurun_out/r2_gputests.log
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 600 gpurun_out/r2_bench_n1.json
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_steps20.json 2> gpurun_out/r2_bench_n1_steps20.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_reference_arm.json
# every launch of a short bench run with its device time.  MRSB_NO_GRAPH=1: ncu does not list the kernels of a graph that contains
# a conditional node, so the library's fall-back path is profiled, which rebuilds the table on EVERY pass (see profiles/r2/README.md)
B="python bench.py --steps 4 --warmup 3 --reps 2 --fast-forward 40 --no-cpu-baseline --no-secondary --no-parity"
MRSB_NO_GRAPH=1 $B > gpurun_out/plain_launches.log 2>&1 && MRSB_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_ncu_launches.log 2>&1
# the steady-state kernels of one tick (after 560 ticks of flight), full metric set with source
T="python tools/time_tick.py"
TICKS=300 MRSB_NO_GRAPH=1 $T > gpurun_out/plain_full.log 2>&1 && TICKS=300 MRSB_NO_GRAPH=1 ncu --set full --clock-control none --import-source on --launch-skip 5040 -c 18 -o gpurun_out/prof_r2 $T > gpurun_out/r2_ncu_full.log 2>&1; tail -1 gpurun_out/r2_ncu_full.log | cut -c1-150
python tools/time_step.py > gpurun_out/r2_time_step.json 2>&1; cat gpurun_out/r2_time_step.json
