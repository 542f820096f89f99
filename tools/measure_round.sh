# tools/measure_round.sh — the single-GPU measurement pass behind profiles/ (run on the GPU box: gpurun -- bash tools/measure_round.sh)
mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r1c_n1.json 2> gpurun_out/bench_r1c_n1.err; tail -c 1500 gpurun_out/bench_r1c_n1.json
python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_r1c_ref.json 2>/dev/null; cut -c1-200 gpurun_out/bench_r1c_ref.json
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && MRSB_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r1c.log 2>&1
MRSB_NO_GRAPH=1 TICKS=300 ncu --set full --clock-control none --import-source on --launch-skip 5400 -c 11 -o gpurun_out/prof_r1c python tools/time_tick.py > gpurun_out/ncu_r1c_full.log 2>&1; tail -1 gpurun_out/ncu_r1c_full.log | cut -c1-150
