#!/bin/bash
# Round evidence in one call: smoke, all -m gpu tests, default bench (+reference arm), bench lines of the other
# workloads, ncu launch lists (config 2, config 4) and --set full captures of the 1-D kernels and of k_pcg.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" > gpurun_out/check.log
timeout -s KILL 300 python __graft_entry__.py smoke >> gpurun_out/check.log 2>&1; echo "smoke rc=$?" >> gpurun_out/check.log
echo "== gpu tests" >> gpurun_out/check.log
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 900 --timeout-method=thread -p no:cacheprovider >> gpurun_out/check.log 2>&1; echo "tests rc=$?" >> gpurun_out/check.log
tail -6 gpurun_out/check.log
timeout -s KILL 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
for w in c5a c5b c3 c4; do
  timeout -s KILL 600 python bench.py --workload $w --steps 3 --no-cpu --no-e2e > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"
done
SHORT="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv $SHORT > gpurun_out/ncu_launches.log 2>&1; echo "launches c2 rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k k1d_pipe -s 8 -c 2 -f -o gpurun_out/prof_pipe_final $SHORT > gpurun_out/ncu_pipe_final.log 2>&1; echo "full c2 rc=$?"
SHORT4="python bench.py --workload c4 --steps 1 --warmup 1 --no-e2e --no-cpu"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv $SHORT4 > gpurun_out/ncu_launches_c4.log 2>&1; echo "launches c4 rc=$?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k k_pcg -s 6 -c 1 -f -o gpurun_out/prof_pcg_final $SHORT4 > gpurun_out/ncu_pcg_final.log 2>&1; echo "full c4 rc=$?"
for f in bench bench_ref bench_c5a bench_c3 bench_c4; do echo "--- $f"; head -c 2500 gpurun_out/$f.json; echo; done
