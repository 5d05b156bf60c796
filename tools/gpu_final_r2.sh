#!/bin/bash
# Round-2 evidence in one call: smoke, all -m gpu tests, default bench (+ reference arm), the bench lines of the other
# workloads, ncu launch lists and --set full captures of the dominant kernels of every workload (each only after the plain
# run of the same command exited 0).  tools/make_profiles.py r02 turns gpurun_out/ into profiles/r02_*.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" > gpurun_out/check.log
timeout -s KILL 300 python __graft_entry__.py smoke >> gpurun_out/check.log 2>&1; echo "smoke rc=$?" >> gpurun_out/check.log
echo "== gpu tests" >> gpurun_out/check.log
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 900 --timeout-method=thread -p no:cacheprovider >> gpurun_out/check.log 2>&1; echo "tests rc=$?" >> gpurun_out/check.log
tail -4 gpurun_out/check.log
timeout -s KILL 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
for w in c5a c5b c3 c4 c2e; do
  timeout -s KILL 900 python bench.py --workload $w --steps 5 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"
done
DFE_SOLVER2D=jacobi timeout -s KILL 900 python bench.py --workload c4 --steps 2 --no-cpu --no-e2e > gpurun_out/bench_c4_jacobi.json 2> gpurun_out/bench_c4_jacobi.err; echo "c4 jacobi rc=$?"
S2="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity --no-sweep"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv $S2 > gpurun_out/ncu_l2.log 2>&1; echo "launches c2 rc=$?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k k1d_pipe -s 8 -c 4 -f -o gpurun_out/prof_pipe_r2 $S2 > gpurun_out/ncu_pipe_r2.log 2>&1; echo "full c2 (+sweep) rc=$?"
S4="python bench.py --workload c4 --steps 1 --warmup 1 --no-e2e --no-cpu"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv $S4 > gpurun_out/ncu_l4.log 2>&1; echo "launches c4 rc=$?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"k_mgpcg|k_assemble_tile" -s 3 -c 2 -f -o gpurun_out/prof_c4_r2 $S4 > gpurun_out/ncu_c4_r2.log 2>&1; echo "full c4 rc=$?"
S5="python bench.py --workload c5b --steps 1 --warmup 1 --no-e2e --no-cpu"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c5b.csv $S5 > gpurun_out/ncu_l5b.log 2>&1; echo "launches c5b rc=$?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"k_band_solve_mma|k_band_rhs_fwd3|k_band_grad3|k_band_factor" -s 4 -c 5 -f -o gpurun_out/prof_c5b_r2 $S5 > gpurun_out/ncu_c5b_r2.log 2>&1; echo "full c5b rc=$?"
S5A="python bench.py --workload c5a --steps 2 --warmup 3 --no-cpu"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c5a.csv $S5A > gpurun_out/ncu_l5a.log 2>&1; echo "launches c5a rc=$?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k k1d_pipe -s 8 -c 2 -f -o gpurun_out/prof_c5a_r2 $S5A > gpurun_out/ncu_c5a_r2.log 2>&1; echo "full c5a rc=$?"
SE="python bench.py --workload c2e --steps 1 --warmup 1 --no-e2e --no-cpu"
timeout -s KILL 600 ncu --set full --clock-control none -k regex:k1d_pe -s 4 -c 4 -f -o gpurun_out/prof_c2e_r2 $SE > gpurun_out/ncu_c2e_r2.log 2>&1; echo "full c2e rc=$?"
timeout -s KILL 300 python examples/convergence_2d.py > gpurun_out/convergence_2d.log 2>&1; echo "convergence example rc=$?"
timeout -s KILL 300 python examples/topopt_heat.py > gpurun_out/topopt_heat.log 2>&1; echo "topology optimisation example rc=$?"
# the summaries are generated HERE as well (gpurun merges at most 64 MiB back: if the captures are larger, the largest
# .ncu-rep files are dropped after their summaries exist)
python tools/make_profiles.py r02 > gpurun_out/make_profiles.log 2>&1; echo "make_profiles rc=$?"
mkdir -p gpurun_out/profiles_r02; cp profiles/r02_* gpurun_out/profiles_r02/
while [ "$(du -sm gpurun_out | cut -f1)" -gt 56 ]; do
  big=$(ls -S gpurun_out/*.ncu-rep 2>/dev/null | head -1); [ -z "$big" ] && break
  echo "dropping $big ($(du -sm $big | cut -f1) MB) to stay under the merge limit"; rm -f "$big"
done
du -sm gpurun_out
for f in bench bench_ref; do echo "--- $f"; head -c 1500 gpurun_out/$f.json; echo; done
