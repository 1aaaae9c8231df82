# One GPU box visit for the numbers and ncu captures quoted in DESIGN.md / profiles/ (round 1, third session).
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 400 gpurun_out/bench_final.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; tail -c 300 gpurun_out/bench_ref_final.json
python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_final.json 2> gpurun_out/bench_c4_final.err; tail -c 300 gpurun_out/bench_c4_final.json
python bench.py --workload c3 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c3_final.json 2> gpurun_out/bench_c3_final.err; tail -c 300 gpurun_out/bench_c3_final.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_final_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stage_strip -s 6 -c 2 -o gpurun_out/prof_final -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_final_b.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 30 -c 30 --csv --log-file gpurun_out/launches_c4_final.csv python bench.py --workload c4 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_final_c.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:stage_strip|general_gradients_ring" -s 9 -c 6 -o gpurun_out/prof_c4_final -f python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_final_d.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
