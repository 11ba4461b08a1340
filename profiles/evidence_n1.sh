set -x
cd $GRAFT_REPO_ROOT
timeout -s KILL 1200 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r2_gputest.txt 2>&1; tail -3 gpurun_out/r2_gputest.txt
timeout -s KILL 120 python __graft_entry__.py smoke > gpurun_out/r2_smoke.txt 2>&1; tail -1 gpurun_out/r2_smoke.txt
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -2 gpurun_out/r2_final_bench.err
timeout -s KILL 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mbx_ -c 300 --csv --log-file gpurun_out/r2_final_launches_bench.csv python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/r2_launches_bench.log 2>&1
for w in cfg2 cfg4 big detect; do timeout -s KILL 280 ncu --set full --clock-control none --import-source on -k regex:mbx_ -s 1 -c 2 -f -o gpurun_out/r2_final_$w python profiles/prof_driver.py $w 2 > gpurun_out/r2_final_ncu_$w.log 2>&1; tail -1 gpurun_out/r2_final_ncu_$w.log; done
timeout -s KILL 200 python profiles/phase_timing.py > gpurun_out/r2_final_phase_timing_match.txt 2>&1
timeout -s KILL 200 python profiles/phase_timing.py --detect > gpurun_out/r2_final_phase_timing_detect.txt 2>&1
timeout -s KILL 200 python profiles/time_detect.py > gpurun_out/r2_final_time_detect.txt 2>&1
timeout -s KILL 200 python profiles/pdl_ab.py > gpurun_out/r2_final_pdl_ab.txt 2>&1
timeout -s KILL 200 python profiles/launch_cost.py > gpurun_out/r2_final_launch_cost.txt 2>&1
for tool in memcheck racecheck synccheck; do timeout -s KILL 400 compute-sanitizer --tool $tool python profiles/sanitize.py > gpurun_out/r2_sanitize_$tool.txt 2>&1; tail -2 gpurun_out/r2_sanitize_$tool.txt; done
ls -la gpurun_out | tail -30
