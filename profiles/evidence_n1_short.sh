# the part of evidence_n1.sh that depends on the host side / all-reduce changes made after the ncu captures
set -x
cd $GRAFT_REPO_ROOT
timeout -s KILL 400 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r2_gputest.txt 2>&1; tail -3 gpurun_out/r2_gputest.txt
timeout -s KILL 120 python __graft_entry__.py smoke > gpurun_out/r2_smoke.txt 2>&1; tail -1 gpurun_out/r2_smoke.txt
timeout -s KILL 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -2 gpurun_out/r2_final_bench.err
timeout -s KILL 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mbx_ -c 300 --csv --log-file gpurun_out/r2_final_launches_bench.csv python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/r2_launches_bench.log 2>&1
timeout -s KILL 100 python profiles/e2e_depth.py > gpurun_out/r2_final_e2e_depth.txt 2>&1
ls -la gpurun_out | tail -12
