# end-of-session evidence: bench line, reference arm, launch list in STEADY STATE (the skip count lands inside the warm-up
# steps that follow the 48-step pre-roll) and full captures of the two kernels of one advance (each after the plain run has
# exited 0); numbers printed under ncu are never bench values
set -x
python bench.py > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err || { tail -20 gpurun_out/bench_now.err; exit 1; }
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_now_ref.json 2> gpurun_out/bench_now_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 80000 -c 400 --csv --log-file gpurun_out/launches_now.csv python bench.py --cpu-seconds 0 --steps 1 --warmup 3 > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tower -s 40000 -c 1 -o gpurun_out/prof_net_now -f python bench.py --cpu-seconds 0 --steps 1 --warmup 3 > gpurun_out/ncu_n.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step -s 40000 -c 1 -o gpurun_out/prof_step_now -f python bench.py --cpu-seconds 0 --steps 1 --warmup 3 > gpurun_out/ncu_s.log 2>&1
tail -2 gpurun_out/bench_now.json | cut -c1-300
