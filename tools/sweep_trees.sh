for cfg in "2 8" "2 16" "2 32" "1 8" "1 16" "1 32" "3 16"; do
  set -- $cfg
  python bench.py --steps 4 --warmup 3 --cpu-seconds 0 --max-free $1 --net-tree-sims $2 > gpurun_out/bench_mf$1_t$2.json 2> gpurun_out/bench_mf$1_t$2.err || tail -5 gpurun_out/bench_mf$1_t$2.err
  python - <<P
import json
d=json.loads(open("gpurun_out/bench_mf$1_t$2.json").read().strip().splitlines()[-1])
print("RESULT mf=$1 inside=$2", round(d["value"]/1e6,3), "Msims/s", round(d["ms_per_step"],1), "ms/step evals", round(d["leaf_evals_per_sec"]/1e6,3), "e2e", round(d["e2e"]["value"]/1e6,3), "net_ms", round(d["roofline"]["ms_per_launch"],4), "tree_ms", round(d["roofline_tree"]["ms_per_launch"],4))
P
done
