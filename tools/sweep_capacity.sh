for cap in 0 131072 98304 65536; do
  python bench.py --steps 4 --warmup 3 --cpu-seconds 0 --node-capacity $cap > gpurun_out/bench_cap$cap.json 2> gpurun_out/bench_cap$cap.err || tail -5 gpurun_out/bench_cap$cap.err
  python - <<P
import json
d=json.loads(open("gpurun_out/bench_cap$cap.json").read().strip().splitlines()[-1])
print("RESULT cap=$cap", d["config"]["node_capacity"], round(d["value"]/1e6,3), "Msims/s", round(d["ms_per_step"],1), "ms/step evals", round(d["leaf_evals_per_sec"]/1e6,3), "e2e", round(d["e2e"]["value"]/1e6,3), "tree_ms", d["roofline_tree"]["ms_per_launch"])
P
done
