"""Small run that touches every kernel of libaz_b200, for compute-sanitizer (memcheck / racecheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "custom-alphazero_b200"))
import numpy as np, torch
from az_b200 import engine, env, selfplay, net as N

torch.manual_seed(0)
for W, H, n, g in [(7, 6, 4, True), (5, 5, 3, False), (9, 9, 5, True)]:
    rules = engine.Rules(W, H, n, g)
    # in-kernel evaluator, host uniforms, small pool -> compaction path
    eng = engine.TreeEngine(rules, n_trees=6, sims_per_move=24, eval_mode="hash", prior_mode="f64", move_mode="host_uniforms",
                            node_capacity=(24 + 2) * ((rules.n_actions + 7) & ~7) * 3 + 64)
    eng.set_uniforms(np.random.RandomState(1).random_sample((6, rules.max_plies)))
    for _ in range(rules.max_plies + 1):
        eng.search(); eng.play()
    torch.cuda.synchronize(); eng.check_status()
    fin = eng.drain_finished()
    assert len(fin["len"]) == 6, fin["len"]
    cells = np.zeros((3, H, W), np.int8)
    out, st = env.env_play(rules, cells, np.asarray([0, 1, 2], np.int32)); env.env_legal(rules, out); env.env_encode(rules, out)
    # fused runner (k_advance, k_play, k_extra, cuDNN tower, decode)
    # routes: whole-net kernel (az_step_gather + az_net_forward_gathered, CTA-pair and single-CTA layouts), az_advance_fused
    # + az_net_tower with extra simulations, three-kernel route; evaluation memo (seqlock) on for two of them
    for whole, fused, extra, memo, pair in ((True, False, 0, 10, "1"), (True, False, 0, 0, "0"), (False, True, 4, 10, "1"),
                                            (False, False, 0, 0, "1")):
        os.environ["AZ_TOWER_PAIR"] = pair
        r = selfplay.SelfPlayRunner(rules, n_trees=8, sims_per_move=12, games_target=12, unroll=2, use_graph=False, fused=fused,
                                    extra_sims=extra, max_free_sims=2, whole_net=whole, eval_cache_log2=memo)
        r.run_until_done(poll_every=32, max_advances=20000)
        s, p, v = r.collect()
        assert len(v) == r.totals()["moves"] and r.totals()["games"] == 12
    print("ok", (W, H, n, g), flush=True)
print("sanitize smoke done")
