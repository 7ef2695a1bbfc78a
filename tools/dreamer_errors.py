"""Per-quantity error of the on-device Dreamer agent vs the float64 oracle on the golden trajectory (GPU box)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import dreamer_policy as dp
from racing_dreamer_b200 import BatchedRaceEnv, EnvConfig, DreamerPolicy
from racing_dreamer_b200.policy import load_dreamer_checkpoint

w = load_dreamer_checkpoint("austria_dreamer")
g = np.load("tests/golden/dreamer_policy_golden.npz")
T, n = g["scans"].shape[:2]
for precision in ("tf32x3", "tf32"):
    env = BatchedRaceEnv(EnvConfig(tracks=("austria",), n_envs=n, action_repeat=4), device="cuda:0")
    env.reset()
    pol = DreamerPolicy(env, w, noise="explicit", precision=precision)
    for t in range(T):
        if t == 0:
            pol.reset()
        else:
            pol.set_state(torch.from_numpy(g["stoch"][t - 1]), torch.from_numpy(g["deter"][t - 1]), torch.from_numpy(g["action"][t - 1]))
        pol.act(torch.from_numpy(g["scans"][t]).cuda(), torch.from_numpy(g["eps_stoch"][t]), torch.from_numpy(g["eps_actor"][t]), debug=True)
        d = pol.diagnostics()
        st, de, ac = pol.get_state()
        out = []
        for k, got in (("deter", de), ("mean", d["mean"]), ("std", d["std"]), ("stoch", st), ("actor_mean", d["actor_mean"]), ("actor_std", d["actor_std"])):
            ref = g[k][t]
            err = np.abs(got.cpu().numpy().astype(np.float64) - ref)
            out.append(f"{k} {err.max():.2e}/{np.abs(ref).max():.1f}")
        print(precision, t, " | ".join(out), "| idx eq", float((d["index"].cpu().numpy().astype(int) == g["index"][t]).mean()))
# float32 numpy for comparison
state = None
for t in range(T):
    if t > 0:
        state = (g["stoch"][t - 1].astype(np.float32), g["deter"][t - 1].astype(np.float32), g["action"][t - 1].astype(np.float32))
    a, s2, d = dp.policy_step(w, g["scans"][t], state, g["eps_stoch"][t], g["eps_actor"][t], np.float32)
    print("numpy-f32", t, f"deter {np.abs(s2[1]-g['deter'][t]).max():.2e} | mean {np.abs(d['mean']-g['mean'][t]).max():.2e} | actor_mean {np.abs(d['actor_mean']-g['actor_mean'][t]).max():.2e}")
