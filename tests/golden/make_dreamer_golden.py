"""Generates tests/golden/dreamer_policy_golden.npz (run in the build container, where /root/reference exists):

    python tests/golden/make_dreamer_golden.py

For the shipped checkpoint ros_agent/checkpoints/austria_dreamer it records, from oracle/dreamer_policy.py in float64,
six consecutive RacingDreamer.action steps of 64 agents in closed loop with the CPU oracle env (austria, random reset,
action_repeat 4) on fixed standard-normal draws (the reference's own draws come from TensorFlow's RNG and cannot be
replayed); the scans of every step are stored, so the GPU test is teacher-forced on them.  It also checks that the packaged
racing_dreamer_b200/data/checkpoints/austria_dreamer.npz holds exactly the arrays of the reference's pickles."""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import dreamer_policy as dp  # noqa: E402
from racing_dreamer_b200.policy import load_dreamer_checkpoint  # noqa: E402

REF = pathlib.Path("/root/reference/ros_agent/checkpoints/austria_dreamer")


def make_env(n):
    from oracle.binding import Oracle, default_config
    from racing_dreamer_b200 import _abi, load_track
    cfg = default_config()
    cfg.n_envs = n
    cfg.action_repeat = 4
    cfg.auto_reset = 1
    cfg.reset_mode = _abi.RESET_RANDOM
    cfg.seed = 7
    env = Oracle(cfg, [load_track("austria")])
    env.reset(mode=_abi.RESET_RANDOM)
    return env


def main():
    w = dp.load_checkpoint(REF)
    packaged = load_dreamer_checkpoint("austria_dreamer")
    assert set(packaged) == set(w) and all(np.array_equal(packaged[k], w[k]) for k in w), "packaged checkpoint differs"
    rng = np.random.RandomState(20260)
    n, steps = 64, 6
    env = make_env(n)
    out = dict(scans=[], eps_stoch=[], eps_actor=[], mean=[], std=[], stoch=[], deter=[], actor_mean=[], actor_std=[],
               action=[], logp=[], index=[], logp_all=[])
    state = None
    for _ in range(steps):
        scan = env.out["lidar"].copy()
        es = rng.standard_normal((n, dp.STOCH)).astype(np.float32)
        ea = rng.standard_normal((n, dp.N_SAMPLES, 2)).astype(np.float32)
        act, state, d = dp.policy_step(w, scan, state, es, ea, np.float64)
        env.step(actions=act.astype(np.float32))
        for k, v in (("scans", scan), ("eps_stoch", es), ("eps_actor", ea), ("mean", d["mean"]), ("std", d["std"]),
                     ("stoch", state[0]), ("deter", state[1]), ("actor_mean", d["actor_mean"]), ("actor_std", d["actor_std"]),
                     ("action", act), ("logp", d["logp"]), ("index", d["index"]), ("logp_all", d["logp_all"])):
            out[k].append(v)
    np.savez_compressed(ROOT / "tests" / "golden" / "dreamer_policy_golden.npz", **{k: np.stack(v) for k, v in out.items()})
    print("written", {k: np.stack(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
