"""Host-side pieces of the dict API that need no GPU: spaces shim, scenario parsing, shard bookkeeping."""
import numpy as np
import pytest

from racing_dreamer_b200 import compat, spaces


def test_spaces_shim():
    b = spaces.Box(np.array([-1.0, -1.0], np.float32), np.array([1.0, 1.0], np.float32))
    assert b.shape == (2,) and b.low.dtype == np.float32
    x = b.sample()
    assert x.shape == (2,) and np.all(x >= -1) and np.all(x <= 1)
    d = spaces.Dict({"A": spaces.Dict({"lidar": spaces.Box(0.0, 15.0, shape=(1080,), dtype=np.float32)})})
    assert d["A"]["lidar"].shape == (1080,) and list(d.spaces) == ["A"]
    assert spaces.Box(-np.inf, np.inf, shape=(1,), dtype=np.float32).sample().shape == (1,)


def test_load_scenario(tmp_path):
    # same shape as the reference's scenario files [REF dreamer/scenarios/max_progress/austria.yml:1-10]
    p = tmp_path / "track.yml"
    p.write_text("world:\n  name: columbia\nagents:\n  - id: A\n    vehicle:\n      name: racecar\n"
                 "      sensors: [lidar, pose, velocity]\n    task:\n      task_name: maximize_progress\n"
                 "      params: {laps: 3, time_limit: 120.0, terminate_on_collision: False, collision_reward: -2.0}\n")
    sc = compat.load_scenario(p)
    assert sc == {"track": "columbia", "task": "maximize_progress", "laps": 3, "time_limit": 120.0,
                  "terminate_on_collision": False, "collision_reward": -2.0, "sensors": ["lidar", "pose", "velocity"]}


def test_scenario_defaults_match_reference_files():
    assert compat.SCENARIO_DEFAULTS["max_progress"]["laps"] == 10     # [REF dreamer/scenarios/max_progress/*.yml]
    assert compat.SCENARIO_DEFAULTS["eval"]["laps"] == 1              # [REF dreamer/scenarios/eval/*.yml]


def test_reference_env_needs_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from racing_dreamer_b200._abi import NativeLibraryError
    with pytest.raises((NativeLibraryError, RuntimeError)):
        compat.ReferenceEnv("austria")


def test_load_multi_agent_scenario(tmp_path):
    # the baselines' four-car files [REF baselines/scenarios/max_progress/austria.yml:1-34]
    p = tmp_path / "austria.yml"
    p.write_text("world:\n  name: austria\nagents:\n"
                 "  - id: A\n    vehicle: {name: racecar, sensors: [lidar, pose]}\n"
                 "    task: {task_name: maximize_progress, params: {laps: 10, time_limit: 180.0}}\n"
                 "  - id: B\n    vehicle: {name: racecar, sensors: [lidar]}\n"
                 "    task: {task_name: n_step_progress, params: {n_steps: 7}}\n")
    sc = compat.load_scenario(p)
    assert sc["agents_per_world"] == 2 and sc["agent_tasks"] == ["maximize_progress", "n_step_progress"]
    assert sc["agent_ids"] == ["A", "B"] and sc["n_step_progress"] == 7 and sc["laps"] == 10
    p.write_text("world:\n  name: austria\nagents:\n" + "".join(
        f"  - id: {c}\n    task: {{task_name: maximize_progress}}\n" for c in "ABCDE"))
    with pytest.raises(ValueError, match="at most"):
        compat.load_scenario(p)
