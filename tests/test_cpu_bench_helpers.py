"""bench.py's bookkeeping that needs no GPU: the workloads are BASELINE.json's, the roofline object states the bytes of
the launch it times (one track's share of the envs when a batch spans several tracks), the committed ncu figures resolve."""
import json
from pathlib import Path

import bench

ROOT = Path(__file__).resolve().parent.parent


def test_workloads_are_the_baseline_configs():
    base = json.loads((ROOT / "BASELINE.json").read_text())
    assert base["metric"].startswith("env") or "env" in json.dumps(base)[:2000]
    sizes = {2: 4096, 3: 16384, 4: 65536, 5: 131072}
    for cid, n in sizes.items():
        wl = bench.workload_of(cid)
        assert wl.envs == n and str(n) in wl.name().replace(" ", "")
    assert bench.workload_of(3).obs == "lidar_occupancy" and bench.workload_of(5).tracks == ("barcelona", "austria")
    assert bench.ALGO_BYTES_PER_ENV_STEP == 224 + 8 + 4 * bench.N_BEAMS + 24 + 48


def _timing(step_ms, lidar_ms, occ_ms=0.0, steps=10, lidar_launches=10, occ_launches=0):
    return {"step_ms": step_ms * steps, "step_launches": steps, "lidar_ms": lidar_ms * lidar_launches,
            "lidar_launches": lidar_launches, "occupancy_ms": occ_ms * max(1, occ_launches), "occupancy_launches": occ_launches}


def test_roofline_counts_the_bytes_of_the_timed_launch():
    # one track: the launch covers every env
    wl = bench.workload_of(2)
    per, roof = bench.roofline_of(wl, wl.envs, _timing(0.023, 0.093), 1.25, 3.3e7)
    assert roof["kernel"] == "k_lidar" and abs(per["k_lidar"] - 0.093) < 1e-12
    assert roof["algorithmic_bytes_per_launch"] == (4 * bench.N_BEAMS + 48) * wl.envs
    assert abs(roof["achieved"] - roof["algorithmic_bytes_per_launch"] / 0.093e-3 / 1e9) < 1e-6
    assert 0 < roof["frac"] < 1 and roof["bound"] == "hbm" and roof["unit"] == "GB/s"
    # two tracks: only the first track's launch is event-bracketed and it covers half of the envs
    wl5 = bench.workload_of(5)
    _, roof5 = bench.roofline_of(wl5, wl5.envs, _timing(0.119, 1.17), 22.5, 5.8e7)
    assert roof5["algorithmic_bytes_per_launch"] == (4 * bench.N_BEAMS + 48) * wl5.envs / 2
    # occupancy config: the dominant kernel is k_occupancy
    wl3 = bench.workload_of(3)
    _, roof3 = bench.roofline_of(wl3, wl3.envs, _timing(0.033, 0.33, 2.38, occ_launches=10), 27.6, 5.9e6)
    assert roof3["kernel"] == "k_occupancy" and roof3["algorithmic_bytes_per_launch"] == (4096 + 48 + 24) * wl3.envs


def test_committed_ncu_figures_resolve():
    t = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())
    for cfg in t.values():
        if not isinstance(cfg, dict):
            continue   # free-text notes
        for entry in cfg.values():
            if isinstance(entry, dict) and "source" in entry:
                assert (ROOT / entry["source"]).exists(), entry["source"]
    assert bench.ncu_traffic(2, 4096, "k_lidar") is not None
    oc = bench.ncu_on_chip(2, 4096, "k_lidar")
    assert oc and 0 < oc["issue_active_pct"] <= 100 and oc["warp_inst_per_beam_group"] < 480
