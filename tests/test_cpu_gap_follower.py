"""Follow-the-gap controller: the numpy restatement (oracle/gap_follower.py) against the golden
commands recorded from the UNMODIFIED reference node, and the C ABI's derived parameters against the restatement's."""
import ctypes as C

import numpy as np
import pytest

from oracle.gap_follower import GapFollowerOracle, GapFollowerParams, percentile_linear
from racing_dreamer_b200 import _abi


def golden_sequences(golden_dir):
    g = np.load(golden_dir / "gap_follower_golden.npz")
    for si in range(int(g["n_seq"])):
        R, s0, s1, nb = (int(x) for x in g[f"seq{si}_meta"])
        yield si, R, s0, s1, nb, g[f"seq{si}_arc_ros"], g[f"seq{si}_cmd"]


def full_scan_ros(arc, s0, s1, nb):
    full = np.zeros(nb, np.float64)
    full[s0:s1 + 1] = arc
    return full


def test_restatement_matches_reference_golden(golden_dir):
    n = 0
    for si, R, s0, s1, nb, arcs, cmds in golden_sequences(golden_dir):
        p = GapFollowerParams(n_beams=nb, dt=R * 0.01)
        assert p.arc() == (s0, s1)
        pol = GapFollowerOracle(p)
        for arc, cmd in zip(arcs, cmds):
            pub, sa, sp, hd = pol(full_scan_ros(arc, s0, s1, nb))
            assert pub == bool(cmd[0])
            assert abs(sa - cmd[1]) <= 1e-12 and abs(sp - cmd[2]) <= 1e-12 and abs(hd - cmd[3]) <= 1e-12
            n += 1
    assert n >= 200


def test_golden_exercises_the_controller(golden_dir):
    """the fixture must contain gap candidates, saturated steering and the speed rules, or it pins nothing"""
    masked, sat, slow = 0, 0, 0
    for si, R, s0, s1, nb, arcs, cmds in golden_sequences(golden_dir):
        pol = GapFollowerOracle(GapFollowerParams(n_beams=nb, dt=R * 0.01))
        for arc in arcs:
            _, _, dbg = pol.heading_of(full_scan_ros(arc, s0, s1, nb))
            masked += int(dbg["mask"].sum())
        sat += int(np.sum(np.abs(cmds[:, 1]) >= np.deg2rad(24) - 1e-12))
        slow += int(np.sum((cmds[:, 2] < 6.0) & (cmds[:, 0] > 0)))
        assert cmds[0, 0] == 0 and cmds[1, 0] == 0 and cmds[2, 0] == 1   # the node's two first-message gates
    assert masked > 50 and sat > 0 and slow > 20


@pytest.mark.parametrize("nb,R", [(1080, 4), (1081, 8), (541, 4), (2048, 2), (360, 4)])
def test_abi_defaults_match_restatement(nb, R):
    lib = _abi.load_library()
    cfg = _abi.default_config()
    cfg.n_beams, cfg.action_repeat = nb, R
    g = _abi.RdGapFollower()
    lib.rd_gap_follower_defaults(C.byref(cfg), C.byref(g))
    p = GapFollowerParams(n_beams=nb, dt=R * 0.01)
    s0, s1 = p.arc()
    _, (lo, hi, gam) = percentile_linear(np.arange(s1 - s0 + 1, dtype=float), p.percentile_q())
    assert (g.arc_first, g.arc_last, g.filter_width, g.pct_lo, g.pct_hi) == (s0, s1, p.filter_width(), lo, hi)
    assert g.pct_gamma == gam and g.lookahead == p.lookahead and g.scan_dt == p.dt
    assert g.angle_min == p.angle_min and g.angle_increment == p.angle_increment and g.range_max == p.range_max
    assert g.vehicle_width == p.vehicle_width and g.max_steering_angle == p.max_steering_angle
    assert (g.kp, g.ki, g.kd, g.max_vehicle_speed) == (p.kp, p.ki, p.kd, p.max_vehicle_speed)
    assert g.minimum_gap_length == p.minimum_gap_length and g.median_dev_threshold == p.median_range_deviation_threshold


def test_live_reference_when_available():
    """in the build container the unmodified node itself is run beside the restatement on fresh scans"""
    from oracle import ref_ftg
    if not ref_ftg.AGENT_FILE.exists():
        pytest.skip("/root/reference not present (GPU box)")
    from oracle import Oracle, default_config
    from racing_dreamer_b200 import load_track
    tm = load_track("barcelona")
    cfg = default_config()
    cfg.n_envs = 1
    orc = Oracle(cfg, [tm], n_threads=2)
    rng = np.random.RandomState(3)
    poses = tm.reset_poses[rng.randint(0, len(tm.reset_poses), 40)].copy()
    poses[:, 2] += rng.uniform(-1.0, 1.0, 40)
    scans = orc.lidar_cast(poses)
    p = GapFollowerParams(dt=0.04)
    ref = ref_ftg.ReferenceGapFollower(p.angle_min, p.angle_increment, p.n_beams, p.range_max, p.dt)
    mine = GapFollowerOracle(p)
    for s in scans:
        ros = s[::-1].astype(np.float64)
        r, m = ref(ros), mine(ros)
        assert r[0] == m[0] and max(abs(r[i] - m[i]) for i in (1, 2, 3)) < 1e-12
