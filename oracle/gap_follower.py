"""float64 NumPy restatement of the reference follow-the-gap controller -- TEST INFRASTRUCTURE ONLY.

Follows ``AgentNode.laserscan_callback`` + ``publish_drive_from_heading`` + ``PID.calculate``
[REF ros_agent/agents/follow_the_gap/src/agent.py:128-193, 200-238, 45-55] statement by statement, with the two SciPy
filters written out as explicit windows (``median_filter(size=w, mode='nearest')`` = element ``w//2`` of the sorted
window ``[i - w//2, i + w - w//2 - 1]`` with clamped indices; ``maximum_filter1d(size=w)`` = max over the same window
with reflected indices) and ``np.percentile`` written out as NumPy's linear-interpolation rule.  Pinned against the
unmodified reference class by tests/golden/gap_follower_golden.npz (generator: tests/golden/make_golden.py).

The checker for the CUDA policy kernel ``k_gap_follower`` (racing_dreamer_b200/csrc/rd_policy.cuh).
"""
from __future__ import annotations

import dataclasses

import numpy as np


@dataclasses.dataclass
class GapFollowerParams:
    """Constants of the reference node [REF agent.py:73-104] plus the scan geometry of the env's LiDAR."""
    n_beams: int = 1080
    fov: float = 270.0 * (np.pi / 180.0)       # [REF dreamer/tools.py:84-86]
    range_max: float = 15.0                    # scan_msg.range_max [REF dreamer/tools.py:274]
    dt: float = 0.04                           # seconds between scans = action_repeat * 0.01
    max_speed: float = 7.0
    max_decel: float = 8.26
    vehicle_width: float = 0.3302 * 1.2
    minimum_gap_length: float = 0.2
    median_range_deviation_threshold: float = 9.0
    kp: float = 1.4
    ki: float = 0.0
    kd: float = 0.1
    max_vehicle_speed: float = 6.0
    max_steering_angle: float = float(np.deg2rad(24))

    # ---- derived exactly as the reference derives them ----
    @property
    def angle_min(self):
        return -0.5 * self.fov

    @property
    def angle_increment(self):
        return self.fov / (self.n_beams - 1)

    @property
    def lookahead(self):
        return 2.0 * (np.square(self.max_speed) / (2.0 * self.max_decel))

    def arc(self):
        """get_lidar_scan_arc(-90 deg, +90 deg) [REF agent.py:117-126] -> (first index, last index)."""
        a = (np.deg2rad(-90.0), np.deg2rad(+90.0))
        sub = np.divide(np.subtract(a, self.angle_min), self.angle_increment).astype(int)
        return int(sub[0]), int(sub[1])

    def filter_width(self):
        return int(np.deg2rad(10.0) / self.angle_increment)

    def percentile_q(self):
        fwd = (np.deg2rad(-90.0), np.deg2rad(+90.0))
        return 100 * (1.0 - (np.deg2rad(30.0) / (fwd[1] - fwd[0])))


def percentile_linear(a, q):
    """np.percentile(a, q) (method 'linear') written out: virtual index q/100*(n-1), lerp with NumPy's two-sided form."""
    s = np.sort(np.asarray(a, dtype=np.float64))
    n = s.shape[0]
    quant = np.true_divide(q, 100)
    vi = n * quant + (1 + quant * (1 - 1 - 1)) - 1      # _compute_virtual_index(n, quant, alpha=1, beta=1)
    lo = int(np.floor(vi))
    g = vi - lo
    lo_c, hi_c = min(max(lo, 0), n - 1), min(max(lo + 1, 0), n - 1)
    x, y = s[lo_c], s[hi_c]
    d = y - x
    return (y - d * (1 - g)) if g >= 0.5 else (x + d * g), (lo_c, hi_c, g)


class GapFollowerOracle:
    """State of one controller (one env): the PID's previous input and the two "first message" gates."""

    def __init__(self, params: GapFollowerParams = None):
        self.p = params or GapFollowerParams()
        self.reset()

    def reset(self):
        self.scans = 0                 # scans seen since reset (gate of laserscan_callback)
        self.headings = 0              # headings seen (gate of publish_drive_from_heading)
        self.prev_input = np.nan       # PID.previous_input_value
        self.integral = 0.0
        self.steering_angle = 0.0
        self.vehicle_speed = 0.0
        self.heading = 0.0
        self.heading_distance = 0.0

    def heading_of(self, scan_ros):
        """scan (ROS order, float64) -> (heading angle, heading distance, debug dict)."""
        p = self.p
        s0, s1 = p.arc()
        inc, amin = p.angle_increment, p.angle_min
        angles = np.arange(s0, s1 + 1).astype(float) * inc + amin
        ranges = np.asarray(scan_ros, dtype=np.float64)[s0:s1 + 1]
        ranges = np.minimum(np.maximum(ranges, 0.0), p.lookahead)
        m = ranges.shape[0]
        diff = np.abs(ranges[1:] - ranges[:-1])
        nd = m - 1
        w = p.filter_width()
        half = w // 2
        mask = np.zeros(nd, dtype=bool)
        for i in range(nd):
            d = diff[i]
            if not d > p.minimum_gap_length:
                continue
            idx = np.arange(i - half, i - half + w)
            refl = np.where(idx < 0, -idx - 1, np.where(idx >= nd, 2 * nd - idx - 1, idx))
            if d != diff[refl].max():
                continue
            near = np.clip(idx, 0, nd - 1)
            med = np.sort(diff[near])[half]
            mask[i] = d > med * p.median_range_deviation_threshold
        adjusted = ranges.copy()
        for i in np.nonzero(mask)[0]:
            theta = angles[i]
            lo = i - 1
            long_side = np.min(ranges[max(lo, 0):i + 2])   # the reference slice is EMPTY for i == 0 (it raises there)
            beta = np.arccos((2.0 * np.square(long_side) - np.square(p.vehicle_width)) / (2.0 * np.square(long_side)))
            g0, g1 = theta - beta, theta + beta
            k0 = int((g0 - angles[0]) / inc)
            k1 = int((g1 - angles[0]) / inc)
            k0, k1 = min(max(k0, 0), m - 1), min(max(k1, 0), m - 1)
            adjusted[k0:k1 + 1] = np.minimum(adjusted[k0:k1 + 1], long_side)
        pct, sel = percentile_linear(adjusted, p.percentile_q())
        chosen = (adjusted >= pct) & (adjusted < p.range_max)     # np.digitize(...) == 2
        heading = np.mean(angles[chosen])
        dist = np.mean(ranges[chosen])
        return float(heading), float(dist), dict(mask=mask, adjusted=adjusted, pct=pct, sel=sel, chosen=chosen)

    def __call__(self, scan_ros):
        """One LaserScan -> (published, steering_angle, speed, heading) exactly as the node would leave them."""
        p = self.p
        self.scans += 1
        if self.scans == 1:            # first scan only arms the timestamp [REF agent.py:132-134]
            return False, self.steering_angle, self.vehicle_speed, self.heading
        heading, dist, _ = self.heading_of(scan_ros)
        self.heading, self.heading_distance = heading, dist
        self.headings += 1
        if self.headings == 1:         # first heading only arms the PID clock [REF agent.py:206-208]
            return False, self.steering_angle, self.vehicle_speed, self.heading
        error = 0.0 - heading          # PID.calculate [REF agent.py:45-55], target_output_value = 0
        P = p.kp * error
        self.integral += p.ki * error * p.dt
        D = 0.0 if np.isnan(self.prev_input) else (p.kd * (self.prev_input - heading) / p.dt)
        self.prev_input = heading
        control = P + self.integral + D
        sa = -control
        sa = min(max(sa, -abs(p.max_steering_angle)), abs(p.max_steering_angle))
        a = abs(sa)
        speed = p.max_vehicle_speed
        if a > np.deg2rad(5):
            speed = p.max_vehicle_speed - (a / p.max_steering_angle) * (p.max_vehicle_speed * 0.30)
        if dist < 5:
            speed = min(speed, dist / 5 * 4)
        speed = max(speed, 1.5)
        self.steering_angle, self.vehicle_speed = float(sa), float(speed)
        return True, self.steering_angle, self.vehicle_speed, self.heading
