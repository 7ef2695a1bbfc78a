"""Run the UNMODIFIED reference follow-the-gap agent in the build container (test infrastructure).

The agent [REF ros_agent/agents/follow_the_gap/src/agent.py] is a ROS node: it needs ``rospy`` and four message
packages at import time, and uses ``np.int`` / ``np.float`` (removed from NumPy).  ``load_reference_agent()`` puts
minimal stand-ins into ``sys.modules`` (a controllable clock, message structs, a publisher that records what is
published), adds the two NumPy aliases, and loads the reference file straight from ``/root/reference`` (nothing is
copied).  ``ReferenceGapFollower`` then feeds it one ``LaserScan`` per env step and returns the drive commands it
publishes.  Used only by ``tests/golden/make_golden.py`` and by tests that are skipped when /root/reference is absent.
"""
from __future__ import annotations

import importlib.util
import sys
import types
import warnings
from pathlib import Path

import numpy as np

REFERENCE_ROOT = Path("/root/reference")
AGENT_FILE = REFERENCE_ROOT / "ros_agent" / "agents" / "follow_the_gap" / "src" / "agent.py"


class _Clock:
    now = 0.0


class _Time:
    def __init__(self, secs=0, nsecs=0):
        self.secs, self.nsecs = int(secs), int(nsecs)

    def is_zero(self):
        return self.secs == 0 and self.nsecs == 0

    def to_sec(self):
        return float(self.secs) + float(self.nsecs) / 1e9

    @staticmethod
    def now():
        return _Time.from_sec(_Clock.now)

    @staticmethod
    def from_sec(t):
        secs = int(t)
        return _Time(secs, int(round((t - secs) * 1e9)))


class _Publisher:
    def __init__(self, name=None, data_class=None, queue_size=1):
        self.name, self.sent = name, []

    def publish(self, msg):
        self.sent.append(msg)


def _ns(**kw):
    return types.SimpleNamespace(**kw)


class LaserScan:
    def __init__(self):
        self.header = _ns(stamp=_Time())
        self.angle_min = self.angle_max = self.angle_increment = 0.0
        self.range_min, self.range_max = 0.0, 0.0
        self.ranges = []


class Float32:
    def __init__(self):
        self.data = 0.0


class AckermannDriveStamped:
    def __init__(self):
        self.header = _ns(stamp=_Time(), frame_id="")
        self.drive = _ns(steering_angle=0.0, speed=0.0)


def load_reference_agent():
    """-> the reference module object (AgentNode, PID)."""
    if not AGENT_FILE.exists():
        raise FileNotFoundError(AGENT_FILE)
    rospy = types.ModuleType("rospy")
    rospy.Time = _Time
    rospy.Subscriber = lambda *a, **k: None
    rospy.Publisher = _Publisher
    rospy.init_node = lambda *a, **k: None
    rospy.spin = lambda: None
    mods = {"rospy": rospy}
    for pkg, names in (("std_msgs", {"Float32": Float32}), ("sensor_msgs", {"LaserScan": LaserScan}),
                       ("nav_msgs", {"Odometry": type("Odometry", (), {})}),
                       ("ackermann_msgs", {"AckermannDriveStamped": AckermannDriveStamped})):
        top = types.ModuleType(pkg)
        sub = types.ModuleType(pkg + ".msg")
        for k, v in names.items():
            setattr(sub, k, v)
        top.msg = sub
        mods[pkg], mods[pkg + ".msg"] = top, sub
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    if not hasattr(np, "int"):
        np.int = int      # noqa: NPY001 -- aliases the reference file still uses
    if not hasattr(np, "float"):
        np.float = float  # noqa: NPY001
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            spec = importlib.util.spec_from_file_location("ref_follow_the_gap_agent", AGENT_FILE)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


class ReferenceGapFollower:
    """One reference AgentNode driven scan by scan.  ``scan_ros`` follows the ROS convention (counter-clockwise,
    index 0 = angle_min): for the env's ``lidar`` observation (index 0 = left) pass ``lidar[::-1]``
    [REF ros_agent/agents/dreamer/src/agent.py:65 uses the same flip in the other direction]."""

    def __init__(self, angle_min, angle_increment, n_beams, range_max, dt):
        import io
        import contextlib
        self.mod = load_reference_agent()
        with contextlib.redirect_stdout(io.StringIO()):
            self.node = self.mod.AgentNode(phase="auto", identifier="ego")
        self.angle_min, self.inc, self.n = float(angle_min), float(angle_increment), int(n_beams)
        self.range_max, self.dt = float(range_max), float(dt)
        self.t = 1.0   # non-zero start so that rospy.Time.is_zero() is False from the first scan on

    def __call__(self, scan_ros):
        """-> (published: bool, steering_angle, speed, heading) after this scan."""
        import io
        import contextlib
        msg = LaserScan()
        msg.angle_min, msg.angle_increment = self.angle_min, self.inc
        msg.angle_max = self.angle_min + self.inc * (self.n - 1)
        msg.range_max = self.range_max
        msg.ranges = [float(x) for x in scan_ros]
        _Clock.now = self.t
        msg.header.stamp = _Time.from_sec(self.t)
        n_before = len(self.node.drive_pub.sent)
        with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            self.node.laserscan_callback(msg)
        self.t += self.dt
        published = len(self.node.drive_pub.sent) > n_before
        return published, float(self.node.steering_angle), float(self.node.vehicle_speed), float(self.node.heading_error)
