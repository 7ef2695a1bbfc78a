"""ctypes binding of oracle/rd_oracle.c (test infrastructure; see oracle/__init__.py)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path
from typing import List, Optional, Sequence

import numpy as np

from racing_dreamer_b200 import _abi  # interface structs only (include/rd_env.h mirror)
from racing_dreamer_b200.maps import TrackMap

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "librd_oracle.so"
_lib = None


def build_oracle(force: bool = False) -> Path:
    src = HERE / "rd_oracle.c"
    hdr = HERE.parent / "include" / "rd_env.h"
    stale = (not LIB.exists()) or LIB.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime)
    if force or stale:
        subprocess.run(["make", "-C", str(HERE)] + (["-B"] if force else []), check=True, capture_output=True)
    return LIB


class _Map(C.Structure):
    _fields_ = [
        ("h", C.c_int32), ("w", C.c_int32), ("col0", C.c_int32), ("row0", C.c_int32), ("full_h", C.c_int32),
        ("dmax", C.c_int32),
        ("resolution", C.c_double), ("origin_x", C.c_double), ("origin_y", C.c_double),
        ("drivable", C.c_void_p), ("dist", C.c_void_p), ("start_poses", C.c_void_p), ("reset_poses", C.c_void_p),
        ("n_start", C.c_int32), ("n_reset", C.c_int32),
        ("ball_next", C.c_void_p),
    ]


class _Outputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "lidar", "occupancy", "pose", "velocity", "speed", "reward", "done", "progress", "lap", "time", "flags",
        "reward64", "rank", "opponents")]


def _load():
    global _lib
    if _lib is None:
        build_oracle()
        lib = C.CDLL(str(LIB))
        assert lib.orc_sizeof_config() == C.sizeof(_abi.RdConfig), "rd_config layout mismatch"
        assert lib.orc_sizeof_map() == C.sizeof(_Map), "orc_map layout mismatch"
        lib.orc_default_config.argtypes = [C.POINTER(_abi.RdConfig)]
        _lib = lib
    return _lib


def default_config() -> _abi.RdConfig:
    cfg = _abi.RdConfig()
    _load().orc_default_config(C.byref(cfg))
    return cfg


class OracleMap:
    """Unpacked y-up arrays of one compiled track, kept alive for the C side."""

    def __init__(self, tm: TrackMap):
        self.tm = tm
        self.drivable = np.ascontiguousarray(tm.drivable[::-1].astype(np.uint8))
        self.dist = np.ascontiguousarray(tm.dist[::-1].astype(np.uint16))
        self.start = np.ascontiguousarray(tm.start_poses, dtype=np.float64)
        self.reset = np.ascontiguousarray(tm.reset_poses, dtype=np.float64)

    def fill(self, m: _Map, cfg=None):
        tm = self.tm
        m.h, m.w, m.col0, m.row0, m.full_h, m.dmax = tm.h, tm.w, tm.c0, tm.cy0, tm.full_shape[0], tm.dmax
        m.resolution, m.origin_x, m.origin_y = tm.resolution, tm.origin[0], tm.origin[1]
        m.drivable = self.drivable.ctypes.data
        m.dist = self.dist.ctypes.data
        m.start_poses = self.start.ctypes.data
        m.reset_poses = self.reset.ctypes.data
        m.n_start, m.n_reset = self.start.shape[0], self.reset.shape[0]
        m.ball_next = None
        if cfg is not None and self.reset.shape[0] > 0:   # random_ball chain (multi-agent resets)
            self.ball_next = np.zeros(self.reset.shape[0], np.int32)
            _load().orc_ball_next(C.byref(cfg), C.byref(m), C.c_void_p(self.ball_next.ctypes.data))
            m.ball_next = self.ball_next.ctypes.data


class Oracle:
    """Stateful CPU env over the oracle C functions: same call surface as the CUDA path's C ABI."""

    def __init__(self, cfg: _abi.RdConfig, tracks: Sequence[TrackMap], env_map_ids: Optional[np.ndarray] = None,
                 n_threads: int = 1):
        self.lib = _load()
        self.cfg = cfg.copy()
        self.n = int(cfg.n_envs)
        self.nb = int(cfg.n_beams)
        self.n_threads = int(n_threads)
        self._maps: List[OracleMap] = [OracleMap(t) for t in tracks]
        self._cmaps = (_Map * len(self._maps))()
        for om, cm in zip(self._maps, self._cmaps):
            om.fill(cm, self.cfg)
        self.f64 = np.zeros((_abi.NF64, self.n), dtype=np.float64)
        self.i32 = np.zeros((_abi.NI32, self.n), dtype=np.int32)
        if env_map_ids is not None:
            self.i32[_abi.I_MAP] = np.asarray(env_map_ids, dtype=np.int32)
        # ring of lap + progress per tick (n_step_progress task); harmless when unused
        self.hist = np.zeros((max(1, int(cfg.n_step_progress)), self.n), dtype=np.float64)
        self.stats = _abi.RdStats()
        self.out = self._alloc_outputs()

    # -- buffers --
    def _alloc_outputs(self):
        n, nb = self.n, self.nb
        occ = bool(self.cfg.obs_flags & _abi.OBS_OCCUPANCY)
        return dict(
            lidar=np.zeros((n, nb), np.float32),
            occupancy=np.zeros((n, 64, 64), np.uint8) if occ else None,
            pose=np.zeros((n, 6), np.float32), velocity=np.zeros((n, 6), np.float32),
            speed=np.zeros(n, np.float32), reward=np.zeros(n, np.float32), done=np.zeros(n, np.uint8),
            progress=np.zeros(n, np.float32), lap=np.zeros(n, np.int32), time=np.zeros(n, np.float32),
            flags=np.zeros(n, np.uint8), reward64=np.zeros(n, np.float64), rank=np.ones(n, np.int32),
            opponents=np.zeros(n, np.uint8))

    def _c_outputs(self) -> _Outputs:
        o = _Outputs()
        for k, v in self.out.items():
            setattr(o, k, v.ctypes.data if v is not None else None)
        return o

    # -- env API --
    def reset(self, mask: Optional[np.ndarray] = None, mode: int = _abi.RESET_GRID):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        o = self._c_outputs()
        self.lib.orc_reset(C.byref(self.cfg), self._cmaps, C.c_void_p(self.f64.ctypes.data),
                           C.c_void_p(self.i32.ctypes.data), C.c_void_p(self.hist.ctypes.data),
                           C.c_void_p(m.ctypes.data if m is not None else None), C.c_int(mode), C.byref(o))
        return self.out

    def step(self, actions: np.ndarray = None, commands: np.ndarray = None):
        """actions: agent-facing float32 [n,2] (clip/rescale applied per cfg);
        commands: sim-facing float64 [n,2] = racecar_gym's {'motor','steering'} (bypasses the action transform)."""
        a = c = None
        if commands is not None:
            c = np.ascontiguousarray(commands, dtype=np.float64).reshape(self.n, 2)
        else:
            a = np.ascontiguousarray(actions, dtype=np.float32).reshape(self.n, 2)
        o = self._c_outputs()
        self.lib.orc_step(C.byref(self.cfg), self._cmaps, C.c_void_p(self.f64.ctypes.data),
                          C.c_void_p(self.i32.ctypes.data), C.c_void_p(self.hist.ctypes.data),
                          C.c_void_p(a.ctypes.data if a is not None else None),
                          C.c_void_p(c.ctypes.data if c is not None else None), C.byref(o),
                          C.byref(self.stats), C.c_int(self.n_threads))
        return self.out

    def read_stats(self, reset: bool = False):
        """Episode statistics accumulated by step() (same fields as rd_read_stats)."""
        d = self.stats.as_dict()
        if reset:
            self.stats = _abi.RdStats()
        return d

    # -- stage functions --
    def lidar_cast(self, poses: np.ndarray, map_ids: Optional[np.ndarray] = None) -> np.ndarray:
        p = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 3)
        ids = None if map_ids is None else np.ascontiguousarray(map_ids, dtype=np.int32)
        out = np.zeros((p.shape[0], self.nb), np.float32)
        self.lib.orc_lidar_cast(C.byref(self.cfg), self._cmaps, C.c_void_p(p.ctypes.data),
                                C.c_void_p(ids.ctypes.data if ids is not None else None), C.c_int(p.shape[0]),
                                C.c_void_p(out.ctypes.data), C.c_int(self.n_threads))
        return out

    def occupancy_obs(self, poses: np.ndarray, map_ids: Optional[np.ndarray] = None) -> np.ndarray:
        p = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 3)
        ids = None if map_ids is None else np.ascontiguousarray(map_ids, dtype=np.int32)
        out = np.zeros((p.shape[0], 64, 64), np.uint8)
        self.lib.orc_occupancy_obs(self._cmaps, C.c_void_p(p.ctypes.data),
                                   C.c_void_p(ids.ctypes.data if ids is not None else None), C.c_int(p.shape[0]),
                                   C.c_void_p(out.ctypes.data), C.c_int(self.n_threads))
        return out

    def dynamics(self, state: np.ndarray, commands: np.ndarray, n_ticks: int) -> np.ndarray:
        s = np.ascontiguousarray(state, dtype=np.float64).copy()
        assert s.shape[0] == 7
        cmd = np.ascontiguousarray(commands, dtype=np.float64).reshape(-1, 2)
        self.lib.orc_dynamics(C.byref(self.cfg), C.c_void_p(s.ctypes.data), C.c_void_p(cmd.ctypes.data),
                              C.c_int(s.shape[1]), C.c_int(n_ticks))
        return s

    def reward_done(self, kin: np.ndarray, steering: Optional[np.ndarray], book_f64: np.ndarray, book_i32: np.ndarray,
                    map_ids: Optional[np.ndarray] = None):
        """One tick of a7/a8 bookkeeping for teacher-forced poses (mirrors rd_reward_done): kin [5, n] = (x, y, yaw, v,
        slip) after the tick, book_f64 [3, n] = (time, progress, last), book_i32 [3, n] = (lap, checkpoint, flags).
        Returns (book_f64', book_i32', reward f64[n], done u8[n])."""
        k = np.ascontiguousarray(kin, dtype=np.float64)
        n = k.shape[1]
        st = None if steering is None else np.ascontiguousarray(steering, dtype=np.float64)
        bf = np.ascontiguousarray(book_f64, dtype=np.float64).copy()
        bi = np.ascontiguousarray(book_i32, dtype=np.int32).copy()
        ids = None if map_ids is None else np.ascontiguousarray(map_ids, dtype=np.int32)
        rew = np.zeros(n, np.float64)
        done = np.zeros(n, np.uint8)
        self.lib.orc_reward_done(C.byref(self.cfg), self._cmaps, C.c_void_p(k.ctypes.data),
                                 C.c_void_p(st.ctypes.data if st is not None else None),
                                 C.c_void_p(ids.ctypes.data if ids is not None else None), C.c_int(n),
                                 C.c_void_p(bf.ctypes.data), C.c_void_p(bi.ctypes.data), C.c_void_p(rew.ctypes.data),
                                 C.c_void_p(done.ctypes.data))
        return bf, bi, rew, done

    def beam_table(self):
        ca = np.zeros(self.nb, np.float64)
        sa = np.zeros(self.nb, np.float64)
        self.lib.orc_beam_table(C.byref(self.cfg), C.c_void_p(ca.ctypes.data), C.c_void_p(sa.ctypes.data))
        return ca, sa

    # -- library-stage pieces for pinning --
    @staticmethod
    def spline_prefilter_2d(a: np.ndarray) -> np.ndarray:
        b = np.ascontiguousarray(a, dtype=np.float64).copy()
        assert b.shape[0] == b.shape[1]
        _load().orc_spline_prefilter_2d(C.c_void_p(b.ctypes.data), C.c_int(b.shape[0]))
        return b

    @staticmethod
    def pil_resize_200_to_64(a: np.ndarray) -> np.ndarray:
        src = np.ascontiguousarray(a, dtype=np.uint8)
        assert src.shape == (200, 200)
        out = np.zeros((64, 64), np.uint8)
        _load().orc_pil_resize_200_to_64(C.c_void_p(src.ctypes.data), C.c_void_p(out.ctypes.data))
        return out
