"""Episode store -- the wire format between the env and Dreamer's replay.

Restates, for a BATCH of envs, what the reference does per env:

* ``Collect.step/reset`` [REF dreamer/wrappers.py:198-250]: every transition is the observation dict plus ``action``,
  ``reward``, ``discount = 1 - done``, ``progress = lap + progress - 1`` and ``time``; the first row of an episode is the
  reset observation with zero action, reward 0, discount 1, progress -1, time 0; floats are cast to float32, signed
  ints to int32, uint8 stays (precision 32).
* ``callbacks.save_episodes`` [REF dreamer/callbacks.py:41-53]: one ``{timestamp}-{uuid}-{length}.npz`` per episode,
  ``np.savez_compressed`` of the stacked rows.
* ``tools.count_episodes / load_episodes`` [REF dreamer/tools.py:224-264]: episode counting from file names and the
  random fixed-length chunk sampler that feeds the learner.

``EpisodeRecorder`` drives any env with the host-facing interface (``reset(mask=None, mode=None) -> dict`` and
``step(actions) -> dict`` of numpy arrays keyed lidar, pose, velocity, speed, reward, done, progress, lap, time
[, occupancy]) -- ``HostSteppedEnv`` on the GPU.  The env must be created with ``auto_reset=False``: the terminal
observation belongs to the finished episode (as in the reference), and the recorder resets exactly the finished envs.
"""
from __future__ import annotations

import datetime
import io
import pathlib
import uuid
from typing import Callable, Dict, Iterator, List, Optional, Sequence

import numpy as np

OBS_KEYS = ("lidar", "pose", "velocity", "speed", "lidar_occupancy")     # order of the reference's obs dict
EXTRA_KEYS = ("action", "reward", "discount", "progress", "time")


def save_episodes(directory, episodes: Sequence[Dict[str, np.ndarray]]) -> List[pathlib.Path]:
    """[REF dreamer/callbacks.py:41-53] -- same file naming and container; returns the paths written."""
    directory = pathlib.Path(directory).expanduser()
    directory.mkdir(parents=True, exist_ok=True)
    timestamp = datetime.datetime.now().strftime("%Y%m%dT%H%M%S")
    out = []
    for episode in episodes:
        identifier = str(uuid.uuid4().hex)
        length = len(episode["reward"])
        filename = directory / f"{timestamp}-{identifier}-{length}.npz"
        with io.BytesIO() as f1:
            np.savez_compressed(f1, **episode)
            f1.seek(0)
            with filename.open("wb") as f2:
                f2.write(f1.read())
        out.append(filename)
    return out


def count_episodes(directory):
    """(episodes, agent steps) from the file names [REF dreamer/tools.py:224-228]."""
    filenames = pathlib.Path(directory).expanduser().glob("*.npz")
    lengths = [int(n.stem.rsplit("-", 1)[-1]) - 1 for n in filenames]
    return len(lengths), sum(lengths)


def load_episodes(directory, rescan, length=None, balance=False, seed=0) -> Iterator[Dict[str, np.ndarray]]:
    """Endless sampler of (chunks of) stored episodes with the reference's RandomState call sequence
    [REF dreamer/tools.py:235-264].  Files are visited in sorted order so that the stream is reproducible."""
    directory = pathlib.Path(directory).expanduser()
    random = np.random.RandomState(seed)
    cache = {}
    while True:
        for filename in sorted(directory.glob("*.npz")):
            if filename not in cache:
                try:
                    with filename.open("rb") as f:
                        episode = np.load(f)
                        episode = {k: episode[k] for k in episode.keys()}
                except Exception as e:  # noqa: BLE001 - a half-written file must not stop training (reference behaviour)
                    print(f"Could not load episode: {e}")
                    continue
                cache[filename] = episode
        keys = list(cache.keys())
        for index in random.choice(len(keys), rescan):
            episode = cache[keys[index]]
            if length:
                total = len(next(iter(episode.values())))
                available = total - length
                if available < 1:
                    print(f"[Info] Skipped short episode of length {available}.")
                    continue
                if balance:
                    index = min(random.randint(0, total), available)
                else:
                    index = int(random.randint(0, available + 1))
                episode = {k: v[index: index + length] for k, v in episode.items()}
            yield episode


def _convert(value: np.ndarray, precision: int = 32) -> np.ndarray:
    """[REF dreamer/wrappers.py:240-250]"""
    value = np.asarray(value)
    if np.issubdtype(value.dtype, np.floating):
        dtype = {16: np.float16, 32: np.float32, 64: np.float64}[precision]
    elif np.issubdtype(value.dtype, np.signedinteger):
        dtype = {16: np.int16, 32: np.int32, 64: np.int64}[precision]
    elif np.issubdtype(value.dtype, np.uint8):
        dtype = np.uint8
    else:
        raise NotImplementedError(value.dtype)
    return value.astype(dtype)


class EpisodeRecorder:
    """Batched ``Collect``: records every env's transitions and hands finished episodes to the callbacks.

    ``callbacks``: callables taking a list with ONE episode dict (the reference calls its callbacks with one episode per
    agent of a single env [REF dreamer/wrappers.py:221-225]; here every finished env yields its own call).
    Storage is a preallocated ``[max_len + 1, n_envs, ...]`` array per key, filled column-wise; ``max_len`` must cover the
    env's TimeLimit.
    """

    def __init__(self, env, max_len: int, callbacks: Sequence[Callable] = (), precision: int = 32,
                 reset_mode: Optional[str] = None, keep_obs: Sequence[str] = OBS_KEYS,
                 agents_per_world: Optional[int] = None):
        """agents_per_world (default: the env's own ``cfg.agents_per_world``): for worlds of several cars the episodes of
        ALL cars of a world end at the step in which any of them is done -- the reference resets the world then
        [REF dreamer/tools.py:178-179] -- and the callbacks get one list with the world's episodes, agent A first, exactly
        what the reference's Collect passes for its single world [REF dreamer/wrappers.py:221-225]."""
        self.env = env
        self.n = int(env.n)
        cfg = getattr(env, "cfg", None) or getattr(getattr(env, "env", None), "cfg", None)
        A = agents_per_world if agents_per_world is not None else int(getattr(cfg, "agents_per_world", 1) or 1)
        self.agents = max(1, int(A))
        # Collect stores the TERMINAL observation of an episode and resets afterwards [REF dreamer/wrappers.py:210-226]; an
        # auto-resetting env has already replaced it with the first observation of the next episode (and advanced the
        # reset counters) by the time the recorder sees the done flag.
        if cfg is not None and int(getattr(cfg, "auto_reset", 0)):
            raise ValueError("EpisodeRecorder needs an env created with auto_reset=False: it resets finished envs itself, "
                             "after storing their terminal observation")
        if self.n % self.agents:
            raise ValueError("n_envs must be a multiple of agents_per_world")
        self.max_len = int(max_len)
        self.callbacks = tuple(callbacks)
        self.precision = precision
        self.reset_mode = reset_mode
        self.keep_obs = tuple(keep_obs)
        self._store: Dict[str, np.ndarray] = {}
        self._len = np.zeros(self.n, np.int64)
        self.episodes_done = 0

    # -- storage --
    def _obs_of(self, out: Dict[str, np.ndarray], reset: bool) -> Dict[str, np.ndarray]:
        obs = {"lidar": out["lidar"], "pose": out["pose"], "velocity": out["velocity"],
               "speed": np.zeros_like(out["speed"]) if reset else out["speed"]}       # [REF dreamer/wrappers.py:66,74]
        if "occupancy" in out:
            obs["lidar_occupancy"] = out["occupancy"]
        return {k: v for k, v in obs.items() if k in self.keep_obs}

    def _put(self, rows: Dict[str, np.ndarray], envs: np.ndarray) -> None:
        """Append one row per env in `envs` (values are [len(envs), ...])."""
        t = self._len[envs]
        if np.any(t > self.max_len):
            raise RuntimeError("episode longer than max_len: size the recorder to the env's TimeLimit")
        for k, v in rows.items():
            v = _convert(v, self.precision)
            if k not in self._store:
                self._store[k] = np.zeros((self.max_len + 1, self.n) + v.shape[1:], v.dtype)
            self._store[k][t, envs] = v
        self._len[envs] = t + 1

    def _flush(self, envs: np.ndarray) -> None:
        A = self.agents
        for e in envs[::A] if A > 1 else envs:      # worlds: `envs` holds whole worlds, one callback per world
            group = range(e, e + A) if A > 1 else (e,)
            episodes = []
            for c in group:
                n = int(self._len[c])
                episodes.append({k: self._store[k][:n, c].copy() for k in self._store})
                self._len[c] = 0
            self.episodes_done += len(episodes)
            for cb in self.callbacks:
                cb(episodes)

    # -- env API --
    def reset(self, mask: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
        out = self.env.reset(mask=mask, mode=self.reset_mode)
        envs = np.arange(self.n) if mask is None else np.flatnonzero(np.asarray(mask))
        self._len[envs] = 0
        m = len(envs)
        rows = {k: v[envs] for k, v in self._obs_of(out, reset=True).items()}
        rows.update(action=np.zeros((m, 2), np.float32), reward=np.zeros(m, np.float32), discount=np.ones(m, np.float32),
                    progress=-np.ones(m, np.float32), time=np.zeros(m, np.float32))   # [REF dreamer/wrappers.py:228-238]
        self._put(rows, envs)
        return out

    def step(self, actions: np.ndarray, auto_reset: bool = True) -> Dict[str, np.ndarray]:
        """One env step for the whole batch; finished envs are flushed to the callbacks and (auto_reset) reset, so the
        dict returned holds, for those envs, the first observation of their next episode -- `done`/`reward` stay the
        finished step's."""
        actions = np.asarray(actions, np.float32)
        out = self.env.step(actions)
        live = np.flatnonzero(self._len > 0)           # envs inside an episode (frozen ones were flushed already)
        done = out["done"].astype(bool)
        rows = {k: v[live] for k, v in self._obs_of(out, reset=False).items()}
        rows.update(action=actions[live], reward=out["reward"][live],
                    discount=(1.0 - done[live].astype(np.float64)),
                    progress=out["lap"][live].astype(np.float64) + out["progress"][live].astype(np.float64) - 1.0,
                    time=out["time"][live])                                           # [REF dreamer/wrappers.py:214-219]
        self._put(rows, live)
        if self.agents > 1:     # a world is over for all of its cars as soon as one is done
            done = np.repeat(done.reshape(-1, self.agents).any(axis=1), self.agents)
        finished = live[done[live]]
        if finished.size:
            reward, done_flags = out["reward"].copy(), out["done"].copy()
            self._flush(finished)
            if auto_reset:
                mask = np.zeros(self.n, np.uint8)
                mask[finished] = 1
                out = dict(self.reset(mask))
                out["reward"], out["done"] = reward, done_flags
        return out
