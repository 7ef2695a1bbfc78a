#!/usr/bin/env python
"""bench.py -- env-steps/s (and LiDAR beams/s) of the batched racing-environment step on B200.

Contract (driver): `python bench.py --gpus N --steps K --warmup W [--impl reference]`; for N > 1 it is launched
under torchrun, one rank per GPU.  Rank 0 prints ONE JSON line.

Workload (BASELINE.json configs[1]): Austria track, 1080-beam LiDAR, 4096 batched envs per GPU, action_repeat 8,
obs 'lidar' f32, reset mode 'random' (seed 1), scripted actions motor=+0.6, steering=0.8*sin(2*pi*k/50 + phi_i),
auto-reset on.  A "step" = one env.step() of the whole batch.

* value       device-resident env-steps/s: actions already in HBM, outputs stay in HBM; CUDA events per step
              on the launching stream, L2 flushed between steps (outside the events); max over ranks.
* e2e         the same metric through the host-facing call (HostSteppedEnv.step: numpy actions in pinned
              memory -> H2D -> kernels -> D2H of every observation/result array -> numpy), wall clock.
* roofline    dominant kernel k_lidar: algorithmic bytes per launch / its mean launch duration (CUDA events
              inside librd_env, rd_enable_timing), against MEASURED_PEAKS.json hbm_gbs.
* cpu_baseline  the CPU oracle (oracle/rd_oracle.c, kind "port") on all host threads, bounded sample.
* --impl reference  the reference arm: the same oracle port timed on the host cores (the reference's env
              arithmetic lives in un-vendored racecar_gym + pybullet, so no reference build exists: DESIGN.md).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

TRACKS = ("austria",)
N_ENVS = 4096
N_BEAMS = 1080
ACTION_REPEAT = 8
PERIOD = 50
ACTIONS = "scripted"   # or "random"
SEED = 1
# BASELINE.json configs[1..4].  The driver's bench line is config 2; the others are measured with
# `--config N` for DESIGN.md / profiles and exercised by the parity tests.
CONFIGS = {
    2: dict(tracks=("austria",), envs=4096, obs="lidar", actions="scripted", seed=1),
    3: dict(tracks=("columbia",), envs=16384, obs="lidar_occupancy", actions="scripted", seed=3),
    4: dict(tracks=("treitlstrasse_v2",), envs=65536, obs="lidar", actions="random", seed=4),
    5: dict(tracks=("barcelona", "austria"), envs=131072, obs="lidar", actions="scripted", seed=5),
}
# state r/w 224 + action 8 + lidar 4320 + scalars 24 + pose/velocity 48
ALGO_BYTES_PER_ENV_STEP = 4624
# k_lidar alone: 1080 f32 ranges written + one 48-byte origin record read per env
LIDAR_BYTES_PER_ENV = N_BEAMS * 4 + 48


def workload(n, obs, cfg_id):
    return (f"config{cfg_id}: {'+'.join(TRACKS)} {N_BEAMS}-beam lidar, {n} envs/GPU, action_repeat={ACTION_REPEAT}, "
            f"obs={obs}, {ACTIONS} actions")


def scripted_actions(n, rank=0):
    """[PERIOD, n, 2] float32: motor +0.6, steering 0.8*sin(2*pi*k/50 + phi_i), phi_i from a seeded stream
    (config 4: uniform random U(-1,1)^2 instead, so that collisions / laps / time limits fire)."""
    rng = np.random.Generator(np.random.Philox(key=2 + 1000 * rank))
    if ACTIONS == "random":
        return rng.uniform(-1.0, 1.0, (PERIOD, n, 2)).astype(np.float32)
    phi = rng.uniform(0.0, 2.0 * np.pi, n)
    k = np.arange(PERIOD)[:, None]
    a = np.empty((PERIOD, n, 2), np.float32)
    a[..., 0] = 0.6
    a[..., 1] = (0.8 * np.sin(2.0 * np.pi * k / PERIOD + phi[None, :])).astype(np.float32)
    return a


def env_config(n_envs, rank=0, obs_type="lidar"):
    from racing_dreamer_b200 import EnvConfig
    return EnvConfig(tracks=TRACKS, n_envs=n_envs, action_repeat=ACTION_REPEAT, obs_type=obs_type, auto_reset=True,
                     reset_mode="random", seed=SEED, env_id_offset=rank * n_envs, time_limit_steps=2000 // ACTION_REPEAT)


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------------------------- CPU arm
def ncu_on_chip(config_id, n_envs, kernel, sm_count=148):
    """The on-chip picture of `kernel` from the committed ncu capture (SURVEY.md §8-d asks for both): issue-slot
    utilisation and the share of the shared-memory pipe its wavefronts use; None without a matching capture."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")) as f:
            ent = json.load(f).get(f"config{config_id}", {}).get(kernel)
        if ent and int(ent["envs"]) == int(n_envs) and "issue_active_pct" in ent:
            return {"bound": "issue slots", "issue_active_pct": ent["issue_active_pct"],
                    "active_lanes_per_warp_inst": ent["active_lanes_per_inst"],
                    "warp_inst_per_beam_group": ent["warp_inst_executed"] / (n_envs * ((N_BEAMS + 31) // 32)),
                    "smem_pipe_frac": ent["smem_wavefronts"] / (sm_count * ent["sm_cycles_elapsed"]),
                    "source": ent["source"] + " (ncu --set full, one launch)"}
    except (OSError, ValueError, KeyError):
        pass
    return None


def ncu_traffic(config_id, n_envs, kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/ncu_traffic.json), or None when no
    capture matches this workload."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")) as f:
            ent = json.load(f).get(f"config{config_id}", {}).get(kernel)
        if ent and int(ent["envs"]) == int(n_envs):
            return int(ent["dram_read_bytes"]) + int(ent["dram_write_bytes"])
    except (OSError, ValueError, KeyError):
        pass
    return None


def cpu_oracle_run(n_envs, steps, warmup, threads, rank=0, obs="lidar"):
    """Times the oracle port on `threads` host threads: `steps` env.step() calls of `n_envs` envs."""
    from oracle import Oracle
    from racing_dreamer_b200 import _abi, load_track
    from racing_dreamer_b200.env import _fill_config
    from oracle import default_config
    cfg = default_config()
    _fill_config(cfg, env_config(n_envs, rank, obs))
    ids = (np.arange(n_envs) % len(TRACKS)).astype(np.int32)
    orc = Oracle(cfg, [load_track(t) for t in TRACKS], ids, n_threads=threads)
    orc.reset(mode=_abi.RESET_RANDOM)
    acts = scripted_actions(n_envs, rank)
    for k in range(warmup):
        orc.step(acts[k % PERIOD])
    t0 = time.perf_counter()
    for k in range(steps):
        orc.step(acts[(warmup + k) % PERIOD])
    dt = time.perf_counter() - t0
    return n_envs * steps / dt, dt


def reference_arm(args):
    """`--impl reference`: the reference's CPU env path on the host cores.  The reference is Python whose env
    arithmetic lives in un-vendored racecar_gym + pybullet (not installable offline), so this times the oracle
    port (oracle/rd_oracle.c) with every host thread, on the same config/metric as our arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    # size each step (env sample) so that steps+warmup finish in about two minutes
    rate, _ = cpu_oracle_run(256, 2, 1, threads, obs=args.obs)
    total_steps = max(1, args.steps + args.warmup)
    n_sample = int(min(N_ENVS, max(threads, rate * 110.0 / total_steps)))
    value, dt = cpu_oracle_run(n_sample, args.steps, args.warmup, threads, obs=args.obs)
    line = {
        "metric": "env_steps_per_s", "value": value, "unit": "env-steps/s", "impl": "reference", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "beams_per_s": value * N_BEAMS,
        "config": {"workload": workload(N_ENVS, args.obs, args.config), "sample": f"{n_sample} of {N_ENVS} envs per step"},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": threads, "kind": "port",
                         "sample": f"{n_sample} envs x {args.steps} steps, oracle/rd_oracle.c, {threads} OpenMP threads"},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json config (default 2)")
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (0 = the config's)")
    ap.add_argument("--obs", default="", choices=["", "lidar", "lidar_occupancy"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-closed-loop", action="store_true", help="skip the on-device policy rollout leg")
    ap.add_argument("--no-multi-agent", action="store_true", help="skip the four-cars-per-world leg (SURVEY §8-f3)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 200)")
    ap.add_argument("--e2e-shards", type=int, default=8, help="stream shards of the host-facing env")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    global TRACKS, N_ENVS, ACTIONS, SEED
    c = CONFIGS[args.config]
    TRACKS, ACTIONS, SEED = c["tracks"], c["actions"], c["seed"]
    N_ENVS = args.envs or c["envs"]
    args.envs = N_ENVS
    args.obs = args.obs or c["obs"]
    if args.impl == "reference":
        return reference_arm(args)

    # stdout carries exactly ONE JSON line: anything native libraries print on fd 1 meanwhile (NCCL's version banner
    # when NCCL_DEBUG is set on the box) goes to stderr instead.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from racing_dreamer_b200 import BatchedRaceEnv
    from racing_dreamer_b200.host import HostSteppedEnv

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.envs
    env = BatchedRaceEnv(env_config(n, rank, args.obs), device=dev)
    acts = torch.from_numpy(scripted_actions(n, rank)).to(dev)      # inputs resident in HBM before timing
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    env.reset()
    for k in range(args.warmup):
        env.step(acts[k % PERIOD])
    barrier()

    # ---- timed region: K steps, each bracketed by CUDA events on the launching stream, L2 flushed in between ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    env.enable_timing(True)
    env.read_timing(reset=True)
    launches0 = env.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record()
        env.step(acts[(args.warmup + k) % PERIOD])      # the public API: BatchedRaceEnv.step() (launches only k_* kernels)
        ev[k][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    launches = env.launch_count - launches0
    timing = env.read_timing(reset=True)
    env.enable_timing(False)
    gpu_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    # back-to-back (no flush, one event pair) for reference
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    b0.record()
    for k in range(args.steps):
        env.step(acts[k % PERIOD])
    b1.record()
    barrier()
    b2b_ms = b0.elapsed_time(b1)
    clocks = sampler.stop() if rank == 0 else None
    stats = env.read_stats()

    t = torch.tensor([gpu_ms, b2b_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gpu_ms_max, b2b_ms_max = float(t[0]), float(t[1])

    # ---- e2e: host numpy actions -> pinned -> H2D -> step -> D2H of all results -> numpy ----
    e2e_steps = args.e2e_steps or min(args.steps, 200)
    henv = HostSteppedEnv(env_config(n, rank, args.obs), device=dev, n_shards=args.e2e_shards, bind_cpu=world > 1)
    hacts = scripted_actions(n, rank)
    henv.reset()
    for k in range(5):
        henv.step(hacts[k % PERIOD])
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        out = henv.step(hacts[(5 + k) % PERIOD])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert out["lidar"].shape == (n, N_BEAMS) and np.isfinite(out["reward"]).all()
    # pinned device->host copy rate of this box: the floor of any host-facing step is d2h_bytes / this
    probe_d = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
    probe_h = torch.empty(64 << 20, dtype=torch.uint8, pin_memory=True)
    probe_h.copy_(probe_d, non_blocking=True)
    torch.cuda.synchronize()
    tp = time.perf_counter()
    for _ in range(5):
        probe_h.copy_(probe_d, non_blocking=True)
    torch.cuda.synchronize()
    d2h_gbs = 5 * (64 << 20) / (time.perf_counter() - tp) / 1e9
    del probe_d, probe_h
    # the same call split in two (step_async / step_wait) over two half-batches: group A's kernels run while group B's
    # results cross PCIe -- the asynchronous vector-env pattern; reported next to the synchronous number, not instead
    pipe_s = None
    if n % 2 == 0:
        from racing_dreamer_b200 import EnvConfig as _EC
        import dataclasses as _dc
        half = n // 2
        base = env_config(half, rank, args.obs)
        groups = [HostSteppedEnv(_dc.replace(base, env_id_offset=rank * n + g * half), device=dev,
                                 n_shards=max(1, args.e2e_shards // 2), bind_cpu=world > 1) for g in range(2)]
        for g in groups:
            g.reset()
        acts2 = [hacts[:, :half], hacts[:, half:]]
        for k in range(5):
            for g in range(2):
                groups[g].step(acts2[g][k % PERIOD])
        barrier()
        t0 = time.perf_counter()
        groups[0].step_async(acts2[0][5 % PERIOD])
        for k in range(e2e_steps):
            groups[1].step_async(acts2[1][(5 + k) % PERIOD])
            o0 = groups[0].step_wait()
            if k + 1 < e2e_steps:
                groups[0].step_async(acts2[0][(6 + k) % PERIOD])
            o1 = groups[1].step_wait()
        pipe_s = time.perf_counter() - t0
        assert o0["lidar"].shape == (half, N_BEAMS) and np.isfinite(o1["reward"]).all()
        for g in groups:
            g.close()
    te = torch.tensor([e2e_s, pipe_s or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n * e2e_steps / float(te[0])
    pipe_value = world * n * e2e_steps / float(te[1]) if pipe_s else None

    # ---- closed loop: an on-device policy drives every env, no host round trip per step ----
    closed = None
    if not args.no_closed_loop:
        from racing_dreamer_b200 import DreamerPolicy, GapFollowerPolicy
        closed = {}
        cl_steps = min(args.steps, 500)
        for pname in ("follow_the_gap", "dreamer"):
            cenv = BatchedRaceEnv(env_config(n, rank, args.obs), device=dev)
            pol = GapFollowerPolicy(cenv) if pname == "follow_the_gap" else DreamerPolicy(cenv, "austria_dreamer", noise="philox")
            cenv.reset()
            pol.rollout(args.warmup)
            cenv.read_stats(reset=True)
            cenv.enable_timing(True)
            cenv.read_timing(reset=True)
            l0 = cenv.launch_count
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            c0.record()
            pol.rollout(cl_steps)
            c1.record()
            barrier()
            tc = torch.tensor([c0.elapsed_time(c1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tc, op=dist.ReduceOp.MAX)
            ct = cenv.read_timing(reset=True)
            cstats = cenv.read_stats()
            policy_ms = ct["policy_ms"] / max(1, ct["policy_launches"])
            leg = {"value": world * n * cl_steps / (float(tc[0]) / 1e3), "unit": "env-steps/s", "steps": cl_steps,
                   "ms_per_step": float(tc[0]) / cl_steps, "launches_per_step": (cenv.launch_count - l0) / cl_steps,
                   "kernel_ms": {"policy": policy_ms, "k_step": ct["step_ms"] / max(1, ct["step_launches"]),
                                 "k_lidar": ct["lidar_ms"] / max(1, ct["lidar_launches"])},
                   "episode_stats_rank0": cstats}
            if pname == "follow_the_gap":
                leg["policy"] = "follow_the_gap on device (k_gap_follower), back-to-back steps"
            else:
                # multiply-accumulates of one RacingDreamer.action: img1 + GRU + obs1 + obs2 + actor (h0..h3, hout)
                macs = 32 * 200 + 2 * 200 * 600 + 1280 * 200 + 200 * 60 + 230 * 400 + 3 * 400 * 400 + 400 * 4
                tf = 2.0 * macs * n / (policy_ms * 1e-3) / 1e12
                peaks = {}
                try:
                    peaks = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")))
                except Exception:
                    pass
                bf16 = float(peaks.get("bf16_tflops", 1590.0))
                leg["policy"] = ("shipped Dreamer agent austria_dreamer on device: k_embed_lidar + 9 x k_dense (tcgen05 kind::tf32, "
                                 "hi/lo x3 passes, float32-grade), Philox draws, back-to-back steps")
                leg["roofline"] = {"bound": "tensor", "kernel": "k_dense (9 launches per agent step)", "achieved": tf, "unit": "TFLOP/s",
                                   "executed_tf32_tflops": 3.0 * tf, "peak": bf16 / 2.0,
                                   "peak_source": ("measured bf16 cuBLAS peak / 2 (TF32 runs at half the bf16 rate)" if peaks else
                                                   "fallback 1.59 PFLOP/s bf16 / 2"),
                                   "frac": 3.0 * tf / (bf16 / 2.0), "flops_per_env_step": 2 * macs,
                                   "note": "launch- and latency-bound at this batch: 10 dependent launches of 32-224 CTAs each"}
            closed[pname] = leg
            cenv.close()

    # ---- multi-agent worlds: the same car count as worlds of four cars that see and hit each other,
    #      tasks of the baselines' scenario files (A maximize_progress, B..D n_step_progress), reset 'random_ball' ----
    multi = None
    if not args.no_multi_agent and n % 4 == 0:
        from racing_dreamer_b200 import EnvConfig
        mec = EnvConfig(tracks=TRACKS, n_envs=n, action_repeat=ACTION_REPEAT, obs_type=args.obs, auto_reset=True,
                        reset_mode="random_ball", seed=SEED, env_id_offset=rank * n, time_limit_steps=2000 // ACTION_REPEAT,
                        agents_per_world=4, agent_tasks=("maximize_progress",) + ("n_step_progress",) * 3)
        menv = BatchedRaceEnv(mec, device=dev)
        menv.reset()
        ma_steps = min(args.steps, 300)
        for k in range(args.warmup):
            menv.step_raw(acts[k % PERIOD].data_ptr())
        menv.read_stats(reset=True)
        menv.enable_timing(True)
        menv.read_timing(reset=True)
        mev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(ma_steps)]
        barrier()
        for k in range(ma_steps):
            flush.zero_()
            mev[k][0].record()
            menv.step_raw(acts[(args.warmup + k) % PERIOD].data_ptr())
            mev[k][1].record()
        barrier()
        tm_ = torch.tensor([float(sum(a.elapsed_time(b) for a, b in mev))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tm_, op=dist.ReduceOp.MAX)
        mt = menv.read_timing(reset=True)
        mstats = menv.read_stats()
        contacts = int((menv.buf["opponents"] != 0).sum())
        multi = {"value": world * n * ma_steps / (float(tm_[0]) / 1e3), "unit": "env-steps/s (cars)", "steps": ma_steps,
                 "ms_per_step": float(tm_[0]) / ma_steps, "worlds_per_gpu": n // 4, "agents_per_world": 4,
                 "kernel_ms": {"k_step_ma": mt["step_ms"] / max(1, mt["step_launches"]),
                               "k_lidar": mt["lidar_ms"] / max(1, mt["lidar_launches"]),
                               "k_occupancy": mt["occupancy_ms"] / max(1, mt["occupancy_launches"])},
                 "cars_in_contact_last_step_rank0": contacts, "episode_stats_rank0": mstats,
                 "note": "worlds of 4 cars (tasks A maximize_progress, B..D n_step_progress), scans see the other cars, "
                         "world-level ActionRepeat/TimeLimit/reset; L2 flushed between steps"}
        menv.close()

    # ---- the only collective of the system: episode statistics gathered across ranks at log cadence ----
    from racing_dreamer_b200.stats import gather_stats
    stats_all, _ = gather_stats(stats, device=dev)

    if rank == 0:
        value = world * n * args.steps / (gpu_ms_max / 1e3)
        peak, peak_src = measured_peak_gbs()
        lidar_ms = timing["lidar_ms"] / max(1, timing["lidar_launches"])
        lidar_gbs = LIDAR_BYTES_PER_ENV * n / (lidar_ms * 1e-3) / 1e9 if lidar_ms > 0 else 0.0
        line = {
            "metric": "env_steps_per_s", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": gpu_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64 dynamics / i32 ray march / f32 ranges", "data": "synthetic",
            "beams_per_s": value * N_BEAMS,
            "sim_ticks_per_s": value * ACTION_REPEAT,
            "config": {"workload": workload(n, args.obs, args.config), "envs_per_gpu": n, "l2": "flushed between steps (256 MiB memset outside the per-step events)",
                       "parallelism": f"env-sharded x{world}, no collective in step"},
            "ms_per_step_back_to_back": b2b_ms_max / args.steps,
            "wall_s_timed_region": wall,
            "roofline": {"bound": "hbm", "kernel": "k_lidar", "achieved": lidar_gbs, "peak": peak, "unit": "GB/s",
                         "frac": lidar_gbs / peak, "traffic": ncu_traffic(args.config, n, "k_lidar"), "peak_source": peak_src,
                         "kernel_ms": lidar_ms, "kernel_share_of_step": timing["lidar_ms"] / max(gpu_ms, 1e-9),
                         "algorithmic_bytes_per_launch": LIDAR_BYTES_PER_ENV * n,
                         "step_algorithmic_gbs": ALGO_BYTES_PER_ENV_STEP * value / world / 1e9,
                         "on_chip": ncu_on_chip(args.config, n, "k_lidar"),
                         "note": "on-chip bound (instruction issue; the map lives in shared memory), see DESIGN.md"},
            "kernel_ms": {"k_step": timing["step_ms"] / max(1, timing["step_launches"]), "k_lidar": lidar_ms,
                          "k_occupancy": timing["occupancy_ms"] / max(1, timing["occupancy_launches"])},
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": henv.h2d_bytes_per_step,
                    "d2h_bytes_per_step": henv.d2h_bytes_per_step, "steps": e2e_steps, "shards": len(henv.shards),
                    "ms_per_step": float(te[0]) / e2e_steps * 1e3, "pinned_d2h_gbs": d2h_gbs,
                    "d2h_floor_ms": henv.d2h_bytes_per_step / d2h_gbs / 1e6,
                    "two_groups_async": None if pipe_value is None else {
                        "value": pipe_value, "unit": "env-steps/s", "ms_per_step": float(te[1]) / e2e_steps * 1e3,
                        "note": "two half-batches through step_async/step_wait (rd_step_host_begin/_end): one group's "
                                "kernels overlap the other's device->host copy; same bytes per env-step"}},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "episode_stats": stats_all,
            "closed_loop": closed,
            "multi_agent": multi,
        }
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            rate, _ = cpu_oracle_run(256, 2, 1, threads, obs=args.obs)
            n_s = int(min(N_ENVS, max(threads, rate * 1.0)))     # <= ~1 s per step
            steps_s = int(min(200, max(12, 1.5 * rate / n_s)))  # ~1.5 s of wall time on every host thread
            v, dt = cpu_oracle_run(n_s, steps_s, 2, threads, obs=args.obs)
            line["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": threads, "kind": "port",
                                    "sample": f"{n_s} envs x {steps_s} steps of the same workload, oracle/rd_oracle.c, "
                                              f"{threads} OpenMP threads, {dt:.1f} s"}
        else:
            line["cpu_baseline"] = None
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    henv.close()
    env.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
