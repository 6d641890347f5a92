#!/usr/bin/env python
"""Benchmark of the env-step hot path (BASELINE.json metric: env-steps/s and agent-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4|c5]

Default workload = BASELINE.json configs[1] ("c2"): flight_easy 3a15t AM0TM0, 4096 envs per kernel launch.
4096 envs are ~2.6 MB of state+outputs, i.e. L2-resident, so the bench keeps 64 independent 4096-env batches
(~168 MB > the 126 MB L2) and steps them round-robin: one bench "step" is one env-step of every batch
(64 launches over 262144 env instances), and by the time a batch comes round again its lines have been evicted.
Under torchrun (N>1) every rank runs the same workload on its own GPU with its own global env ids (weak
scaling, no data-path collective); the only collective is the NCCL all-reduce of the 8-double episode-statistics
vector after the timed region.

The JSON line carries: value (device-resident inputs, CUDA-graph replay of the step launches), e2e (HOST
buffers through cs_flight_step_host: H2D actions + kernel + D2H reward/terminated/win/obs/state every step),
roofline of the step kernel against MEASURED_PEAKS.json, the CPU baseline (oracle port on the host cores) and
clocks sampled while the GPU was under this load.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

TEMPLATE = {  # flight_targets.txt of the reference (parsed as main.py:19-32 does); synthetic layout = the reference's own
    "x": [5, 2, 7.5, 2.8, 6.9, 5.5, 5.3, 1.8, 3, 4.5, 6.3, 8, 0.9, 9.4, 4.2],
    "y": [9.1, 7.5, 7, 8, 8.5, 8, 6.6, 6.8, 5.7, 5, 5.7, 6.7, 8.7, 9, 9.3],
    "deter": ["f", "t", "f", "f", "t", "f", "t", "t", "f", "f", "t", "f", "f", "t", "f"],
    "priority": [3, 3, 3, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1],
    "dx": [0.2, 0.3, 0.3, 0.27, 0.25, 0.25, 0.1, 0.28, 0.18, 0.23, 0.31, 0.29, 0.15, 0.21, 0.34],
    "dy": [0.2, 0.3, 0.26, 0.27, 0.25, 0.25, 0.12, 0.28, 0.18, 0.25, 0.30, 0.28, 0.16, 0.21, 0.33],
}

WORKLOADS = {
    # name: kind, n_agents, agent_mode, envs per launch, number of batches, description
    "c2": dict(kind="flight_easy", n=3, am=0, envs=4096, batches=64, grouped=True, lanes=1,
               desc="flight_easy 3a15t AM0TM0, 4096 batched envs per handle (BASELINE.json configs[1]); 64 independent handles "
                    "(rollout workers, 168 MB of state > L2) advance in ONE grouped launch (cs_flight_group_step)"),
    "c2s": dict(kind="flight_easy", n=3, am=0, envs=4096, batches=64,
                desc="flight_easy 3a15t AM0TM0, the same 64 handles x 4096 envs with one launch per handle (8 streams in a CUDA graph)"),
    "c3": dict(kind="flight_easy", n=5, am=2, envs=65536, batches=4,
               desc="flight_easy 5a15t AM2TM0, 65536 envs per launch (configs[2])"),
    "c4": dict(kind="flight", n=3, am=0, envs=16384, batches=1,
               desc="flight (probability map) 3a15t AM0TM0, 16384 envs per GPU (configs[3])"),
    "c2w": dict(kind="flight_easy", n=3, am=0, envs=1048576, batches=1,
                desc="flight_easy 3a15t AM0TM0, 1048576 envs in ONE launch (throughput asymptote of the c2 kernel)"),
    "c5": dict(kind="search", n=64, am=0, envs=16384, batches=1,
               desc="search_env 64 agents / 1000 targets / map 64, 16384 envs per launch (configs[4] per-launch slice)"),
}


def flight_args(kind, n, am, m=15, M=50, R=7, T=200):
    return types.SimpleNamespace(env=kind, map_size=M, target_num=m, target_mode=0, agent_mode=am, n_agents=n, view_range=R,
                                 time_limit=T, detect_prob=0.9, safe_dist=1, agent_velocity=1, force_dist=3,
                                 turn_limit=np.pi / 4, wrong_alarm_prob=0.1)


def search_args(n=64, m=1000, M=64, R=7):
    return types.SimpleNamespace(env="search", map_size=M, target_num=m, target_mode=0, target_dir="./targets/", agent_mode=0,
                                 n_agents=n, view_range=R)


def algorithmic_bytes(kind, n, m=15, M=50, R=7, touched_per_step=0.0):
    """SURVEY.md section 8d contract figures (fp32 state, u8 actions), bytes per env-step."""
    if kind == "flight_easy":
        return 57 * n + 20 * m + 34
    if kind == "flight":
        return 57 * n + 20 * m + 34 + 8.0 * touched_per_step
    S = 2 * R - 1
    return 4 * n * (S * S + 2) + 8 * M * M + 16 * n + 8 * n + 3 * M * M // 8 + 8 * n


# ----------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.remove(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ----------------------------------------------------------------------------------------------------------------
def _py_port_worker(job):
    """Reference-style loop on the Python oracle: reset, then per step get_obs, get_state, step (rollout.py:43-63)."""
    kind, n, am, seconds, wid = job
    from oracle.py_envs import FlightOracle, FlightSpec, SearchOracle, SearchSpec
    rng = np.random.default_rng(1000 + wid)
    steps = 0
    t_end = time.perf_counter() + seconds
    t0 = time.perf_counter()
    if kind in ("flight_easy", "flight"):
        spec = FlightSpec(n_agents=n, agent_mode=am, variant="easy" if kind == "flight_easy" else "probmap")
        env = FlightOracle(spec, TEMPLATE, 42, wid)
        while time.perf_counter() < t_end:
            env.reset(init=False)
            done = False
            while not done and time.perf_counter() < t_end:
                env.get_obs(); env.get_state()
                _, done, _ = env.step(rng.integers(0, 3, size=n))
                steps += 1
    else:
        spec = SearchSpec(n_agents=n, target_num=1000 if n == 64 else 15, map_size=64 if n == 64 else 50)
        env = SearchOracle(spec, 42, wid)
        while time.perf_counter() < t_end:
            env.reset()
            done, k = False, 0
            while not done and k < 500 and time.perf_counter() < t_end:
                env.get_obs(); env.get_state()
                acts = [int(rng.choice(np.nonzero(env.get_avail_agent_actions(i))[0])) for i in range(n)]
                _, done, _ = env.step(acts)
                steps += 1; k += 1
    return steps, time.perf_counter() - t0


def cpu_port_throughput(kind, n, am, seconds, procs):
    """env-steps/s of the Python oracle port with `procs` independent processes (the reference is single-threaded;
    independent env instances are its embarrassingly-parallel upper bound)."""
    import multiprocessing as mp
    jobs = [(kind, n, am, seconds, w) for w in range(procs)]
    if procs == 1:
        res = [_py_port_worker(jobs[0])]
    else:
        with mp.get_context("spawn").Pool(procs) as pool:
            res = pool.map(_py_port_worker, jobs)
    return sum(s / dt for s, dt in res), sum(s for s, _ in res)


def c_port_throughput(kind, n, am, seconds, threads):
    """env-steps/s of the C oracle (oracle/coopsearch_oracle.c) incl. obs/state emission, `threads` host threads."""
    from oracle import c_oracle
    from oracle.py_envs import FlightSpec, SearchSpec
    c_oracle.set_threads(threads)
    E = 4096 * max(1, threads // 2)
    steps, t0 = 0, time.perf_counter()
    if kind in ("flight_easy", "flight"):
        spec = FlightSpec(n_agents=n, agent_mode=am, variant="easy" if kind == "flight_easy" else "probmap")
        if kind == "flight":
            E = 64 * threads
        b = c_oracle.FlightBatch(spec, TEMPLATE, 42, 0, E, auto_reset=True)
        b.reset(init=True)
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            b.step(None); b.obs_state()
            steps += E
    else:
        spec = SearchSpec(n_agents=n, target_num=1000 if n == 64 else 15, map_size=64 if n == 64 else 50)
        E = 16 * threads
        b = c_oracle.SearchBatch(spec, 42, 0, E)
        b.reset()
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            b.step(None); b.views()
            steps += E
    dt = time.perf_counter() - t0
    c_oracle.set_threads(1)
    return steps / dt, steps


def cpu_baseline(kind, n, am, seconds_py, seconds_c):
    cores = os.cpu_count() or 1
    v_all, steps_all = cpu_port_throughput(kind, n, am, seconds_py, cores)
    v_one, _ = cpu_port_throughput(kind, n, am, min(seconds_py, 4.0), 1)
    try:
        v_c, _ = c_port_throughput(kind, n, am, seconds_c, cores)
        v_c1, _ = c_port_throughput(kind, n, am, min(seconds_c, 2.0), 1)
    except Exception as exc:  # the C oracle needs gcc/make on the box
        v_c, v_c1 = None, None
        print("C oracle unavailable: %s" % exc, file=sys.stderr)
    return {
        "value": v_all, "unit": "env-steps/s", "cores": cores, "kind": "port",
        "sample": "Python oracle port (oracle/py_envs.py, the reference's scalar float64 loop): %d independent processes x %.0f s of "
                  "reset/get_obs/get_state/step with uniform-random actions, %d env-steps in total" % (cores, seconds_py, steps_all),
        "single_core": v_one,
        "c_port": {"value": v_c, "single_core": v_c1, "cores": cores,
                   "note": "plain-C oracle (oracle/coopsearch_oracle.c), same arithmetic, batched; not what the reference runs"},
        "cpu_model": _cpu_model(),
    }


def _cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------
def make_envs(cs, w, device, rank, seed=42, count_touched=False):
    envs = []
    lpe = int(os.environ.get("CS_BENCH_LPE", str(w.get("lanes", 0))))      # 0 = the library's own choice
    for b in range(w["batches"]):
        base = (rank * w["batches"] + b) * w["envs"]
        if w["kind"] == "flight_easy":
            e = cs.VecFlightEasyEnv(flight_args("flight_easy", w["n"], w["am"]), TEMPLATE, num_envs=w["envs"], device=device,
                                    seed=seed, env_id_base=base, auto_reset=True, lanes_per_env=lpe)
        elif w["kind"] == "flight":
            e = cs.VecFlightEnv(flight_args("flight", w["n"], w["am"]), TEMPLATE, num_envs=w["envs"], device=device, seed=seed,
                                env_id_base=base, auto_reset=True, count_touched=count_touched, lanes_per_env=lpe)
        else:
            e = cs.VecSearchEnv(search_args(), num_envs=w["envs"], device=device, seed=seed, env_id_base=base, auto_reset=True)
        envs.append(e)
    return envs


def silence(fn, *a, **k):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def run_gpu_workload(cs, torch, name, steps, warmup, device, rank, world, want_e2e=True, burn_in_s=0.3):
    """Returns dict(ms_per_step, env_steps_per_step, e2e_value, ...) for one workload on this rank."""
    w = WORKLOADS[name]
    n = w["n"]
    A = 4 if w["kind"] == "search" else 3
    envs = silence(make_envs, cs, w, device, rank)
    POOL = 8
    gen = torch.Generator(device=device).manual_seed(1234 + rank)
    if w["kind"] == "search":
        actions = None          # legal moves depend on positions: the random legal policy is drawn in-kernel
    else:
        actions = [[torch.randint(0, A, (w["envs"], n), generator=gen, device=device, dtype=torch.uint8) for _ in envs]
                   for _ in range(POOL)]

    side = torch.cuda.Stream(device=device)
    nstreams = max(1, min(int(os.environ.get("CS_BENCH_STREAMS", "8")), len(envs)))
    forks = [torch.cuda.Stream(device=device) for _ in range(nstreams)] if nstreams > 1 else []
    grouped = None
    if w.get("grouped") and len(envs) > 1 and os.environ.get("CS_BENCH_GROUPED", "1") != "0":
        grouped = cs.DeviceStepper(envs)          # all batches advance in ONE launch (cs_flight_group_step)
        nstreams, forks = 1, []

    def one_step(k):
        """One env-step of every batch.  The batches are independent env sets (rollout workers): either one grouped
        launch steps them all, or their launches are forked over `nstreams` streams (captured as parallel branches of
        the CUDA graph) and joined again."""
        if grouped is not None:
            with torch.cuda.stream(side):
                grouped.step(actions[k % POOL])
            return
        if forks:
            for f in forks:
                f.wait_stream(side)
        for b, e in enumerate(envs):
            with torch.cuda.stream(forks[b % nstreams] if forks else side):
                if actions is None:
                    e.step_random(1)
                else:
                    e.step(actions[k % POOL][b])
        for f in forks:
            side.wait_stream(f)

    torch.cuda.synchronize(device)
    graphs = []
    lib = cs.load_library()
    with torch.cuda.stream(side):
        k0 = lib.cs_launch_count()
        one_step(0)
        kernels_per_step = int(lib.cs_launch_count() - k0)     # kernels of OUR library one step launches (graph replays repeat them)
        torch.cuda.synchronize(device)
        for k in range(POOL if actions is not None else 1):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                one_step(k)
            graphs.append(g)
    torch.cuda.synchronize(device)

    gmulti = None
    if grouped is not None and len(graphs) > 1:
        # one grouped launch is ~57 us of GPU work: a one-node graph per step leaves ~10 us of launch gap between them, so
        # the POOL steps are also captured back to back in one graph
        gmulti = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(gmulti, stream=side):
                for k in range(len(graphs)):
                    one_step(k)
        torch.cuda.synchronize(device)

    def replay(k):
        graphs[k % len(graphs)].replay()

    def replay_many(count, start=0):
        """`count` consecutive bench steps starting with action set `start`."""
        k = start
        if gmulti is not None:
            while k % len(graphs) != 0 and count > 0:
                replay(k); k += 1; count -= 1
            while count >= len(graphs):
                gmulti.replay(); k += len(graphs); count -= len(graphs)
        while count > 0:
            replay(k); k += 1; count -= 1

    # clock burn-in: same work, untimed, so that the sampler sees the GPU under THIS load and clocks have ramped
    t_end = time.perf_counter() + burn_in_s
    k = 0
    while time.perf_counter() < t_end:
        replay(k); k += 1
        if k % 64 == 0:
            torch.cuda.synchronize(device)
    for k in range(warmup):
        replay(k)
    torch.cuda.synchronize(device)
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize(device)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    replay_many(steps)
    ev1.record()
    torch.cuda.synchronize(device)
    if world > 1:
        torch.distributed.barrier()
    ms = ev0.elapsed_time(ev1)
    launches = steps * (1 if grouped is not None else len(envs))
    out = dict(grouped=grouped is not None, ms_total=ms, ms_per_step=ms / steps, env_steps_per_step=w["envs"] * len(envs), launches=launches, streams=nstreams,
               kernels=kernels_per_step * steps,
               us_per_launch=1000.0 * ms / launches, lanes_per_env=getattr(envs[0], "lanes_per_env", None))

    # ---- end to end through the host-buffer C-ABI call -----------------------------------------------------
    if want_e2e:
        e2e_steps = max(3, min(steps, 50))
        rng = np.random.default_rng(7)
        if w["kind"] == "search":
            host_actions = None
        else:
            host_actions = [[rng.integers(0, A, size=(w["envs"], n), dtype=np.uint8) for _ in envs] for _ in range(2)]
        for e in envs:
            e.host_buffers()

        streams = [torch.cuda.Stream(device=device) for _ in range(min(4, len(envs)))]
        steppers = None
        if w["kind"] != "search":
            # this step's actions are in PINNED host memory (two alternating sets); one library call steps every batch
            pinned = [[torch.from_numpy(a).pin_memory() for a in host_actions[k]] for k in range(2)]
            use_graph = os.environ.get("CS_BENCH_E2E_GRAPH", "1") != "0"
            steppers = [cs.HostStepper(envs, streams, actions=pinned[k], graph=use_graph) for k in range(2)]

        def host_step(k):
            if w["kind"] == "search":
                for b, e in enumerate(envs):
                    av = e.host_buffers()["avail"].numpy()
                    # first legal move of every agent, computed on the host from the previous step's D2H avail mask
                    acts = av.argmax(axis=2).astype(np.uint8)
                    e.step_host(acts)
                return
            # independent env batches are pipelined over a few streams: H2D / kernel / result traffic of different
            # batches overlap; the call returns when every batch's results are in host memory
            steppers[k % 2].step()

        if w["kind"] == "search":
            for e in envs:
                e.host_buffers()["avail"].copy_(e.get_avail_actions().cpu())
        for k in range(3):
            host_step(k)
        torch.cuda.synchronize(device)
        if world > 1:
            torch.distributed.barrier()
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            host_step(k)
        torch.cuda.synchronize(device)
        dt = time.perf_counter() - t0
        hb = envs[0].host_buffers()
        h2d = sum(hb["actions"].numel() for _ in envs)
        if "slab" in hb:
            d2h = sum(e.host_buffers()["slab"].numel() for e in envs)      # one D2H copy of the output slab per batch
        else:
            d2h = sum(sum(v.numel() * v.element_size() for kname, v in e.host_buffers().items() if kname != "actions") for e in envs)
        out.update(e2e_s_per_step=dt / e2e_steps, e2e_steps=e2e_steps, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h)
    # statistics of the finished episodes on this GPU (all-reduced by the caller)
    stats = torch.zeros(8, dtype=torch.float64, device=device)
    for e in envs:
        stats += torch.tensor(list(e.stats().values()), dtype=torch.float64, device=device)
    out["stats"] = stats
    return out


def measure_obs_full(cs, torch, device, peak):
    """Reference-shaped flight observation (flight_env.py:223-230): 4*M*M read + 4n(M*M+4) written per env."""
    w = dict(WORKLOADS["c4"])
    env = silence(make_envs, cs, w, device, 0)[0]
    env.step_random(3)
    out = {}
    for kind in ("tma", "plain"):
        env.set_obs_kernel(kind)
        for _ in range(5):
            env.get_obs()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            env.get_obs()
        e1.record()
        torch.cuda.synchronize(device)
        us = 1000.0 * e0.elapsed_time(e1) / 20
        M2 = env.map_size ** 2
        nbytes = env.num_envs * (4 * M2 + 4 * env.n_agents * (M2 + 4))
        out[kind] = {"us_per_launch": us, "achieved": nbytes / (us * 1e-6) / 1e9, "frac": nbytes / (us * 1e-6) / 1e9 / peak,
                     "bytes_per_env": nbytes // env.num_envs}
    return out


def measure_policy(cs, torch, device, E=4096, n=3, iters=50):
    """Batched agent network + action choice (csrc/policy.cu; SURVEY 8f rank 1, first version): rows/s of one
    choose_actions launch on random-init weights of the reference's architecture (in 10 -> 64 -> GRU 64 -> 64 -> 3)."""
    torch.manual_seed(0)
    in_dim = 4 + 3 + n
    sd = {"fc1.weight": torch.randn(64, in_dim) * 0.1, "fc1.bias": torch.zeros(64), "rnn.weight_ih": torch.randn(192, 64) * 0.1,
          "rnn.weight_hh": torch.randn(192, 64) * 0.1, "rnn.bias_ih": torch.zeros(192), "rnn.bias_hh": torch.zeros(192),
          "fc2.0.weight": torch.randn(64, 64) * 0.1, "fc2.0.bias": torch.zeros(64), "fc2.2.weight": torch.randn(3, 64) * 0.1,
          "fc2.2.bias": torch.zeros(3)}
    agents = cs.BatchedRNNAgents(sd, num_envs=E, n_agents=n, device=device)
    obs = torch.rand(E, n, 4, device=device) * 2 - 1
    for _ in range(5):
        agents.choose_actions(obs)
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        agents.choose_actions(obs)
    e1.record()
    torch.cuda.synchronize(device)
    us = 1000.0 * e0.elapsed_time(e1) / iters
    rows = E * n
    flop = 2 * (in_dim * 64 + 2 * 64 * 192 + 64 * 64 + 64 * 3)
    return {"workload": "agent network + greedy action choice for %d envs x %d agents (random-init weights, synthetic obs)" % (E, n),
            "rows": rows, "us_per_launch": us, "agent_steps_per_s": rows / (us * 1e-6), "gflops": rows * flop / (us * 1e-6) / 1e9,
            "note": "first version on CUDA cores (one warp per four rows); not part of the headline metric"}


def measure_touched(cs, torch, device, steps=200):
    """Exact mean number of probability-map cells updated per env-step on the c4 workload (separate, untimed pass)."""
    w = dict(WORKLOADS["c4"]); w["envs"] = 2048
    envs = silence(make_envs, cs, w, device, 0, count_touched=True)
    e = envs[0]
    base = e.stats()["map_cells_touched"]
    e.step_random(steps)
    torch.cuda.synchronize(device)
    return (e.stats()["map_cells_touched"] - base) / (steps * w["envs"])


def kernel_name(w, lanes, grouped=False):
    """The dominant kernel of a workload (csrc/flight.cu, csrc/search.cu)."""
    if w["kind"] == "search":
        return "search_kernel<STEP>"
    if grouped:
        return "flight_tpe_group_kernel<N=%d,K=%d>" % (w["n"], lanes)
    if w["kind"] == "flight":
        return ("flight_fused_kernel<N=%d,STEP> (step + belief map of an env in one launch, %d lanes per env)" % (w["n"], lanes)) if lanes == 8 else \
            "flight_kernel<LPE=%s,STEP> + flight_map_generic_kernel" % lanes
    return ("flight_tpe_kernel<N=%d,K=%d,STEP>" % (w["n"], lanes)) if lanes and lanes <= 4 else "flight_kernel<LPE=%s,STEP>" % lanes


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic(key):
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get(key)
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-extra", action="store_true", help="skip the supplementary workloads / CPU baseline")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        # The reference's own CPU implementation of the path: the Python oracle port on every host core.
        if rank != 0:
            return 0
        t0 = time.perf_counter()
        per_step_seconds = max(0.5, min(10.0, 60.0 / max(1, args.steps + args.warmup)))
        for _ in range(min(args.warmup, 3)):
            cpu_port_throughput(w["kind"], w["n"], w["am"], 0.3, os.cpu_count() or 1)
        vals, total_steps = [], 0
        for _ in range(min(args.steps, 20)):
            v, s = cpu_port_throughput(w["kind"], w["n"], w["am"], per_step_seconds, os.cpu_count() or 1)
            vals.append(v); total_steps += s
            if time.perf_counter() - t0 > 150:
                break
        v = float(np.mean(vals))
        line = {
            "impl": "reference", "metric": "env-steps/s", "value": v, "unit": "env-steps/s", "n_gpus": args.gpus,
            "steps": len(vals), "warmup": min(args.warmup, 3), "ms_per_step": 1000.0 * per_step_seconds, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["desc"], "note": "each step = %.1f s of the Python port of the reference env loop on all host cores" % per_step_seconds},
            "agent_steps_per_s": v * w["n"],
            "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": os.cpu_count() or 1, "kind": "port",
                             "sample": "%d env-steps of reset/get_obs/get_state/step under uniform-random actions" % total_steps,
                             "cpu_model": _cpu_model()},
            "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    import torch
    import coopsearch_b200 as cs
    from coopsearch_b200 import dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: the product has no CPU path")
    rank, local, world = dist.init_from_env("nccl" if world > 1 else None)
    device = torch.device("cuda", local if world > 1 else 0)
    torch.cuda.set_device(device)
    lib = cs.load_library()
    launches0 = lib.cs_launch_count()

    sampler = ClockSampler(device.index)
    if rank == 0:
        sampler.start()
    res = run_gpu_workload(cs, torch, args.workload, args.steps, args.warmup, device, rank, world)
    clocks = sampler.stop() if rank == 0 else None

    ms_max = dist.max_over_ranks(res["ms_total"], device=device)
    e2e_max = dist.max_over_ranks(res["e2e_s_per_step"], device=device)
    t0 = time.perf_counter()
    stats = dist.allreduce_stats(res["stats"])          # NCCL all-reduce of the episode statistics
    torch.cuda.synchronize(device)
    allreduce_us = 1e6 * (time.perf_counter() - t0)
    if rank != 0:
        if world > 1:
            torch.distributed.barrier()
            torch.distributed.destroy_process_group()
        return 0

    env_steps = res["env_steps_per_step"] * world * args.steps
    value = env_steps / (ms_max / 1000.0)
    e2e_value = res["env_steps_per_step"] * world / e2e_max
    peak, peak_src = load_peaks()
    touched = None
    if w["kind"] == "flight":
        touched = measure_touched(cs, torch, device)
    per_unit = algorithmic_bytes(w["kind"], w["n"], touched_per_step=touched or 0.0) if w["kind"] != "search" else \
        algorithmic_bytes("search", 64, m=1000, M=64, R=7)
    bytes_per_launch = per_unit * w["envs"] * (w["batches"] if res["grouped"] else 1)
    sec_per_launch = (res["ms_total"] / 1000.0) / res["launches"]
    achieved = bytes_per_launch / sec_per_launch / 1e9
    line = {
        "metric": "env-steps/s", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": w["desc"], "envs_per_handle": w["envs"], "handles_per_gpu": w["batches"],
            "envs_per_launch": w["envs"] * (w["batches"] if res["grouped"] else 1), "launches_per_step": 1 if res["grouped"] else w["batches"],
            "env_instances_per_gpu": w["envs"] * w["batches"], "auto_reset": True,
            "actions": "pre-generated uniform-random u8 tensors resident in HBM" if w["kind"] != "search" else "uniform-random legal policy drawn in-kernel (Philox)",
            "launch": ("CUDA-graph replay (8 steps per graph) of one grouped launch per step that steps every handle (cs_flight_group_step)" if res["grouped"] else
                       "CUDA-graph replay of the step launches; the independent batches are forked over %d streams inside the graph" % res["streams"]), "l2": "working set of all batches exceeds the 126 MB L2; batches are revisited round-robin, no flush",
            "lanes_per_env": res["lanes_per_env"], "parallelism": "dp%d (env instances sharded by global id, no data-path collective)" % world,
        },
        "agent_steps_per_s": value * w["n"],
        "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": res["h2d_bytes_per_step"],
                "d2h_bytes_per_step": res["d2h_bytes_per_step"], "agent_steps_per_s": e2e_value * w["n"],
                "steps": res["e2e_steps"],
                "path": "HostStepper.step() = cs_flight_step_host_many captured in one CUDA graph; per batch: pinned host actions -> H2D -> step kernel -> one D2H of reward/terminated/win/target_find/obs/state into the pinned slab; batches pipelined over 4 streams, all synchronised every step (search: cs_search_step_host per batch)"},
        "gpu_launches": int(res["kernels"]),
        "gpu_launches_process_total": int(lib.cs_launch_count() - launches0),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": load_traffic(args.workload), "peak_source": peak_src,
                     "kernel": kernel_name(w, res["lanes_per_env"], res["grouped"]),
                     "algorithmic_bytes_per_env_step": per_unit, "env_steps_per_launch": w["envs"] * (w["batches"] if res["grouped"] else 1),
                     "us_per_launch": 1e6 * sec_per_launch, "map_cells_touched_per_env_step": touched,
                     "note": "duration = timed region / launches; launches of independent batches overlap when streams > 1"},
        "clocks": clocks,
        "episode_stats_allreduced": dict(zip(cs._lib.STAT_NAMES, [float(x) for x in stats.tolist()])),
        "stats_allreduce_us": allreduce_us,
    }
    if world == 1 and not args.no_extra:
        extra = {}
        for name in ("c2s", "c2w", "c3", "c4", "c5"):
            if name == args.workload:
                continue
            try:
                k = 60 if name != "c5" else 30
                r = run_gpu_workload(cs, torch, name, k, 5, device, 0, 1, want_e2e=True, burn_in_s=0.1)
                ww = WORKLOADS[name]
                tch = measure_touched(cs, torch, device) if ww["kind"] == "flight" else None
                pu = algorithmic_bytes(ww["kind"], ww["n"], touched_per_step=tch or 0.0) if ww["kind"] != "search" else \
                    algorithmic_bytes("search", 64, m=1000, M=64, R=7)
                v = r["env_steps_per_step"] / (r["ms_per_step"] / 1000.0)
                gbs = pu * ww["envs"] * (ww["batches"] if r["grouped"] else 1) / (1e-6 * r["us_per_launch"]) / 1e9
                extra[name] = {"workload": ww["desc"], "kernel": kernel_name(ww, r["lanes_per_env"], r["grouped"]), "value": v, "unit": "env-steps/s", "agent_steps_per_s": v * ww["n"],
                               "us_per_launch": r["us_per_launch"], "e2e_value": r["env_steps_per_step"] / r["e2e_s_per_step"],
                               "roofline": {"achieved": gbs, "peak": peak, "frac": gbs / peak, "unit": "GB/s",
                                            "algorithmic_bytes_per_env_step": pu, "map_cells_touched_per_env_step": tch,
                                            "traffic": load_traffic(name)}}
            except Exception as exc:  # supplementary only: never lose the headline line
                extra[name] = {"error": repr(exc)}
        try:
            extra["c4_obs"] = {"workload": "flight get_obs(): prob_map || features for 16384 envs x 3 agents (656 MB per call)",
                               "unit": "GB/s", "peak": peak, **measure_obs_full(cs, torch, device, peak)}
        except Exception as exc:
            extra["c4_obs"] = {"error": repr(exc)}
        try:
            extra["policy"] = measure_policy(cs, torch, device)
        except Exception as exc:
            extra["policy"] = {"error": repr(exc)}
        line["extra"] = extra
        line["cpu_baseline"] = cpu_baseline(w["kind"], w["n"], w["am"], args.cpu_seconds, 4.0)
    print(json.dumps(line))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
