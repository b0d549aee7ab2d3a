#!/usr/bin/env python
"""bench.py -- headline benchmark of the K-CBS low-level hot path on B200.

  python bench.py --gpus N --steps K --warmup W            our arm (sm_100a kernels via the C ABI)
  python bench.py --impl reference --gpus N --steps K ...   the reference's CPU path (oracle restatement,
                                                            all host threads) on the same workload

Workload (BASELINE.json configs[1]): Empty32x32, 10 robots, 2nd-order cars, batched propagation with
65,536 controls per expansion.  One STEP = every robot's low-level planner issuing `--expansions`
tree expansions (one C-ABI call = one kernel launch = one 64K-control expansion from one tree node),
each robot on its own CUDA stream.  metric = propagated states/s; one unit = one 0.1 s propagation
step of one control sample (10 RK4 sub-steps + theta wrap + validity).

Printed JSON (one line, rank 0): the driver contract keys plus `roofline` (dominant kernel:
k_propagate, FP32-pipe bound), `cpu_baseline`, `e2e` (host buffers through the reference-facing
host entry point, H2D + D2H inside the timed region), `clocks`, `gpu_launches`, and `secondary`
(validity and conflict-scan throughput with their HBM rooflines).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

SCENE = "Empty32x32_10robots"
N_CONTROLS = 65536
MAX_STEPS = 10
FLOPS_PER_STATE = 980.0      # 860 arithmetic + 120 transcendental FP32 ops per propagated state
BYTES_PER_STATE = 20.0       # 5 f32 written per propagated state
BYTES_PER_SAMPLE = 16.0      # 2 f32 controls + 1 i32 steps read + 1 i32 valid count written per sample
CONFLICT_BYTES_PER_TRAJ_STATE = 12.0   # x, y, theta f32 read once per plan
VALIDITY_BYTES_PER_STATE = 20.125      # 5 f32 read + 1 bit written
FP32_NOMINAL_TFLOPS = 74.4   # 148 SM x 128 lanes x 2 x 1.965 GHz


def workload_config(n_robots=10, n_exp=25, fixed_steps=None):
    """`config` of the JSON line: the same keys and values in both arms (ours / --impl reference), so the driver can
    tell that they ran the same workload; arm-specific facts live in `arm`."""
    return {"workload": "Empty32x32 10 robots, batched propagation 64K controls/expansion",
            "scene": SCENE, "robots": n_robots, "controls_per_expansion": N_CONTROLS, "max_steps": MAX_STEPS,
            "durations": "U{1..10}" if fixed_steps is None else f"fixed {fixed_steps}",
            "parents": "one shared tree node per expansion",
            "unit_definition": "one 0.1 s propagation step of one control sample (10 RK4 sub-steps + wrap + validity)",
            "l2": "GPU arm: every step writes distinct segment buffers of > 126 MB per GPU (no flush needed); CPU arm: n/a",
            "parallelism": "independent per-robot expansions (replicas per GPU), no data-path collective"}


def read_traffic():
    """DRAM traffic per launch from the committed ncu captures (profiles/traffic_r1.json); None when absent."""
    path = os.path.join(ROOT, "profiles", "traffic_r2.json")
    if not os.path.exists(path):
        path = os.path.join(ROOT, "profiles", "traffic_r1.json")
    if not os.path.exists(path):
        return {}
    with open(path) as f:
        return json.load(f)


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, torch_device_index: int, period_s: float = 0.005):
        self.samples, self.reasons, self.power = [], set(), []
        self.period = period_s
        self._stop = threading.Event()
        self._t = None
        self.max_mhz = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            h = None
            try:
                uuid = str(torch.cuda.get_device_properties(torch_device_index).uuid)
                if not uuid.startswith("GPU-"):
                    uuid = "GPU-" + uuid
                h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                idx = int(vis.split(",")[torch_device_index]) if vis and vis.split(",")[0].isdigit() else torch_device_index
                h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.h = h
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # NVML unavailable: clocks are reported as unknown
            self.nv = None
            self.err = repr(e)

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for b, name in self.REASONS.items():
                    if bits & b:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        if self._t is not None:
            self._stop.set()
            self._t.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0,
                    "note": getattr(self, "err", "no samples")}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples), "power_w_max": max(self.power) if self.power else None}


# ---------------------------------------------------------------------------------------------
def make_expansions(workloads, rank: int, n_robots: int, n_exp: int, fixed_steps=None):
    """Inputs of one step for this rank: per robot `n_exp` expansions, each from its own tree node
    (one shared parent) with N_CONTROLS sampled controls and durations."""
    per_robot = []
    for r in range(n_robots):
        exps = []
        for e in range(n_exp):
            seed = 1 + 1000003 * rank + 1009 * r + e
            parents, controls, steps = workloads.expansion_batch(SCENE, N_CONTROLS, seed=seed, shared_parent=True,
                                                                 interior=1.5, fixed_steps=fixed_steps)
            exps.append((parents, controls, steps))
        per_robot.append(exps)
    return per_robot


def cpu_baseline_leg(args, pkg, orc, workloads, budget_s: float, n_threads: int):
    """Oracle (port of the reference's CPU path) on a bounded sample of the same workload."""
    world = pkg.scenes.world(SCENE)
    oracle = orc.Oracle(world)
    units, t_used, n_exp = 0, 0.0, 0
    t_end = time.perf_counter() + budget_s
    while True:
        parents, controls, steps = workloads.expansion_batch(SCENE, N_CONTROLS, seed=77 + n_exp, shared_parent=True,
                                                             interior=1.5)
        t0 = time.perf_counter()
        _, _, _, calls = oracle.propagate_batch(0, parents, controls, steps, want_seg=True, n_threads=n_threads)
        t_used += time.perf_counter() - t0
        units += calls
        n_exp += 1
        if time.perf_counter() >= t_end or n_exp >= 4096:
            break
    return {"value": units / t_used, "unit": "states/s", "cores": n_threads, "kind": "port",
            "sample": f"{n_exp} expansions x {N_CONTROLS} controls ({units} propagated states, {t_used:.1f} s)"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The reference cannot be
    built here (OMPL K-CBS fork + Boost absent), so this is the oracle port with every host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg = graft.load_package()
    orc = graft.load_oracle()
    orc.build()
    from kcbs_b200 import workloads
    world = pkg.scenes.world(SCENE)
    oracle = orc.Oracle(world)
    n_threads = os.cpu_count() or 1
    sample = 4  # expansions of 64K controls per step (bounded sample of the 10 x --expansions GPU step)

    def one_step(seed0):
        units, t = 0, 0.0
        for e in range(sample):
            parents, controls, steps = workloads.expansion_batch(SCENE, N_CONTROLS, seed=seed0 + e, shared_parent=True,
                                                                 interior=1.5)
            t0 = time.perf_counter()
            _, _, _, calls = oracle.propagate_batch(0, parents, controls, steps, want_seg=True, n_threads=n_threads)
            t += time.perf_counter() - t0
            units += calls
        return units, t

    for w in range(args.warmup):
        one_step(9000 + 10 * w)
    units, t = 0, 0.0
    for k in range(args.steps):
        u, dt = one_step(100 + 10 * k)
        units += u
        t += dt
    value = units / t
    # the reference itself is single-threaded end to end: the same port on one thread, one expansion
    parents, controls, steps = workloads.expansion_batch(SCENE, N_CONTROLS, seed=4242, shared_parent=True, interior=1.5)
    t0 = time.perf_counter()
    _, _, _, calls1 = oracle.propagate_batch(0, parents, controls, steps, want_seg=True, n_threads=1)
    single = calls1 / (time.perf_counter() - t0)
    line = {
        "impl": "reference", "metric": "propagated states/s", "value": value, "unit": "states/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(),
        "arm": {"step_sample": f"{sample} expansions per step (bounded sample of the GPU arm's step)"},
        "cpu_baseline": {"value": value, "unit": "states/s", "cores": n_threads, "kind": "port",
                         "sample": f"{sample} expansions x {N_CONTROLS} controls per step, {args.steps} steps",
                         "single_thread": {"value": single, "unit": "states/s", "sample": f"1 expansion x {N_CONTROLS} controls"}},
        "e2e": {"value": value, "unit": "states/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference = CPU oracle port (OMPL K-CBS fork + Boost are not installable here); the reference is "
                "single-threaded, this arm uses every host thread",
    }
    emit(line)


# ---------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly one JSON line.  Libraries print banners there (NCCL's version line, for one), so file
    descriptor 1 is pointed at stderr for the whole run and the line goes to a private duplicate of the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--expansions", type=int, default=25, help="expansions per robot per step")
    ap.add_argument("--fixed-steps", type=int, default=None, help="use a fixed duration instead of U{1..10}")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU baseline work")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue the step's launches from Python instead of replaying a CUDA graph")
    ap.add_argument("--layout", default="packed", choices=["packed", "dense"],
                    help="segment layout of the headline: packed = valid states only (CSR), dense = [5][10][n] zero padded")
    ap.add_argument("--expansion-mode", type=int, default=2, choices=[0, 1, 2],
                    help="kcbs_set_expansion_mode of the headline contexts (2 = throughput shape: ten planners launch concurrently)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to stdout at NCCL_DEBUG=VERSION; stdout carries the one JSON line only
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world_size == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world_size == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    pkg = graft.load_package()
    from kcbs_b200 import workloads
    world = pkg.scenes.world(SCENE)
    n_robots = len(world.robot_dims)
    n_exp = args.expansions
    hbm_peak, peak_src = read_peaks()

    # one context (= one stream) per robot planner
    ctxs = [pkg.Context(world, device=local_rank) for _ in range(n_robots)]
    streams = [torch.cuda.Stream() for _ in range(n_robots)]
    main_stream = torch.cuda.current_stream()

    # ---- inputs resident in HBM ---------------------------------------------------------------
    host_in = make_expansions(workloads, rank, n_robots, n_exp, args.fixed_steps)
    dev_in = [[(torch.from_numpy(p).cuda(), torch.from_numpy(c).cuda(), torch.from_numpy(s).cuda()) for p, c, s in exps]
              for exps in host_in]
    # outputs: every expansion of a step has its own segment buffer (13.1 MB each), so one step writes
    # n_robots * n_exp * 13.1 MB (3.3 GB by default) >> 126 MB L2: no L2 flush needed between steps
    # (packed layout: the same capacity per expansion, of which only the valid rows -- about 45 % -- are written)
    packed_layout = args.layout == "packed"
    cap = MAX_STEPS * N_CONTROLS
    seg = torch.empty((n_robots, n_exp, 5, MAX_STEPS, N_CONTROLS), dtype=torch.float32, device="cuda")
    fin = torch.empty((n_robots, n_exp, 5, N_CONTROLS), dtype=torch.float32, device="cuda")
    valid = torch.zeros((n_robots, n_exp, N_CONTROLS), dtype=torch.int32, device="cuda")
    offsets = torch.zeros((n_robots, n_exp, N_CONTROLS + 1), dtype=torch.int32, device="cuda")
    for c in ctxs:
        c.set_expansion_mode(args.expansion_mode)

    def gpu_step(packed_layout=packed_layout):
        """One step: every robot planner issues its expansions on its own stream (fork from / join into the
        current stream, so the whole step can be captured into one CUDA graph)."""
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(cur)
        for r in range(n_robots):
            streams[r].wait_event(ev)
            sp = streams[r].cuda_stream
            for e in range(n_exp):
                p, c, s = dev_in[r][e]
                if packed_layout:
                    ctxs[r].propagate_packed_dev(r, N_CONTROLS, 1, p, c, s, None, MAX_STEPS, seg[r, e], cap, offsets[r, e],
                                                 fin[r, e], valid[r, e], sp)
                else:
                    ctxs[r].propagate_batch_dev(r, N_CONTROLS, 1, p, c, s, None, MAX_STEPS, seg[r, e], fin[r, e], valid[r, e], sp)
        for r in range(n_robots):
            done = torch.cuda.Event()
            done.record(streams[r])
            cur.wait_event(done)

    fp32_peak_before = ctxs[0].measure_fp32_peak()

    for _ in range(max(args.warmup, 1)):
        gpu_step()
    barrier()
    # units of one step: propagate() calls the reference would have made = min(steps, valid + 1)
    steps_all = torch.stack([torch.stack([s for _, _, s in exps]) for exps in dev_in])
    units_step = float(torch.minimum(steps_all, valid + 1).sum().item())
    samples_step = n_robots * n_exp * N_CONTROLS
    written_step = float(valid.sum().item())
    # steps the kernel really integrates: v and phi are linear in time, so the step at which they leave their bounds
    # is decided analytically and never integrated (propagate.cu); the other failures (x, y) cost their integration
    par_all = torch.stack([torch.stack([p for p, _, _ in exps]) for exps in dev_in])     # [r, e, 5, 1]
    ctl_all = torch.stack([torch.stack([c for _, c, _ in exps]) for exps in dev_in])     # [r, e, 2, n]
    t1 = (torch.arange(1, MAX_STEPS + 1, device="cuda", dtype=torch.float32) * np.float32(0.1))[None, None, :, None]
    vn = par_all[:, :, 2:3, :] + ctl_all[:, :, 0:1, :] * t1
    pn = par_all[:, :, 3:4, :] + ctl_all[:, :, 1:2, :] * t1
    inside = ((vn.abs() <= 1.0) & (pn.abs() <= float(np.pi / 3))).to(torch.int32)
    eff = torch.minimum(torch.cumprod(inside, dim=2).sum(dim=2), steps_all)
    integrated_step = float(torch.minimum(eff, valid + 1).sum().item())
    del par_all, ctl_all, vn, pn, inside, eff
    launches_per_step = n_robots * n_exp

    # The step is launch-bound from Python (250 C-ABI calls of ~20 us of GPU work each), so it is captured once
    # into a CUDA graph (the same kernels, streams and dependencies) and replayed.
    run_step = gpu_step
    graphed = False
    if not args.no_graph:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            gpu_step()
        run_step = graph.replay
        graphed = True
        for _ in range(max(args.warmup, 1)):
            run_step()
        barrier()
        assert float(torch.minimum(steps_all, valid + 1).sum().item()) == units_step

    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main_stream)
    for _ in range(args.steps):
        run_step()
    e1.record(main_stream)
    barrier()
    clocks = sampler.stop()
    launches = launches_per_step * args.steps
    elapsed_ms = max_over_ranks(e0.elapsed_time(e1))
    total_units = sum_over_ranks(units_step) * args.steps
    value = total_units / (elapsed_ms * 1e-3)
    # the other segment layout through the same loop (a few steps), for the record: dense = zero-padded [5][10][n] segments
    # (2.2 x the bytes, no row bookkeeping), packed = valid rows only
    other = {}
    if not args.no_graph:
        other_packed = not packed_layout
        og = torch.cuda.CUDAGraph()
        with torch.cuda.graph(og):
            gpu_step(other_packed)
        for _ in range(3):
            og.replay()
        barrier()
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o_steps = max(3, min(args.steps, 5))
        o0.record(main_stream)
        for _ in range(o_steps):
            og.replay()
        o1.record(main_stream)
        barrier()
        o_ms = max_over_ranks(o0.elapsed_time(o1))
        other = {"layout": "packed" if other_packed else "dense", "steps": o_steps,
                 "value": sum_over_ranks(units_step) * o_steps / (o_ms * 1e-3), "ms_per_step": o_ms / o_steps}
        del og
    fp32_peak_after = ctxs[0].measure_fp32_peak()
    fp32_peak = max(fp32_peak_before, fp32_peak_after)

    # kernel-level roofline (this rank): all launches of the timed region are k_propagate
    kern_ms = e0.elapsed_time(e1) / max(launches, 1)      # average launch duration (streams overlap: busy time / launches)
    ach_tflops = FLOPS_PER_STATE * units_step * args.steps / (e0.elapsed_time(e1) * 1e-3) * 1e-12
    ach_gbs = (BYTES_PER_STATE * written_step + BYTES_PER_SAMPLE * samples_step) * args.steps / (e0.elapsed_time(e1) * 1e-3) * 1e-9
    traffic = read_traffic()
    tp = traffic.get("k_propagate")
    # ncu capture is one launch of 655,360 samples; a bench launch is one expansion of 65,536 samples of the same mix
    if packed_layout:
        tp = traffic.get("k_propagate_packed")
        k1_traffic = None if not tp else (tp["dram_read_bytes"] + max(tp["dram_write_bytes"], tp.get("output_bytes", 0))) * N_CONTROLS / tp["samples"]
    else:
        k1_traffic = None if not tp else (tp["dram_read_bytes"] + tp["segment_buffer_bytes"]) * N_CONTROLS / tp["samples"]
    roofline = {
        "kernel": "k_propagate", "bound": "fp32", "achieved": ach_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
        "frac": ach_tflops / fp32_peak if fp32_peak else None,
        "units": units_step, "units_integrated": integrated_step, "units_written": written_step,
        "frac_integrated": ach_tflops * integrated_step / units_step / fp32_peak if fp32_peak else None,
        "frac_written": ach_tflops * written_step / units_step / fp32_peak if fp32_peak else None,
        "peak_detail": {"probe": fp32_peak, "nominal": FP32_NOMINAL_TFLOPS,
                        "note": "MEASURED_PEAKS.json has no FP32 entry: peak = FFMA saturation probe run in this process"},
        "traffic": k1_traffic,
        "layout": args.layout,
        "traffic_unit": "bytes per launch: ncu dram__bytes_read of one 655,360-sample launch + the bytes that launch writes (ncu's write "
                        "counter stops while most of them are still dirty in L2), scaled to 65,536 samples; profiles/traffic_r2.json",
        "algorithmic_bytes_per_launch": (BYTES_PER_STATE * written_step + BYTES_PER_SAMPLE * samples_step) / launches_per_step,
        # the call also returns the final state (20 B) and, packed, the row offset (4 B) of every sample
        "algorithmic_bytes_with_per_sample_outputs": (BYTES_PER_STATE * written_step + (BYTES_PER_SAMPLE + 24) * samples_step) / launches_per_step,
        "peak_source": "FFMA saturation probe run in this process (kcbs_measure_fp32_peak); nominal 74.4",
        "frac_of_nominal": ach_tflops / FP32_NOMINAL_TFLOPS,
        "algorithmic_flops_per_state": FLOPS_PER_STATE, "avg_launch_ms": kern_ms,
        "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak, "peak_source": peak_src},
    }
    if other:
        other["frac"] = FLOPS_PER_STATE * units_step / (other["ms_per_step"] * 1e-3) * 1e-12 / fp32_peak if fp32_peak else None
        roofline["other_layout"] = other
    if k1_traffic:
        roofline["traffic_over_algorithmic"] = k1_traffic / roofline["algorithmic_bytes_per_launch"]
        roofline["traffic_over_algorithmic_with_per_sample_outputs"] = k1_traffic / roofline["algorithmic_bytes_with_per_sample_outputs"]

    # ---- e2e: host buffers through the reference-facing host entry point --------------------------
    e2e_exp = min(n_exp, 4)
    ctx0 = ctxs[0]
    pin = []
    for r in range(n_robots):
        row = []
        for e in range(e2e_exp):
            p, c, s = host_in[r][e]
            pp, pc, ps = ctx0.pinned_empty(p.shape, np.float32), ctx0.pinned_empty(c.shape, np.float32), ctx0.pinned_empty(s.shape, np.int32)
            pp[...], pc[...], ps[...] = p, c, s
            row.append((pp, pc, ps))
        pin.append(row)
    # one pinned result buffer set per robot planner: the planners are independent host threads, each driving its
    # own context through the synchronous host entry point (ctypes releases the GIL during the call); inside a call
    # the batch is cut into chunks whose upload / kernel / download overlap (api_propagate.cu)
    from concurrent.futures import ThreadPoolExecutor
    outs = [(ctxs[r].pinned_empty((5, MAX_STEPS, N_CONTROLS), np.float32), ctxs[r].pinned_empty((5, N_CONTROLS), np.float32),
             ctxs[r].pinned_empty((N_CONTROLS,), np.int32), ctxs[r].pinned_empty((N_CONTROLS + 1,), np.int32)) for r in range(n_robots)]
    pool = ThreadPoolExecutor(max_workers=n_robots)
    rows_seen = [[0] * e2e_exp for _ in range(n_robots)]

    def robot_calls(r, kind):
        torch.cuda.set_device(local_rank)
        seg_b, fin_b, val_b, off_b = outs[r]
        for e in range(e2e_exp):
            pp, pc, ps = pin[r][e]
            if kind == "packed":      # whole segments, valid states only: rows + offsets are the complete result
                res = ctxs[r].propagate_packed(r, pp, pc, ps, max_steps=MAX_STEPS, out=(seg_b.reshape(5, -1), off_b, None, val_b))
                rows_seen[r][e] = res[4]
            else:
                ctxs[r].propagate_batch(r, pp, pc, ps, max_steps=MAX_STEPS, out=(seg_b if kind == "dense" else None, fin_b, val_b))

    def e2e_run(kind):
        # every robot planner thread makes its calls of ALL timed steps one after the other: the planners are independent
        # (no barrier between steps -- they only meet at the start and at the end of the timed region)
        run = lambda k: list(pool.map(lambda r: [robot_calls(r, kind) for _ in range(k)], range(n_robots)))  # noqa: E731
        run(1)
        barrier()
        t0 = time.perf_counter()
        run(e2e_steps)
        barrier()
        return max_over_ranks(time.perf_counter() - t0)

    e2e_steps = max(2, min(args.steps, 10))
    calls = n_robots * e2e_exp
    e2e_units = float(torch.minimum(steps_all[:, :e2e_exp], valid[:, :e2e_exp] + 1).sum().item())
    h2d = calls * (5 * 4 + N_CONTROLS * 12)
    # three timed regions of e2e_steps steps each, the median counts (a region lasts ~60 ms: one descheduled planner thread
    # on a shared host shows; all three are in the line)
    e2e_regions = sorted(e2e_run("packed") for _ in range(3))
    e2e_s = e2e_regions[1]
    rows = float(sum(sum(x) for x in rows_seen))
    rows_valid = float(valid[:, :e2e_exp].sum().item())
    assert rows_valid <= rows <= 1.25 * rows_valid + 1024 * calls, "the host call returned other rows than the device call"
    d2h_packed = int(20 * rows + 8 * (N_CONTROLS + 4) * calls)   # rows of five planes + offsets and valid counts (4 B + 4 B per sample)
    e2e = {"value": sum_over_ranks(e2e_units) * e2e_steps / e2e_s, "unit": "states/s",
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_packed,
           "calls_per_step": calls, "steps": e2e_steps, "regions_s": [round(t, 5) for t in e2e_regions],
           "layout": "packed (valid states only) + row offsets",
           "pcie_d2h_gbs": d2h_packed * e2e_steps / e2e_s * 1e-9,
           "note": "kcbs_propagate_packed with pinned host buffers, one host thread per robot planner, every call pipelined in "
                   "chunks; whole segments = the rows of every sample (20 B each, 1.05 per valid state) + 8 B of row offset and valid "
                   "count per sample copied back", "rows_per_valid_state": rows / rows_valid}
    # the zero-padded dense layout through kcbs_propagate_batch (224 B per sample), and the same calls asking only for
    # what propagateWhileValid(state, control, steps, result) returns (final state + step count, 24 B per sample)
    d2h_dense = calls * (N_CONTROLS * (5 * MAX_STEPS * 4 + 5 * 4 + 4))
    dense_s = e2e_run("dense")
    e2e["dense_layout"] = {"value": sum_over_ranks(e2e_units) * e2e_steps / dense_s, "unit": "states/s",
                           "d2h_bytes_per_step": d2h_dense, "pcie_d2h_gbs": d2h_dense * e2e_steps / dense_s * 1e-9}
    final_s = e2e_run("final")
    e2e["final_state_only"] = {"value": sum_over_ranks(e2e_units) * e2e_steps / final_s, "unit": "states/s",
                               "d2h_bytes_per_step": calls * N_CONTROLS * 24}
    # the ceiling of the e2e leg: plain pinned copies of 256 MB, all ranks at once (at N > 1 the ranks share the host's memory
    # system and root complex: the box is one NUMA node, there is nothing to pin to)
    probe_dev = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
    probe_host = torch.empty(64 << 20, dtype=torch.float32).pin_memory()
    probe = {}
    for name, dst, src in (("d2h_gbs", probe_host, probe_dev), ("h2d_gbs", probe_dev, probe_host)):
        dst.copy_(src, non_blocking=True)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(4):
            dst.copy_(src, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        probe[name] = 4 * probe_dev.numel() * 4 / (max_over_ranks(c0.elapsed_time(c1)) * 1e-3) * 1e-9
    del probe_dev, probe_host
    e2e["pcie_probe"] = dict(probe, note="per GPU, 256 MB pinned cudaMemcpyAsync, all ranks concurrently")
    e2e["d2h_fraction_of_probe"] = e2e["pcie_d2h_gbs"] / probe["d2h_gbs"]
    # one planner alone (one host thread, one call at a time): what the chunk pipeline inside a call buys
    t0 = time.perf_counter()
    for _ in range(4):
        robot_calls(0, "packed")
    one_s = time.perf_counter() - t0
    e2e["single_caller"] = {"value": float(torch.minimum(steps_all[0, :e2e_exp], valid[0, :e2e_exp] + 1).sum().item()) * 4 / one_s,
                            "unit": "states/s", "note": "one host thread, sequential kcbs_propagate_packed calls"}

    # ---- secondary metrics: validity (K2) and conflict scan (K3), device resident ------------------
    secondary = {}
    if not args.no_secondary:
        secondary = run_secondary(pkg, workloads, local_rank, hbm_peak, peak_src, max_over_ranks, sum_over_ranks, barrier, rank, traffic)

    cpu = None
    if rank == 0 and world_size == 1 and not args.no_cpu:
        orc = graft.load_oracle()
        orc.build()
        cpu = cpu_baseline_leg(args, pkg, orc, workloads, args.cpu_budget, os.cpu_count() or 1)
        # the reference itself is single-threaded end to end: the same port on one thread, shorter sample
        one = cpu_baseline_leg(args, pkg, orc, workloads, max(2.0, args.cpu_budget / 4), 1)
        cpu["single_thread"] = {"value": one["value"], "unit": one["unit"], "sample": one["sample"]}

    if rank == 0:
        # what the partitioned paths do at this N, first in the line (the driver keeps the head of it)
        cs = secondary.get("conflict_strong", {})
        multi = {
            "conflict_strong_P256_pairs_per_s": cs.get("P256", {}).get("value"),
            "conflict_strong_P256_speedup_vs_one_gpu": cs.get("P256", {}).get("speedup_vs_one_gpu"),
            "conflict_strong_P2048_pairs_per_s": cs.get("P2048", {}).get("value"),
            "conflict_strong_P2048_speedup_vs_one_gpu": cs.get("P2048", {}).get("speedup_vs_one_gpu"),
            "replans20_s": secondary.get("replans", {}).get("median_s"),
            "e2e_states_per_s": e2e["value"],
            "kcbs_solve_30robots_median_s": secondary.get("kcbs_solve", {}).get("scenes", {}).get("Benchmark_Empty32x32_30robots", {}).get("median_s"),
        }
        line = {
            "metric": "propagated states/s", "value": value, "unit": "states/s", "n_gpus": world_size,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "multi_gpu": multi,
            "config": workload_config(n_robots, n_exp, args.fixed_steps),
            "arm": {"expansions_per_robot_per_step": n_exp, "streams": n_robots, "segment_layout": args.layout,
                    "expansion_mode": args.expansion_mode,
                    "launch": "one CUDA graph per step (captured C-ABI launches)" if graphed else "eager C-ABI calls from Python",
                    "states_per_step_per_gpu": units_step,
                    "segment_bytes_per_step_per_gpu": (20 * written_step + 4 * samples_step) if packed_layout else samples_step * 200},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "gpu_launches": int(launches),
            "secondary": secondary,
        }
        emit(line)
    pool.shutdown()
    for c in ctxs:
        c.close()
    if world_size > 1:
        dist.destroy_process_group()


def run_secondary(pkg, workloads, dev, hbm_peak, peak_src, max_over_ranks, sum_over_ranks, barrier, rank, traffic):
    import torch
    out = {}
    # ---- validity: Congested10x10, 2^24 states; Empty32x32 (no obstacles) beside it ----------
    n = 1 << 24
    for key, scene in (("validity", "Congested10x10_7robots"), ("validity_no_obstacles", "Empty32x32_10robots")):
        with pkg.Context(pkg.scenes.world(scene), device=dev) as ctx:
            sets = [torch.from_numpy(workloads.validity_states(scene, n, seed=1 + s + 100 * rank)).cuda() for s in range(2)]
            mask = torch.empty((n + 31) // 32, dtype=torch.int32, device="cuda")
            st = torch.cuda.current_stream()
            for i in range(3):
                ctx.validity_batch_dev(0, n, sets[i % 2], mask, st)
            barrier()
            reps = 10
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(reps):
                ctx.validity_batch_dev(0, n, sets[i % 2], mask, st)   # 2 x 336 MB inputs alternate: > L2
            e1.record()
            barrier()
            ms = max_over_ranks(e0.elapsed_time(e1))
            rate = sum_over_ranks(float(n)) * reps / (ms * 1e-3)
            gbs = VALIDITY_BYTES_PER_STATE * n * reps / (e0.elapsed_time(e1) * 1e-3) * 1e-9
            has_traffic = key == "validity" and "k_validity" in traffic
            out[key] = {"metric": "validity-checked states/s", "value": rate, "scene": scene, "states": n,
                        "roofline": {"kernel": "k_validity", "bound": "hbm", "achieved": gbs, "peak": hbm_peak,
                                     "unit": "GB/s", "frac": gbs / hbm_peak, "peak_source": peak_src,
                                     "traffic": traffic["k_validity"]["dram_read_bytes"] + traffic["k_validity"]["dram_write_bytes"]
                                     if has_traffic else None,
                                     "algorithmic_bytes_per_launch": VALIDITY_BYTES_PER_STATE * n}}
            del sets, mask
    # ---- the expansion kernel on the obstacle scenes (configs 1 and 3 of BASELINE.json: Corridor10x10, Congested10x10):
    # the headline workload has no obstacles; here every step also looks its cell up and runs the exact rectangle test in
    # "maybe" cells.  655,360 samples per launch (own parents), packed segments, units = min(steps, valid + 1).
    prop_obs = {}
    n_s = 655360
    for scene in ("Corridor10x10_4robots", "Congested10x10_7robots"):
        with pkg.Context(pkg.scenes.world(scene), device=dev) as ctx:
            p, c, s = workloads.expansion_batch(scene, n_s, seed=3 + rank, interior=1.5)
            dp, dc, ds = torch.from_numpy(p).cuda(), torch.from_numpy(c).cuda(), torch.from_numpy(s).cuda()
            seg = torch.empty((5, 10 * n_s), device="cuda")
            fin = torch.empty((5, n_s), device="cuda")
            val = torch.zeros(n_s, dtype=torch.int32, device="cuda")
            off = torch.zeros(n_s + 1, dtype=torch.int32, device="cuda")
            st = torch.cuda.current_stream()
            for _ in range(3):
                ctx.propagate_packed_dev(0, n_s, n_s, dp, dc, ds, None, 10, seg, 10 * n_s, off, fin, val, st)
            barrier()
            reps = 10
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                ctx.propagate_packed_dev(0, n_s, n_s, dp, dc, ds, None, 10, seg, 10 * n_s, off, fin, val, st)
            e1.record()
            barrier()
            ms = max_over_ranks(e0.elapsed_time(e1))
            units = float(torch.minimum(ds, val + 1).sum().item())
            prop_obs[scene] = {"value": sum_over_ranks(units) * reps / (ms * 1e-3), "unit": "states/s", "samples": n_s,
                               "us_per_launch": 1e3 * ms / reps, "algorithmic_tflops": units * reps * FLOPS_PER_STATE / (ms * 1e-3) * 1e-12}
            del dp, dc, ds, seg, fin, val, off
    out["propagate_obstacle_scenes"] = dict(prop_obs, metric="propagated states/s (scenes with obstacles)",
                                            note="same kernel with the fused obstacle test; not the headline workload")
    # ---- conflict scan: Empty32x32 30 robots, 435 pairs x 10K steps, 256 plans (config 5) --------------
    # N > 1: every rank scans its own 256-plan batch (weak); one batch sharded by plan is reported as well (strong).
    from kcbs_b200 import sharding
    scene = "Benchmark_Empty32x32_30robots"
    P, R, T = 256, 30, 10000
    world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
    with pkg.Context(pkg.scenes.world(scene), device=dev) as ctx:
        # every rank validates its own batch of 256 candidate plans (e.g. plans of different constraint-tree nodes);
        # the only exchange is the all-gather of the 16-byte verdicts
        traj = workloads.cell_plans(scene, P, T, seed=1 + rank, xp=torch)   # 922 MB per GPU >> L2, conflict free
        res = torch.empty((P, 4), dtype=torch.int32, device="cuda")
        st = torch.cuda.current_stream()

        def scan(n_plans=P):
            ctx.first_conflict_dev(n_plans, R, T, traj, None, res[:n_plans], st)
            return res[:n_plans]

        for _ in range(3):
            allres = scan()
        barrier()
        assert int((allres[:, 0] >= 0).sum().item()) == 0, "throughput plans must be conflict free"
        reps = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            scan()
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        pairs = P * (R * (R - 1) // 2) * T
        rate = float(pairs) * world * reps / (ms * 1e-3)
        gbs = CONFLICT_BYTES_PER_TRAJ_STATE * P * R * T * reps / (e0.elapsed_time(e1) * 1e-3) * 1e-9
        # single-plan latency (435 pairs x 10K steps, L2 resident), rank-local
        one = traj[:1].contiguous()
        for _ in range(3):
            ctx.first_conflict_dev(1, R, T, one, None, res[:1], st)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(50):
            ctx.first_conflict_dev(1, R, T, one, None, res[:1], st)
        e1.record()
        torch.cuda.synchronize()
        one_us = 1e3 * e0.elapsed_time(e1) / 50
        # the same single-plan scan replayed from a CUDA graph (20 launches per replay): kernel + launch latency without
        # the ~9 us a ctypes call from Python costs
        side = torch.cuda.Stream()
        side.wait_stream(st)
        with torch.cuda.stream(side):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for _ in range(20):
                    ctx.first_conflict_dev(1, R, T, one, None, res[:1], side)
            for _ in range(3):
                g.replay()
            side.synchronize()
            e0.record(side)
            for _ in range(10):
                g.replay()
            e1.record(side)
            side.synchronize()
        one_graph_us = 1e3 * e0.elapsed_time(e1) / 200
        del g
        out["conflict"] = {"metric": "conflict-checked state-pairs/s", "value": rate, "scene": scene, "plans": P,
                           "pairs": R * (R - 1) // 2, "timesteps": T,
                           "scaling": "weak (an independent 256-plan batch per rank, no exchange)",
                           "single_plan_latency_us": one_graph_us, "single_plan_latency_python_us": one_us,
                           "roofline": {"kernel": "k_conflict_keys", "bound": "hbm", "achieved": gbs, "peak": hbm_peak,
                                        "unit": "GB/s", "frac": gbs / hbm_peak, "peak_source": peak_src,
                                        "traffic": None if "k_conflict_keys" not in traffic else
                                        traffic["k_conflict_keys"]["dram_read_bytes"] + traffic["k_conflict_keys"]["dram_write_bytes"],
                                        "algorithmic_bytes_per_launch": CONFLICT_BYTES_PER_TRAJ_STATE * P * R * T,
                                        "note": "per GPU; algorithmic 12 B per trajectory state (x, y, theta), the scan itself reads 8 B"}}
    out["conflict_strong"] = run_conflict_strong(pkg, workloads, sharding, dev, rank, world, barrier, max_over_ranks)
    # ---- per-robot replans sharded over the ranks (config 4: Empty32x32 20 robots) ------------------------------
    scene = "Empty32x32_20robots"
    with pkg.Context(pkg.scenes.world(scene), device=dev) as ctx:
        starts, goals = pkg.scenes.start_states(scene), pkg.scenes.goals(scene)
        R20, max_T = starts.shape[1], 1024
        res = torch.empty((1, 4), dtype=torch.int32, device="cuda")
        times = []
        for rep in range(4):
            barrier()
            t0 = time.perf_counter()
            plan, lengths, ok = sharding.plan_replans_sharded(ctx, starts, goals, rank, world, max_T, seed=11 + rep)
            T = int(lengths.max().item())
            Tp = (T + 3) // 4 * 4
            traj3 = plan[:, [0, 1, 4], :Tp].contiguous().unsqueeze(0)
            ctx.first_conflict_dev(1, R20, Tp, traj3, lengths.unsqueeze(0).contiguous(), res, torch.cuda.current_stream())
            barrier()
            times.append(max_over_ranks(time.perf_counter() - t0))
            assert ok
        out["replans"] = {"metric": "20 per-robot replans + all-gather + plan validation", "scene": scene, "unit": "s",
                          "higher_is_better": False, "median_s": float(np.median(times[1:])), "runs_s": [round(t, 4) for t in times],
                          "scaling": "strong (replans dealt round-robin over ranks, trajectories merged with one all-reduce)"}

    # ---- K-CBS solve time (Empty32x32): the whole planner (kcbs_kcbs_solve), rank-local -------------------------
    solve = {}
    for scene in ("Empty32x32_10robots", "Benchmark_Empty32x32_30robots"):
        with pkg.Context(pkg.scenes.world(scene), device=dev) as ctx:
            starts, goals = pkg.scenes.start_states(scene), pkg.scenes.goals(scene)
            times, ok, detail = [], 0, []
            # one untimed solve first: the planner lanes allocate their workspaces on first use
            t0 = time.perf_counter()
            ctx.kcbs_solve(starts, goals, ctx.kcbs_params(low_level=ctx.rrt_params(seed=99), time_limit_s=60.0))
            first_s = time.perf_counter() - t0
            for seed in range(1, 9):   # eight seeds: the search (10-55 constraint-tree nodes) varies a lot with the seed
                l0 = ctx.launch_count()
                t0 = time.perf_counter()
                _, _, st = ctx.kcbs_solve(starts, goals, ctx.kcbs_params(low_level=ctx.rrt_params(seed=seed), time_limit_s=60.0))
                times.append(time.perf_counter() - t0)
                ok += int(st.solved)
                detail.append({"nodes_expanded": int(st.nodes_expanded), "low_level_plans": int(st.low_level_calls),
                               "plans_validated": int(st.conflict_checks), "merges": int(st.merges),
                               "root_s": round(float(st.root_seconds), 4), "kernel_launches": int(ctx.launch_count() - l0)})
            solve[scene] = {"robots": int(starts.shape[1]), "median_s": float(np.median(times)), "runs_s": [round(t, 4) for t in times],
                            "solved": ok, "of": len(times), "runs": detail, "first_solve_incl_allocation_s": round(first_s, 4)}
    out["kcbs_solve"] = {"metric": "K-CBS solve time (Empty32x32)", "unit": "s", "higher_is_better": False, "scenes": solve,
                         "baseline": "none possible (the reference's planners live in the un-vendored OMPL K-CBS fork; only its "
                                     "180 s budget per solve is known, demo_*.cpp:168)",
                         "note": "batched GPU RRT low level + K-CBS high level"}
    return out


def run_conflict_strong(pkg, workloads, sharding, dev, rank, world, barrier, max_over_ranks):
    """BASELINE config 5 at N GPUs, STRONG scaling: ONE plan batch (identical on every rank = the all-gathered plans) is
    validated by all ranks together through kcbs_first_conflict_sharded_dev -- time slabs per rank, the per-plan keys
    exchanged inside the scan kernel over NVLink peer memory, every rank ends with every record.  Timed as CUDA-graph
    replays of `reps` back-to-back calls (events on the launching stream, max over ranks); the ranks run in lock step
    through the mailboxes.  L2: consecutive calls scan different batches (4 x 922 MB in rotation at P = 256; one
    7.4 GB batch at P = 2048), so a rank's slab is never L2 resident from the previous call."""
    import torch
    scene = "Benchmark_Empty32x32_30robots"
    R, T = 30, 10000
    pairs = R * (R - 1) // 2
    res = {"metric": "conflict-checked state-pairs/s (one batch, all ranks)", "scene": scene, "pairs": pairs, "timesteps": T,
           "scaling": "strong", "exchange": "per-plan 8-byte keys stored into every peer's mailbox from inside the scan kernel "
           "(NVLink P2P, CUDA IPC); no NCCL call and no second launch on the critical path"}
    with pkg.Context(pkg.scenes.world(scene), device=dev) as ctx:
        sharding.connect_peers(ctx, rank, world, max_plans=2048)
        st = torch.cuda.Stream()
        for P, n_sets, reps in ((256, 4, 8), (2048, 1, 2)):
            sets = [torch.cat([workloads.cell_plans(scene, 256, T, seed=1000 * P + 16 * q + c, xp=torch) for c in range(P // 256)])
                    for q in range(n_sets)]
            out = torch.empty((P, 4), dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()

            def timed(fn, n_replays=5):
                """fn(i) issues call i on stream st; returns microseconds per call (this rank)."""
                with torch.cuda.stream(st):
                    for i in range(2 * n_sets):
                        fn(i)
                    st.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=st):
                        for i in range(reps * n_sets):
                            fn(i)
                    g.replay()
                    st.synchronize()
                    barrier()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(st)
                    for _ in range(n_replays):
                        g.replay()
                    e1.record(st)
                    st.synchronize()
                    barrier()
                del g
                return 1e3 * e0.elapsed_time(e1) / (n_replays * reps * n_sets)

            us = max_over_ranks(timed(lambda i: ctx.first_conflict_sharded_dev(P, R, T, sets[i % n_sets], None, out, st)))
            ctx.peer_status()
            assert int((out[:, 0] != -1).sum().item()) == 0, "throughput plans must be conflict free on every rank"
            entry = {"plans": P, "us_per_batch": us, "value": P * pairs * T / (us * 1e-6), "unit": "state-pairs/s",
                     "slab_MB_per_rank": 12.0 * P * R * T / world / 1e6}
            if world > 1:
                # the same batch on ONE GPU (rank 0 alone, the others wait): the denominator of the speed-up, same run
                one_us = 0.0
                if rank == 0:
                    with torch.cuda.stream(st):
                        for i in range(2 * n_sets):
                            ctx.first_conflict_dev(P, R, T, sets[i % n_sets], None, out, st)
                        st.synchronize()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(st)
                        for i in range(reps * n_sets):
                            ctx.first_conflict_dev(P, R, T, sets[i % n_sets], None, out, st)
                        e1.record(st)
                        st.synchronize()
                    one_us = 1e3 * e0.elapsed_time(e1) / (reps * n_sets)
                barrier()
                one_us = max_over_ranks(one_us)
                entry["single_gpu_us_per_batch"] = one_us
                entry["speedup_vs_one_gpu"] = one_us / us
            res[f"P{P}"] = entry
            del sets, out
        barrier()
        ctx.peer_disconnect()
    return res


if __name__ == "__main__":
    main()
