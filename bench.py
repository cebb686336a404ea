#!/usr/bin/env python
"""bench.py -- aligned scan-pairs/s of the batched projective 2D registration path on B200.

Workload (BASELINE.json configs[2], "config 3" of SURVEY.md 8d): 4096 scan pairs x 1081 beams (Hokuyo
UTM-30LX shape), 10 Gauss-Newton/ICP iterations, the reference's tracking parameter set
(point_distance 0.5, normal_cos 0.9, Cauchy 0.01; LASER_0.json:598-611,76-81,498), canvas 1081 columns
over +-3.14159 rad.  A "step" is one pass of the hot path over the whole batch.

  value      device-resident throughput: clouds already in HBM, one fused-ICP launch per step, CUDA events
             on the launching stream.  The 142 MB of clouds exceed the 126 MB L2, so every step re-reads
             them from HBM (config.l2: "inputs_larger_than_l2").
  e2e        the same metric through the C ABI with HOST buffers (ls2d_align_pairs_host): pinned-memory
             H2D of both cloud sets and the initial guesses, the kernel, D2H of the 80-byte results, all
             inside the timed region.
  roofline   achieved = 34,672 algorithmic bytes per pair (16*1081 + 16*1081 + 16 + 64) x pairs
             per launch / mean launch duration, against the measured HBM copy bandwidth.
  cpu_baseline  the CPU oracle (oracle/ls2d_oracle.c, the restatement of the reference's aligner -- the
             reference itself cannot be built here) on the box's host cores.
  sub-records of the default line (each a measurement of its own, same batch unless it says otherwise):
    roofline_issue       the bound that limits the 10-iteration kernel: warp instructions (committed ncu capture) / s
    score_pass           ls2d_score_batch, the single-linearisation pass (score_kernel): HBM and issue rooflines
    sustained            --sustain seconds of back-to-back launches: time per launch and clocks under load
    e2e_pageable         e2e from pageable numpy buffers (what a caller without pinned staging sees)
    e2e_track            ls2d_track_batch: RAW scans (4 B/beam) -> pre-process -> clip -> align, pinned host buffers
    latency_single_pair  one MultiAligner2D::compute() per call through the C++ plugin class vs the 1-thread oracle
    verify               config 4 through ls2d_verify_sharded_nccl: 65,536 distinct candidate clouds x 8 guesses x 30
                         iterations, strong scaling over the ranks, mean executed iterations, winner

N > 1 (torchrun): one process per GPU, every rank aligns its own 4096-pair batch (independent pairs, no
data-path collective: weak scaling); time = max over ranks.   --impl reference times the oracle only.
--workload verify runs the sharded loop-closure verification (config 4 shape) instead.
--workload allpairs runs the large-map all-pairs loop-closure search (config 5 shape): candidate pairs grouped by query
map, groups sharded over the ranks, one all-gather of the per-map best records.
--workload multi times the MULTI.json-shaped aligner (two laser slices + odometry prior in one 3x3 system).
--workload track times the tracker's frame step from RAW scans (ls2d_track_batch: pre-process -> clip -> align).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

A_PAIR_BYTES = 16 * 1081 + 16 * 1081 + 16 + 64  # compulsory bytes per aligned pair
TRACK = dict(canvas_cols=1081, point_distance=0.5, normal_cos=0.9, cauchy_chi_threshold=0.01, max_iterations=10)
LOOP = dict(canvas_cols=1081, point_distance=1.414, normal_cos=0.8, cauchy_chi_threshold=0.05, max_iterations=30)


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(kernel: str):
    """dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/), or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(kernel)
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def dist_env():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    return rank, int(os.environ.get("LOCAL_RANK", rank)), world


N_BEAMS = 1081  # --beams (align workload only): 721 = the scan shape of the shipped configurations


def make_workload(n_pairs: int, seed: int, device: str, loop: bool = False):
    from srrg2_laser_slam_2d_b200.synthetic import make_scan_pairs
    if loop:
        return make_scan_pairs(n_pairs, seed=seed, device=device, motion_xy=0.4, motion_theta=0.2,
                               init_noise_xy=0.2, init_noise_theta=0.08)
    return make_scan_pairs(n_pairs, n_beams=N_BEAMS, seed=seed, device=device)


def bind_to_gpu_numa_node(torch, local_rank: int) -> str:
    """N > 1: run this rank on the CPUs next to its GPU (sysfs local_cpulist of the GPU's PCI function), so that the
    pinned host buffers it allocates are first-touched on that NUMA node and eight uploads do not share one
    socket's memory controllers.  Returns what it did; never fails the bench."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        addr = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open("/sys/bus/pci/devices/%s/local_cpulist" % addr) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "unchanged (no local cpus in the affinity mask)"
        os.sched_setaffinity(0, cpus)
        return "cpus %s of %s" % (spec, addr)
    except (OSError, AttributeError, ValueError) as e:
        return "unchanged (%s)" % type(e).__name__


def bench_config(n_pairs: int) -> dict:
    """the `config` object of the JSON line -- the SAME keys and values in both arms (ours / --impl reference)"""
    return {"workload": "batched scan-to-local-map registration: %d pairs x %d beams, 10 GN iterations, tracking "
                        "parameter set (config 3)" % (n_pairs, N_BEAMS),
            "pairs_per_step": n_pairs, "beams": N_BEAMS, "canvas_cols": N_BEAMS, "iterations": 10,
            "parameters": "point_distance 0.5, normal_cos 0.9, Cauchy 0.01 (LASER_0.json:598-611,76-81,498)",
            "seed": "0xC0FFEE + rank", "l2": "inputs_larger_than_l2 (142 MB of clouds per step vs 126 MB)"}


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's own CPU implementation of the path.  It cannot be compiled here (needs Eigen3 and three
    un-vendored srrg2 packages), so this arm times the oracle port of it with every host thread."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    prm = ob.default_params(**TRACK)
    cores = host_cores()  # not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1
    n_pairs = args.pairs
    sp = make_workload(n_pairs, 0xC0FFEE, "cpu")
    run = lambda: ob.align_batch(prm, sp.fixed_pts, sp.fixed_off, sp.moving_pts, sp.moving_off, sp.init_xyt,
                                 n_threads=cores, want_iters=False)
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = time.perf_counter() - t0
    value = n_pairs * args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": "aligned scan-pairs/sec (%d beams, 10 GN iters)" % N_BEAMS, "value": value,
        "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": bench_config(n_pairs),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": "%d pairs per step (the whole config-3 batch), OpenMP over pairs on %d threads; the "
                                   "reference itself cannot be built here (no Eigen / srrg2_*): in-repo oracle port"
                                   % (n_pairs, cores)},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------------------------- our arm
def timed_steps(torch, stream, step, n, barrier):
    """n calls of step() on `stream`, bracketed by barrier + synchronize; returns (total ms, per-call ms list)"""
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record(stream)
    for i in range(n):
        step()
        ev[i + 1].record(stream)
    barrier()
    return ev[0].elapsed_time(ev[-1]), [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]


def measure_e2e(torch, he, bufs, hout, args, barrier, pinned: bool):
    fp, fo, mp_, mo, init = bufs
    for _ in range(args.warmup):
        he.align_pairs_host(fp, fo, mp_, mo, init, hout)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        he.align_pairs_host(fp, fo, mp_, mo, init, hout)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    return dt


def measure_track(torch, local_rank, n, args):
    """the tracker's frame step with RAW scans as the wire format (ls2d_track_batch: pre-process -> clip -> align):
    4 B/beam + ids + poses up, 80 B/frame down; the local maps are resident"""
    from srrg2_laser_slam_2d_b200 import Handle, default_params
    from srrg2_laser_slam_2d_b200._abi import RESULT_DTYPE, default_scan_params
    from srrg2_laser_slam_2d_b200.synthetic import make_raw_scans
    raw = make_raw_scans(n, seed=0xC0FFEE, device="cuda:%d" % local_rank)
    sp_map = default_scan_params(angle_min=raw.angle_min, angle_max=raw.angle_max, voxelize_resolution=0.0)
    sp = default_scan_params(angle_min=raw.angle_min, angle_max=raw.angle_max, voxelize_resolution=args.voxel)
    h = Handle(local_rank, default_params(**TRACK))
    h.preprocess_scans_to_set(2, sp_map, raw.moving_ranges)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    ranges, ids = pin(raw.fixed_ranges), pin(np.arange(n, dtype=np.int32))
    robots, init = pin(np.zeros((n, 3), np.float32)), pin(np.zeros((n, 3), np.float32))
    out = torch.zeros(n * RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory().numpy().view(RESULT_DTYPE)
    for _ in range(args.warmup):
        h.track_batch(sp, ranges, 2, ids, robots, init, out)
    torch.cuda.synchronize()
    l0 = h.launch_count
    t0 = time.perf_counter()
    for _ in range(args.steps):
        h.track_batch(sp, ranges, 2, ids, robots, init, out)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    err = np.abs(np.stack([out["x"], out["y"], out["theta"]], 1) - raw.gt_xyt)
    rec = {"value": n * args.steps / dt, "unit": "frames/s", "ms_per_step": 1e3 * dt / args.steps,
           "h2d_bytes_per_step": int(ranges.nbytes + ids.nbytes + robots.nbytes + init.nbytes),
           "d2h_bytes_per_step": int(out.nbytes), "frames_per_step": n, "voxelize_resolution": args.voxel,
           "gpu_launches_per_step": int(h.launch_count - l0) // args.steps,
           "success_rate": float((out["status"] == 0).mean()),
           "median_abs_pose_error": [float(v) for v in np.median(err, 0)],
           "note": "ls2d_track_batch: raw 1081-beam scans (4 B/beam) -> pre-process -> clip of the resident local map -> "
                   "10 GN iterations; the same metric with the raw-range wire format instead of 16 B/point clouds"}
    h.close()
    return rec


def measure_latency(sp, n_calls=200):
    """single-pair latency of MultiAligner2D::compute() through the C++ shim (plugin_test latency), the reference's
    real tracker use (apps/visual_test_tracker_2d.cpp:167-179), next to the 1-thread oracle on the same pairs"""
    import struct
    import tempfile
    exe = os.path.join(ROOT, "srrg2_laser_slam_2d_b200", "plugin_test")
    cfg = os.path.join(ROOT, "configs", "laser_aligner_b200.json")
    if not (os.path.exists(exe) and os.path.exists(cfg)):
        return {"unavailable": "plugin_test not built"}
    n = 16
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
        f.write(struct.pack("<ii3f", n, 1, 0.0, 0.0, 0.0))
        for p in range(n):
            fx = sp.fixed_pts[sp.fixed_off[p]:sp.fixed_off[p + 1]]
            mv = sp.moving_pts[sp.moving_off[p]:sp.moving_off[p + 1]]
            f.write(struct.pack("<ii", len(fx), len(mv)))
            f.write(np.ascontiguousarray(fx, np.float32).tobytes())
            f.write(np.ascontiguousarray(mv, np.float32).tobytes())
            f.write(np.ascontiguousarray(sp.init_xyt[p], np.float32).tobytes())
        path = f.name
    try:
        r = subprocess.run([exe, "latency", cfg, "aligner_tracking", path, str(n_calls)], capture_output=True, text=True,
                           timeout=300)
        if r.returncode != 0:
            return {"unavailable": "plugin_test latency failed: " + r.stderr.strip()[-200:]}
        rec = json.loads(r.stdout.strip().splitlines()[-1])
    finally:
        os.unlink(path)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    prm = ob.default_params(**dict(TRACK, with_sensor=1))
    t0 = time.perf_counter()
    ob.align_batch(prm, sp.fixed_pts, sp.fixed_off[:n + 1], sp.moving_pts, sp.moving_off[:n + 1], sp.init_xyt[:n],
                   n_threads=1, want_iters=False)
    rec["oracle_1_thread_us"] = 1e6 * (time.perf_counter() - t0) / n
    rec["note"] = ("one MultiAligner2D::compute() per call through the plugin class: stage + upload 2 x 1081 points, "
                   "aligner, download result and iteration records (slice->correspondences() is fetched from the device on first "
                   "access, which the timed call does not make); "
                   "pageable host memory, nothing batched")
    return rec


def run_ours(args):
    import torch
    import torch.distributed as dist

    from srrg2_laser_slam_2d_b200 import Handle, default_params
    from srrg2_laser_slam_2d_b200._abi import LS2D_FIXED, LS2D_MOVING, RESULT_DTYPE

    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else "unchanged (one rank)"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_pairs = args.pairs
    sp = make_workload(n_pairs, 0xC0FFEE + rank, str(dev))
    h = Handle(local_rank, default_params(**TRACK))
    stream = torch.cuda.Stream(device=dev)   # kernels and timing events share this (non-default) stream
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    h.set_stream(stream.cuda_stream)
    words = RESULT_DTYPE.itemsize // 4

    # ---- device-resident: clouds live in HBM, one launch per step
    fp, fo = torch.from_numpy(sp.fixed_pts).to(dev), torch.from_numpy(sp.fixed_off).to(dev)
    mp, mo = torch.from_numpy(sp.moving_pts).to(dev), torch.from_numpy(sp.moving_off).to(dev)
    init = torch.from_numpy(sp.init_xyt).to(dev)
    out = torch.zeros(n_pairs * words, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()  # ls2d_set_clouds_dev reads the offsets: they must be complete
    h.set_clouds_dev(LS2D_FIXED, fp.data_ptr(), fo.data_ptr(), n_pairs, N_BEAMS)
    h.set_clouds_dev(LS2D_MOVING, mp.data_ptr(), mo.data_ptr(), n_pairs, N_BEAMS)

    def step():
        h.align_batch_dev(None, None, init.data_ptr(), n_pairs, out.data_ptr())

    for _ in range(args.warmup):
        step()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = h.launch_count
    total_ms, per_launch_ms = timed_steps(torch, stream, step, args.steps, barrier)
    launches = h.launch_count - launches0
    res = np.frombuffer(out.cpu().numpy().tobytes(), dtype=RESULT_DTYPE)
    ok_rate = float((res["status"] == 0).mean())

    # ---- the single-linearisation scoring pass (ls2d_score_batch): the regime closest to the HBM roofline
    out_s = torch.zeros(n_pairs * words, dtype=torch.int32, device=dev)
    score = lambda: h.score_batch_dev(None, None, init.data_ptr(), n_pairs, out_s.data_ptr())
    for _ in range(args.warmup):
        score()
    score_total, _ = timed_steps(torch, stream, score, args.steps, barrier)
    score_ms = score_total / args.steps

    # ---- sustained: the same step back to back for >= args.sustain seconds (the timed region above is a few ms of
    # burst clock; an issue-bound kernel slows down when the SM clock settles under load)
    sustained = None
    if args.sustain > 0:
        reps = max(args.steps, int(args.sustain * 1e3 / (total_ms / args.steps)) + 1)
        sclk = ClockSampler(local_rank)
        if rank == 0:
            sclk.start()
        sus_ms, _ = timed_steps(torch, stream, step, reps, barrier)
        sc = sclk.stop() if rank == 0 else None
        sustained = {"launches": reps, "seconds": sus_ms * 1e-3, "ms_per_launch": sus_ms / reps,
                     "pairs_per_s_per_gpu": n_pairs * reps / (sus_ms * 1e-3), "clocks": sc}

    # ---- end to end: host buffers -> C ABI -> host results (pinned = the headline e2e; pageable = what a caller that
    # never pinned anything gets)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    host = (sp.fixed_pts, sp.fixed_off, sp.moving_pts, sp.moving_off, sp.init_xyt)
    pinned = tuple(pin(a) for a in host)
    hout = torch.zeros(n_pairs * RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory().numpy().view(RESULT_DTYPE)
    he = Handle(local_rank, default_params(**TRACK))
    e2e_s = measure_e2e(torch, he, pinned, hout, args, barrier, True)
    e2e_launches = he.launch_count
    assert hout.tobytes() == res.tobytes(), "end-to-end results differ from the device-resident run"
    pageable_s = None
    if rank == 0 and world == 1:
        pout = np.zeros(n_pairs, RESULT_DTYPE)
        pageable_s = measure_e2e(torch, he, host, pout, args, barrier, False)
        assert pout.tobytes() == res.tobytes()
    clk = clocks.stop() if rank == 0 else None

    # ---- max over ranks
    t = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t[0]), float(t[1])

    # ---- the path that shards: loop-closure verification (config 4), strong scaling over the ranks
    verify = measure_verify(args, torch, dist, rank, local_rank, world) if args.verify_candidates > 0 else None

    if rank == 0:
        value = world * n_pairs * args.steps / (total_ms * 1e-3)
        e2e_value = world * n_pairs * args.steps / (e2e_ms * 1e-3)
        peak, peak_src = hbm_peak()
        mean_launch_s = float(np.mean(per_launch_ms)) * 1e-3
        achieved = A_PAIR_BYTES * n_pairs / mean_launch_s / 1e9
        h2d = sum(a.nbytes for a in pinned)
        line = {
            "metric": "aligned scan-pairs/sec (%d beams, 10 GN iters)" % N_BEAMS, "value": value, "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(n_pairs),
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(hout.nbytes), "ms_per_step": e2e_ms / args.steps,
                    "host_memory": "pinned (caller-provided)", "gpu_launches": int(e2e_launches)},
            "gpu_launches": int(launches),  # timed region of `value`; the sub-records count their own
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": recorded_traffic("icp_fused2_kernel"), "peak_source": peak_src,
                         "kernel": "icp_fused2_kernel", "algorithmic_bytes_per_launch": A_PAIR_BYTES * n_pairs,
                         "mean_launch_ms": mean_launch_s * 1e3,
                         "note": "10 fused iterations per pair make this kernel issue-bound, not HBM-bound "
                                 "(SURVEY.md 8d); see roofline_issue and DESIGN.md; the HBM-bound regime is score_pass"},
            "clocks": clk,
            "info": {"success_rate": ok_rate, "host_binding_rank0": numa},
        }
        inst = recorded_traffic("icp_fused2_kernel_warp_instructions")
        if inst and clk and clk.get("sm_mhz"):
            issue_peak = 148 * 4 * clk["sm_mhz"] * 1e6          # warp instructions / s: 4 schedulers per SM
            line["roofline_issue"] = {"bound": "issue", "achieved": inst / mean_launch_s, "peak": issue_peak,
                                      "unit": "warp-inst/s", "frac": inst / mean_launch_s / issue_peak,
                                      "warp_instructions_per_launch": inst,
                                      "note": "the bound that actually limits the 10-iteration kernel; instruction count "
                                              "from the committed ncu capture (profiles/)"}
        score_gbs = A_PAIR_BYTES * n_pairs / (score_ms * 1e-3) / 1e9
        line["score_pass"] = {"ms_per_launch": score_ms, "pairs_per_s": n_pairs / (score_ms * 1e-3),
                              "roofline": {"bound": "hbm", "achieved": score_gbs, "peak": peak, "unit": "GB/s",
                                           "frac": score_gbs / peak, "traffic": recorded_traffic("score_kernel"),
                                           "kernel": "score_kernel"},
                              "frac_of_hbm_peak": score_gbs / peak,
                              "note": "ls2d_score_batch: fixed image + one projection / linearisation per pair, same bytes"}
        s_inst = recorded_traffic("score_kernel_warp_instructions")
        if s_inst and clk and clk.get("sm_mhz"):
            issue_peak = 148 * 4 * clk["sm_mhz"] * 1e6
            line["score_pass"]["roofline_issue"] = {"bound": "issue", "achieved": s_inst / (score_ms * 1e-3),
                                                    "peak": issue_peak, "unit": "warp-inst/s",
                                                    "frac": s_inst / (score_ms * 1e-3) / issue_peak,
                                                    "warp_instructions_per_launch": s_inst}
        if sustained:
            line["sustained"] = sustained
        if pageable_s is not None:
            line["e2e_pageable"] = {"value": n_pairs * args.steps / pageable_s, "unit": "pairs/s",
                                    "ms_per_step": 1e3 * pageable_s / args.steps,
                                    "note": "the same call from pageable numpy buffers"}
        if world == 1:
            line["e2e_track"] = measure_track(torch, local_rank, n_pairs, args)
            line["latency_single_pair"] = measure_latency(sp)
        if verify:
            line["verify"] = verify
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(sp, res)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def candidate_block(block: int, device: str):
    """candidates [block * 4096, (block + 1) * 4096) of the verification workload: every block has its own seed, so any
    sharding of the candidate list over ranks sees the same clouds"""
    return make_workload(4096, 0xBEEF + block, device, loop=True)


def measure_verify(args, torch, dist, rank, local_rank, world):
    """Sharded loop-closure verification (BASELINE.json configs[3], "config 4"): one query local map against n_cand
    DISTINCT candidate local maps x n_guess initial guesses, 30 GN iterations, loop-closure parameter set; the
    candidates are split contiguously over the ranks (strong scaling: the total is fixed), the only exchange is the
    all-gather of the ranks' 48-byte ls2d_best records inside ls2d_verify_sharded_nccl (the C ABI's collective, over a
    raw ncclComm_t), then the same deterministic best-of on every rank."""
    from srrg2_laser_slam_2d_b200 import Gates, Handle, default_params
    from srrg2_laser_slam_2d_b200._abi import LS2D_FIXED, LS2D_MOVING, RESULT_DTYPE
    from srrg2_laser_slam_2d_b200.nccl_comm import NcclComm
    from srrg2_laser_slam_2d_b200.sharding import shard_range

    dev = torch.device("cuda", local_rank)
    n_cand, n_guess = args.verify_candidates, args.guesses
    lo, hi = shard_range(n_cand, rank, world)
    t_gen = time.perf_counter()
    blocks = range(lo // 4096, (max(hi, lo + 1) - 1) // 4096 + 1)
    parts, query = [], None
    b0 = candidate_block(0, str(dev))               # the query = fixed cloud 0 of block 0; candidate 0 is its match
    query = torch.from_numpy(b0.fixed_pts[:1081].copy()).to(dev)
    gt0 = b0.gt_xyt[0]
    for b in blocks:
        blk = b0 if b == 0 else candidate_block(b, str(dev))
        a, z = max(lo, b * 4096) - b * 4096, min(hi, (b + 1) * 4096) - b * 4096
        parts.append(torch.from_numpy(blk.moving_pts[a * 1081:z * 1081].copy()).to(dev))
    cands = torch.cat(parts) if parts else torch.zeros((0, 4), dtype=torch.float32, device=dev)
    del parts
    n_local = hi - lo
    off = torch.arange(n_local + 1, dtype=torch.int32, device=dev) * 1081
    qoff = torch.tensor([0, 1081], dtype=torch.int32, device=dev)
    rng = np.random.default_rng(1)
    guesses = (gt0[None, None, :] + rng.uniform(-0.15, 0.15, (n_cand, n_guess, 3))).astype(np.float32)
    gs = torch.from_numpy(guesses[lo:hi].copy()).to(dev)
    gen_s = time.perf_counter() - t_gen
    h = Handle(local_rank, default_params(**LOOP))
    stream = torch.cuda.current_stream(dev)
    h.set_stream(stream.cuda_stream)
    torch.cuda.synchronize()
    h.set_clouds_dev(LS2D_FIXED, query.data_ptr(), qoff.data_ptr(), 1, 1081)
    h.set_clouds_dev(LS2D_MOVING, cands.data_ptr(), off.data_ptr(), n_local, 1081)
    gates = Gates(300, 0.1, 0.8)
    comm, how = None, "ls2d_verify_sharded_nccl (raw ncclComm_t, all-gather of %d x 48 B)" % world
    try:
        comm = NcclComm(rank, world, local_rank)
    except Exception as e:                                   # a box without a usable NCCL: single rank only
        if world > 1:
            raise
        how = "ls2d_verify_dev (NCCL unavailable: %s)" % type(e).__name__
    best_dev = torch.zeros(12, dtype=torch.int32, device=dev)
    winner = {}

    def step():
        if comm is not None:
            b = h.verify_sharded_nccl(0, None, n_local, gs.data_ptr(), n_guess, gates, lo, comm.ptr, world)
            winner.update(candidate=int(b["candidate"]), guess=int(b["guess"]), n_inliers=int(b["n_inliers"]))
        else:
            h.verify_dev(0, None, n_local, gs.data_ptr(), n_guess, gates, lo, best_dev.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    steps = max(args.steps, 20)
    for _ in range(max(args.warmup, 3)):
        step()
    l0 = h.launch_count
    ms, _ = timed_steps(torch, stream, step, steps, barrier)
    launches = h.launch_count - l0
    # executed iterations (the kernel leaves early on a failure status): one untimed pass that keeps every result
    allr = torch.zeros(max(n_local * n_guess, 1) * (RESULT_DTYPE.itemsize // 4), dtype=torch.int32, device=dev)
    h.verify_dev(0, None, n_local, gs.data_ptr(), n_guess, gates, lo, best_dev.data_ptr(), allr.data_ptr())
    torch.cuda.synchronize()
    r = np.frombuffer(allr.cpu().numpy().tobytes(), dtype=RESULT_DTYPE)[:n_local * n_guess]
    stats = torch.tensor([ms, float(r["iterations"].sum()), float((r["status"] == 0).sum()), gen_s], dtype=torch.float64,
                         device=dev)
    mx = stats.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    if comm is not None:
        comm.close()
    h.close()
    n_align = n_cand * n_guess
    ms = float(mx[0])
    peak, _ = hbm_peak()
    a_bytes = 16 * 1081 + n_cand * 16 * 1081 + n_align * (16 + 80)     # the query is counted once
    return {"metric": "verified candidate alignments/sec (1081 beams, 30 GN iters)",
            "value": n_align * steps / (ms * 1e-3), "unit": "alignments/s", "n_gpus": world, "steps": steps,
            "ms_per_step": ms / steps, "scaling": "strong",
            "config": {"workload": "loop-closure verification: 1 query x %d candidates x %d guesses, 30 GN iterations, "
                                   "loop-closure parameter set (config 4)" % (n_cand, n_guess),
                       "distinct_candidate_clouds": n_cand, "candidate_bytes": int(n_cand * 1081 * 16),
                       "collective": how},
            "mean_executed_iterations": float(stats[1]) / n_align, "success_rate": float(stats[2]) / n_align,
            "winner": winner, "gpu_launches": int(launches),
            "fixed_part": "per step and rank: 1 aligner launch + best_of_kernel (1 CTA) + the %d x 48 B all-gather + one "
                          "stream synchronize and a %d-byte D2H of the gathered records" % (world, 48 * world),
            "roofline": {"bound": "hbm", "achieved": a_bytes / (ms / steps * 1e-3) / 1e9 / world, "peak": peak,
                         "unit": "GB/s", "frac": a_bytes / (ms / steps * 1e-3) / 1e9 / world / peak, "traffic": None,
                         "kernel": "icp_fused2_kernel", "note": "per GPU; 30 iterations per alignment: issue-bound"},
            "generation_seconds_max_rank": float(mx[3])}


def parity_against_cpu(gpu, cpu):
    """BASELINE.json's parity gate on the very batch that was timed: the oracle here sums H/b in the REFERENCE's
    sequential order, so this is tolerance parity (1e-5 m, 1e-6 rad, 1e-4 relative chi2), not the bit-exact
    comparison of the test-suite."""
    ints = np.ones(len(gpu), bool)
    for f in ("status", "n_corr", "n_inliers", "n_kernelized", "iterations"):
        ints &= gpu[f] == cpu[f]
    dth = np.abs((gpu["theta"] - cpu["theta"] + np.pi) % (2 * np.pi) - np.pi)
    chi = np.abs(gpu["chi_inliers"] - cpu["chi_inliers"]) <= 1e-4 * np.maximum(np.abs(cpu["chi_inliers"]), 1e-12)
    pose = (np.abs(gpu["x"] - cpu["x"]) <= 1e-5) & (np.abs(gpu["y"] - cpu["y"]) <= 1e-5) & (dth <= 1e-6)
    return {"pairs": int(len(gpu)), "counts_and_status_equal": float(ints.mean()),
            "pose_within_1e-5m_1e-6rad": float(pose.mean()), "chi2_within_1e-4_rel": float(chi.mean()),
            "all_three": float((ints & pose & chi).mean()),
            "against": "the CPU oracle summing in the reference's sequential order, same 4096 pairs"}


def cpu_baseline(sp, gpu_results=None):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    prm = ob.default_params(**TRACK)
    cores = host_cores()  # not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1
    n = sp.n_pairs
    best = float("inf")
    cpu = None
    for _ in range(3):
        t0 = time.perf_counter()
        cpu = ob.align_batch(prm, sp.fixed_pts, sp.fixed_off, sp.moving_pts, sp.moving_off, sp.init_xyt,
                             n_threads=cores, want_iters=False)
        best = min(best, time.perf_counter() - t0)
    m = min(n, 256)
    t0 = time.perf_counter()
    ob.align_batch(prm, sp.fixed_pts, sp.fixed_off[:m + 1], sp.moving_pts, sp.moving_off[:m + 1], sp.init_xyt[:m],
                   n_threads=1, want_iters=False)
    single = m / (time.perf_counter() - t0)
    out = {"value": n / best, "unit": "pairs/s", "cores": cores, "kind": "port",
           "sample": "the same %d pairs, OpenMP over pairs on %d threads, best of 3" % (n, cores),
           "single_thread_pairs_per_s": single}
    if gpu_results is not None and cpu is not None:
        cpu_res = cpu[0] if isinstance(cpu, tuple) else cpu
        out["parity"] = parity_against_cpu(gpu_results, cpu_res)
    return out


# ----------------------------------------------------------------------------------------------- verification
def run_verify(args):
    """--workload verify: the sharded loop-closure verification alone (the `verify` sub-record of the default line)"""
    import torch
    import torch.distributed as dist
    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rec = measure_verify(args, torch, dist, rank, local_rank, world)
    if rank == 0:
        rec.update({"warmup": max(args.warmup, 3), "higher_is_better": True, "vs_baseline": None, "dtype": "f32",
                    "data": "synthetic"})
        print(json.dumps(rec))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------- all-pairs search
def run_allpairs(args):
    """Large-map all-pairs loop-closure search (BASELINE.json configs[4], "config 5" of SURVEY.md 8d): n_maps local
    maps of map_points points each, all resident on every GPU; candidate pairs from a seeded radius query over the map
    positions, grouped by query map; ranks own contiguous runs of groups with equal pair counts; the only exchange
    is one all-gather of the per-map ls2d_best records (n_maps x 48 B per rank)."""
    import torch
    import torch.distributed as dist
    from scipy.spatial import cKDTree

    from srrg2_laser_slam_2d_b200 import Gates, Handle, default_params
    from srrg2_laser_slam_2d_b200._abi import BEST_DTYPE, LS2D_FIXED, LS2D_MOVING
    from srrg2_laser_slam_2d_b200.sharding import shard_groups
    from srrg2_laser_slam_2d_b200.synthetic import make_scan_pairs

    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_maps, n_pts, uniq = args.maps, args.map_points, min(args.maps, args.unique)
    # unique local maps: 360-degree scans of seeded rooms with n_pts beams; map m holds cloud m % uniq, materialised
    # n_maps times so the working set is the real one (20,000 x 8192 x 16 B = 2.6 GB per GPU)
    sp = make_scan_pairs(uniq, n_beams=n_pts, seed=0xA11, device=str(dev), fov=6.2, motion_xy=0.4, motion_theta=0.2)
    reps = (n_maps + uniq - 1) // uniq
    pts = torch.from_numpy(sp.moving_pts).to(dev).view(uniq, n_pts, 4).repeat(reps, 1, 1)[:n_maps].contiguous()
    off = torch.arange(n_maps + 1, dtype=torch.int32, device=dev) * n_pts
    fpts = torch.from_numpy(sp.fixed_pts).to(dev).view(uniq, n_pts, 4).repeat(reps, 1, 1)[:n_maps].contiguous()
    # candidate pairs: maps sit on a seeded random walk; every earlier map within `radius` of a query map is a candidate
    rng = np.random.default_rng(5)
    pos = np.cumsum(rng.normal(0.0, 1.0, (n_maps, 2)), 0)
    pairs = cKDTree(pos).query_pairs(args.radius, output_type="ndarray")          # (i < j)
    pairs = pairs[pairs[:, 1] - pairs[:, 0] > 8]                                   # not the immediate predecessors
    order = np.lexsort((pairs[:, 0], pairs[:, 1]))
    fid_all, mid_all = pairs[order, 1].astype(np.int32), pairs[order, 0].astype(np.int32)
    group_off = np.searchsorted(fid_all, np.arange(n_maps + 1)).astype(np.int32)   # one group per query map
    g_lo, g_hi = shard_groups(group_off, rank, world)
    p_lo, p_hi = int(group_off[g_lo]), int(group_off[g_hi])
    n_local = p_hi - p_lo
    guesses = (sp.gt_xyt[mid_all % uniq] + rng.uniform(-0.1, 0.1, (len(mid_all), 3))).astype(np.float32)
    h = Handle(local_rank, default_params(**dict(LOOP, canvas_cols=args.canvas)))
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    h.set_stream(stream.cuda_stream)
    torch.cuda.synchronize()
    h.set_clouds_dev(LS2D_FIXED, fpts.data_ptr(), off.data_ptr(), n_maps, n_pts)
    h.set_clouds_dev(LS2D_MOVING, pts.data_ptr(), off.data_ptr(), n_maps, n_pts)
    fid = torch.from_numpy(fid_all[p_lo:p_hi].copy()).to(dev)
    mid = torch.from_numpy(mid_all[p_lo:p_hi].copy()).to(dev)
    gs = torch.from_numpy(guesses[p_lo:p_hi].copy()).to(dev)
    goff = torch.from_numpy((group_off[g_lo:g_hi + 1] - p_lo).astype(np.int32)).to(dev)
    BW = BEST_DTYPE.itemsize // 4
    best = torch.zeros(n_maps * BW, dtype=torch.int32, device=dev)
    best.view(n_maps, BW)[:, 6:8] = -1                    # candidate / guess = -1 outside this rank's groups
    gathered = torch.zeros(world * n_maps * BW, dtype=torch.int32, device=dev)
    owner = torch.zeros(n_maps, dtype=torch.int64, device=dev)
    for r in range(world):
        lo, hi = shard_groups(group_off, r, world)
        owner[lo:hi] = r
    final = torch.zeros(n_maps, BW, dtype=torch.int32, device=dev)
    gates = Gates(300, 0.1, 0.8)
    rows = torch.arange(n_maps, device=dev)

    def step():
        h.verify_pairs_dev(fid.data_ptr(), mid.data_ptr(), gs.data_ptr(), n_local, goff.data_ptr(), g_hi - g_lo, gates,
                           best.data_ptr() + BEST_DTYPE.itemsize * g_lo)
        if world > 1:
            dist.all_gather_into_tensor(gathered, best)
        else:
            gathered.copy_(best)
        final.copy_(gathered.view(world, n_maps, BW)[owner, rows])   # every group has exactly one owner

    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clk = clocks.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    rec = np.frombuffer(final.cpu().numpy().tobytes(), dtype=BEST_DTYPE)
    if rank == 0:
        n_pairs = len(fid_all)
        a_bytes = 16 * n_pts * n_pairs + 16 * n_pts * int((np.diff(group_off) > 0).sum()) + 80 * n_pairs
        s_per_step = float(ms[0]) * 1e-3 / args.steps
        peak, peak_src = hbm_peak()
        print(json.dumps({
            "metric": "verified candidate pairs/sec (%d-point local maps, 30 GN iters)" % n_pts,
            "value": n_pairs / s_per_step, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "large-map all-pairs loop-closure search: %d local maps x %d points, %d candidate "
                                   "pairs (radius %.1f), 30 GN iterations, loop-closure parameter set (config 5 shape)"
                                   % (n_maps, n_pts, n_pairs, args.radius),
                       "unique_map_clouds": uniq, "resident_bytes_per_gpu": int(2 * pts.numel() * 4),
                       "canvas_cols": args.canvas, "collective": "all_gather of %d x %d x 48 B" % (world, n_maps)},
            "accepted_maps": int((rec["candidate"] >= 0).sum()),
            "checksum": int(np.bitwise_xor.reduce(rec.view(np.uint32).ravel())),
            "roofline": {"bound": "hbm", "achieved": a_bytes / s_per_step / 1e9 / world, "peak": peak, "unit": "GB/s",
                         "frac": a_bytes / s_per_step / 1e9 / world / peak, "peak_source": peak_src,
                         "kernel": "icp_stream_kernel", "traffic": None,
                         "note": "per GPU; algorithmic bytes = 16*N per candidate map + 16*N per query map + 80 per pair"},
            "gpu_launches": int(h.launch_count), "clocks": clk,
        }))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------- multi-slice aligner
def run_multi(args):
    """BASELINE.json configs[1] (stage_segway_double_config_MULTI.json:700-730): two rangefinders + wheel odometry
    fused in one SE(2) aligner, batched: 721-beam scans, Cauchy 0.01 on laser_0, no robustifier on laser_1,
    min_num_correspondences 5, 10 iterations, prior information diag(100, 100, 400)."""
    import torch

    from srrg2_laser_slam_2d_b200 import Handle, default_params
    from srrg2_laser_slam_2d_b200._abi import RESULT_DTYPE, make_prior
    from srrg2_laser_slam_2d_b200.synthetic import make_multi_sensor_pairs

    rank, local_rank, world = dist_env()
    if rank != 0:
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    n = args.pairs
    msp = make_multi_sensor_pairs(n, n_beams=721, seed=0xD0C, device=str(dev))
    base = dict(canvas_cols=721, max_iterations=10, min_num_correspondences=5, with_sensor=1, point_distance=0.5)
    sl = [default_params(normal_cos=0.9, cauchy_chi_threshold=0.01, sensor_in_robot=tuple(float(v) for v in msp.sensors[0]), **base),
          default_params(normal_cos=0.8, cauchy_chi_threshold=-1.0, sensor_in_robot=tuple(float(v) for v in msp.sensors[1]), **base)]
    prior = make_prior((100.0, 0.0, 0.0, 100.0, 0.0, 400.0))
    h = Handle(local_rank)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    h.set_stream(stream.cuda_stream)
    h.upload_clouds(0, msp.fixed_pts[0], msp.fixed_off[0])
    h.upload_clouds(2, msp.fixed_pts[1], msp.fixed_off[1])
    h.upload_clouds(1, msp.moving_pts, msp.moving_off)
    init = torch.from_numpy(msp.init_xyt).to(dev)
    z = torch.from_numpy(msp.odom_xyt).to(dev)
    out = torch.zeros(n * (RESULT_DTYPE.itemsize // 4), dtype=torch.int32, device=dev)

    def step():
        h.align_multi_dev(sl, [0, 2], [1, 1], init.data_ptr(), n, out.data_ptr(), prior=prior, prior_z_ptr=z.data_ptr())

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    l0 = h.launch_count
    clocks = ClockSampler(local_rank)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    clk = clocks.stop()
    ms = e0.elapsed_time(e1) / args.steps
    res = np.frombuffer(out.cpu().numpy().tobytes(), dtype=RESULT_DTYPE)
    err = np.abs(np.stack([res["x"], res["y"], res["theta"]], 1) - msp.gt_xyt)
    bytes_per_pair = 16 * (int(msp.fixed_off[0][1]) + int(msp.fixed_off[1][1]) + int(msp.moving_off[1])) + 12 + 12 + 64
    peak, peak_src = hbm_peak()
    print(json.dumps({
        "metric": "aligned multi-sensor pairs/sec (2 x 721 beams + odometry prior, 10 GN iters)",
        "value": n / (ms * 1e-3), "unit": "pairs/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "multi-slice registration (MULTI.json tracking aligner shape): %d pairs, two laser slices of "
                               "721 beams with their own sensor_in_robot + odometry prior" % n,
                   "success_rate": float((res["status"] == 0).mean()),
                   "median_abs_pose_error": [float(v) for v in np.median(err, 0)]},
        "roofline": {"bound": "hbm", "achieved": bytes_per_pair * n / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": bytes_per_pair * n / (ms * 1e-3) / 1e9 / peak, "peak_source": peak_src,
                     "kernel": "icp_multi_kernel", "traffic": None, "algorithmic_bytes_per_pair": bytes_per_pair},
        "gpu_launches": int(h.launch_count - l0), "clocks": clk,
    }))


# ----------------------------------------------------------------------------------------------- tracker step
def run_track(args):
    """--workload track: the tracker's frame step from RAW scans alone (the `e2e_track` sub-record of the default line)"""
    import torch
    rank, local_rank, world = dist_env()
    if rank != 0:
        return
    torch.cuda.set_device(local_rank)
    rec = measure_track(torch, local_rank, args.pairs, args)
    rec.update({"metric": "tracked frames/sec (raw 1081-beam scan -> pose; pre-process + clip + 10 GN iters)", "n_gpus": 1,
                "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic"})
    print(json.dumps(rec))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=["align", "verify", "track", "allpairs", "multi"], default="align")
    ap.add_argument("--pairs", type=int, default=4096)
    ap.add_argument("--beams", type=int, default=1081, help="align: beams per scan = canvas columns (headline: 1081)")
    ap.add_argument("--verify-candidates", type=int, default=65536,
                    help="candidates of the sharded verification sub-record (0: skip it)")
    ap.add_argument("--guesses", type=int, default=8)
    ap.add_argument("--unique", type=int, default=20000, help="allpairs: distinct map clouds (default: every map its own)")
    ap.add_argument("--sustain", type=float, default=2.0, help="seconds of back-to-back launches for the sustained sub-record")
    ap.add_argument("--maps", type=int, default=20000, help="allpairs: local maps")
    ap.add_argument("--map-points", type=int, default=8192, help="allpairs: points per local map")
    ap.add_argument("--radius", type=float, default=2.0, help="allpairs: candidate radius over the map positions")
    ap.add_argument("--canvas", type=int, default=1081, help="allpairs: projector canvas_cols")
    ap.add_argument("--voxel", type=float, default=0.02, help="track: voxelize_resolution of the pre-processor")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.beams != 1081:  # side measurement on another scan shape; the headline stays 1081 beams
        global N_BEAMS, A_PAIR_BYTES
        N_BEAMS, A_PAIR_BYTES = args.beams, 16 * args.beams + 16 * args.beams + 16 + 64
        TRACK["canvas_cols"] = args.beams
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "verify":
        run_verify(args)
    elif args.workload == "allpairs":
        run_allpairs(args)
    elif args.workload == "multi":
        run_multi(args)
    elif args.workload == "track":
        run_track(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
