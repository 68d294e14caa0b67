#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native scan->pointcloud path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload k1|k2]

Metric (BASELINE.json): Mpoints/s of 128x2048 dual-return range->XYZ (+ destaggered range), and
achieved HBM GB/s of the dominant kernel against the measured copy bandwidth.

A "step" is one pass of the hot path over one batch of `frames_per_step` synthetic frames
(default 64 frames = 33.5 Mpoints, 1.07 GB of algorithmic traffic for K1 -- larger than the
126 MB L2, so consecutive steps cannot be served from cache).

  value : device-resident inputs/outputs, one fused launch per step, CUDA-event timed.
  e2e   : the same batch through the C ABI with HOST (pinned) buffers: H2D of the range
          images and D2H of XYZ + destaggered range are inside the timed region.
  --impl reference : the reference's CPU algorithm (oracle port; the reference itself cannot be
          compiled here -- needs Eigen3) on all host cores, bounded sample per step.

Multi-GPU (torchrun, one rank per GPU): independent sensor streams shard across ranks with no
data-path collective (weak scaling); the LUT is broadcast once from rank 0 over NCCL before
the timed region.
"""
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

H, W, R = 128, 2048, 2                 # OS1-128 2048x128 dual return (BASELINE configs[1])
POINTS_PER_FRAME = H * W * R
K1_BYTES_PER_FRAME_F32 = 16_777_216     # 64 B/px = 8 (range) + 24 (LUT) + 24 (xyz) + 8 (rd)
K2_BYTES_PER_FRAME_F32 = 23_917_696
SHIFTS = np.tile(np.array([48, 32, 16, 0], np.int32), H // 4)  # OS1-128 1024-mode shifts x2


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names)
                   if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def synth_pool(n_frames, seed=42):
    """Range images as the reference's benchmark generator draws them
    (tests/benchmarks/benchmark_utils.h:93-110): ~50 % zeros (RANGE2 80 %), valid returns uniform
    in [1, 2^19-1] (19-bit RNG19 field)."""
    rs = np.random.default_rng(seed)
    rng = rs.integers(1, 1 << 19, size=(n_frames, R, H, W), dtype=np.uint32)
    rng[:, 0][rs.random((n_frames, H, W)) < 0.5] = 0
    rng[:, 1][rs.random((n_frames, H, W)) < 0.8] = 0
    return rng


def synth_lut(seed=43):
    """Random LUT as tests/benchmarks/benchmark_utils.h:112-126 (dir U(0.5,1.5), off U(0,0.01))."""
    rs = np.random.default_rng(seed)
    d = (rs.random((H * W, 3)) + 0.5).astype(np.float32)
    o = (rs.random((H * W, 3)) * 0.01).astype(np.float32)
    return d, o


def cpu_reference_pass(orc, rng_frames, d, o, threads):
    """destagger<uint32_t>() + cartesianT<float>() per frame and return, frames spread over a
    thread pool (streams are independent; ctypes releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor

    def one(f):
        for r in range(R):
            orc.destagger(rng_frames[f, r], SHIFTS)
            orc.cartesian(rng_frames[f, r], d, o)

    t0 = time.perf_counter()
    if threads <= 1:
        for f in range(rng_frames.shape[0]):
            one(f)
    else:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(one, range(rng_frames.shape[0])))
    return time.perf_counter() - t0


def run_reference(args, rank, world):
    """--impl reference: CPU arm (oracle port of the reference loops) on all host cores."""
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    cores = os.cpu_count() or 1
    sample_frames = max(cores, 16)
    rng = synth_pool(sample_frames)
    d, o = synth_lut()
    for _ in range(args.warmup):
        cpu_reference_pass(orc, rng, d, o, cores)
    ts = [cpu_reference_pass(orc, rng, d, o, cores) for _ in range(args.steps)]
    t = float(np.sum(ts))
    val = sample_frames * POINTS_PER_FRAME * args.steps / t / 1e6
    line = {
        "impl": "reference", "metric": "Mpoints/s 128x2048 dual-return range->XYZ", "value": val,
        "unit": "Mpoints/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "OS1-128 2048x128 dual-return: destagger<u32>+cartesianT<float> on CPU",
                   "frames_per_step": sample_frames},
        "cpu_baseline": {"value": val, "unit": "Mpoints/s", "cores": cores, "kind": "port",
                         "sample": f"{sample_frames} frames/step x {args.steps} steps, one thread per frame"},
        "e2e": {"value": val, "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def bind_to_gpu_numa(local_rank):
    """Multi-GPU runs: pin this rank to the CPU cores NVML reports as local to its GPU, before any
    page-locked buffer is allocated, so the e2e staging memory is NUMA-local to the PCIe root of the
    GPU.  Best effort; returns the number of cores bound to (0 = left alone)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="k1", choices=["k1", "k2"])
    ap.add_argument("--frames", type=int, default=64, help="frames per step (per GPU)")
    ap.add_argument("--streams-per-gpu", type=int, default=1,
                    help="k2: independent sensor streams (own LUT each) batched per launch per GPU; "
                         "1 = BASELINE configs[2], 8 (x8 GPUs = 64 streams) = configs[3]")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel-only", action="store_true", help="tuning aid: device-resident timing only")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    numa_cores = bind_to_gpu_numa(local_rank) if world > 1 else 0
    args.numa_cores = numa_cores
    import torch
    import __graft_entry__ as graft
    graft.build()
    ob = graft.load_package()
    if ob.device_count() <= 0:
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if not os.environ.get("OB_KEEP_NCCL_DEBUG"):
            os.environ.pop("NCCL_DEBUG", None)   # keep stdout to the single JSON line
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    if args.workload == "k2":
        from bench_k2 import run_k2
        run_k2(args, ob, torch, dist, rank, local_rank, world, ClockSampler, measured_peaks, ROOT)
        return

    F = args.frames
    # ---- inputs: each rank owns F independent frames (one "sensor stream shard") ----
    rng_host = synth_pool(F, seed=42 + rank)
    d, o = synth_lut()
    # the only collective: one LUT broadcast from rank 0, outside the timed region
    t_dir, t_off = ob.sharding.broadcast_lut(d, o, dist, src=0, device=dev)
    lut = ob.XYZLutT.from_arrays(t_dir, t_off, H, W, device=local_rank)
    t_rng = torch.from_numpy(rng_host.view(np.int32)).to(dev)
    t_xyz = torch.empty((F, R, H * W, 3), dtype=torch.float32, device=dev)
    t_rd = torch.empty((F, R, H, W), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    obs = ob.Stream(local_rank, cuda_stream=stream.cuda_stream)

    def step():
        ob.scan_to_cloud(lut, SHIFTS, t_rng, xyz=t_xyz, range_destaggered=t_rd, stream=obs)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    l0 = ob.kernel_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record(stream)
    for i in range(args.steps):
        step()
        ev[i + 1].record(stream)
    barrier()
    launches = ob.kernel_launch_count() - l0
    clocks = sampler.stop()
    ms_total = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    t_ms = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_total_max = float(t_ms.item())
    value = world * F * POINTS_PER_FRAME * args.steps / (ms_total_max * 1e-3) / 1e6

    if args.kernel_only:
        if rank == 0:
            peak, _ = measured_peaks()
            avg = float(np.mean(per_launch_ms)) * 1e-3
            print(json.dumps({"value": value, "ms_per_step": ms_total_max / args.steps,
                              "gbps": K1_BYTES_PER_FRAME_F32 * F / avg / 1e9,
                              "frac": K1_BYTES_PER_FRAME_F32 * F / avg / 1e9 / peak,
                              "env": {k: v for k, v in os.environ.items() if k.startswith("OB_")},
                              "clocks": clocks}))
        return

    # device results of frame 0, kept for the e2e self-check and the cpu_baseline leg's parity check
    xyz0 = t_xyz[0].cpu().numpy() if rank == 0 else None
    rd0 = t_rd[0].cpu().numpy().view(np.uint32) if rank == 0 else None
    parity = None

    # ---- e2e: host (pinned) buffers through the C ABI, copies inside the timed region ----
    CH = 8                                   # frames per call
    n_chunks = F // CH
    NS = 3                                   # streams in flight: H2D / kernel / D2H overlap
    h_rng = ob.pinned_empty((F, R, H, W), np.uint32)
    h_rng[...] = rng_host
    h_xyz = ob.pinned_empty((F, R, H * W, 3), np.float32)
    h_rd = ob.pinned_empty((F, R, H, W), np.uint32)
    tstreams = [torch.cuda.Stream(device=dev) for _ in range(NS)]
    ostreams = [ob.Stream(local_rank, cuda_stream=s.cuda_stream) for s in tstreams]

    def e2e_step():
        for c in range(n_chunks):
            sl = slice(c * CH, (c + 1) * CH)
            ob.scan_to_cloud(lut, SHIFTS, h_rng[sl], xyz=h_xyz[sl], range_destaggered=h_rd[sl],
                             stream=ostreams[c % NS])

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for s in tstreams:
        s.wait_stream(stream)
    for _ in range(e2e_steps):
        e2e_step()
    for s in tstreams:
        stream.wait_stream(s)
    e1.record(stream)
    barrier()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_val = world * F * POINTS_PER_FRAME * e2e_steps / (float(e2e_ms.item()) * 1e-3) / 1e6
    e2e_ok = bool(np.array_equal(h_xyz[0, 0], xyz0[0])) if rank == 0 else None

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- the reference's default XYZLut is double: same launch with a float64 LUT, for the record ----
    f64 = None
    try:
        lut64 = ob.XYZLutT.from_arrays(t_dir.double(), t_off.double(), H, W, device=local_rank)
        t_xyz64 = torch.empty((F, R, H * W, 3), dtype=torch.float64, device=dev)
        for _ in range(3):
            ob.scan_to_cloud(lut64, SHIFTS, t_rng, xyz=t_xyz64, range_destaggered=t_rd, stream=obs)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a0.record(stream)
        for _ in range(5):
            ob.scan_to_cloud(lut64, SHIFTS, t_rng, xyz=t_xyz64, range_destaggered=t_rd, stream=obs)
        a1.record(stream)
        torch.cuda.synchronize()
        s64 = a0.elapsed_time(a1) / 5 * 1e-3
        f64 = {"value_mpoints_s": F * POINTS_PER_FRAME / s64 / 1e6, "bytes_per_frame": 29_360_128,
               "achieved_gbps": F * 29_360_128 / s64 / 1e9}
        del t_xyz64, lut64
    except Exception as ex:  # supplementary figure only
        f64 = {"error": str(ex)}

    # ---- roofline of the dominant kernel (the fused launch IS the step) ----
    peak, peak_kind = measured_peaks()
    avg_launch_s = float(np.mean(per_launch_ms)) * 1e-3
    achieved = K1_BYTES_PER_FRAME_F32 * F / avg_launch_s / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")

    # ---- CPU baseline in the same run: oracle port on the host cores, bounded sample ----
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        from oracle import oracle as orc   # test infrastructure: used only in this CPU-baseline leg
        parity = all(np.array_equal(xyz0[r], orc.cartesian(rng_host[0, r], d, o)) and
                     np.array_equal(rd0[r], orc.destagger(rng_host[0, r], SHIFTS)) for r in range(R))
        cores = os.cpu_count() or 1
        nf = max(16, cores)
        sample = rng_host[:nf] if nf <= F else synth_pool(nf)
        cpu_reference_pass(orc, sample[:2], d, o, 1)
        t1 = cpu_reference_pass(orc, sample[:8], d, o, 1)
        reps = 3
        tN = min(cpu_reference_pass(orc, sample, d, o, cores) for _ in range(reps))
        cpu = {"value": sample.shape[0] * POINTS_PER_FRAME / tN / 1e6, "unit": "Mpoints/s",
               "cores": cores, "kind": "port",
               "sample": f"{sample.shape[0]} frames, destagger<u32>+cartesianT<float> per return, "
                         f"one thread per frame, best of {reps}",
               "single_thread_value": 8 * POINTS_PER_FRAME / t1 / 1e6}

    line = {
        "metric": "Mpoints/s 128x2048 dual-return range->XYZ", "value": value, "unit": "Mpoints/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "OS1-128 2048x128 dual-return fused destagger+cartesian (K1), "
                               "LUT tiles staged in smem via TMA",
                   "frames_per_step_per_gpu": F, "points_per_frame": POINTS_PER_FRAME,
                   "l2_policy": f"inputs+outputs per step {F * K1_BYTES_PER_FRAME_F32 / 1e6:.0f} MB > 126 MB L2",
                   "parallelism": f"{world} independent stream shards, LUT broadcast only",
                   "numa_bound_cores_per_rank": numa_cores},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_kind": peak_kind,
                     "kernel": "cloud_tma_kernel<float,2>",
                     "algorithmic_bytes_per_launch": K1_BYTES_PER_FRAME_F32 * F,
                     "avg_launch_ms": avg_launch_s * 1e3,
                     "frac_of_ncu_dram_traffic": (traffic / avg_launch_s / 1e9 / peak) if traffic else None,
                     "note": "algorithmic bytes count the 6.3 MB LUT once per frame (SURVEY 8d); it is "
                             "L2-resident across the frames of a launch, so DRAM traffic (ncu) is lower "
                             "and frac can exceed 1"},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_val, "unit": "Mpoints/s",
                "h2d_bytes_per_step": int(F * R * H * W * 4),
                "d2h_bytes_per_step": int(F * R * H * W * (12 + 4)),
                "frames_per_call": CH, "streams": NS, "matches_device_path": e2e_ok},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "parity_vs_oracle": parity,
        "f64_lut": f64,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
