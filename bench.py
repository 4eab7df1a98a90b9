#!/usr/bin/env python
"""bench.py -- headline benchmark of the render() -> ray_color() -> hit()/scatter() hot path on B200.

Metric (BASELINE.json): Mrays/s on scene_random_spheres, 1920x1080, 1000 spp, 50 bounces, Float32.
A "ray" is one ray segment = one execution of hit(world, r, ...) (src/ray_color.jl:19), counted on the device.
A "step" is one full render of that workload (all rows, all samples).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--spp S] [--width W] [--depth D]
  N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
             bench.py --gpus N --steps K --warmup W
         one rank per GPU; image rows are interleaved over ranks (row r -> rank r mod N); the only data-path
         collective is one NCCL gather of the finished row tiles to rank 0.

Prints ONE JSON line on rank 0 (see the keys at the bottom).  `value` is measured with the scene resident in
HBM and the image left in HBM; `e2e` goes through the host-buffer C-ABI call (scene H2D + image D2H inside).
PyTorch is plumbing only (device buffers, streams/events, torch.distributed); the product is librtw_b200.so.
The oracle (oracle/) is used only for the cpu_baseline leg and for --impl reference.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

METRIC = "Mrays/s (1920x1080x1000spp random_spheres)"
UNIT = "Mrays/s"
FP32_INSTR_PER_TEST = 11  # 3 FADD + 2 FMUL + 6 FFMA per ray-sphere test (src/hit.jl:13-18), SURVEY.md 8(d)
FLOP_PER_TEST = 17


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--spp", type=int, default=1000)
    ap.add_argument("--depth", type=int, default=50)
    ap.add_argument("--cpu-spp", type=int, default=0, help="spp of the bounded CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the N = 1 side measurements (grid, cfg5, Float64, latency)")
    ap.add_argument("--gather", default="", choices=["", "peer", "nccl"],
                    help="framebuffer gather inside the multi-device rtw_render_scene call (e2e arm, N > 1): "
                         "peer copies (library default) or NCCL send/recv")
    ap.add_argument("--half-extent", type=int, default=11,
                    help="grid half extent of scene_random_spheres: 11 = the reference scene (484 spheres), "
                         "158 = BASELINE configs[4] (~100k spheres, TMA-streamed sweep)")
    ap.add_argument("--mode", default="linear", choices=["linear", "grid"],
                    help="linear = the sphere-list sweep of the reference (default, the benchmarked path); grid = "
                         "RTW_MODE_GRID, the same image from a uniform-grid traversal (not comparable with the roofline)")
    return ap.parse_args()


def workload_name(args):
    return (f"scene_random_spheres (reseed!(); {{n}} spheres), camera t_cam1, {args.width}x{(args.width * 9) // 16}, "
            f"{args.spp} spp, max_depth {args.depth}, Float32, seed 1")


def build_scene(half_extent: int = 11):
    import rtw_b200 as R

    R.reseed()
    scene = R.flatten_scene(R.scene_random_spheres(half_extent=half_extent))
    return R, scene, R.t_cam1()


# ------------------------------------------------------------------------------------------- clocks sampling
class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.device_index = device_index
        self.rows = []
        self._stop = threading.Event()
        self._thread = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device_index), f"--query-gpu={self.QUERY}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def stop(self) -> dict:
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=6)
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------- CPU legs (oracle)
def host_cpu_info() -> dict:
    """What the process can actually use: the affinity mask and the cgroup CPU quota (cpu.max), next to nproc."""
    info = {"nproc": os.cpu_count()}
    try:
        info["affinity"] = len(os.sched_getaffinity(0))
    except Exception:
        info["affinity"] = None
    quota = None
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = Path(path).read_text().split()
            if path.endswith("cpu.max"):
                quota = None if txt[0] == "max" else float(txt[0]) / float(txt[1])
            else:
                q = float(txt[0])
                quota = None if q <= 0 else q / float(Path("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read_text())
            break
        except Exception:
            continue
    info["cgroup_cpus"] = quota
    return info


_ORACLE_TIMING = None


def oracle_for_timing():
    """The oracle binding used for the timed CPU legs: the same C source rebuilt ON THIS HOST with -O3 -march=native
    (still -ffp-contract=off, so the image bits do not change -- asserted on a small render against the portable -O2
    build that the parity tests use).  Falls back to the portable build, and says so, when gcc is unavailable."""
    global _ORACLE_TIMING
    if _ORACLE_TIMING is not None:
        return _ORACLE_TIMING
    from oracle import binding as O

    note = "C oracle, portable build (-O2 -mavx2 -mfma)"
    try:
        r = subprocess.run(["make", "-C", os.fspath(ROOT / "oracle"), "-B", "native"], capture_output=True, text=True,
                           timeout=300)
        native = ROOT / "oracle" / "_native" / "librtw_oracle_native.so"
        if r.returncode == 0 and native.exists():
            import rtw_b200 as R

            g, m, k = R.flatten_scene(R.scene_4_spheres())
            cam = R.t_default_cam().as_array()
            a, _, sa = O.render(g, m, k, cam, 64, 4, max_depth=8, seed=3, n_threads=1)
            O.use_library(native)
            b, _, sb = O.render(g, m, k, cam, 64, 4, max_depth=8, seed=3, n_threads=1)
            if np.array_equal(a, b) and sa["ray_segments"] == sb["ray_segments"]:
                note = "C oracle rebuilt on this host with -O3 -march=native -ffp-contract=off (image bits equal to the -O2 build)"
            else:
                O.use_library(O.LIB_PATH_DEFAULT)
                note += "; the -march=native build changed image bits and was rejected"
    except Exception as e:  # no compiler on the box etc.
        note += f"; native rebuild unavailable ({type(e).__name__})"
    _ORACLE_TIMING = (O, note)
    return _ORACLE_TIMING


def cpu_threads() -> int:
    info = host_cpu_info()
    n = info["affinity"] or info["nproc"] or 1
    if info["cgroup_cpus"]:
        n = max(1, min(n, int(round(info["cgroup_cpus"]))))
    return n


def cpu_sample(args, scene, cam, spp: int, threads: int = 0) -> dict:
    """Times the CPU oracle (the port of the reference's algorithm; Julia itself is not installed) on a bounded
    sample of the same workload: same scene/camera/width/depth, fewer samples per pixel (cost is linear in spp).
    Threads = the CPUs this process may use (affinity mask, capped by the cgroup quota)."""
    O, note = oracle_for_timing()
    g, m, k = scene
    threads = threads or cpu_threads()
    _, _, st = O.render(g, m, k, cam.as_array(), args.width, spp, max_depth=args.depth, seed=1, n_threads=threads)
    return {"value": st["ray_segments"] / st["seconds"] / 1e6, "unit": UNIT, "cores": st["threads"], "kind": "port",
            "sample": f"same scene/camera/{args.width}x{(args.width * 9) // 16}/depth {args.depth} at {spp} spp "
                      f"({st['paths']} paths, {st['ray_segments']} ray segments, {st['seconds']:.2f} s); "
                      f"{note}; Philox stream; 16-pixel blocks dealt round-robin to {st['threads']} pthreads",
            "host": host_cpu_info(),
            "seconds": st["seconds"], "ray_segments": st["ray_segments"], "paths": st["paths"]}


def auto_cpu_spp(args) -> int:
    if args.cpu_spp > 0:
        return args.cpu_spp
    # ~0.4 Mrays/s per core (native build, 484-sphere scene) => a few spp of 1080p is 10-30 s of CPU work
    cores = cpu_threads()
    pixels = args.width * ((args.width * 9) // 16)
    target_s = 15.0
    rays = target_s * 0.4e6 * cores  # cost is linear in n_spheres
    if args.half_extent != 11:
        rays *= 484.0 / (4.0 * args.half_extent * args.half_extent)
    return max(1, min(args.spp, int(rays / (pixels * 4.1))))


def julia_probe() -> dict:
    """If a `julia` binary exists on this box (it does not in the build image), time the reference's own render() with
    tools/dump_reference_golden.jl -- the only way to pin the oracle's stream against Julia."""
    import shutil

    exe = shutil.which("julia")
    if not exe:
        return {"found": False}
    out = {"found": True, "path": exe}
    try:
        r = subprocess.run([exe, "--threads=1", os.fspath(ROOT / "tools" / "dump_reference_golden.jl"),
                            os.fspath(ROOT / "gpurun_out" / "reference_golden_cfg1.bin")], capture_output=True, text=True,
                           timeout=900)
        out["returncode"] = r.returncode
        out["tail"] = (r.stdout + r.stderr)[-400:]
    except Exception as e:
        out["error"] = str(e)
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is Julia and cannot
    be installed offline (no julia binary, registry packages), so this arm times the oracle port on all host
    cores, on a bounded sample of the same workload.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    R, scene, cam = build_scene(args.half_extent)
    spp = auto_cpu_spp(args)
    spp = max(1, spp // max(1, args.steps + args.warmup)) if args.cpu_spp == 0 else spp
    for _ in range(min(args.warmup, 1)):
        cpu_sample(args, scene, cam, 1)
    vals, secs = [], []
    last = None
    for _ in range(args.steps):
        last = cpu_sample(args, scene, cam, spp)
        vals.append(last["value"])
        secs.append(last["seconds"])
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args).format(n=len(scene[2])), "note": "each step = bounded sample of the workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": "port", "sample": last["sample"],
                         "host": last["host"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "julia": julia_probe(),
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- GPU arm
def image_sha256(host_array) -> str:
    import hashlib

    return hashlib.sha256(np.ascontiguousarray(host_array).tobytes()).hexdigest()


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cpu_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        # host-side barrier for the phases in which rank 0 alone drives all GPUs (an NCCL barrier would park a spinning
        # kernel on every other GPU)
        cpu_group = dist.new_group(backend="gloo")

    R, scene, cam = build_scene(args.half_extent)
    n_spheres = len(scene[2])
    W, spp, depth = args.width, args.spp, args.depth
    H = R.image_height(W)
    G = world
    rows_pad = R.sharding.rows_pad(H, G)

    r = R.Renderer([local_rank])
    if args.mode == "grid":
        r.set_option(R.RTW_OPT_MODE, R.RTW_MODE_GRID)
    r.set_scene(scene)
    # a dedicated (non-default) stream: the library enqueues on exactly this stream, so torch CUDA events see it
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    tile = torch.zeros((rows_pad, W, 3), dtype=torch.float32, device=dev)
    image = torch.zeros((W, H, 3), dtype=torch.float32, device=dev)  # Julia column-major H x W x RGB
    gathered = torch.zeros((G, rows_pad, W, 3), dtype=torch.float32, device=dev) if (G > 1 and rank == 0) else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    host_image = torch.empty((W, H, 3), dtype=torch.float32).pin_memory() if rank == 0 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def host_barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier(group=cpu_group)

    def one_step_resident():
        """scene already in HBM, image stays in HBM: trace + resolve (+ gather + assemble for G > 1)"""
        if G == 1:
            r.render_rows_device(cam, W, spp, image.data_ptr(), max_depth=depth, seed=1, column_major=True,
                                 stream=stream.cuda_stream)
        else:
            r.render_rows_device(cam, W, spp, tile.data_ptr(), max_depth=depth, seed=1, row_start=rank, row_stride=G,
                                 stream=stream.cuda_stream)
            dist.gather(tile, list(gathered.unbind(0)) if rank == 0 else None, dst=0)
            if rank == 0:
                r.assemble_tiles_device(gathered.data_ptr(), G, W, image.data_ptr(), stream=stream.cuda_stream)

    def timed_steps(step_fn, k):
        """k steps, each bracketed by CUDA events on the launching stream; L2 flushed between steps"""
        per_step = []
        segs = 0
        barrier()
        t_wall0 = time.perf_counter()
        for _ in range(k):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            extra = step_fn()
            e1.record(stream)
            e1.synchronize()
            per_step.append(e0.elapsed_time(e1))
            segs += r.stats()["ray_segments"] if extra is None else extra
        barrier()
        wall = time.perf_counter() - t_wall0
        return per_step, segs, wall

    # ---- warm-up
    for _ in range(max(args.warmup, 0)):
        one_step_resident()
    barrier()

    # ---- FP32 roofline denominator, measured live on this GPU (not in MEASURED_PEAKS.json)
    fp32_peak, _ = r.measure_fp32_peak(0)
    fp32_sweep_mix, _ = r.measure_fp32_peak(2)

    # ---- timed region: resident
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    per_step, my_segs, _ = timed_steps(lambda: (one_step_resident(), None)[1], args.steps)
    clocks = sampler.stop() if rank == 0 else None
    stats = r.stats()
    t_ms = torch.tensor([sum(per_step)], dtype=torch.float64, device=dev)
    seg_t = torch.tensor([float(my_segs)], dtype=torch.float64, device=dev)
    trace_ms = torch.tensor([stats["ms_trace"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(seg_t, op=dist.ReduceOp.SUM)
        dist.all_reduce(trace_ms, op=dist.ReduceOp.MAX)
    total_ms = float(t_ms.item())
    total_segs = float(seg_t.item())
    ms_per_step = total_ms / args.steps
    value = total_segs / (total_ms * 1e-3) / 1e6
    segs_per_step = total_segs / args.steps
    # the image of the resident arm (NCCL-gathered for G > 1), hashed on the host: bit-identity across N is visible in
    # the driver's own SCALE records
    resident_sha = None
    if rank == 0:
        host_image.copy_(image)
        torch.cuda.synchronize(dev)
        resident_sha = image_sha256(host_image.numpy())

    # ---- end to end through the product's own host-buffer call: ONE rtw_render_scene per step (scene H2D, trace on
    # all G devices of ONE context owned by rank 0, framebuffer gather onto device 0, image D2H into pinned host memory).
    # This is the call the Julia `render` method binds to; for G > 1 the other ranks idle at a host-side barrier.
    scene_bytes = int(sum(a.nbytes for a in scene)) + 88
    image_bytes = W * H * 3 * 4
    e2e = None
    host_barrier()
    if rank == 0:
        rr = r if G == 1 else R.Renderer(list(range(G)))
        if G > 1 and args.mode == "grid":
            rr.set_option(R.RTW_OPT_MODE, R.RTW_MODE_GRID)
        if G > 1 and args.gather:
            rr.set_option(R.RTW_OPT_GATHER, {"peer": R.RTW_GATHER_PEER, "nccl": R.RTW_GATHER_NCCL}[args.gather])
        try:
            rr.render(cam, W, spp, max_depth=depth, seed=1, scene=scene, out=host_image.numpy())  # warm-up (buffers, peers)
            e2e_steps = args.steps
            e2e_segs = 0
            dev_ms = 0.0
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                rr.render(cam, W, spp, max_depth=depth, seed=1, scene=scene, out=host_image.numpy())
                e2e_segs += rr.last_stats["ray_segments"]
                dev_ms += rr.last_stats["ms_total"]
            e2e_wall = time.perf_counter() - t0
            e2e = {"value": e2e_segs / e2e_wall / 1e6, "unit": UNIT, "h2d_bytes_per_step": scene_bytes * G,
                   "d2h_bytes_per_step": image_bytes, "steps": e2e_steps, "ms_per_step": e2e_wall / e2e_steps * 1e3,
                   "device_ms_per_step": dev_ms / e2e_steps,
                   "call": f"one rtw_render_scene per step on a {G}-device context (host buffers in and out; "
                           f"gather = {rr.gather_name()})",
                   "image_sha256": image_sha256(host_image.numpy())}
        finally:
            if rr is not r:
                rr.close()
    host_barrier()

    # ---- N = 1: the row-tile decomposition the multi-GPU runs use (4 interleaved tiles + assemble) gives the same bits
    tiles_sha = None
    if world == 1:
        Gt = 4
        rp = R.sharding.rows_pad(H, Gt)
        tiles = torch.zeros((Gt, rp, W, 3), dtype=torch.float32, device=dev)
        tspp = min(spp, 8)
        for g in range(Gt):
            r.render_rows_device(cam, W, tspp, tiles[g].data_ptr(), max_depth=depth, seed=1, row_start=g, row_stride=Gt,
                                 stream=stream.cuda_stream)
        r.assemble_tiles_device(tiles.data_ptr(), Gt, W, image.data_ptr(), stream=stream.cuda_stream)
        stream.synchronize()
        a = image.cpu().numpy().copy()
        r.render_rows_device(cam, W, tspp, image.data_ptr(), max_depth=depth, seed=1, column_major=True,
                             stream=stream.cuda_stream)
        stream.synchronize()
        b = image.cpu().numpy()
        tiles_sha = {"spp": tspp, "tiles": Gt, "assembled": image_sha256(a), "single": image_sha256(b)}
        if tiles_sha["assembled"] != tiles_sha["single"]:
            raise SystemExit("row-tile render differs from the single-device render")
        del tiles

    # ---- extras (N = 1, default mode only), reported beside -- never instead of -- the headline
    extras = {}
    if world == 1 and args.mode == "linear" and not args.no_extras:
        extras = run_extras(args, R, r, scene, cam, stream, timed_steps, one_step_resident, n_spheres, fp32_peak)

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        c = cpu_sample(args, scene, cam, auto_cpu_spp(args))
        cpu = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample", "host")}

    if rank == 0:
        tests_per_step = segs_per_step * n_spheres
        trace_s = float(trace_ms.item()) * 1e-3  # dominant kernel: fused_trace2_kernel, per launch (max over ranks)
        achieved_instr = tests_per_step / G * FP32_INSTR_PER_TEST / trace_s / 1e12  # per GPU
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": G, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args).format(n=n_spheres), "l2": "flushed between steps (256 MiB write)",
                       "parallelism": f"rows interleaved over {G} GPU(s), one NCCL gather" if G > 1 else "1 GPU",
                       "kernel": "fused persistent trace, unified tail (packed FP32x2 mask sweep, cooperative rejection "
                                 "sampling) + resolve"},
            "paths_per_step": W * H * spp, "ray_segments_per_step": segs_per_step,
            "segments_per_path": segs_per_step / (W * H * spp), "Mpaths_per_s": W * H * spp / (ms_per_step * 1e-3) / 1e6,
            "e2e": e2e,
            "image_sha256": e2e["image_sha256"] if e2e else None,
            "image_sha256_resident_arm": resident_sha,
            # kernels of librtw_b200.so launched inside the timed region: per rank and step the u/v table kernel, the
            # fused trace kernel and the resolve kernel (rtw_stats.kernel_launches), plus one assemble on rank 0 for G > 1
            "gpu_launches": (int(stats["kernel_launches"]) * G + (1 if G > 1 else 0)) * args.steps,
            "roofline": {
                "bound": "fp32", "kernel": "fused_trace2_kernel",
                "achieved": achieved_instr, "peak": fp32_peak / 1e12,
                "unit": "T FP32 instr/s per GPU (11 per ray-sphere test)",
                "frac": achieved_instr / (fp32_peak / 1e12),
                "frac_of_nominal": achieved_instr / (148 * 128 * 1.965e9 / 1e12),
                "peak_source": "measured live: rtw_measure_fp32_peak variant 0 (independent FFMA chains, all SMs); "
                               "MEASURED_PEAKS.json holds no FP32 CUDA-core figure; nominal = 148 SMs x 128 lanes x 1.965 GHz",
                "sweep_mix_ceiling": fp32_sweep_mix / 1e12,
                "achieved_tflops": tests_per_step * FLOP_PER_TEST / trace_s / 1e12 / G,
                # DRAM bytes per launch of the trace kernel from the committed ncu --set full capture (a 1080p slice;
                # it is the 66 MB fixed-point accumulator, independent of spp) -- the kernel is FP32-bound, not HBM-bound
                "traffic": measured_dram_traffic(),
                "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu)",
                "algorithmic_hbm_bytes": W * H * 4 * 8 if G == 1 else None,
            },
            "clocks": clocks,
        }
        if resident_sha and e2e and resident_sha != e2e["image_sha256"]:
            raise SystemExit("the NCCL-gathered image differs from the single-call multi-device image")
        if tiles_sha:
            line["row_tiles_check"] = tiles_sha
        # secondary roofline: the kernel's measured DRAM traffic against the measured HBM copy bandwidth -- shows how far
        # from HBM-bound this path is (the sphere list lives in shared memory, rays in registers)
        traffic = line["roofline"]["traffic"]
        hbm_peak, hbm_src = hbm_peak_gbs()
        if traffic:
            line["roofline"]["hbm"] = {"achieved": traffic / trace_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                       "frac": traffic / trace_s / 1e9 / hbm_peak, "peak_source": hbm_src}
        if args.mode == "grid":
            # the grid traversal skips most ray-sphere tests: the linear-sweep work model does not apply
            line["mode"] = "grid (RTW_MODE_GRID: uniform-grid traversal, same image bits as the linear sweep)"
            line["config"]["kernel"] = "fused persistent trace, unified tail, uniform-grid closest hit + resolve"
            line["roofline"] = {"bound": "latency/divergence (per-lane grid traversal); no linear-sweep work model",
                                "achieved": None, "peak": fp32_peak / 1e12, "unit": "T FP32 instr/s", "frac": None,
                                "equivalent_linear_sweep_T_instr_s": achieved_instr, "traffic": None,
                                "grid_fallback_rays_per_step": stats.get("grid_fallback_rays")}
        line.update(extras)
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    r.close()
    if world > 1:
        dist.destroy_process_group()


def grid_work_model(st) -> dict:
    """RTW_MODE_GRID has no linear-sweep work model; its own: cells walked and sphere tests made per ray segment (device
    counters), against the n_spheres tests per segment of the linear sweep."""
    segs = max(1, st["ray_segments"])
    return {"cells_per_segment": st["grid_cells"] / segs, "tests_per_segment": st["grid_tests"] / segs,
            "linear_sweep_tests_per_segment": st["n_spheres"],
            "tests_saved_factor": st["n_spheres"] * segs / max(1, st["grid_tests"]),
            "Mtests_per_s": st["grid_tests"] / (st["ms_trace"] * 1e-3) / 1e6 if st["ms_trace"] else None}


def run_extras(args, R, r, scene, cam, stream, timed_steps, one_step_resident, n_spheres, fp32_peak) -> dict:
    """N = 1 side measurements on the same GPU: RTW_MODE_GRID on the headline workload, BASELINE configs[4] (100k
    spheres, 1920x1080x256 spp) through the grid, the Float64 instantiation, and the small-render latency."""
    import torch

    out = {}
    W, depth = args.width, args.depth
    try:  # the same workload through RTW_MODE_GRID -- same image bits
        r.set_option(R.RTW_OPT_MODE, R.RTW_MODE_GRID)
        one_step_resident()
        g_ms, g_segs, _ = timed_steps(lambda: (one_step_resident(), None)[1], 2)
        out["grid_mode"] = {"value": g_segs / (sum(g_ms) * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": sum(g_ms) / 2, "steps": 2,
                            "grid_fallback_rays_per_step": r.stats().get("grid_fallback_rays"),
                            "grid_loose_cells_per_step": r.stats().get("grid_loose_cells"),
                            "work_model": grid_work_model(r.stats()),
                            "note": "RTW_MODE_GRID: uniform-grid traversal instead of the linear sweep, bit-identical image; "
                                    "not the benchmarked path (no linear-sweep roofline applies)"}
    except Exception as e:  # the headline must not depend on the optional mode
        out["grid_mode"] = {"error": str(e)}
    finally:
        r.set_option(R.RTW_OPT_MODE, R.RTW_MODE_FUSED)
    try:  # BASELINE configs[4]: ~100k spheres, 1920x1080, 256 spp (full), grid mode
        R.reseed()
        t0 = time.perf_counter()
        big = r.generate_random_spheres(158, install=False)  # the host loop's list, bit for bit, built on the device
        gen = "device (rtw_scene_random_spheres)"
        if big is None:
            R.reseed()
            big = R.flatten_scene(R.scene_random_spheres(half_extent=158))
            gen = "host loop"
        t_gen = time.perf_counter() - t0
        r.set_option(R.RTW_OPT_MODE, R.RTW_MODE_GRID)
        r.set_scene(big)
        img = torch.zeros((W, R.image_height(W), 3), dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))

        def step5():
            r.render_rows_device(cam, W, 256, img.data_ptr(), max_depth=depth, seed=1, column_major=True, stream=stream.cuda_stream)

        step5()
        ms5, segs5, _ = timed_steps(lambda: (step5(), None)[1], 2)
        st5 = r.stats()
        out["cfg5_grid"] = {"value": segs5 / (sum(ms5) * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": sum(ms5) / 2, "steps": 2,
                            "n_spheres": len(big[2]), "spp": 256, "scene_generation_s": t_gen, "scene_generator": gen,
                            "grid_fallback_rays_per_step": st5.get("grid_fallback_rays"),
                            "grid_loose_cells_per_step": st5.get("grid_loose_cells"),
                            "work_model": grid_work_model(st5),
                            "equivalent_linear_sweep_T_instr_s": segs5 / 2 * len(big[2]) * FP32_INSTR_PER_TEST / (sum(ms5) / 2 * 1e-3) / 1e12,
                            "workload": f"BASELINE configs[4]: {len(big[2])} spheres, {W}x{R.image_height(W)}, 256 spp, depth {depth}, RTW_MODE_GRID"}
        # the same list through the LINEAR sweep (streamed tiles of 1024 spheres, TMA double buffer) on a 4-spp slice --
        # the full 256 spp would take ~2 minutes; a short slice also pays the drain of the persistent kernel (the last
        # 50-bounce paths keep whole warps busy), 16 spp reaches 0.83 -- against the FP32 roofline, and the grid image at
        # the same spp
        lspp = 4
        r.render_rows_device(cam, W, lspp, img.data_ptr(), max_depth=depth, seed=1, column_major=True, stream=stream.cuda_stream)
        stream.synchronize()
        sha_grid = image_sha256(img.cpu().numpy())
        r.set_option(R.RTW_OPT_MODE, R.RTW_MODE_FUSED)

        def step5l():
            r.render_rows_device(cam, W, lspp, img.data_ptr(), max_depth=depth, seed=1, column_major=True, stream=stream.cuda_stream)

        r.render_rows_device(cam, W, 1, img.data_ptr(), max_depth=depth, seed=1, column_major=True, stream=stream.cuda_stream)  # warm-up
        ms5l, segs5l, _ = timed_steps(lambda: (step5l(), None)[1], 1)
        st5l = r.stats()
        sha_lin = image_sha256(img.cpu().numpy())
        rate5 = st5l["sphere_tests"] * FP32_INSTR_PER_TEST / (st5l["ms_trace"] * 1e-3)
        out["cfg5_linear"] = {"value": segs5l / (sum(ms5l) * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": sum(ms5l), "steps": 1,
                              "n_spheres": len(big[2]), "spp": lspp, "T_fp32_instr_s": rate5 / 1e12,
                              "frac_of_measured_fp32_peak": rate5 / fp32_peak,
                              "extrapolated_s_at_256_spp": sum(ms5l) * 1e-3 * 256 / lspp,
                              "image_sha256": sha_lin, "image_sha256_grid_same_spp": sha_grid,
                              "workload": f"BASELINE configs[4] slice: {len(big[2])} spheres, {W}x{R.image_height(W)}, {lspp} of 256 spp, "
                                          f"depth {depth}, linear sweep (streamed tiles)"}
        if sha_lin != sha_grid:
            out["cfg5_linear"]["error"] = "grid and linear images differ"
    except Exception as e:
        out.setdefault("cfg5_grid", {"error": str(e)})
        out.setdefault("cfg5_linear", {"error": str(e)})
    finally:
        r.set_option(R.RTW_OPT_MODE, R.RTW_MODE_FUSED)
        r.set_scene(scene)
    try:  # Float64 instantiation (the reference's own test and published timings are Float64)
        R.reseed()
        s64 = R.flatten_scene(R.scene_random_spheres(elem_type=np.float64), np.float64)
        cam64 = R.t_cam1(np.float64)
        spp64 = min(args.spp, 100)
        r.render(cam64, W, 8, max_depth=depth, scene=s64)
        r.render(cam64, W, spp64, max_depth=depth, scene=s64)
        st = r.last_stats
        rate = st["sphere_tests"] * FP32_INSTR_PER_TEST / (st["ms_trace"] * 1e-3) / 1e12
        dfma_peak, _ = r.measure_fp32_peak(9)  # independent DFMA chains on every SM, measured live
        out["float64"] = {"value": st["ray_segments"] / (st["ms_trace"] * 1e-3) / 1e6, "unit": UNIT, "spp": spp64,
                          "ms_trace": st["ms_trace"], "T_fp64_instr_s": rate, "frac_of_nominal_dfma": rate / (148 * 64 * 1.965e9 / 1e12),
                          "measured_dfma_peak_T_instr_s": dfma_peak / 1e12, "frac_of_measured_dfma": rate / (dfma_peak / 1e12),
                          "workload": f"scene_random_spheres(Float64), t_cam1, {W}x{R.image_height(W)}, {spp64} spp, depth {depth}"}
    except Exception as e:
        out["float64"] = {"error": str(e)}
    try:  # the reference's own small renders through the host-buffer call rtw_render_scene (wall clock incl. the Python
        # binding, best of 50); they take the single-launch latency path (csrc/rtw_small.cu)
        lat = {}
        cases = [("scene_2_spheres_96x54x1", R.flatten_scene(R.scene_2_spheres()), R.t_default_cam(), 96, 1),
                 ("scene_2_spheres_96x54x16", R.flatten_scene(R.scene_2_spheres()), R.t_default_cam(), 96, 16),
                 ("random_spheres_96x54x1", scene, cam, 96, 1)]
        for name, sc, cm, w, s in cases:
            buf = np.empty((w, R.image_height(w), 3), dtype=np.float32)
            best = 1e9
            for _ in range(50):  # one rtw_render_scene per call: scene arrays, camera and image are host buffers
                t0 = time.perf_counter()
                r.render(cm, w, s, scene=sc, out=buf)
                best = min(best, time.perf_counter() - t0)
            lat[name] = best * 1e6
            lat[name + "_kernel_launches"] = r.last_stats["kernel_launches"]
        s64 = R.flatten_scene(R.scene_2_spheres(elem_type=np.float64), np.float64)
        cam64 = R.t_default_cam(np.float64)
        for name, s in (("scene_2_spheres_96x54x16_float64", 16), ("scene_2_spheres_96x54x1_float64", 1)):
            best = 1e9
            for _ in range(50):  # test/runtests.jl:190-194 as the reference runs it: Float64
                t0 = time.perf_counter()
                r.render(cam64, 96, s, scene=s64)
                best = min(best, time.perf_counter() - t0)
            lat[name] = best * 1e6
        lat["reference_us"] = {"scene_2_spheres_96x54x1": 101, "scene_2_spheres_96x54x16": 951, "random_spheres_96x54x1": 2040,
                               "source": "src/proto/proto.jl:87-89, :64-66, :142-144 (Ryzen 3700X, 16 threads, Float64 scenes)"}
        out["latency_us"] = lat
    except Exception as e:
        out["latency_us"] = {"error": str(e)}
    finally:
        r.set_scene(scene)
    return out


def hbm_peak_gbs():
    """Measured HBM copy bandwidth of this pool's B200s (driver-written MEASURED_PEAKS.json), else the profiling
    recipe's stated fallback."""
    try:
        d = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        for key in ("hbm_gbs", "hbm_GBps", "hbm"):
            if key in d:
                return float(d[key]), "MEASURED_PEAKS.json (measured)"
    except Exception:
        pass
    return 6650.0, "B200_PROFILING.md fallback (MEASURED_PEAKS.json absent)"


def measured_dram_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed `ncu --set full` capture
    (profiles/r02_trace_kernel_full_metrics.csv, else round 1's; per launch, 1920x1080 slice).  None when missing."""
    try:
        vals = {}
        path = ROOT / "profiles" / "r02_trace_kernel_full_metrics.csv"
        if not path.exists():
            path = ROOT / "profiles" / "r01_unified_trace_kernel_full_metrics.csv"
        for line in path.read_text().splitlines():
            parts = line.split(",")
            if len(parts) == 3 and parts[0].startswith("dram__bytes_"):
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(parts[1], None)
                if scale is not None:
                    vals[parts[0]] = float(parts[2]) * scale
        if "dram__bytes_read.sum" in vals and "dram__bytes_write.sum" in vals:
            return vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"]
    except Exception:
        pass
    return None


def main():
    args = parse_args()
    # keep stdout clean for the ONE JSON line: libraries (e.g. NCCL's version banner) print to fd 1, so everything
    # else is routed to stderr and the line is written to the real stdout
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
