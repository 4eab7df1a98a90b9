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
def cpu_sample(args, scene, cam, spp: int, threads: int = 0) -> dict:
    """Times the CPU oracle (the port of the reference's algorithm; Julia itself is not installed) on a bounded
    sample of the same workload: same scene/camera/width/depth, fewer samples per pixel (cost is linear in spp)."""
    from oracle import binding as O

    g, m, k = scene
    _, _, st = O.render(g, m, k, cam.as_array(), args.width, spp, max_depth=args.depth, seed=1, n_threads=threads)
    return {"value": st["ray_segments"] / st["seconds"] / 1e6, "unit": UNIT, "cores": st["threads"], "kind": "port",
            "sample": f"same scene/camera/{args.width}x{(args.width * 9) // 16}/depth {args.depth} at {spp} spp "
                      f"({st['paths']} paths, {st['ray_segments']} ray segments, {st['seconds']:.2f} s); "
                      f"C oracle, -O2, Philox stream, row-interleaved pthreads",
            "seconds": st["seconds"], "ray_segments": st["ray_segments"], "paths": st["paths"]}


def auto_cpu_spp(args) -> int:
    if args.cpu_spp > 0:
        return args.cpu_spp
    # ~1.1 Mrays/s per 8 cores measured in the build container => a few spp of 1080p is 10-30 s of CPU work
    cores = os.cpu_count() or 8
    pixels = args.width * ((args.width * 9) // 16)
    target_s = 15.0
    rays = target_s * 0.19e6 * cores  # ~0.19 Mrays/s per core on the 484-sphere scene; cost is linear in n_spheres
    if args.half_extent != 11:
        rays *= 484.0 / (4.0 * args.half_extent * args.half_extent)
    return max(1, min(args.spp, int(rays / (pixels * 4.1))))


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
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": "port", "sample": last["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- GPU arm
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
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

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

    # ---- end to end through the host-buffer API (scene H2D + image D2H inside the timed region)
    scene_bytes = int(sum(a.nbytes for a in scene)) + 88
    image_bytes = W * H * 3 * 4

    def one_step_e2e():
        if G == 1:
            r.render(cam, W, spp, max_depth=depth, seed=1, scene=scene, out=host_image.numpy())
            return r.last_stats["ray_segments"]
        r.set_scene(scene)  # H2D on every rank
        one_step_resident()
        if rank == 0:
            host_image.copy_(image, non_blocking=True)
        stream.synchronize()
        return r.stats()["ray_segments"]

    e2e_steps = max(1, min(args.steps, 2))
    barrier()
    t0 = time.perf_counter()
    e2e_segs = 0
    for _ in range(e2e_steps):
        e2e_segs += one_step_e2e()
    barrier()
    e2e_wall = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_wall], dtype=torch.float64, device=dev)
    e2e_s = torch.tensor([float(e2e_segs)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_s, op=dist.ReduceOp.SUM)
    e2e_value = float(e2e_s.item()) / float(e2e_t.item()) / 1e6

    # ---- extra (N = 1, default mode only): the same workload through RTW_MODE_GRID -- same image bits, reported beside
    # the headline, never instead of it (the benchmarked path is the reference's linear sweep)
    grid_extra = None
    if world == 1 and args.mode == "linear":
        try:
            r.set_option(R.RTW_OPT_MODE, R.RTW_MODE_GRID)
            one_step_resident()
            g_ms, g_segs, _ = timed_steps(lambda: (one_step_resident(), None)[1], 2)
            grid_extra = {"value": g_segs / (sum(g_ms) * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": sum(g_ms) / 2, "steps": 2,
                          "note": "RTW_MODE_GRID: uniform-grid traversal instead of the linear sweep, bit-identical image; "
                                  "not the benchmarked path (no linear-sweep roofline applies)"}
        except Exception as e:  # the headline must not depend on the optional mode
            grid_extra = {"error": str(e)}
        finally:
            r.set_option(R.RTW_OPT_MODE, R.RTW_MODE_FUSED)

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        c = cpu_sample(args, scene, cam, auto_cpu_spp(args))
        cpu = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        tests_per_step = segs_per_step * n_spheres
        trace_s = float(trace_ms.item()) * 1e-3  # dominant kernel: fused_trace_kernel, per launch (max over ranks)
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
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": scene_bytes * G,
                    "d2h_bytes_per_step": image_bytes, "steps": e2e_steps},
            # kernels of librtw_b200.so launched inside the timed region: per rank and step the u/v table kernel, the
            # fused trace kernel and the resolve kernel (rtw_stats.kernel_launches), plus one assemble on rank 0 for G > 1
            "gpu_launches": (int(stats["kernel_launches"]) * G + (1 if G > 1 else 0)) * args.steps,
            "roofline": {
                "bound": "fp32", "kernel": "fused_trace2_kernel",
                "achieved": achieved_instr, "peak": fp32_peak / 1e12,
                "unit": "T FP32 instr/s per GPU (11 per ray-sphere test)",
                "frac": achieved_instr / (fp32_peak / 1e12),
                "peak_source": "measured live: rtw_measure_fp32_peak variant 0 (independent FFMA chains, all SMs); "
                               "MEASURED_PEAKS.json holds no FP32 CUDA-core figure",
                "sweep_mix_ceiling": fp32_sweep_mix / 1e12,
                "achieved_tflops": tests_per_step * FLOP_PER_TEST / trace_s / 1e12 / G,
                # DRAM bytes per launch of fused_trace_kernel from the committed ncu --set full capture (a 1080p slice;
                # it is the 66 MB fixed-point accumulator, independent of spp) -- the kernel is FP32-bound, not HBM-bound
                "traffic": measured_dram_traffic(),
                "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu)",
                "algorithmic_hbm_bytes": W * H * 4 * 8 if G == 1 else None,
            },
            "clocks": clocks,
        }
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
                                "equivalent_linear_sweep_T_instr_s": achieved_instr, "traffic": None}
        if grid_extra is not None:
            line["grid_mode"] = grid_extra
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    r.close()
    if world > 1:
        dist.destroy_process_group()


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
    (profiles/r01_unified_trace_kernel_full_metrics.csv; per launch, 1920x1080 slice).  None when the file is missing."""
    try:
        vals = {}
        for line in (ROOT / "profiles" / "r01_unified_trace_kernel_full_metrics.csv").read_text().splitlines():
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
