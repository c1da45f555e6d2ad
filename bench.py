#!/usr/bin/env python
"""bench.py -- panorama pixels/s (and ray-steps/s) of the B200 ray march, one process per GPU.

    python bench.py --gpus N --steps K --warmup W [--workload c5] [--impl reference]

A "step" is one full render of the workload's panorama (stage A terrain profiles, stage B ray paths,
stage C march + colouring + metadata) with the packed terrain already resident in HBM. With N > 1
(torchrun) the panorama is sharded by contiguous column blocks, one block per rank; the total work is
fixed, so scaling is "strong". `value` = W*H / (max-over-ranks device time per step).
`e2e` is the same metric through the public host API with HOST buffers: every step uploads the decoded
DTED tiles from pinned memory (rank 0), broadcasts the packed terrain (NCCL), renders, gathers the
shards to rank 0 and reads RGB + per-pixel metadata back to pinned host memory.

`--impl reference` times the CPU restatement of the reference (oracle/, OpenMP on all host cores): the
reference itself is Rust and cannot be built in this image. Each step is a bounded
sample: every S-th column and row of the same panorama; the three stages are extrapolated separately
(terrain x S, paths x S, pixels x S^2) to the full job, which is what pixels/s is quoted on.
"""
import argparse
import json
import os
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "panorama pixels/s"
# Algorithmic work per unit (DESIGN.md "Kernels and rooflines"): FP64 instructions counted from the
# reference arithmetic each stage restates, not from the SASS.
STAGE_UNITS = {
    "terrain": ("k_terrain_profile", "terrain samples"),
    "paths": ("k_ray_paths_macro", "path steps"),
    "march": ("k_march", "ray steps"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=["c1", "c2", "c3_flat", "c3_sph", "c4", "c5"])
    ap.add_argument("--scale", type=float, default=1.0, help="image scale (1.0 = the BASELINE size; anything else is a dry run)")
    ap.add_argument("--march-mode", type=int, default=0, help="0 product default, 1 brute force (every step), 2 hierarchical march always")
    ap.add_argument("--cpu-stride", type=int, default=0, help="column/row stride of the bounded CPU sample (0 = auto)")
    ap.add_argument("--generator", default="Fast", choices=["Fast", "Rectilinear"],
                    help="output.generator of the workload (diagnosis; the BASELINE configs use the default Fast generator)")
    ap.add_argument("--emulate-ranks", type=int, default=0,
                    help="diagnosis on ONE GPU: render only rank 0's column block of an N-rank frame (what each GPU of an N-GPU run does, without the exchange)")
    ap.add_argument("--sweep-bands", type=int, default=0, help="row bands of the horizon sweep (0 = automatic; tuning)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def build_workload(name, scale, with_posts, generator="Fast"):
    from atm_raytracer_b200 import config, scenes
    from atm_raytracer_b200.terrain import Terrain  # no native library behind it: the reference arm never loads the CUDA .so

    cfg, grid = scenes.make_scene(name, scale=scale)
    cfg["output"]["generator"] = generator
    params = config.into_params(cfg)
    objects, textures = config.lower_objects(cfg)
    lat0, lon0, nlat, nlon = grid
    keys = [(lat0 + i, lon0 + j) for i in range(nlat) for j in range(nlon)]
    if with_posts:
        from atm_raytracer_b200 import synth

        with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
            posts = list(ex.map(lambda k: synth.make_tile(k[0], k[1], 1), keys))
        terrain = Terrain.from_arrays([(k[0], k[1], p) for k, p in zip(keys, posts)])
    else:
        shape = np.empty((1201, 1201), np.int16)
        terrain = Terrain([(Terrain.desc(k[0], k[1], shape), None) for k in keys])
    return cfg, params, terrain, objects, textures


def describe(name, params, terrain, extra=None):
    d = {
        "workload": f"{name}: {params.width}x{params.height} fov {params.fov:g} deg, "
                    f"{'straight' if params.straight_rays else 'refracted (US-76)'} rays, "
                    f"{'FlatDistorted' if params.earth_model == 1 else 'Spherical R=%g km' % (params.radius / 1e3)}, "
                    f"{len(terrain.tiles)} synthetic DTED L1 tiles, max_distance {params.max_distance / 1e3:g} km, step {params.simulation_step:g} m",
        "generator": "Rectilinear" if params.generator == 1 else "Fast",
        "l2_policy": "inputs larger than L2 (profile caches >> 126 MB are rewritten and re-read every step)" if params.generator == 0
                     else "L2 flushed by the render itself only when the terrain exceeds it; the Rectilinear generator keeps no caches",
    }
    d.update(extra or {})
    return d


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------
# CPU baseline (oracle; rank 0 only)
# ---------------------------------------------------------------------------------------------
def cpu_sample(params, terrain, objects, textures, stride):
    """One bounded CPU sample; returns (pixels/s of the full job, ray-steps/s, description, timing)."""
    import oracle

    oracle.use_native()  # the `baseline` build of the port (-O3 -march=native, compiled on this host), as BASELINE.md states
    # all host cores, whatever OMP_NUM_THREADS says (torchrun sets it to 1 for every rank)
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    r = oracle.render(params, terrain.tiles, objects, textures, stride_x=stride, stride_y=stride, meta=True, steps=False, threads=threads)
    threads = oracle.num_threads()
    tm = r["timing"]
    # the reference's three stages scale differently with the image: extrapolate each one
    t_full = tm["s_terrain"] * stride + tm["s_paths"] * stride + tm["s_pixels"] * stride * stride
    wl = params.x1 - params.x0
    pixels = wl * params.height
    steps_full = r["stats"]["ray_steps"] * stride * stride
    desc = (f"every {stride}th column and row ({r['rgb'].shape[1]}x{r['rgb'].shape[0]} px), all three stages; "
            f"stage times extrapolated to the full image (terrain x{stride}, paths x{stride}, pixels x{stride * stride}); "
            f"sample took {tm['s_total']:.2f} s")
    desc += f"; oracle build: {oracle.BUILD}"
    return pixels / t_full, steps_full / t_full, desc, tm, threads


def auto_stride(params):
    # aim at ~10-30 s of CPU work: stage A dominates the sample (~2.5 us per terrain sample per core)
    n_t = params.max_distance / params.simulation_step
    cores = os.cpu_count() or 1
    for s in (1, 2, 4, 8, 16, 32, 64, 128):
        cols = (params.x1 - params.x0) / s
        est = cols * n_t * 2.5e-6 / cores + (cols * params.height / s) * n_t * 2.0e-9 / cores
        if params.generator == 1:  # Rectilinear: one RK4 step + one TerrainData per ray step (~0.7 us per step per core measured)
            est = (cols * params.height / s) * n_t * 0.7e-6 / cores
        if est < 12.0:
            return s
    return 128


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg, params, terrain, objects, textures = build_workload(args.workload, args.scale, True, args.generator)
    stride = args.cpu_stride or auto_stride(params)
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_sample(params, terrain, objects, textures, stride * 2)
    vals, steps_rate, desc, threads = [], [], "", 1
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, sr, desc, tm, threads = cpu_sample(params, terrain, objects, textures, stride)
        vals.append(v)
        steps_rate.append(sr)
    wall = time.perf_counter() - t0
    value = float(np.mean(vals))
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pixels/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * params.width * params.height / value, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # the configuration of the GPU arm for the same flags, key for key (the driver compares the two arms' `config`)
        "config": describe(args.workload, params, terrain, {"parallelism": f"column blocks x{max(1, args.gpus)}"}),
        "ray_steps_per_s": float(np.mean(steps_rate)),
        "cpu_baseline": {"value": value, "unit": "pixels/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "pixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of the reference (C++/OpenMP oracle); the Rust reference cannot be built here. "
                "ms_per_step is the extrapolated full-image time; wall time of the sampled steps was %.1f s" % wall,
    }
    emit(out)


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    from atm_raytracer_b200 import parallel, runtime

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    # host-side rendezvous for the end-to-end leg: a rank waiting in an NCCL barrier keeps a spinning kernel on its
    # GPU, which the one process that drives all N GPUs there would have to time-slice with
    standby_group = dist.new_group(backend="gloo") if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def standby():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=standby_group)

    cfg, params, terrain, objects, textures = build_workload(args.workload, args.scale, rank == 0, args.generator)
    my = parallel.shard_params(params, rank, world)
    if args.emulate_ranks > 1 and world == 1:
        my = parallel.shard_params(params, 0, args.emulate_ranks)
    wl, H, W = my.x1 - my.x0, params.height, params.width

    ctx = runtime.Context(local)
    ctx.set_march_mode(args.march_mode)
    ctx.set_sweep_bands(args.sweep_bands)
    # terrain: rank 0 uploads + retiles, everyone receives the packed copy over NCCL (once, untimed)
    nbytes = ctx.packed_bytes(terrain)
    packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    if rank == 0:
        ctx.pack_terrain(terrain, packed.data_ptr())
    parallel.broadcast_terrain(packed)
    ctx.bind_terrain(terrain if rank == 0 else terrain.descriptors_only(), packed.data_ptr())
    ctx.set_params(my)
    ctx.set_objects(objects, textures)

    rgb = torch.empty((H, wl, 3), dtype=torch.uint8, device=dev)
    meta = torch.empty((H, wl, 4), dtype=torch.float64, device=dev)
    # a dedicated (non-default) stream: the library forks its stage streams from it and joins back,
    # so CUDA events recorded on it bracket every kernel of a render
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def step():
        ctx.render_device(rgb.data_ptr(), meta.data_ptr(), 0, stream)

    for _ in range(max(args.warmup, 0)):
        step()
    barrier()
    ctx.stage_times()  # drop the warm-up renders from the stage averages
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1) / args.steps
    kernel_ms = ctx.kernel_times()  # per hot kernel, CUDA events on the stream it is launched on, averaged over the timed steps
    stage = ctx.stage_times()
    t = torch.tensor([ms, stage["ms_terrain"], stage["ms_paths"], stage["ms_march"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_a, ms_b, ms_c = t.tolist()

    # one extra (untimed) render for the counters
    st = ctx.render_device(rgb.data_ptr(), meta.data_ptr(), 0, stream, want_stats=True)
    st = parallel.reduce_stats(st, dev)
    launches_per_step = st["kernel_launches"]
    fp = ctx.fp64_peak() if rank == 0 else None

    # ---- e2e: host buffers in, host buffers out ------------------------------------------------
    # What `gen` does for this workload: decoded DTED tiles in host memory -> device (H2D + retile),
    # broadcast, render, gather to rank 0, image (and, when the workload says --output-meta, the
    # per-pixel metadata) back to host memory. Reported twice: as the workload's CLI line produces it
    # (`e2e`) and with the 32 B/pixel metadata always read back (`e2e_with_meta`).
    e2e = e2e_meta = e2e_gen = None
    if not args.no_e2e:
        # What `atm-raytracer gen --gpus N` does between the decoded tiles and the image: ONE process drives the N GPUs of
        # the box through the library's group API (atmrt_group_*: one context and one host thread per GPU). Every step
        # uploads the decoded DTED tiles from page-locked host memory -- each GPU a slice over its own PCIe link, the
        # slices all-gathered over NVLink -- retiles, renders the column blocks and lands every block in the host's
        # row-major image (and metadata) over each GPU's own link. Under torchrun rank 0 is that process; the other
        # ranks stand by at a host-side (gloo) barrier with their GPUs idle.
        del rgb, meta
        torch.cuda.empty_cache()
        barrier()
        standby()
        if rank == 0:
            group = runtime.Group(world)
            tiles_pinned = []
            for d, posts in terrain.tiles:
                a = runtime.host_array(posts.shape, np.int16)
                a[...] = posts
                tiles_pinned.append((d, a))
            terrain_pinned = runtime.Terrain(tiles_pinned)
            group.set_params(params)
            group.set_objects(objects, textures)
            host = {"rgb": runtime.host_array((H, W, 3), np.uint8), "meta": runtime.host_array((H, W), runtime.META_DTYPE)}

            def e2e_time(with_meta):
                e2e_steps = max(1, min(args.steps, 3))

                def one():
                    # H2D of the decoded tiles (sliced) + retile + all-gather, the render and the image's way back: one call
                    group.render(rgb=True, meta=with_meta, steps=False, out=host, terrain=terrain_pinned)

                one()
                t0 = time.perf_counter()
                for _ in range(e2e_steps):
                    one()
                dt = (time.perf_counter() - t0) / e2e_steps
                return {"value": W * H / dt, "unit": "pixels/s", "h2d_bytes_per_step": int(terrain.bytes),
                        "d2h_bytes_per_step": int(W * H * 3 + (W * H * 32 if with_meta else 0)), "ms_per_step": dt * 1e3, "steps": e2e_steps,
                        "api": f"atmrt_group_render_tiles on {world} GPU(s) from one process (the path of `atm-raytracer gen --gpus {world}`): "
                               "decoded tiles in page-locked host memory in, rgb" + (" and per-pixel metadata" if with_meta else "") + " in page-locked host memory out"}

            wants_meta = bool(cfg["output"].get("file_metadata"))
            e2e_meta = e2e_time(True)
            e2e = e2e_meta if wants_meta else e2e_time(False)
            assert host["rgb"].any()
            group.close()
            if world == 1:
                e2e_gen = time_gen_executable()
        standby()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -------------------------------------------------------
    # EXECUTED work against the measured peaks: FP64-pipe instructions of one launch (counted on the SASS page of the
    # committed ncu capture of these very sources) / the live duration of that launch, against the live DFMA issue rate;
    # DRAM bytes of one launch (same capture) / live duration against MEASURED_PEAKS.json's copy bandwidth. What the
    # REFERENCE would execute for the same units is reported separately, as `algorithmic`.
    n_t = st["n_terrain"]
    units = {"terrain": wl * n_t, "paths": st["path_steps"] / world if world > 1 else st["path_steps"], "march": st["ray_steps"] / world}
    roofs = {k: kernel_roofline(k, v, fp, args, world) for k, v in kernel_ms.items()}
    dom = max(kernel_ms, key=kernel_ms.get)
    roof = dict(roofs[dom])
    stage_of = {"k_terrain_profile": "terrain", "k_ray_paths_macro": "paths", "k_sweep_bits": "march", "k_hit_normals": "march",
                "k_shade_tiles": "march", "k_march": "march", "k_rectilinear": "march"}
    stage_ms = {"terrain": ms_a, "paths": ms_b, "march": ms_c}
    algorithmic = {}
    for sname, t_ms in stage_ms.items():
        per_unit = fp64_instr_per_unit(sname, params)
        rate = per_unit * units[sname] / (t_ms * 1e-3) / 1e9 if t_ms > 0 else 0.0
        peak = (fp or {}).get("dfma_gflops", 0.0) / 2.0
        algorithmic[sname] = {"reference_fp64_instr_per_unit": per_unit, "units_per_step": units[sname], "unit_name": STAGE_UNITS[sname][1],
                              "reference_equivalent_ginstr_per_s": rate, "algorithmic_speedup": rate / peak if peak else None,
                              "hbm_algorithmic_bytes_per_unit": hbm_bytes_per_unit(sname, params)}
    roof["stage"] = stage_of.get(dom)
    roof["algorithmic"] = algorithmic.get(stage_of.get(dom))

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        stride = args.cpu_stride or auto_stride(params)
        v, sr, desc, tm, threads = cpu_sample(params, terrain, objects, textures, stride)
        cpu = {"value": v, "unit": "pixels/s", "cores": threads, "kind": "port", "sample": desc, "ray_steps_per_s": sr,
               "stage_s_in_sample": {k: tm[k] for k in ("s_terrain", "s_paths", "s_pixels")}}

    out = {
        "metric": METRIC, "value": W * H / (ms * 1e-3), "unit": "pixels/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": describe(args.workload, params, terrain, {"parallelism": f"column blocks x{world}"}),
        "march_mode": {0: "horizon sweep (opaque terrain, no objects) / crossing march (translucent terrain, objects)", 1: "brute force", 2: "hierarchical march"}[args.march_mode],
        "ray_steps_per_s": st["ray_steps"] / (ms * 1e-3),
        "ray_steps_per_step": st["ray_steps"],
        "stage_ms": {"terrain_profile": ms_a, "ray_paths": ms_b, "march": ms_c, "note": "terrain and paths overlap on two streams; max over ranks"},
        "clocks": clocks.summary(),
        "e2e": e2e,
        "e2e_with_meta": e2e_meta,
        "e2e_gen": e2e_gen,
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roof,
        "roofline_kernels": roofs,
        "algorithmic": algorithmic,
        "kernel_ms": kernel_ms,
        "cpu_baseline": cpu,
        "fp64_peak_measured": fp,
        "pixels_hit": st["pixels_hit"], "step_overflows": st["step_overflows"],
    }
    emit(out)
    if world > 1:
        dist.destroy_process_group()


def time_gen_executable():
    """Wall time of the `atm-raytracer gen` executable on BASELINE config 2 (1920x1080, 2x2 DTED tiles on disk, with
    --output-meta): process start, CUDA initialisation, DTED decode, upload, render, PNG encode and the metadata sidecar --
    everything SURVEY section 8d(ii) counts as end-to-end `gen`. The executable's own stage stamps are reported next to it."""
    import re
    import subprocess
    import tempfile

    from atm_raytracer_b200 import host, synth

    with tempfile.TemporaryDirectory() as tmp:
        folder = os.path.join(tmp, "terrain")
        os.mkdir(folder)
        synth.write_tile_grid(folder, 45, 5, 2, 2, level=1)
        png, dat = os.path.join(tmp, "out.png"), os.path.join(tmp, "out.dat")
        argv = [host.EXECUTABLE, "gen", "-t", folder, "-l", "45.05", "-g", "6.0", "-a", "1800", "-d", "0", "-f", "10", "-m", "200", "--step", "50",
                "-w", "1920", "-h", "1080", "--output", png, "--output-meta", dat]
        best, stamps = None, {}
        for _ in range(2):
            t0 = time.perf_counter()
            r = subprocess.run(argv, capture_output=True, text=True)
            dt = time.perf_counter() - t0
            if r.returncode != 0:
                return {"error": r.stderr.strip()[-300:]}
            if best is None or dt < best:
                best = dt
                stamps = {m.group(2).strip(" ."): float(m.group(1)) for m in re.finditer(r"^([0-9.]+): ([A-Za-z ]+)", r.stdout, re.M)}
        return {"workload": "c2 via `atm-raytracer gen` (CLI flags, DTED L1 files on disk -> PNG + metadata sidecar)", "wall_s": best,
                "pixels_per_s": 1920 * 1080 / best, "stage_stamps_s": stamps, "png_bytes": os.path.getsize(png), "meta_bytes": os.path.getsize(dat)}


def _source_sha():
    sys.path.insert(0, os.path.join(ROOT, "profiles"))
    try:
        from source_sha import source_sha

        return source_sha()
    except Exception:
        return None


def kernel_roofline(kernel, ms, fp, args, world):
    """`roofline` of one kernel. `bound` is the FP64 pipe: the path is f64 arithmetic on data that is read once (SURVEY
    section 8d), no dense contraction. achieved = executed FP64-pipe thread instructions per launch / live launch duration;
    peak = the DFMA issue rate measured live (atmrt_fp64_peak; MEASURED_PEAKS.json has no FP64 entry); frac = achieved /
    peak. The executed counts come from profiles/ncu_summary.json and are used only when that capture was taken from the
    sources this run executes (source_sha) on this workload at N = 1; otherwise frac is null and the entry says why."""
    hbm_peak = 6548.2
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    peak = (fp or {}).get("dfma_gflops", 0.0) / 2.0  # G FP64 thread instr/s
    out = {"kernel": kernel, "bound": "fp64", "launch_ms": ms, "peak": peak, "unit": "G FP64 thread instr/s", "achieved": None, "frac": None, "traffic": None,
           "peak_source": "measured live (atmrt_fp64_peak: 8 independent DFMA chains per thread); MEASURED_PEAKS.json has no FP64 entry",
           "hbm": {"peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs", "achieved_gbs": None, "frac": None}}
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
    except Exception:
        table = {}
    e = table.get(f"{args.workload}:{kernel}")
    if e is None and kernel == "k_march":
        # translucent terrain / objects: the march bracket holds the crossing march (k_thresholds + k_cross_march)
        parts = [table.get(f"{args.workload}:{k}") for k in ("k_cross_march", "k_thresholds")]
        if all(parts) and len({q.get("source_sha") for q in parts}) == 1:
            e = dict(parts[0])
            for f in ("fp64_warp_instructions", "warp_instructions", "dram_read_bytes", "dram_write_bytes", "gpu_time_ms"):
                e[f] = sum(q[f] for q in parts)
            out["kernel"] = "k_thresholds + k_cross_march"
    sha = _source_sha()
    if e is None or args.scale != 1.0 or world != 1 or args.emulate_ranks > 1 or args.generator != "Fast":
        out["note"] = "no ncu capture of this kernel for this workload / shard under profiles/"
        return out
    if e.get("source_sha") != sha:
        out["note"] = f"profiles/ncu_summary.json was captured from other sources (sha {e.get('source_sha')}, running {sha}): executed counts withheld"
        out["stale_ncu"] = {k: e.get(k) for k in ("gpu_time_ms", "fp64_pipe_pct", "issue_active_pct", "source")}
        return out
    fp64_thread = e["fp64_warp_instructions"] * 32.0
    dram = e["dram_read_bytes"] + e["dram_write_bytes"]
    out["achieved"] = fp64_thread / (ms * 1e-3) / 1e9 if ms > 0 else None
    out["frac"] = out["achieved"] / peak if peak and out["achieved"] is not None else None
    out["traffic"] = dram
    out["hbm"].update(achieved_gbs=dram / (ms * 1e-3) / 1e9 if ms > 0 else None)
    out["hbm"]["frac"] = out["hbm"]["achieved_gbs"] / hbm_peak if out["hbm"]["achieved_gbs"] else None
    out["issue"] = {"warp_instructions_per_launch": e["warp_instructions"],
                    "achieved_g_warp_instr_per_s": e["warp_instructions"] / (ms * 1e-3) / 1e9 if ms > 0 else None}
    out["ncu"] = {k: e.get(k) for k in ("gpu_time_ms", "fp64_pipe_pct", "issue_active_pct", "warps_active_pct", "dram_throughput_pct", "registers",
                                        "top_stalls_warps_per_issue_cycle", "source", "source_sha")}
    out["note"] = ("frac = FP64-pipe instructions this kernel EXECUTES per launch (SASS page of the ncu capture of these sources) / live launch "
                   "duration / live DFMA issue rate; `ncu.fp64_pipe_pct` is the hardware counter of the same capture (kernel alone, cold cache)")
    return out


def fp64_instr_per_unit(stage, params):
    # DESIGN.md, "Kernels and rooflines": arithmetic the reference performs per unit, transcendentals
    # costed at the instruction count of CUDA's f64 libm paths.
    if stage == "march" and params.generator == 1:
        # Rectilinear: every ray step is one RK4 step + one TerrainData::from_lat_lon + the sign test
        return (25.0 if params.straight_rays else 2600.0) + (1500.0 if params.earth_model == 0 else 700.0) + 3.0
    if stage == "march":
        return 3.0  # 1 DADD (diff), 1 DMUL (product), 1 DSETP (sign test) per ray step (utils.rs:220-222)
    if stage == "paths":
        return 25.0 if params.straight_rays else 2600.0
    return 1500.0 if params.earth_model == 0 else 700.0


def hbm_bytes_per_unit(stage, params):
    if stage == "march" and params.generator == 1:
        return 40.0  # five bilinear taps of 4 i16 posts per ray step as the reference computes it (8 B here: normals at hits only)
    if stage == "march":
        return 8.0 * (1.0 / params.height + 1.0 / max(1, params.x1 - params.x0))
    if stage == "paths":
        return 24.0  # PathElem{dist, elev, path_length} per step as the reference stores it (16 B here: dist is row-independent)
    return 48.0 + 40.0  # six f64 outputs + five bilinear taps of 4 i16 posts


def emit(obj):
    """The one JSON line, on the process's ORIGINAL stdout. File descriptor 1 itself is pointed at stderr for
    the whole run (see below), so that nothing a library prints from native code (NCCL's version banner at
    communicator creation, whatever NCCL_DEBUG the box sets) can land on stdout next to it."""
    _JSON_OUT.write(json.dumps(obj) + "\n")
    _JSON_OUT.flush()


if __name__ == "__main__":
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
