#!/usr/bin/env python
"""bench.py — paths/s (depth 4) and ms/frame at 1080p on the named BASELINE.json workload.

  python bench.py --gpus N --steps K --warmup W            # CUDA arm (libvoxelrt, sm_100a)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle restatement of
                                                           # the reference path on the host cores

Workload (config.workload = "config3"): BASELINE.json configs[2] — synthetic dense random 256^3
voxel grid (~50 % occupancy), 1920x1080, depth 4, physical sky + clouds, sun (1,1,1); this is the
configuration the north-star target (>= 1 Gpaths/s per B200) is quoted on and it fits one GPU.
A "step" is one accumulate() batch of --spp samples per pixel over the whole frame (default 64: config 3 is quoted at ">= 64 spp
per timing"; until profiles/r04l the default was 8, which measures 3.7 % lower because the per-launch costs weigh more).

N > 1 (launched by torch.distributed.run, one rank per GPU): sample sharding — rank r renders
sample indices r, r+N, ... of every pixel (weak scaling: per-GPU work fixed). Per step the partial
accumulation buffers are merged by the fused reduce-scatter + tonemap kernel over NVLink peer
memory (voxel_rt2_b200/parallel.py FusedMerge: every rank merges 1/N of the pixels, reading the
peers' buffers through CUDA-IPC mappings, and stores the tonemapped pixels into rank 0's image
buffer); the only NCCL collective left is a 4-byte all-reduce used as a stream-ordered barrier.
Batches alternate between two accumulation slots so batch k+1 renders while batch k is merged.

Timing: W untimed warm-up steps, then exactly K steps between barrier + synchronize, CUDA events
on the launching stream, max over ranks. The sky tables (2 x 236 MB as float4) exceed the 126 MB
L2 and every step touches ~100 MB of colour/occupancy/accumulation data, so inputs are larger than
L2 (config.l2 = "inputs>L2").

After the timed region (and outside it) the line gains:
  parity         per-pixel agreement of the GPU with the CPU oracle at FULL size (1920x1080, 256^3,
                 same 48 sample indices: the oracle image is the one the cpu_baseline leg renders anyway)
  other_configs  short legs of BASELINE configs 1, 2, 4 (N = 1) and 5 (N = 8), each with its own clocks
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SUN = ((1, 1, 1), 0.025, (1.0 * 1.3, 0.949 * 1.3, 0.937 * 1.3))  # example6.py:10 colour, default direction
GOLD = os.path.join(ROOT, "tests", "golden")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--spp", type=int, default=64,
                    help="samples per pixel per step (per GPU). 64 = the batch BASELINE.json config 3 names (\">= 64 spp per timing\") and the default "
                         "VRT_SPP of Scene.finish; one k_path launch per step. Measured device rate by batch: 8 spp 3.50, 16 spp 3.56, 32 spp 3.58, "
                         "64 spp 3.63 G paths/s (profiles/r04l_spp_per_launch.log): per-launch set-up, the tail of the tile queue and the one "
                         "accumulation read-modify-write per pixel are amortised over more samples")
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--res", default="1920x1080")
    ap.add_argument("--sky-res", type=int, default=3840)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the config 1/2/4/5 legs after the timed region")
    ap.add_argument("--merge", default="fused", choices=["fused", "allreduce"],
                    help="N > 1: fused = peer-memory reduce-scatter + tonemap (default); allreduce = one NCCL all-reduce of the buffer per step")
    ap.add_argument("--workload", default="config3", choices=["config3", "config2", "config4", "config4_example3"],
                    help="config3 (default, the bench line): dense random grid; config2: example6 fixture scene, sky + clouds; "
                         "config4: ReSTIR mode (render + temporal + spatial resampling per frame) on the example6 scene; "
                         "config4_example3: the same on example3 (emissive ceiling, black sun)")
    ap.add_argument("--cpu-spp", type=int, default=48, help="samples per pixel of the bounded CPU-baseline sample")
    return ap.parse_args()


# ----------------------------------------------------------------------------- workloads
class Workload:
    """Scene + renderer settings of one BASELINE.json configuration."""

    def __init__(self, name, R):
        self.name, self.R = name, R
        self.restir = name.startswith("config4")
        self.sky = True
        self.exposure = 2.0
        self.voxel_edges = 0.06
        if name == "config3":
            import scenes

            self.mat, self.col = scenes.random_grid(R, 0.5, 1234)
            self.desc = "config3: dense random %d^3 (50%% occupancy), {W}x{H}, depth 4, physical sky + clouds, sun (1,1,1)" % R
        elif name in ("config2", "config4"):
            z = np.load(os.path.join(GOLD, "example6_seed0.npz"))  # example6.py run through the shim, seed 0
            self.mat, self.col, self.R = z["material"], z["color"], 128
            self.voxel_edges = 0.0
            self.desc = ("config2: example6 scene (shim seed 0) 128^3, {W}x{H}, depth 4, physical sky + clouds" if name == "config2" else
                         "config4: ReSTIR mode (render + temporal reuse + spatial GRIS 32 taps per frame) on the example6 scene 128^3, {W}x{H}")
        elif name == "config4_example3":
            z = np.load(os.path.join(GOLD, "example3_seed0.npz"))  # example3.py run through the shim
            self.mat, self.col, self.R = z["material"], z["color"], 128
            self.sky, self.exposure, self.voxel_edges = False, 30.0, 0.0
            self.desc = "config4: ReSTIR mode (render + temporal reuse + spatial GRIS 32 taps per frame) on the example3 scene (emissive ceiling, black sun) 128^3, {W}x{H}"
        else:
            raise ValueError(name)

    def configure(self, r, sky=None):
        sky = self.sky if sky is None else (sky and self.sky)
        r.set_voxels(self.mat, self.col)
        if self.name == "config3":
            r.set_floor(-1e5, (1.0, 1.0, 1.0))  # floor disabled as example9.py:4 does
            r.set_directional_light(*SUN)
        elif self.name == "config4_example3":
            r.set_floor(0.0, (1.0, 1.0, 1.0))                      # example3.py:7
            r.set_directional_light((1, 1, 1), 0.1, (0.0, 0.0, 0.0))  # scene.py:127 default: black sun
        else:
            r.set_floor(-0.85, (1.0, 1.0, 1.0))                   # example6.py:8
            r.set_directional_light((1, 1, -1), SUN[1], SUN[2])    # example6.py:10
        if sky:
            r.set_use_physical_sky(True, True)


class ClockSampler:
    """SM clock and throttle reasons sampled during a timed region. In-process NVML (two cheap queries every 50 ms from a
    thread); an `nvidia-smi -lms 50` child — the obvious way — is NOT used: its periodic multi-field query stalls kernel
    launches on the GPU it watches and cost the launch-heavy ReSTIR legs a factor 2-3 (2.14 -> 3.85-6.35 ms/frame at
    N = 4, profiles/r03h_clock_sampler_perturbation.log) and the 5 ms path-kernel steps ~0.8 %."""

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.t, self.err = index, [], threading.Event(), None, None

    def start(self):
        if os.environ.get("VRT_BENCH_NO_CLOCKS"):  # diagnostic: is the sampler itself perturbing a leg?
            self.err = "disabled by VRT_BENCH_NO_CLOCKS"
            return self
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def loop():
                while not self.stop_flag.is_set():
                    try:
                        self.samples.append((float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                                             int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))))
                    except Exception as e:  # noqa: BLE001
                        self.err = repr(e)
                        return
                    self.stop_flag.wait(0.05)

            self.pynvml = pynvml
            self.t = threading.Thread(target=loop, daemon=True)
            self.t.start()
            t0 = time.time()
            while not self.samples and self.err is None and time.time() - t0 < 2.0:
                time.sleep(0.005)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
        return self

    def stop(self):
        if self.t is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampler unavailable: %s" % self.err], "samples": 0}
        self.stop_flag.set()
        self.t.join(timeout=1.0)
        nv = self.pynvml
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted(nm for nm, bit in names.items() if any(r & bit for _, r in self.samples))
        sm = [c for c, _ in self.samples]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm),
                "how": "in-process NVML, 50 ms period"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def kernel_source_digest():
    """sha256 over the sources of the path kernel: ties a committed ncu traffic figure to the kernel it was taken from."""
    h = hashlib.sha256()
    for f in ("vrt_render.cu", "vrt_trace.cuh", "vrt_bsdf.cuh", "vrt_sky.cuh", "vrt_common.cuh"):
        h.update(open(os.path.join(ROOT, "voxel_rt2_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def measured_traffic(spp, sky_res, sky_format):
    """DRAM bytes per k_path launch from the committed ncu capture (profiles/k_path_traffic.json), or None when
    the capture was taken from another kernel source / another launch shape: a stale constant is worse than none."""
    tp = os.path.join(ROOT, "profiles", "k_path_traffic.json")
    if not os.path.exists(tp):
        return None, "no capture committed"
    tj = json.load(open(tp))
    if tj.get("spp") != spp or tj.get("sky_res") != sky_res or tj.get("sky_format", "f32") != sky_format:
        return None, "capture is for another launch shape"
    if tj.get("kernel_source_digest") != kernel_source_digest():
        return None, "capture predates the current kernel sources"
    return tj.get("dram_bytes_per_launch"), "ncu --set full capture %s" % tj.get("report", "")


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def oracle_threads(lib):
    """torch.distributed.run exports OMP_NUM_THREADS=1; the CPU legs use every core the process may run on."""
    lib.orc_set_num_threads(host_cores())
    return int(lib.orc_num_threads())


def cpu_sample(args, wl, W, H, sky_tables, cpu_spp):
    """Oracle on all host cores over a bounded sample of the same workload: the full 1080p frame,
    same camera / scene / sky tables (downloaded from the GPU: the 3840^2 precompute is ~1e12
    inner iterations, not a CPU job), `cpu_spp` samples per pixel. Returns the image as well: the
    GPU renders the same sample indices for the full-size parity figure."""
    from oracle.binding import OracleRenderer, load
    from voxel_rt2_b200.materials import material_table

    lib = load()
    cores = oracle_threads(lib)
    o = OracleRenderer(dx=2.0 / wl.R, image_res=(W, H), grid_res=wl.R, sky_res=args.sky_res if sky_tables else 0, exposure=wl.exposure, seed=1,
                       voxel_edges=wl.voxel_edges, materials=material_table())
    wl.configure(o, sky=sky_tables is not None)
    if sky_tables is not None:
        o.set_sky_tables(*sky_tables)
    o.prepare_data()
    o.accumulate(cpu_spp, stats=True)
    c = o.counters()
    ms = o.last_ms()
    return {"paths": c["paths"], "ms": ms, "paths_per_s": c["paths"] / (ms * 1e-3), "counters": c, "cores": cores,
            "spp": cpu_spp, "hdr": o.fetch_hdr()}


def parity_figures(a, b):
    """GPU image a vs oracle image b (float4 [H, W, 4] means): SURVEY.md §8c rel-RMSE on the linear HDR buffer and
    the fraction of pixels within 1e-3 relative (same sampler on both sides => per-pixel comparison)."""
    from util import rel_rmse

    err = np.abs(a[..., :3] - b[..., :3]).max(axis=-1)
    scale = np.maximum(np.abs(b[..., :3]).max(axis=-1), 1e-3)
    return {"rel_rmse": rel_rmse(a, b), "frac_within_1e-3": float(np.mean(err <= 1e-3 * scale + 1e-5)),
            "worst_rel": float((err / scale).max()), "mean_ratio": float(a[..., :3].mean() / max(b[..., :3].mean(), 1e-12)),
            "pixels": int(a.shape[0] * a.shape[1]), "samples_equal": bool((a[..., 3] == b[..., 3]).all())}


def _upsample_table(tab, S):
    """Bilinear resampling (wrap-around, texel centres) of an s x s x 3 sky table to S x S x 3."""
    s = tab.shape[0]
    x = (np.arange(S) + 0.5) * s / S - 0.5
    i0 = np.floor(x).astype(np.int64)
    f = (x - i0).astype(np.float32)
    a = tab[i0 % s] * (1.0 - f)[:, None, None] + tab[(i0 + 1) % s] * f[:, None, None]
    return np.ascontiguousarray(a[:, i0 % s] * (1.0 - f)[None, :, None] + a[:, (i0 + 1) % s] * f[None, :, None], dtype=np.float32)


def run_reference(args, rank, world):
    """CPU arm: the oracle restatement (the reference itself is Taichi/Vulkan and cannot run here)."""
    if rank != 0:
        return
    W, H = [int(x) for x in args.res.split("x")]
    wl = Workload(args.workload, args.grid)
    R = wl.R
    from oracle.binding import OracleRenderer, load
    from voxel_rt2_b200.materials import material_table

    lib = load()
    cores = oracle_threads(lib)  # all host cores, whatever OMP_NUM_THREADS the launcher exported
    # sky tables: the CPU cannot run the 3840^2 precompute in bounded time (~1e12 inner iterations), so
    # the oracle computes a 64^2 table itself and it is resampled to the GPU arm's table size: every
    # lookup then walks the same 2 x 177 MB footprint (same cache behaviour) through the same code.
    S = 64 if wl.sky else 0
    o = OracleRenderer(dx=2.0 / R, image_res=(W, H), grid_res=R, sky_res=S, cloud_passes=2, exposure=wl.exposure, seed=1,
                       voxel_edges=wl.voxel_edges, materials=material_table())
    wl.configure(o)
    if wl.restir:
        o.set_restir_temporal(True)
    o.prepare_data()
    if wl.sky and args.sky_res > S:
        tabs = [_upsample_table(t, args.sky_res) for t in o.get_sky_tables()]
        S = o.sky_res = args.sky_res
        o.set_sky_tables(*tabs)
    # each step = 1 spp over every 4th 8x4 tile (1/4 of the frame) so K+W steps end in minutes
    n = 4
    o.set_tile_shard(0, n)
    times, paths = [], 0
    for i in range(args.warmup + args.steps):
        before = o.counters()["paths"]
        if wl.restir:
            o.set_tile_shard(0, 1)  # the resampling passes read neighbours: whole frames
            o.accumulate_restir(1)
            done = W * H
        else:
            o.accumulate(1, stats=True)
            done = o.counters()["paths"] - before
        if i >= args.warmup:
            times.append(o.last_ms())
            paths += done
    tot = sum(times) * 1e-3
    v = paths / tot
    sample = "whole 1080p frame per step (ReSTIR passes read neighbours)" if wl.restir else "every 4th 8x4 tile of the frame per step (1/4 frame, 1 spp)"
    line = {"impl": "reference", "metric": "paths_per_sec_depth4_1080p", "value": v, "unit": "paths/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / max(len(times), 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.desc.format(W=W, H=H), "spp_per_step": 1, "sample": sample, "sky_res": S},
            "cpu_baseline": {"value": v, "unit": "paths/s", "cores": cores, "kind": "port",
                             "sample": "oracle C++/OpenMP restatement, %d threads; per step 1 spp, %s" % (cores, sample)},
            "e2e": {"value": v, "unit": "paths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_REAL_STDOUT = None


def _protect_stdout():
    """Everything that libraries print to fd 1 during the run (NCCL's version banner, for one) goes to
    stderr; the one JSON line is written to the real stdout at the end."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


# ----------------------------------------------------------------------------- other configurations
def _leg(fn):
    """Run one short leg; a failure is reported in the line, never hidden and never fatal for the main number."""
    try:
        return fn()
    except Exception as e:  # noqa: BLE001
        return {"failed": repr(e)}


def leg_path_or_restir(vrt, torch, name, W, H, device, sky_res, spp_per_step, steps):
    """configs 2 / 4: ms per frame of accumulate(spp) (path tracing) or accumulate_restir (ReSTIR) at N = 1."""
    wl = Workload(name, 128)
    r = vrt.Renderer(dx=2.0 / wl.R, image_res=(W, H), grid_res=wl.R, sky_res=sky_res if wl.sky else 0, exposure=wl.exposure, seed=1,
                     voxel_edges=wl.voxel_edges, device=device)
    wl.configure(r)
    r.prepare_data()
    if wl.restir:
        r.set_restir_temporal(True)  # config 4: temporal + spatial resampling per frame
    run = (lambda: r.accumulate_restir(spp_per_step)) if wl.restir else (lambda: r.accumulate(spp_per_step))
    for _ in range(3):
        run()
    r.synchronize()
    r.stats()
    clocks = ClockSampler(device).start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    render_ms = gris_ms = temporal_ms = 0.0
    for _ in range(steps):
        run()
        if wl.restir:  # ReSTIR calls are synchronous; their per-phase device times come from the library's events
            s = r.stats()
            render_ms, gris_ms, temporal_ms = render_ms + s["last_render_ms"], gris_ms + s["last_gris_ms"], temporal_ms + s["last_temporal_ms"]
    r.synchronize()
    wall = time.perf_counter() - t0
    st = r.stats()
    clk = clocks.stop()
    frames = steps * spp_per_step
    if wl.restir:
        dev_ms = render_ms + gris_ms + temporal_ms
    else:
        dev_ms = st["render_ms_sum"]
    img = r.fetch_image()
    out = {"workload": wl.desc.format(W=W, H=H), "frames": frames, "ms_per_frame": dev_ms / frames, "paths_per_s": W * H * frames / (dev_ms * 1e-3),
           "wall_ms_per_frame": 1e3 * wall / frames, "timing": "CUDA events of the library around its kernels, summed over the frames",
           "mean_ldr": float(img[..., :3].mean()), "clocks": clk}
    if wl.restir:
        out["phases_ms_per_frame"] = {"path+reservoir": render_ms / frames, "temporal": temporal_ms / frames, "spatial_gris": gris_ms / frames}
    r.close()
    return out


def leg_config1(vrt, torch, device):
    """config 1: example1.py scene, 640x640, primary ray + sun shadow ray on the cone axis, hit-buffer dump;
    the buffer is compared with the committed fixture (bit-exact) and the dump is timed end to end."""
    import hashlib as hl

    z = np.load(os.path.join(GOLD, "example1_seed0.npz"))
    gold = np.load(os.path.join(GOLD, "hits_example1_640.npz"))
    r = vrt.Renderer(dx=2.0 / 128, image_res=(640, 640), grid_res=128, sky_res=0, jitter=False, voxel_edges=float(z["voxel_edges"]),
                     exposure=float(z["exposure"]), device=device)
    r.set_voxels(z["material"], z["color"])
    r.set_floor(float(z["floor_height"]), z["floor_color"], int(z["floor_material"]))
    r.set_directional_light(z["light_dir"], float(z["light_noise"]), z["light_color"])
    r.set_background_color(z["background"])
    r.prepare_data()
    h = r.trace_primary()
    clocks = ClockSampler(device).start()
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        h = r.trace_primary()
    dt = (time.perf_counter() - t0) / n
    clk = clocks.stop()
    ok = hl.sha256(h.tobytes()).hexdigest() == str(gold["sha256"])
    r.close()
    return {"workload": "config1: example1 scene 640x640, primary hit + sun shadow ray, hit-buffer dump to host", "ms_per_dump_e2e": 1e3 * dt,
            "rays_per_s_e2e": 2 * 640 * 640 / dt, "bit_exact_vs_committed_oracle_fixture": bool(ok), "clocks": clk}


def sphere_scene(R=128, n=60):
    """example4.py:13-17 (the example with the most occupied voxels, 319 489): x.x < n*n/2 inside (-n, n)^3."""
    i = np.arange(-R // 2, R // 2)
    x, y, z = np.meshgrid(i, i, i, indexing="ij")
    inside = (x * x + y * y + z * z < n * n * 0.5) & (np.abs(x) < n) & (np.abs(y) < n) & (np.abs(z) < n)
    mat = inside.astype(np.int8)
    col = np.zeros((R, R, R, 3), np.uint8)
    col[inside] = (229, 76, 76)  # u8(0.9*255), u8(0.3*255)
    return mat, col


def leg_config5(vrt, torch, dist, rank, world, device, spp=1024):
    """config 5: 3840x2160, 1024 spp, example4's sphere, tile-sharded over the ranks (interleaved 8x4 tiles), merged by
    ONE gather: every rank's tonemap-and-merge kernel covers 1/N of the pixels and stores into rank 0's image
    buffer (the partial buffers have disjoint support, so the peer-memory sum IS the gather, W*H/N float4 per rank)."""
    from voxel_rt2_b200 import parallel

    W, H = 3840, 2160
    r = vrt.Renderer(dx=1 / 64, image_res=(W, H), grid_res=128, sky_res=0, exposure=1.0, seed=5, device=device)
    stream = torch.cuda.Stream()
    r.set_stream(stream.cuda_stream)
    mat, col = sphere_scene()
    r.set_voxels(mat, col)
    r.set_directional_light((1, 1, 1), 0.1, (1, 1, 1))      # example4.py:6
    r.set_background_color((0.3, 0.4, 0.6))                  # example4.py:7
    parallel.shard_tiles(r, rank, world)
    r.prepare_data()
    fm = parallel.FusedMerge(r) if world > 1 else None
    host = torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True).numpy() if rank == 0 else None
    with torch.cuda.stream(stream):
        if fm:
            fm.begin(1)
        r.accumulate(8)  # warm-up, incl. one merge: the first collective on a stream pays NCCL's set-up
        if fm:
            fm.merge()
            fm.barrier()
        r.reset_framebuffer()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(device).start() if rank == 0 else None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        if fm:
            fm.begin(0)
        done = 0
        while done < spp:
            n = min(64, spp - done)
            r.accumulate(n)
            done += n
        if fm:
            fm.merge()
            fm.finish(host)
        elif rank == 0:
            r.fetch_image_async(host)
            r.wait_image()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    out = None
    if rank == 0:
        out = {"workload": "config5: 3840x2160, %d spp, example4 sphere (%d voxels), tile-sharded x%d, fused peer-memory gather + tonemap, image to pinned host"
                           % (spp, int((mat > 0).sum()), world),
               "seconds": dt, "paths_per_s": W * H * spp / dt, "mean_ldr": float(host[..., :3].mean()), "alpha_min": float(host[..., 3].min()),
               "clocks": clocks.stop()}
    if fm:
        fm.close()
    r.close()
    return out


def leg_config4_rows(vrt, torch, dist, rank, world, device, W, H, sky_res, frames=120):
    """config 4 on N GPUs as ONE reservoir chain: row strips with a 24-pixel halo (vrt_set_row_shard), temporal + spatial
    resampling per frame, the strips merged per frame by the fused peer-memory kernel. Strong scaling of the ReSTIR frame."""
    from voxel_rt2_b200 import parallel

    wl = Workload("config4", 128)
    r = vrt.Renderer(dx=2.0 / wl.R, image_res=(W, H), grid_res=wl.R, sky_res=sky_res, exposure=wl.exposure, seed=1, voxel_edges=wl.voxel_edges,
                     device=device)
    stream = torch.cuda.Stream()
    r.set_stream(stream.cuda_stream)
    wl.configure(r)
    r.set_sky_shard(rank, world)
    r.prepare_data()
    cuts = parallel.shard_rows(r, rank, world)  # strips balanced by geometry pixels (sky rows are nearly free)
    r.set_restir_temporal(True)
    fm = parallel.FusedMerge(r)
    with torch.cuda.stream(stream):
        for k in range(4):  # warm-up: the chain reaches its steady state; the first collective on a stream pays NCCL's set-up
            fm.begin(k, reset=False)
            r.accumulate_restir(1)
            fm.merge()
        fm.barrier()
    torch.cuda.synchronize()
    dist.barrier()
    clocks = ClockSampler(device).start() if rank == 0 else None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for k in range(frames):
            fm.begin(k, reset=False)  # progressive accumulation: a reset would also drop the reservoir history
            r.accumulate_restir(1)
            fm.merge()
        fm.barrier()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    out = None
    if rank == 0:
        out = {"workload": wl.desc.format(W=W, H=H) + "; ONE chain over %d GPUs (row strips + 24-pixel halo, fused merge per frame)" % world,
               "frames": frames, "ms_per_frame": 1e3 * dt / frames, "paths_per_s": W * H * frames / dt, "tile_row_cuts": cuts,
               "timing": "wall clock around the frame loop, max over ranks", "clocks": clocks.stop()}
    fm.close()
    r.close()
    return out


# ----------------------------------------------------------------------------- main
def main():
    args = parse()
    _protect_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import voxel_rt2_b200 as vrt
    from voxel_rt2_b200 import parallel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    W, H = [int(x) for x in args.res.split("x")]
    wl = Workload(args.workload, args.grid)
    R = wl.R
    spp = 1 if wl.restir else args.spp  # ReSTIR mode renders one sample per frame

    r = vrt.Renderer(dx=2.0 / R, image_res=(W, H), grid_res=R, sky_res=args.sky_res if wl.sky else 0, exposure=wl.exposure, seed=1,
                     voxel_edges=wl.voxel_edges, device=local_rank)
    stream = torch.cuda.Stream()
    r.set_stream(stream.cuda_stream)
    wl.configure(r)
    if wl.restir:
        r.set_restir_temporal(True)  # config 4: temporal + spatial resampling per frame
    r.set_sample_shard(rank, world)
    if world > 1:
        r.set_sky_shard(rank, world)  # every rank computes 1/N of the sky-table rows, one all-gather per table
    t_prep = time.time()
    r.prepare_data()
    t_prep = time.time() - t_prep
    sky_ms = r.stats()["sky_precompute_ms"]

    # untimed counter pass (algorithmic bytes per path)
    r.accumulate(1, stats=True)
    st = r.stats()
    per = {k: st[k] / max(st["paths"], 1) for k in ("rays", "steps", "queries", "hits", "sky_escapes", "nee_visible", "vertices")}
    b_path = 4 * per["queries"] + 4 * per["hits"] + 96 * per["sky_escapes"] + 48 * per["nee_visible"] + 32
    r.reset_framebuffer()
    r.current_spp = 0

    fused = world > 1 and args.merge == "fused" and not wl.restir
    fm, merge_note = None, None
    if fused:
        try:
            fm = parallel.FusedMerge(r)  # raises on every rank or on none
        except RuntimeError as e:       # no CUDA-IPC peer access on this box: the NCCL all-reduce path still works
            fused, merge_note = False, "fused merge unavailable (%s): NCCL all-reduce used" % e
            r.set_accum_slot(0)
    accum = r.accum_tensor() if (world > 1 and not fused) else None
    counter = [0]

    def render():
        if wl.restir:
            r.accumulate_restir(spp)
        else:
            r.accumulate(spp)

    def step():
        with torch.cuda.stream(stream):
            if fm:
                fm.begin(counter[0])
                render()
                fm.merge()
            else:
                if world > 1:
                    r.reset_framebuffer()
                render()
                if world > 1:
                    dist.all_reduce(accum)  # --merge allreduce: one NCCL all-reduce of the accumulation buffer per step
            counter[0] += 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    r.stats()  # clears the "since the last query" sums
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    restir_ms = 0.0
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
        if wl.restir:
            s = r.stats()
            restir_ms += s["last_render_ms"] + s["last_gris_ms"] + s["last_temporal_ms"]
    if fm:
        with torch.cuda.stream(stream):
            fm.barrier()  # every rank's last merge has landed in rank 0's image buffer
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    clk = clocks.stop() if rank == 0 else None
    st = r.stats()
    launches = int(st["launches_total"]) if not wl.restir else 4 * spp * args.steps
    k_ms = (st["render_ms_sum"] / max(st["render_launches"], 1)) if not wl.restir else restir_ms / args.steps
    if world > 1:
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    paths_total = W * H * spp * args.steps * world
    value = paths_total / (ms_total * 1e-3)

    # ---- e2e: the call a user makes per displayed frame (scene.py:233-262): camera upload,
    # accumulate(spp), fetch_image into host memory. Host buffers, copies inside the timed region.
    pos, view, proj = vrt.default_camera_matrices(W, H)
    host_imgs = [torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True).numpy() for _ in range(2)] if rank == 0 else None
    n_img = [0]

    def e2e_step():
        with torch.cuda.stream(stream):
            r.set_view_proj(pos, view, proj)
            if fm:
                fm.begin(counter[0])
                render()
                fm.merge()
                if rank == 0 and n_img[0] > 0:
                    fm.copy_previous(host_imgs[n_img[0] % 2])  # image of the previous batch: complete since this step's barrier
            else:
                if world > 1:
                    r.reset_framebuffer()
                render()
                if world > 1:
                    dist.all_reduce(accum)
                if rank == 0:
                    # the displayed frame goes to pinned host memory through the pipelined fetch: tonemap on the render
                    # stream, D2H on the copy engine while the next step renders; the call first waits for the previous
                    # step's image, the last one is waited for before the clock stops
                    r.fetch_image_async(host_imgs[n_img[0] % 2])
            counter[0] += 1
            n_img[0] += 1

    def e2e_finish():
        with torch.cuda.stream(stream):
            if fm:
                fm.finish(host_imgs[n_img[0] % 2] if rank == 0 else None)
            elif rank == 0:
                r.wait_image()

    for _ in range(2):
        e2e_step()
    e2e_finish()
    barrier()
    n_img[0] = 0
    t0 = time.perf_counter()
    n_e2e = max(3, min(args.steps, 40))  # the pipeline drains once at the end: enough steps that the drain is not what is measured
    for _ in range(n_e2e):
        e2e_step()
    e2e_finish()
    barrier()
    e2e_s = time.perf_counter() - t0
    assert rank != 0 or float(host_imgs[0][..., :3].max()) > 0.0 and float(host_imgs[1][..., :3].max()) > 0.0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = W * H * spp * n_e2e * world / e2e_s
    h2d = 3 * 4 + 2 * 64  # camera position + view + projection (the TAA jitter is computed on the device)
    d2h = W * H * 16

    line = None
    if rank == 0:
        peak, how = measured_peak()
        achieved = (W * H * spp) * b_path / (k_ms * 1e-3) / 1e9
        traffic, traffic_src = measured_traffic(spp, args.sky_res, r.sky_format) if args.workload == "config3" else (None, "not captured for this workload")
        if world == 1:
            par = "1 GPU"
        elif fused:
            par = "sample-shard x%d + fused peer-memory reduce-scatter/tonemap per step (4-byte NCCL barrier)" % world
        else:
            par = "sample-shard x%d + 1 NCCL all-reduce/step" % world + ("; " + merge_note if merge_note else "")
        line = {
            "metric": "paths_per_sec_depth4_1080p", "value": value, "unit": "paths/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "ms_per_frame": ms_total / args.steps / spp,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.desc.format(W=W, H=H), "spp_per_step_per_gpu": spp, "sky_res": args.sky_res if wl.sky else 0,
                       "sky_format": r.sky_format, "parallelism": par,
                       "l2": "inputs>L2 (sky tables 2x%d MB + colour %d MB)" % (args.sky_res ** 2 * 16 // 2 ** 20, R ** 3 * 4 // 2 ** 20),
                       "sky_precompute_ms": sky_ms, "prepare_s": t_prep},
            "rays_per_sec": value * per["rays"], "per_path": per,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_source": traffic_src,
                         "peak_source": how + " (MEASURED_PEAKS.json hbm_gbs)" if how == "measured" else "fallback 6650",
                         "bytes_per_path": b_path, "kernel": "k_path", "kernel_ms": k_ms,
                         "note": "latency/issue bound by construction: working set is L2-resident except the sky tables"},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "paths/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s / n_e2e,
                    "call": ("set_view_proj + accumulate(spp) + fused merge + image of the previous batch -> pinned host on rank 0 (every image is complete before the clock stops)"
                             if fm else
                             "set_view_proj + accumulate(spp) [+ all-reduce] + fetch_image_async -> pinned host on rank 0 (D2H of step k overlaps step k+1; every image is complete before the clock stops)")},
            "gpu_launches": launches,
        }
        if wl.restir:
            line["roofline"]["kernel"] = "k_path<restir> + k_temporal + k_rc_sky + k_gris"
            line["roofline"]["note"] = "bytes_per_path counts the path kernel only; the resampling passes add 33 x 96 B of (L2-resident) tap reads per pixel"
    if fm:
        fm.close()

    # ---- full-size parity + CPU baseline (rank 0, N = 1): one oracle render serves both
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not wl.restir:
        try:
            sky_tables = r.get_sky_tables() if wl.sky else None
            cb = cpu_sample(args, wl, W, H, sky_tables, args.cpu_spp)
            line["cpu_baseline"] = {"value": cb["paths_per_s"], "unit": "paths/s", "cores": cb["cores"], "kind": "port",
                                    "sample": "oracle (C++/OpenMP restatement), full 1080p frame of the same workload, %d spp, %d paths, %.1f s"
                                              % (cb["spp"], cb["paths"], cb["ms"] * 1e-3),
                                    "per_path": {k: cb["counters"][k] / max(cb["paths"], 1) for k in
                                                 ("rays", "steps", "queries", "hits", "sky_escapes", "nee_visible")}}
            # the same sample indices on the GPU, float sky tables (the oracle reads float tables), default camera
            g = vrt.Renderer(dx=2.0 / R, image_res=(W, H), grid_res=R, sky_res=args.sky_res if wl.sky else 0, exposure=wl.exposure, seed=1,
                             voxel_edges=wl.voxel_edges, device=local_rank, sky_format="f32")
            wl.configure(g)
            if sky_tables is not None:
                g.set_sky_tables(*sky_tables)
            g.prepare_data()
            g.accumulate(args.cpu_spp)
            par_f32 = parity_figures(g.fetch_hdr(), cb["hdr"])
            line["parity"] = dict(par_f32, what="GPU vs CPU oracle, full frame %dx%d, %d^3 grid, sample indices 0..%d, float sky tables" % (W, H, R, args.cpu_spp - 1))
            if r.sky_format != "f32":
                r.reset_framebuffer()
                r.current_spp = 0
                r.set_sample_shard(0, 1)
                r.accumulate(args.cpu_spp)
                line["parity_bench_format"] = dict(parity_figures(r.fetch_hdr(), cb["hdr"]), what="the same with the sky-table format the timed region used (%s)" % r.sky_format)
            g.close()
        except Exception as e:  # the baseline is reported, never required for the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "paths/s", "cores": host_cores(), "kind": "port", "sample": "failed: %r" % (e,)}
    elif rank == 0 and wl.restir:
        line["cpu_baseline"] = {"value": None, "unit": "paths/s", "cores": host_cores(), "kind": "port",
                                "sample": "ReSTIR workloads: run `bench.py --impl reference --workload %s` for the CPU arm" % args.workload}
    r.close()

    # ---- the other BASELINE configurations, outside the timed region, each with its own clock record.
    # A watchdog guards the main line: if a leg hangs (a rank lost in a collective, say), every rank gives up after
    # 240 s, rank 0 prints the line measured so far with the failure noted, and the processes exit.
    if args.workload == "config3" and not args.no_other_configs:
        def give_up():
            if rank == 0:
                line["other_configs"] = {"failed": "a leg did not finish within 240 s; the main measurement above is unaffected"}
                emit(line)
            os._exit(0)

        watchdog = threading.Timer(240.0, give_up)
        watchdog.daemon = True
        watchdog.start()
        others = {}
        if world == 1:
            others["config1"] = _leg(lambda: leg_config1(vrt, torch, local_rank))
            others["config2"] = _leg(lambda: leg_path_or_restir(vrt, torch, "config2", W, H, local_rank, args.sky_res, 8, 8))
            others["config4_example6"] = _leg(lambda: leg_path_or_restir(vrt, torch, "config4", W, H, local_rank, args.sky_res, 1, 24))
            others["config4_example3"] = _leg(lambda: leg_path_or_restir(vrt, torch, "config4_example3", W, H, local_rank, args.sky_res, 1, 24))
        if world == 8:
            c5 = _leg(lambda: leg_config5(vrt, torch, dist, rank, world, local_rank))
            c4 = _leg(lambda: leg_config4_rows(vrt, torch, dist, rank, world, local_rank, W, H, args.sky_res))
            if rank == 0:
                others["config5"] = c5
                others["config4_example6_one_chain_over_8_gpus"] = c4
        watchdog.cancel()
        if rank == 0 and others:
            line["other_configs"] = others
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
