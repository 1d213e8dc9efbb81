#!/usr/bin/env python
"""bench.py — semi-dense mapping throughput on B200 (BASELINE.json metric) with roofline + CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1..5] [--kf KF_PER_GPU]

One "step" = one full SemiDenseLoop (pass 1: epipolar search + fusion [+ intra checks], halo exchange,
pass 2: inter-keyframe check + point set) over this rank's shard of a synthetic fr3_long_office-like
trajectory.  Default workload (--config 2) = BASELINE configs[1]: 200 keyframes per GPU, 640x480, TUM fr3
intrinsics, 6 covisible neighbours, intra + inter depth checks; N>1 (torchrun): weak scaling, contiguous shards,
halo (rho,sigma) planes pulled from peers over NVLink between the passes, ordered on the devices (sdm_exchange:
no host synchronisation or barrier inside a step).  The other BASELINE configs:
  --config 1  configs[0]: 10 keyframes (the reference's CPU-runnable correctness case)
  --config 3  configs[2]: 1000 keyframes in total, sharded over the GPUs (strong scaling; contiguous shards of equal estimated
              cost - candidate pixels x search-interval width, shard.balanced_bounds - unless --no-balance: 121-134 per GPU on 8)
  --config 4  configs[3]: 1280x960, 10 neighbours, wide inverse-depth range (long epipolar scans), 48 keyframes per GPU
  --config 5  configs[4]: 4096 keyframes in total (strong scaling), CPU arm on a 64-keyframe subset

`value`   : candidate pixels (edge & GradImg>8; each is searched against all N neighbours and fused)
            per second, planes resident in HBM, CUDA-event timed on the library's compute stream.
`e2e`     : the same through the C-ABI with pinned HOST buffers: H2D of every input plane, both passes,
            D2H of depth_map_/depth_sigma_/depth_map_checked_/SemiDensePointSets_ inside the timed region
            (sdm_run_loop: the library's own pipelined SemiDenseLoop); `e2e_class` = the same loop through the
            drop-in C++ class, ProbabilityMapping::SemiDenseLoop() (tests/cpp/test_shim.cpp --time); `e2e_online` = the class in
            online mode, wall clock per arriving keyframe (--time-online).  Further variants in the same line, each checked
            against the dense planes: e2e_blocks / e2e_points_blocks (sdm_loop.sparse_download = 1 / 2), e2e_scatter,
            e2e_point_export, e2e_image_in_points_out.
`--impl reference` : the CPU path (oracle/, a restatement of the reference's ProbabilityMapping; the
            reference itself cannot be compiled here) on all host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "eao-slam_b200", "python"))

import numpy as np  # noqa: E402

from sdmb200 import shard, synth  # noqa: E402

_REAL_STDOUT = sys.stdout
UNIT = "px/s"
W, H = 640, 480  # set per config in main()
L2_BYTES = 126 * 2 ** 20

# BASELINE.json configs (1-based like BASELINE.md / VERDICT.md).  kf = keyframes per GPU (weak scaling) or total = keyframes of
# the whole job divided over the GPUs (strong scaling)
CONFIGS = {
    1: dict(tag="configs[0]", W=640, H=480, kf=10, nbr=6, seed=1, wide=False, cpu_kf=10,
            what="10 synthetic 640x480 keyframes, 6 covisible neighbours (the reference's CPU-runnable correctness case)"),
    2: dict(tag="configs[1]", W=640, H=480, kf=200, nbr=6, seed=2, wide=False, cpu_kf=12,
            what="200-keyframe synthetic fr3_long_office-like trajectory per GPU"),
    3: dict(tag="configs[2]", W=640, H=480, total=1000, nbr=6, seed=3, wide=False, cpu_kf=12,
            what="1000 keyframes sharded over the GPUs, neighbour depth maps exchanged over NVLink"),
    4: dict(tag="configs[3]", W=1280, H=960, kf=48, nbr=10, seed=4, wide=True, cpu_kf=12, contrast=1.8,
            what="1280x960 keyframes, 10 neighbours, wide inverse-depth search range (long epipolar scans); texture contrast 1.8 "
                 "so that 24 % of the pixels are candidates like in configs[1] (the 0.6 of the VGA scenes leaves 2 % at this resolution)"),
    5: dict(tag="configs[4]", W=640, H=480, total=4096, nbr=6, seed=5, wide=False, cpu_kf=64,
            what="4096-keyframe offline re-densification sweep"),
}


def metric_name():
    return f"semi-dense pixels/sec (searched+fused) at {W}x{H}"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE config, 1-based (default 2 = configs[1])")
    ap.add_argument("--kf", type=int, default=None, help="keyframes per GPU (overrides the config)")
    ap.add_argument("--nbr", type=int, default=None)
    ap.add_argument("--intra", type=int, default=1, help="IntraKeyFrameDepthChecking/Growing on (config[1]: full checks)")
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--cpu-sample-kf", type=int, default=None, help="keyframes of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-variants", action="store_true", help="skip e2e_scatter / e2e_point_export / e2e_image_in_points_out / e2e_class")
    ap.add_argument("--no-balance", action="store_true", help="strong-scaling configs at N > 1: equal keyframe counts per rank instead of equal estimated cost")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the boundary-keyframe check against the oracle")
    ap.add_argument("--no-hot-spin", action="store_true", help="profiling runs: skip the ~0.6 s of extra untimed steps")
    return ap.parse_args()


def apply_config(a, world):
    """fills a.kf / a.nbr / a.seed / a.cpu_sample_kf from the config, sets the global image size"""
    global W, H
    c = CONFIGS[a.config]
    W, H = c["W"], c["H"]
    a.scaling = "strong" if "total" in c and a.kf is None else "weak"
    if a.kf is None:
        if "total" in c:
            if c["total"] % world:
                raise SystemExit(f"config {a.config}: {c['total']} keyframes do not divide over {world} GPUs")
            a.kf = c["total"] // world
        else:
            a.kf = c["kf"]
    a.nbr = a.nbr if a.nbr is not None else c["nbr"]
    a.seed = a.seed if a.seed is not None else c["seed"]
    a.cpu_sample_kf = a.cpu_sample_kf if a.cpu_sample_kf is not None else c["cpu_kf"]
    a.wide = c["wide"]
    a.contrast = c.get("contrast", 0.6)
    return c


# ------------------------------------------------------------------------------------------------
# scene (cached on local disk so the two arms of one box share the render)
# ------------------------------------------------------------------------------------------------
def load_scene(n, first, nbr_idx, seed, tag, wide=False, contrast=0.6):
    cache = os.path.join(os.environ.get("SDM_SCENE_CACHE", "/tmp/sdm_scene_cache"),
                         f"{tag}_s{seed}_f{first}_n{n}_{W}x{H}_N{nbr_idx.shape[1]}{'_wide' if wide else ''}_c{contrast}")
    names = ("im", "grad", "theta", "Tcw", "min_depth", "max_depth")
    if os.path.isdir(cache) and all(os.path.exists(os.path.join(cache, k + ".npy")) for k in names):
        a = {k: np.load(os.path.join(cache, k + ".npy")) for k in names}
        K = tuple(float(np.float32(v * (W / 640.0))) for v in synth.TUM3_K)
        return synth.Scene(im=a["im"], grad=a["grad"], theta=a["theta"], edge=None, K=K, Tcw=a["Tcw"],
                           nbr_idx=nbr_idx, rot=np.zeros(nbr_idx.shape, np.float32),
                           min_depth=a["min_depth"], max_depth=a["max_depth"], meta={"cache": cache})
    world = int(os.environ.get("WORLD_SIZE", "1"))
    workers = max(1, min(16, (os.cpu_count() or 8) // world))
    sc = synth.make_scene(n, W, H, nbr_idx.shape[1], seed=seed, first=first, workers=workers, nbr_idx=nbr_idx, wide_range=wide,
                          contrast=contrast)
    try:
        os.makedirs(cache, exist_ok=True)
        for k in names:
            np.save(os.path.join(cache, k + ".npy"), getattr(sc, k))
    except OSError:
        pass
    return sc


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks DURING the timed region")
# ------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None
        # nvidia-smi delivers a row every ~100 ms at best, i.e. one or two inside a 5-step timed region: NVML is polled
        # directly as well (every 10 ms, same fields) and its rows are the ones reported when it works
        self.nvml_rows, self.nvml_on = [], True
        threading.Thread(target=self._poll_nvml, daemon=True).start()

    def _poll_nvml(self):
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.idx)
            mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
            get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = (N.nvmlClocksThrottleReasonHwSlowdown, N.nvmlClocksThrottleReasonHwThermalSlowdown,
                    N.nvmlClocksThrottleReasonSwThermalSlowdown, N.nvmlClocksThrottleReasonSwPowerCap)
            while self.nvml_on:
                r = get_reasons(h)
                row = [str(self.idx), str(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)), str(mx),
                       str(N.nvmlDeviceGetPowerUsage(h) / 1000.0)] + ["Active" if r & b else "Not Active" for b in bits]
                self.nvml_rows.append((time.time(), row))
                time.sleep(0.01)
        except Exception as e:  # noqa: BLE001  (no NVML, unsupported query: the nvidia-smi rows remain)
            self.nvml_error = repr(e)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        self.nvml_on = False
        if self.proc:
            self.proc.terminate()

    def summary(self, t0, t1):
        source = "nvml, polled every 10 ms during the timed region"
        rows = [r for t, r in getattr(self, "nvml_rows", []) if t0 <= t <= t1]
        if len(rows) < 2:
            source = "nvidia-smi -lms 100"
            rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.15 and len(r) >= 8] or [r for _, r in self.rows if len(r) >= 8]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[1]) for r in rows]
        reasons = [n for k, n in ((4, "hw_slowdown"), (5, "hw_thermal_slowdown"), (6, "sw_thermal_slowdown"),
                                  (7, "sw_power_cap")) if any(r[k].lower().startswith("active") for r in rows)]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][2]), "reasons": reasons,
                "samples": len(rows), "power_w_max": max(float(r[3]) for r in rows), "source": source}


# ------------------------------------------------------------------------------------------------
# byte model (DESIGN.md "Roofline")
# ------------------------------------------------------------------------------------------------
def bytes_pass1(n_kf, N):
    return (17 + 9 * N) * W * H * n_kf


def bytes_total(n_kf, N, intra):
    return (41 + 17 * N + (16 if intra else 0)) * W * H * n_kf


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(config):
    """dram bytes per k_pass1_lane launch from the committed ncu --set full capture of this same command
    (profiles/traffic.json is written by tools/ncu_summary.py from the capture, not by hand)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t.get(f"config{config}", {}).get("k_pass1_dram_bytes_per_launch")
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# CPU path (oracle) — used ONLY as cpu_baseline and as the --impl reference arm
# ------------------------------------------------------------------------------------------------
def cpu_run(sample_kf, nbr, seed, intra, steps, warmup, wide=False, contrast=0.6):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as O
    # launchers (torchrun) export OMP_NUM_THREADS=1 to their workers: the CPU arm uses every core of the box, as the
    # reference does (CMakeLists.txt:47-52), whatever the launcher set
    O.lib("fast").oracle_set_num_threads(os.cpu_count() or 1)
    sample_kf = max(sample_kf, nbr + 1)
    nb = synth.neighbours(sample_kf, nbr)
    sc = load_scene(sample_kf, 0, nb, seed, "cpu", wide, contrast)
    times, cands = [], 0
    for it in range(warmup + steps):
        osc = O.OracleScene(sc, "fast")
        p = O.default_params("fast", intra_check=intra, intra_grow=intra)
        t = osc.run(params=p)
        cands = osc.stats.as_dict()["candidates"]
        if it >= warmup:
            times.append(t)
    cores = O.lib("fast").oracle_num_threads()
    sec = sum(times) / len(times)
    return {"value": cands / sec, "sec_per_step": sec, "cores": cores, "candidates": cands, "sample_kf": sample_kf,
            "sample": f"full SemiDenseLoop (pass 1 + {'intra + ' if intra else ''}pass 2) over the first {sample_kf} "
                      f"keyframes of the same synthetic trajectory ({W}x{H}), {nbr} neighbours, OpenMP over {cores} threads, "
                      f"gcc -O3 -march=x86-64-v3; ms/KF = {1e3 * sec / sample_kf:.1f} (per-keyframe cost is independent of the "
                      f"trajectory length, so the job's CPU time extrapolates linearly)"}


def main_reference(a, rank):
    if rank != 0:
        return
    r = cpu_run(a.cpu_sample_kf, a.nbr, a.seed, a.intra, max(1, a.steps), max(0, a.warmup), a.wide, a.contrast)
    line = {
        "impl": "reference", "metric": metric_name(), "value": r["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * r["sec_per_step"], "higher_is_better": True, "scaling": a.scaling,
        "vs_baseline": None, "dtype": "f32 (f64 where OpenCV accumulates in double)", "data": "synthetic",
        "config": workload_config(a, a.kf),
        "reference_arm_note": f"each step = a bounded sample ({r['sample_kf']} keyframes) of the workload; CPU restatement of the "
                              "reference path (the reference needs OpenCV/Eigen/Boost/CGAL headers absent here)",
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if a.config == 2:
        src = reference_source_timing(a)
        if src:
            line["reference_source"] = src
    print(json.dumps(line), file=_REAL_STDOUT, flush=True)


def reference_source_timing(a):
    """How conservative is the port as a baseline?  oracle/_ref/libref_pm.so is the reference's OWN
    src/ProbabilityMapping.cc (compiled against the stand-in cv::Mat of oracle/refshim/; covisN is compiled in as 7 and
    the intra checks are commented out of its loop, so it cannot run the 6-neighbour workload itself).  Time its
    SemiDenseLoop() and the port on the same 12-keyframe, 7-neighbour sample."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_py as O
        import ref_py
        if not ref_py.available(build=False):
            return None
        n = max(12, a.cpu_sample_kf)
        nb = synth.neighbours(n, 7)
        sc = synth.make_scene(n, W, H, 7, seed=a.seed, nbr_idx=nb, workers=1)
        lib = O.lib("fast")
        for i in range(n):
            lo, hi = C.c_float(), C.c_float()
            lib.oracle_stereo_search_constraints(O.fptr(sc.inv_depths[i]), len(sc.inv_depths[i]), C.byref(lo), C.byref(hi))
            sc.min_depth[i], sc.max_depth[i] = lo.value, hi.value
        t = time.perf_counter()
        ref = ref_py.run_reference_loop(sc, np.full(n, 50.0, np.float32))
        t_ref = time.perf_counter() - t
        osc = O.OracleScene(sc, "fast")
        t_port = osc.run(params=O.default_params("fast"))
        cands = osc.stats.as_dict()["candidates"]
        same = bool(((ref["checked"] > 0) == (osc.checked > 0)).mean() > 0.999)
        return {"value": cands / t_ref, "unit": UNIT, "ms_per_keyframe": 1e3 * t_ref / n, "port_ms_per_keyframe": 1e3 * t_port / n,
                "port_speedup_over_reference_source": t_ref / t_port, "same_accepted_set": same,
                "sample": f"{n} keyframes, 7 neighbours (the reference's compiled-in covisN), no intra stage, "
                          f"{O.lib('fast').oracle_num_threads()} OpenMP threads; cv::Mat is a stand-in, so this is the reference's "
                          "control flow and per-pixel allocation pattern, not OpenCV's allocator"}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": str(e)[:200]}


def workload_config(a, kf_per_gpu, note=None):
    cfg = CONFIGS[a.config]
    in_mb = W * H * 26 / 1e6  # packed planes a keyframe's scan reads per neighbour slot: texel 16 + intensity 2 + skip 8 B/px
    c = {"workload": f"BASELINE {cfg['tag']}: {cfg['what']}; {kf_per_gpu} keyframes per GPU x {a.gpus} GPU(s), {W}x{H}, TUM fr3 "
                     f"intrinsics{' x2' if W == 1280 else ''}, {a.nbr} covisible neighbours, epipolar search + fusion"
                     f"{' + intra-keyframe checks' if a.intra else ''} + inter-keyframe check + point set",
         "baseline_config": a.config, "keyframes_per_gpu": kf_per_gpu, "keyframes_total": kf_per_gpu * a.gpus,
         "neighbours": a.nbr, "intra": bool(a.intra), "seed": a.seed, "image": f"{W}x{H}", "wide_depth_range": bool(a.wide),
         "l2": f"inputs larger than L2: {kf_per_gpu} keyframes x {in_mb:.1f} MB of packed planes read per step vs 126 MB L2"
               if kf_per_gpu * in_mb > 126 else
               f"inputs ({kf_per_gpu} keyframes x {in_mb:.1f} MB) smaller than the 126 MB L2: a 256 MB buffer is overwritten between timed steps",
         "parallelism": f"keyframe shards x{a.gpus}, halo (rho,sigma) planes pulled over NVLink (CUDA IPC peer copies ordered by "
                        "flags in peer memory: no host barrier inside a step)" if a.gpus > 1 else "single GPU"}
    if note:
        c["note"] = note
    return c


# ------------------------------------------------------------------------------------------------
def pinned(api_lib, shape, dtype, keep):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    rc = api_lib.sdm_host_alloc(C.byref(p), n)
    if rc:
        raise RuntimeError("sdm_host_alloc failed: " + api_lib.sdm_last_error().decode())
    keep.append(p)
    buf = (C.c_char * n).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def boundary_parity(a, plan, rank, world, out, owned):
    """N > 1: the keyframes at the ends of this rank's shard are the ones whose pass 2 read planes pulled from a peer.
    Recompute each of them with the oracle (canonical build) on a window of the global trajectory around it - its
    neighbours' pass 1 needs the inputs of THEIR neighbours, i.e. +-N keyframes - and compare the four planes bit for
    bit.  Returns (keyframes checked, differing 32-bit words)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as O
    O.lib("canonical").oracle_set_num_threads(max(1, (os.cpu_count() or 1) // world))
    G = a.kf * world
    nb_global = synth.neighbours(G, a.nbr)
    checked, bad = 0, 0
    for g in sorted({plan.own_lo, plan.own_hi - 1}):
        if not any(not (plan.own_lo <= int(v) < plan.own_hi) for v in nb_global[g]):
            continue  # (ends of the whole trajectory: no remote neighbour)
        lo, hi = max(0, g - a.nbr - 1), min(G, g + a.nbr + 2)
        nb_win = np.clip(nb_global[lo:hi] - lo, 0, hi - lo - 1).astype(np.int32)  # out-of-window entries: unused keyframes only
        sc = synth.make_scene(hi - lo, W, H, a.nbr, seed=a.seed, first=lo, nbr_idx=nb_win, wide_range=a.wide, workers=1,
                              contrast=a.contrast)
        osc = O.OracleScene(sc)
        p = O.default_params("canonical", intra_check=a.intra, intra_grow=a.intra)
        need1 = sorted({g} | {int(v) for v in nb_global[g]})
        for k in need1:  # pass 1 of the keyframe and of its neighbours (their lists lie inside the window)
            assert all(lo <= int(v) < hi for v in nb_global[k])
            osc.run(params=p, first=k - lo, count=1, pass_mask=1)
        osc.run(params=p, first=g - lo, count=1, pass_mask=2)
        j = owned.index(g - plan.lo)
        for name, ref in (("depth", osc.depth), ("sigma", osc.sigma), ("checked", osc.checked), ("points", osc.points)):
            bad += int((out[name][j].view(np.uint32) != ref[g - lo].view(np.uint32)).sum())
        checked += 1
    return checked, bad


def class_e2e(a, sc, n_loc, mode="--time"):
    """ProbabilityMapping::SemiDenseLoop() of the drop-in C++ class on the same keyframes (planes in pinned host memory,
    everything uploaded again in every repetition): tests/cpp/test_shim.cpp --time, compiled here with g++ -std=c++11."""
    import struct
    import tempfile
    tmp = tempfile.mkdtemp(prefix="sdm_class_")
    exe, scene_path, out_path = (os.path.join(tmp, f) for f in ("test_shim", "scene.bin", "out.bin"))
    libdir = os.path.join(ROOT, "eao-slam_b200", "lib")
    r = subprocess.run(["g++", "-std=c++11", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_shim.cpp"),
                        "-o", exe, "-L", libdir, "-lsdm_b200", f"-Wl,-rpath,{libdir}", "-lpthread"], capture_output=True, text=True)
    if r.returncode:
        return {"unavailable": "g++: " + r.stderr[-200:]}
    with open(scene_path, "wb") as f:
        f.write(struct.pack("9i", n_loc, W, H, a.nbr, 0, 0, sc.nbr_idx.shape[1], 1, 11))  # (neighbour lists index < n_loc)
        f.write(np.asarray(sc.K, np.float32).tobytes())
        for i in range(n_loc):
            f.write(np.ascontiguousarray(sc.Tcw[i], np.float32).tobytes())
            f.write(sc.im[i].tobytes()); f.write(sc.grad[i].tobytes()); f.write(sc.theta[i].tobytes())
            # StereoSearchConstraints input that reproduces this keyframe's search bounds exactly is not kept by the cached
            # scene: two inverse depths whose mean -+ 2 sigma give them back to within rounding (timing only)
            lo, hi = 1.0 / float(sc.min_depth[i]), 1.0 / float(sc.max_depth[i])  # mean - 2s, mean + 2s
            m, sd = 0.5 * (lo + hi), 0.25 * (hi - lo)
            inv = np.array([m - sd, m + sd], np.float32)
            f.write(struct.pack("i", 2)); f.write(inv.tobytes())
            f.write(np.ascontiguousarray(sc.nbr_idx[i], np.int32).tobytes())
            f.write(struct.pack("i", 0))
    r = subprocess.run([exe, mode, scene_path, out_path, str(max(2, a.steps))], capture_output=True, text=True)
    for f in (scene_path, out_path, exe):
        try:
            os.remove(f)
        except OSError:
            pass
    if r.returncode:
        return {"unavailable": (r.stdout + r.stderr)[-300:]}
    return json.loads(r.stdout.strip().splitlines()[-1])


def balance_bounds(a, rank, world):
    """Strong scaling: cut the trajectory into contiguous ranges of equal estimated COST instead of equal keyframe counts.
    Every rank renders the keyframes of its equal-count range, reports per keyframe the candidate pixels and the width of
    the inverse-depth search interval (what the scan length is proportional to); the ranks swap these through files in the
    node's tmp directory (before CUDA / NCCL are initialised, like the render itself) and all compute the same bounds."""
    import tempfile
    G = a.kf * world
    own_lo = rank * a.kf
    sc = load_scene(a.kf, own_lo, np.zeros((a.kf, a.nbr), np.int32), a.seed, "gpu_own", a.wide, a.contrast)
    cands = (sc.grad > 8).reshape(a.kf, -1).sum(1).astype(np.float64)
    width = (1.0 / sc.max_depth.astype(np.float64) - 1.0 / sc.min_depth.astype(np.float64))
    d = os.path.join(tempfile.gettempdir(), f"sdm_balance_{os.getppid()}_{os.environ.get('MASTER_PORT', '0')}_c{a.config}_s{a.seed}_{G}_{world}")
    os.makedirs(d, exist_ok=True)
    tmp = os.path.join(d, f".r{rank}.npy")
    np.save(tmp, np.stack([cands, width]))
    os.replace(tmp, os.path.join(d, f"r{rank}.npy"))
    parts, t0 = [], time.time()
    for r in range(world):
        f = os.path.join(d, f"r{r}.npy")
        while not os.path.exists(f):
            if time.time() - t0 > 900:
                raise RuntimeError(f"rank {rank}: no shard weights from rank {r} after 900 s")
            time.sleep(0.05)
        parts.append(np.load(f))
    allc = np.concatenate([p[0] for p in parts])
    allw = np.concatenate([p[1] for p in parts])
    # 70 % of the scan kernel's instructions are in the column loop (proportional to the search range), the rest per candidate
    weights = allc * (0.3 + 0.7 * allw / max(allw.mean(), 1e-30))
    return shard.balanced_bounds(weights, world), weights


def main_ours(a, rank, world, local_rank):
    dist = None
    nb_global = synth.neighbours(a.kf * world, a.nbr)
    bounds, shard_note = None, None
    if world > 1 and a.scaling == "strong" and not a.no_balance:
        bounds, wts = balance_bounds(a, rank, world)
        cost = [float(wts[bounds[r]:bounds[r + 1]].sum()) for r in range(world)]
        eq = [float(wts[r * a.kf:(r + 1) * a.kf].sum()) for r in range(world)]
        shard_note = {"bounds": [int(b) for b in bounds], "estimated_cost_max_over_mean": max(cost) / (sum(cost) / world),
                      "equal_count_shards_would_be": max(eq) / (sum(eq) / world),
                      "weights": "candidate pixels x (0.3 + 0.7 x width of the inverse-depth search interval / its mean)"}
    plan = shard.make_plan(nb_global, a.kf, rank, world, bounds)
    nb_local = np.where(plan.nbr_local >= 0, plan.nbr_local, 0).astype(np.int32)
    sc = load_scene(plan.n_local, plan.lo, nb_local, a.seed, "gpu", a.wide, a.contrast)  # before CUDA init (forks workers)

    try:  # bind this rank to the CPUs / NUMA node next to its GPU before any pinned buffer is touched
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
    except Exception as e:  # noqa: BLE001
        print(f"rank {rank}: no NUMA affinity ({e})", file=sys.stderr)

    import torch
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from sdmb200 import api
    os.environ.setdefault("SDM_SCATTER_THREADS", str(max(2, min(6, (os.cpu_count() or 8) // world - 1))))
    lib = api.load()  # raises if the CUDA library is absent: there is no fallback
    ctx = api.Context(width=W, height=H, max_keyframes=plan.n_local, intra_check=a.intra, intra_grow=a.intra,
                      device=local_rank)
    keep = []
    n_loc, owned = plan.n_local, list(plan.owned_local)
    do_e2e = not a.no_e2e and a.config != 5  # 4096 keyframes: 40 GB of pinned output planes for nothing new
    h_im = pinned(lib, (n_loc, H, W), np.uint8, keep); h_im[:] = sc.im
    h_g = pinned(lib, (n_loc, H, W), np.float32, keep); h_g[:] = sc.grad
    h_t = pinned(lib, (n_loc, H, W), np.float32, keep); h_t[:] = sc.theta
    sc.im, sc.grad, sc.theta = h_im, h_g, h_t
    n_out = len(owned) if (do_e2e or world > 1) else 0
    out = {k: pinned(lib, (max(n_out, 1), H, W) + ((3,) if k == "points" else ()), np.float32, keep)
           for k in ("depth", "sigma", "checked", "points")}

    host_exchange = bool(os.environ.get("SDM_BENCH_HOST_EXCHANGE"))  # A/B: round 1's synchronize + barrier + sdm_pull_halo
    if world > 1:  # exchange CUDA-IPC handles once
        handles = [None] * world
        if host_exchange:
            dist.all_gather_object(handles, ctx.export_arena())
            for r in set(int(x) for x in plan.halo_rank):
                ctx.import_peer_arena(r, handles[r])
        else:
            dist.all_gather_object(handles, ctx.export_peer_handle(rank))
            for r in set(int(x) for x in plan.halo_rank):
                ctx.import_peer(r, handles[r])
            ctx.set_halo(plan.halo_local, plan.halo_rank, plan.halo_peer_slot)
            dist.barrier()  # every import (= puller registration) done before the first exchange

    items = api.make_items(owned, sc.nbr_idx, sc.rot, sc.min_depth, sc.max_depth)

    def barrier():
        if world > 1:
            dist.barrier()

    def upload():
        descs = ctx.upload_descs(sc, range(n_loc))
        ctx.upload_keyframes(descs)

    def exchange():
        if world > 1:
            if host_exchange:
                ctx.synchronize(); barrier()
                ctx.pull_halo(plan.halo_local, plan.halo_rank, plan.halo_peer_slot)
            else:
                ctx.exchange()  # flags in peer memory order the pulls: nothing for the host to wait for

    flush = None
    if n_loc * W * H * 26 < 2 * L2_BYTES:  # small workloads: overwrite a buffer larger than L2 between timed steps
        flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device=f"cuda:{local_rank}")

    dbg = bool(os.environ.get("SDM_BENCH_DEBUG"))

    def step():
        if dbg:
            ctx.mark(2); t = [time.perf_counter()]
        ctx.pass1(items)
        if dbg:
            ctx.mark(3); t.append(time.perf_counter())
        exchange()
        if dbg:
            ctx.mark(4); t.append(time.perf_counter())
        ctx.pass2(items)
        if dbg:
            ctx.mark(5); t.append(time.perf_counter())
            ctx.synchronize()
            print(f"rank {rank} step: host issue ms (pass1, exchange, pass2) {[round(1e3 * (b - a_), 2) for a_, b in zip(t, t[1:])]}, "
                  f"device ms {[round(ctx.elapsed_ms(i, i + 1), 2) for i in (2, 3, 4)]}", file=sys.stderr)
        if world > 1 and host_exchange:
            ctx.synchronize(); barrier()

    # e2e: the whole loop through sdm_run_loop (the library issues uploads, passes and downloads in chunks so that its
    # three streams overlap: H2D of chunk k+1 under pass 1 of chunk k, depth_map_/depth_sigma_ of a chunk back as soon as
    # its pass 1 is queued, pass 2 of a keyframe as soon as pass 1 of all its neighbours is queued; work orders that read
    # halo planes wait for the exchange).  One sdm_upload_desc / sdm_download_desc per keyframe, built once.
    CH = int(os.environ.get("SDM_BENCH_CHUNK", "4"))  # 4 / 6 / 10 / 16 keyframes per chunk: 33.9 / 34.4 / 34.6 / 35.1 ms per loop
    up_desc = ctx.upload_descs(sc, range(n_loc))
    chk = ctx._chk
    dl1 = dl2 = dl4 = None
    if n_out:
        dl1 = (api.DownloadDesc * len(owned))()   # depth_map_, depth_sigma_   (after pass 1)
        dl2 = (api.DownloadDesc * len(owned))()   # depth_map_checked_, SemiDensePointSets_   (after pass 2)
        dl4 = (api.DownloadDesc * len(owned))()   # all four, for sdm_scatter_keyframes
        for j, s in enumerate(owned):
            dl1[j].kf = dl2[j].kf = dl4[j].kf = s
            dl1[j].depth, dl1[j].depth_step = out["depth"][j].ctypes.data, 4 * W
            dl1[j].sigma, dl1[j].sigma_step = out["sigma"][j].ctypes.data, 4 * W
            dl2[j].checked, dl2[j].checked_step = out["checked"][j].ctypes.data, 4 * W
            dl2[j].points, dl2[j].points_step = out["points"][j].ctypes.data, 12 * W
            dl4[j].depth, dl4[j].depth_step = dl1[j].depth, 4 * W
            dl4[j].sigma, dl4[j].sigma_step = dl1[j].sigma, 4 * W
            dl4[j].checked, dl4[j].checked_step = dl2[j].checked, 4 * W
            dl4[j].points, dl4[j].points_step = dl2[j].points, 12 * W
    loop = api.Loop()
    loop.n_upload, loop.upload = n_loc, up_desc
    loop.n_pass1, loop.pass1, loop.down1 = len(items), items, dl1
    loop.n_pass2, loop.pass2, loop.down2 = len(items), items, dl2
    loop.chunk, loop.exchange = CH, int(world > 1 and not host_exchange)

    def e2e_step(sparse=False, blocks=True):
        if not sparse and not host_exchange:
            # blocks: the destination planes are zero-initialised like KeyFrame's (KeyFrame.cc:78-81), so only the 16-pixel
            # blocks that hold a candidate cross PCIe (sdm_loop.sparse_download); else the dense DMA of every plane
            loop.sparse_download = int(blocks)
            chk(lib.sdm_run_loop(ctx.h, C.byref(loop)))
        else:  # variants that need a call between the passes: the same order issued from here
            ctx.upload_keyframes(up_desc)
            ctx.pass1(items)
            if not sparse:
                ctx.download_keyframes(dl1)
            exchange()
            ctx.pass2(items)
            chk(lib.sdm_scatter_keyframes(ctx.h, len(owned), dl4) if sparse else lib.sdm_download_keyframes(ctx.h, len(owned), dl2))
        ctx.synchronize()
        if world > 1 and host_exchange:
            barrier()

    upload(); ctx.synchronize()
    pack_ms = ctx.last_pack_ms()
    cands = sum(ctx.candidate_count(s) for s in owned)
    # warm-up: W (>= 3) untimed steps, then keep stepping for ~0.6 s so that the nvidia-smi sampler (started first)
    # is delivering rows and the GPU goes into the timed region hot, with no idle gap in between
    clocks = Clocks(local_rank)
    if rank == 0:
        clocks.start()
    tw = time.perf_counter()
    for _ in range(max(3, a.warmup)):
        step()
    ctx.synchronize()
    per_step = max(1e-4, (time.perf_counter() - tw) / max(3, a.warmup))
    extra = [0 if a.no_hot_spin else int(min(200, max(0, 0.6 / per_step)))]
    if world > 1:
        dist.broadcast_object_list(extra, src=0)
    for _ in range(extra[0]):
        step()
    ctx.synchronize(); barrier(); torch.cuda.synchronize()
    l0, t0 = ctx.launch_count(), time.time()
    if flush is None:
        ctx.mark(0)
        for _ in range(a.steps):
            step()
        ctx.mark(1)
        ctx.synchronize(); torch.cuda.synchronize(); barrier()
        ms = ctx.elapsed_ms(0, 1) / a.steps
    else:  # L2 flushed between steps: each step timed on its own (marks 0 / 1), the flush in between is not
        tot = 0.0
        for _ in range(a.steps):
            flush.add_(1); torch.cuda.synchronize()
            ctx.mark(0); step(); ctx.mark(1)
            ctx.synchronize()
            tot += ctx.elapsed_ms(0, 1)
        barrier()
        ms = tot / a.steps
    t1 = time.time()
    clocks.nvml_on = False  # the poller only serves the timed region above: keep it out of the host-timed e2e legs
    launches = ctx.launch_count() - l0
    timing = ctx.last_timing()
    timing["pack_ms"] = pack_ms
    stats = ctx.stats()

    # ---- multi-GPU parity on the hardware: boundary keyframes (pass 2 read planes that crossed NVLink) against the oracle
    parity = None
    if world > 1 and not a.no_parity:
        ctx.download_keyframes(dl4); ctx.synchronize()  # results of the timed steps
        parity = boundary_parity(a, plan, rank, world, out, owned)

    # ---- end to end through the C-ABI with host buffers
    e2e = None
    if do_e2e:
        # the headline: every plane leaves by DMA, no assumption about the destination planes
        e2e_step(blocks=False)
        barrier(); tt = time.perf_counter()
        for _ in range(a.steps):
            e2e_step(blocks=False)
        barrier(); e2e_s = (time.perf_counter() - tt) / a.steps
        e2e = {"sec": e2e_s, "h2d": int(n_loc * W * H * 9), "d2h": int(len(owned) * W * H * 24)}
    if do_e2e and not a.no_e2e_variants:
        # variant: block-sparse download into planes that start zero-initialised like KeyFrame's (KeyFrame.cc:78-81)
        ref_planes = {k: out[k][::5].copy() for k in out}
        for k in out:
            out[k][:] = 0
        e2e_step()
        barrier(); tt = time.perf_counter()
        for _ in range(a.steps):
            e2e_step()
        barrier(); e2e["blocks_sec"] = (time.perf_counter() - tt) / a.steps
        e2e["blocks_same"] = all(np.array_equal(out[k][::5].view(np.uint32), ref_planes[k].view(np.uint32)) for k in out)
        if not e2e["blocks_same"]:
            print("WARNING: block-sparse download differs from the dense download", file=sys.stderr)
        e2e["blocks_d2h"] = int(ctx.candidate_blocks(owned) * 16 * 24)
        # variant: dense DMA of the three 4-byte planes, block-sparse kernel for SemiDensePointSets_ only (sparse_download = 2)
        for k in out:
            out[k][:] = 0
        e2e_step(blocks=2)
        barrier(); tt = time.perf_counter()
        for _ in range(a.steps):
            e2e_step(blocks=2)
        barrier(); e2e["hybrid_sec"] = (time.perf_counter() - tt) / a.steps
        e2e["hybrid_same"] = all(np.array_equal(out[k][::5].view(np.uint32), ref_planes[k].view(np.uint32)) for k in out)
        if not e2e["hybrid_same"]:
            print("WARNING: hybrid download differs from the dense download", file=sys.stderr)
        e2e["hybrid_d2h"] = int(len(owned) * W * H * 12 + e2e["blocks_d2h"] // 2)
        del ref_planes
    if do_e2e and not a.no_e2e_variants:
        # the same loop with sdm_scatter_keyframes: the planes start zero-initialised like KeyFrame.cc:78-81 leaves them
        # and only the candidate pixels' records cross PCIe.  Checked against the dense result on every 7th keyframe.
        ref = {k: out[k][::7].copy() for k in out}
        for k in out:
            out[k][:] = 0
        e2e_step(sparse=True)
        same = all(np.array_equal(out[k][::7].view(np.uint32), ref[k].view(np.uint32)) for k in out)
        if not same:
            print("WARNING: sdm_scatter_keyframes planes differ from sdm_download_keyframes planes", file=sys.stderr)
        barrier(); tt = time.perf_counter()
        for _ in range(a.steps):
            e2e_step(sparse=True)
        barrier()
        e2e["sparse_sec"] = (time.perf_counter() - tt) / a.steps
        e2e["sparse_same"] = bool(same)
        e2e["sparse_d2h"] = int(28 * cands)
        # the same loop for consumers that only need the surviving points (SaveSemiDensePoints / DrawSemiDense / CARV:
        # sigma <= 0.02 and checked > 1e-6): upload + both passes + sdm_export_points instead of the dense downloads
        owned_arr = np.ascontiguousarray(owned, np.int32)
        pts_buf = pinned(lib, (len(owned) * max(40000, W * H // 6),), np.dtype([("x", "f4"), ("y", "f4"), ("z", "f4"), ("pixel", "u4")]), keep)
        tot = C.c_uint64()
        xloop = api.Loop()
        xloop.n_pass1, xloop.pass1, xloop.n_pass2, xloop.pass2 = len(items), items, len(items), items
        xloop.chunk, xloop.exchange = int(os.environ.get("SDM_BENCH_XCHUNK", "50")), loop.exchange
        up_img = ctx.upload_descs(sc, range(n_loc), images_only=True)  # grad = theta = NULL: planes made on the device

        def export_step(up):
            xloop.n_upload, xloop.upload = n_loc, up
            if host_exchange:
                ctx.upload_keyframes(up); ctx.pass1(items); exchange(); ctx.pass2(items)
            else:
                chk(lib.sdm_run_loop(ctx.h, C.byref(xloop)))
            chk(lib.sdm_export_points(ctx.h, owned_arr.size, owned_arr.ctypes.data_as(C.POINTER(C.c_int32)), 0.02,
                                      pts_buf.ctypes.data, pts_buf.size, None, C.byref(tot)))
            ctx.synchronize()  # (peers may still pull: the next pass 1 waits for their acknowledgements on the device)
            if host_exchange:
                barrier()
        for name, up in (("export", up_desc), ("image", up_img)):
            export_step(up)
            barrier(); tt = time.perf_counter()
            for _ in range(a.steps):
                export_step(up)
            e2e[name + "_sec"] = (time.perf_counter() - tt) / a.steps
            e2e[name + "_points"] = int(tot.value)
        if e2e["image_points"] != e2e["export_points"]:
            print(f"WARNING: device-produced planes gave {e2e['image_points']} points, uploaded planes {e2e['export_points']}", file=sys.stderr)
        if world == 1 and not host_exchange:
            # the reference's complete loop: it detects the edge map of every keyframe inside pass 1 (ProbabilityMapping.cc:394,
            # LineDetector::DetectEdgeMap -> closed-source EDLib.a) and only edge pixels are candidates (:454).  Here: im_ up,
            # sdm_edge_drawing with the routing walks on the device, the masks handed to the upload as device planes
            # (sdm_ed_device_edge_plane), both passes, the point cloud back.
            ed_imgs = (api.EdImage * n_loc)()
            for i in range(n_loc):
                ed_imgs[i].im, ed_imgs[i].im_step = h_im[i].ctypes.data, h_im[i].strides[0]
            up_edge = ctx.upload_descs(sc, range(n_loc), images_only=True)
            h_edge = pinned(lib, (n_loc, H, W), np.int32, keep)
            ed_ms = []

            def edge_step(on_device):
                # on_device 1: routing walks in k_ed_route, masks stay on the device; 2: walks on host threads, masks scattered on
                # the device from the chains (n_loc <= 256: one batch); 0: host threads, masks through pinned host planes
                for i in range(n_loc):
                    ed_imgs[i].edge_index, ed_imgs[i].edge_step = (None, 0) if on_device else (h_edge[i].ctypes.data, 4 * W)
                res = C.c_void_p()
                chk(lib.sdm_edge_drawing(ctx.h, n_loc, ed_imgs, 36, 8, 0, C.byref(res)))
                lib.sdm_ed_free(res)
                ed_ms.append(ctx.last_edge_drawing_ms()["wall_ms"])
                for i in range(n_loc):
                    up_edge[i].edge, up_edge[i].edge_step = (ctx.ed_device_edge_plane(i) if on_device else h_edge[i].ctypes.data), 4 * W
                export_step(up_edge)
            for on_device in ((1, 0, 2) if n_loc <= 256 else (1, 0)):
                ctx.set_edge_drawing_route(on_device)
                edge_step(on_device)
                key = {1: "edge_dev", 0: "edge_host", 2: "edge_hostdev"}[on_device]
                e2e[key + "_candidates"] = int(sum(ctx.candidate_count(s) for s in owned))
                ed_ms.clear()
                tt = time.perf_counter()
                for _ in range(a.steps):
                    edge_step(on_device)
                e2e[key + "_sec"] = (time.perf_counter() - tt) / a.steps
                e2e[key + "_points"] = int(tot.value)
                e2e[key + "_ed_ms"] = float(np.mean(ed_ms))
                if on_device == 1:
                    e2e["edge_fallbacks"] = ctx.last_edge_drawing_fallbacks()
            ctx.set_edge_drawing_route(False)
            upload(); ctx.synchronize()  # back to the bench's own planes and mask for what follows
    if rank == 0:
        clocks.stop()

    tot_cands, ms_max, e2e_max = cands, ms, (e2e["sec"] if e2e else 0.0)
    dense_max = e2e.get("blocks_sec", 0.0) if e2e else 0.0  # (block-sparse variant)
    blocks_ok = 1.0 if (e2e and e2e.get("blocks_same")) else 0.0
    sparse_max = e2e.get("sparse_sec", 0.0) if e2e else 0.0
    sparse_same = 1.0 if (e2e and e2e.get("sparse_same")) else 0.0
    hyb_max = e2e.get("hybrid_sec", 0.0) if e2e else 0.0
    par_n, par_bad = parity if parity else (0, 0)
    if world > 1:
        v = torch.tensor([float(cands), float(par_n), float(par_bad), 1.0 if parity else 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(v); tot_cands, par_n, par_bad, par_ranks = (int(x) for x in v.tolist())
        xs = [e2e.get("export_sec", 0.0), e2e.get("image_sec", 0.0)] if e2e else [0.0, 0.0]
        m = torch.tensor([ms, e2e_max, sparse_max, -sparse_same] + xs + [dense_max, -blocks_ok, hyb_max], device="cuda", dtype=torch.float64)
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
        ms_max, e2e_max, sparse_max, sparse_same, x_sec, i_sec, dense_max, blocks_ok, hyb_max = m.tolist()
        sparse_same, blocks_ok = -sparse_same, -blocks_ok
        if e2e and "export_sec" in e2e:
            pv = torch.tensor([float(e2e["export_points"]), float(e2e["image_points"])], device="cuda", dtype=torch.float64)
            dist.all_reduce(pv)
            e2e["export_sec"], e2e["image_sec"] = x_sec, i_sec
            e2e["export_points"], e2e["image_points"] = int(pv[0].item()), int(pv[1].item())
        pk = torch.tensor([pack_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(pk, op=dist.ReduceOp.MAX); timing["pack_ms"] = pk.item()
    if rank != 0:
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    n_own = len(owned)
    p1_bytes = bytes_pass1(n_own, a.nbr)
    ach = p1_bytes / (timing["pass1_scan_ms"] * 1e-3) / 1e9
    G_total = a.kf * world
    whole_bytes = bytes_total(G_total, a.nbr, a.intra)
    whole = whole_bytes / (ms_max * 1e-3) / 1e9 / world
    whole_c = whole_bytes / ((ms_max + timing["pack_ms"]) * 1e-3) / 1e9 / world
    line = {
        "metric": metric_name(), "value": tot_cands / (ms_max * 1e-3), "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": max(3, a.warmup), "ms_per_step": ms_max, "higher_is_better": True, "scaling": a.scaling,
        "vs_baseline": None, "dtype": "f32 (f64 where OpenCV accumulates in double)", "data": "synthetic",
        "config": workload_config(a, a.kf),
        "candidates_per_step": tot_cands, "keyframes_per_s": G_total / (ms_max * 1e-3),
        "image_px_per_s": G_total * W * H / (ms_max * 1e-3), "us_per_keyframe_per_gpu": 1e3 * ms_max / a.kf,
        "fused_per_step_rank0": stats["fused"], "checked_per_step_rank0": stats["checked"],
        "kernel_ms_rank0": timing,
        "roofline": {"bound": "hbm", "kernel": "k_pass1_lane (epipolar scan + hypothesis fusion)", "achieved": ach,
                     "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": ncu_traffic(a.config),
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": p1_bytes,
                     "model": "(17 + 9N) * W*H bytes per keyframe x keyframes per launch (SURVEY.md 8d)",
                     "whole_path_achieved_per_gpu": whole, "whole_path_frac": whole / peak,
                     "whole_path_frac_with_compaction": whole_c / peak,
                     "whole_path_note": "whole path = (41 + 17N [+16 intra]) * W*H bytes per keyframe / ms_per_step; with_compaction adds the "
                                        "device time of the packing + candidate-compaction kernels of one upload of the shard "
                                        "(kernel_ms_rank0.pack_ms: k_pack, k_skip; run at upload, outside the resident step)"},
        "clocks": clocks.summary(t0, t1),
        "gpu_launches": int(launches),
    }
    if shard_note:
        line["sharding"] = shard_note
    if a.config == 3:
        budget_ms = 0.5 ** -1 * whole_bytes / world / (peak * 1e9) * 1e3
        line["north_star"] = {"target": "1000 keyframes / 8 GPUs at >= 50 % of the HBM roofline",
                              "budget_ms_per_loop_at_50pct": budget_ms, "measured_ms_per_loop": ms_max,
                              "whole_path_frac": whole / peak,
                              "note": "the path is bound by instruction issue, not by HBM (DESIGN.md section 5): SURVEY.md 8(d) predicted "
                                      "5-10 % for this scan length"}
    if world > 1:
        line["parity_checked_ranks"] = par_ranks if parity is not None else 0
        line["parity_boundary_keyframes"] = par_n
        line["parity_boundary_mismatch_words"] = par_bad
        line["exchange"] = "host-ordered (synchronize + barrier + sdm_pull_halo)" if host_exchange else "device-ordered (sdm_exchange)"
    if e2e:
        line["e2e"] = {"value": tot_cands / e2e_max, "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"],
                       "d2h_bytes_per_step": e2e["d2h"], "ms_per_step": 1e3 * e2e_max,
                       "api": "sdm_run_loop (uploads, sdm_pass1, sdm_exchange, sdm_pass2, downloads pipelined inside the library) on "
                              "pinned host planes, chunks of " + str(CH) + " keyframes; every plane leaves by DMA"}
        if "blocks_sec" in e2e:
            line["e2e_blocks"] = {"value": tot_cands / dense_max, "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"],
                                  "d2h_bytes_per_step": e2e["blocks_d2h"], "ms_per_step": 1e3 * dense_max,
                                  "identical_to_dense_download": bool(blocks_ok > 0.5),
                                  "api": "the same call with sparse_download = 1: the result planes start zero-initialised as "
                                         "KeyFrame.cc:78-81 leaves them and a kernel writes only the 16-pixel blocks that hold a "
                                         "candidate pixel over PCIe (rank 0's byte count)"}
        if "hybrid_sec" in e2e:
            line["e2e_points_blocks"] = {"value": tot_cands / hyb_max, "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"],
                                         "d2h_bytes_per_step": e2e["hybrid_d2h"], "ms_per_step": 1e3 * hyb_max,
                                         "identical_to_dense_download": bool(e2e["hybrid_same"]),
                                         "api": "the same call with sparse_download = 2: depth_map_, depth_sigma_, depth_map_checked_ leave "
                                                "by DMA, SemiDensePointSets_ (zero-initialised like KeyFrame.cc:81) receives only its "
                                                "16-pixel blocks that hold a candidate, written by a kernel on a second stream"}
        if "sparse_sec" in e2e:
            # The sparse path moves 3.7x fewer bytes over PCIe but its host-side scatter touches nearly every cache line of
            # the planes at this candidate density (23 %), so it is host-memory bound and slower here; a variant.
            line["e2e_scatter"] = {"value": tot_cands / sparse_max, "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"],
                                   "d2h_bytes_per_step": e2e["sparse_d2h"], "ms_per_step": 1e3 * sparse_max,
                                   "identical_to_dense_download": bool(sparse_same > 0.5),
                                   "scatter_threads": int(os.environ.get("SDM_SCATTER_THREADS", "6")),
                                   "api": "sdm_upload_keyframes / sdm_pass1 / sdm_pass2 / sdm_scatter_keyframes: candidate records cross "
                                          "PCIe, library worker threads write them into zero-initialised planes (KeyFrame.cc:78-81)"}
        if "export_sec" in e2e:
            line["e2e_point_export"] = {"value": tot_cands / e2e["export_sec"], "unit": UNIT, "ms_per_step": 1e3 * e2e["export_sec"],
                                        "points_per_step": e2e["export_points"], "d2h_bytes_per_step": 16 * e2e["export_points"],
                                        "note": "same loop, but the result leaves as the compacted point cloud of "
                                                "sdm_export_points (sigma <= 0.02, checked > 1e-6) instead of four dense planes"}
        if "image_sec" in e2e:
            line["e2e_image_in_points_out"] = {
                "value": tot_cands / e2e["image_sec"], "unit": UNIT, "ms_per_step": 1e3 * e2e["image_sec"],
                "points_per_step": e2e["image_points"], "h2d_bytes_per_step_per_rank": int(n_loc * W * H),
                "d2h_bytes_per_step": 16 * e2e["image_points"],
                "note": "as e2e_point_export, but only im_ is uploaded: GradImg / GradTheta (KeyFrame.cc:69-74) are "
                        "produced on the device by k_pack_image"}
        if "edge_dev_sec" in e2e:
            best = min((k for k in ("edge_dev", "edge_host", "edge_hostdev") if k + "_sec" in e2e), key=lambda k: e2e[k + "_sec"])
            line["e2e_image_in_edge_drawing_points_out"] = {
                "ms_per_step": 1e3 * e2e[best + "_sec"], "edge_drawing_ms_per_step": e2e[best + "_ed_ms"],
                "routing": {"edge_dev": "device (k_ed_route)", "edge_host": "host threads, masks through host planes",
                            "edge_hostdev": "host threads, masks scattered on the device from the chains (k_ed_mask)"}[best],
                "value": e2e[best + "_candidates"] / e2e[best + "_sec"], "unit": UNIT, "candidates_per_step": e2e[best + "_candidates"],
                "points_per_step": e2e[best + "_points"], "d2h_bytes_per_step": 16 * e2e[best + "_points"],
                "routing_on_device": {"ms_per_step": 1e3 * e2e["edge_dev_sec"], "edge_drawing_ms_per_step": e2e["edge_dev_ed_ms"],
                                      "h2d_bytes_per_step": int(n_loc * W * H), "host_routed_keyframes": e2e["edge_fallbacks"],
                                      "same_result": bool(e2e["edge_dev_points"] == e2e["edge_host_points"] and
                                                          e2e["edge_dev_candidates"] == e2e["edge_host_candidates"])},
                "routing_on_host_threads": {"ms_per_step": 1e3 * e2e["edge_host_sec"], "edge_drawing_ms_per_step": e2e["edge_host_ed_ms"],
                                            "h2d_bytes_per_step": int(5 * n_loc * W * H), "d2h_bytes_extra_per_step": int(3 * n_loc * W * H)},
                "routing_on_host_threads_masks_on_device": (
                    {"ms_per_step": 1e3 * e2e["edge_hostdev_sec"], "edge_drawing_ms_per_step": e2e["edge_hostdev_ed_ms"],
                     "same_result": bool(e2e["edge_hostdev_points"] == e2e["edge_host_points"] and
                                         e2e["edge_hostdev_candidates"] == e2e["edge_host_candidates"])} if "edge_hostdev_sec" in e2e else None),
                "note": "the reference's complete loop, detector included (DetectEdgeMap inside pass 1, ProbabilityMapping.cc:394: only edge "
                        "pixels are candidates, :454): im_ up, sdm_edge_drawing, kf->mEdgeIndex to the packing kernel (a device plane when the "
                        "routing walks run on the device, a pinned host plane when they run on host threads), both passes, sdm_export_points; "
                        "the faster routing mode at this batch size is the line's value"}
    line["scan_generation"] = ctx.scan_generation()
    line["scan_long_build"] = ctx.last_scan_long()
    if world == 1 and do_e2e and not a.no_e2e_variants:
        # what the reference does per keyframe inside pass 1 before the pixel loop (LineDetector::DetectEdgeMap,
        # ProbabilityMapping.cc:394 -> LineDetector.cc:843-881, closed-source EDLib.a): sdm_edge_drawing over this rank's images
        n_ed = min(n_loc, 256)
        ims = sc.im[:n_ed]
        ctx.edge_drawing(ims[:32], edge_index=False)
        best = None
        for _ in range(3):
            offs, _pix, _edge = ctx.edge_drawing(ims, edge_index=True)
            t = ctx.last_edge_drawing_ms()
            if best is None or t["wall_ms"] < best["wall_ms"]:
                best = t
        k_s = best["kernel_ms"] * 1e-3 / n_ed
        line["edge_drawing"] = {"keyframes": n_ed, "wall_ms_per_keyframe": best["wall_ms"] / n_ed,
                                "stage1_kernel_us_per_keyframe": 1e6 * k_s,
                                "stage1_roofline": {"bound": "hbm", "achieved": 4.0 * W * H / k_s / 1e9, "peak": peak, "unit": "GB/s",
                                                    "frac": 4.0 * W * H / k_s / 1e9 / peak},
                                "routing_thread_ms_per_keyframe": best["route_thread_ms"] / n_ed,
                                "chains": int(sum(len(o) - 1 for o in offs)),
                                "device_route": None,
                                "api": "sdm_edge_drawing: smoothing / Sobel / direction / anchors on the device (k_ed_planes4, 4 algorithmic "
                                       "bytes per pixel), the sequential routing walk on host threads; chains identical to the reference's EDLib.a"}
        ctx.set_edge_drawing_route(True)  # stage 2 in k_ed_route (one warp per keyframe) instead of host threads
        ctx.edge_drawing(ims[:32], edge_index=False)
        best = None
        for _ in range(3):
            ctx.edge_drawing(ims, edge_index=False)
            t = ctx.last_edge_drawing_ms()
            if best is None or t["wall_ms"] < best["wall_ms"]:
                best = t
        line["edge_drawing"]["device_route"] = {"wall_ms_per_keyframe": best["wall_ms"] / n_ed, "route_kernel_ms": best["route_thread_ms"],
                                                "host_routed_keyframes": ctx.last_edge_drawing_fallbacks(),
                                                "note": "chains only (the mask stays on the device); one launch walks all keyframes at once, "
                                                        "its duration is that of the slowest image"}
    ctx.close()
    if world == 1 and do_e2e and not a.no_e2e_variants:
        r = class_e2e(a, sc, n_loc)
        if "semidense_loop_ms_mean" in r:
            line["e2e_class"] = {"value": tot_cands / (r["semidense_loop_ms_mean"] * 1e-3), "unit": UNIT,
                                 "ms_per_step": r["semidense_loop_ms_mean"], "ms_best": r["semidense_loop_ms_best"],
                                 "keyframes_finished": r["finished"],
                                 "api": "ProbabilityMapping::SemiDenseLoop() of the drop-in C++ class (eao-slam_b200/host/"
                                        "ProbabilityMapping.h; harness tests/cpp/test_shim.cpp --time): gating, work orders, one "
                                        "sdm_run_loop, synchronise; planes in pinned host memory, all uploaded again every repetition"}
        else:
            line["e2e_class"] = r
        # online mode (#define OnlineLoop, ProbabilityMapping.cc:42/:223-234): keyframes arrive one at a time
        n_on = min(n_loc, 64)
        nb_on = np.clip(sc.nbr_idx[:n_on], 0, n_on - 1).astype(np.int32)  # (the last keyframes never reach pass 2: not timed)
        sc_on = synth.Scene(im=sc.im[:n_on], grad=sc.grad[:n_on], theta=sc.theta[:n_on], edge=None, K=sc.K, Tcw=sc.Tcw[:n_on],
                            nbr_idx=nb_on, rot=sc.rot[:n_on], min_depth=sc.min_depth[:n_on], max_depth=sc.max_depth[:n_on])
        r = class_e2e(a, sc_on, n_on, "--time-online")
        if "ms_per_arrival_median" in r:
            r["api"] = ("drop-in C++ class in online mode (harness tests/cpp/test_shim.cpp --time-online): one keyframe is added to the "
                        "map, then SemiDenseLoop() + UpdateAllSemiDensePointSet() as the reference's Run() does per iteration; steady "
                        "state = arrivals that finish exactly one keyframe (upload of the new planes, pass 1 of one keyframe, pass 2 "
                        "of another, their planes back to pinned host memory); wall clock per arrival")
        line["e2e_online"] = r
    if world == 1 and not a.no_cpu_baseline:
        r = cpu_run(a.cpu_sample_kf, a.nbr, a.seed, a.intra, 1, 0, a.wide, a.contrast)
        line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}
    print(json.dumps(line), file=_REAL_STDOUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and a.gpus > 1:
        # not under torchrun: relaunch ourselves the way the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29531"),
               os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    # the contract is ONE JSON line on stdout; libraries (NCCL's version banner) write there too, so park the real
    # stdout and point fd 1 at stderr until the line is printed
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    a.gpus = world
    apply_config(a, world)
    if a.impl == "reference":
        main_reference(a, rank)
    else:
        main_ours(a, rank, world, local_rank)


if __name__ == "__main__":
    main()
