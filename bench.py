#!/usr/bin/env python
"""bench.py — HopperRender hot path on B200: interpolated frames/s.

One "step" = one source frame of a stream through the hot path:
    update (pack) -> calculateOpticalFlow -> the warps the filter's pacing rule asks for
(vf_HopperRender.c:371-374,481; SURVEY.md Appendix D: 25 outputs per 10 source frames at 24->60, 19 per 4 at 24->144).

  value        interpolated (= delivered, every one is a warp output) frames/s, device-resident: the source frames
               already sit in HBM (a ring larger than L2), output stays in HBM. The K steps of --steps are repeated
               `reps` times inside ONE event pair so that the timed region lasts >= 0.5 s whatever K is.
  e2e          the same metric through the reference-facing C host layer with pinned HOST planes: updateFrame (H2D) /
               calculateOpticalFlow / warpFrames / downloadFrame (D2H) per output, every call blocking.
  e2e_pageable the same with malloc'd planes (what mpv's image pool hands a filter); e2e_pinned_pool: pageable source
  frames, the output image from allocHostPlanes (the filter's pool with patches/0004); e2e_zero_copy: device planes in,
               device planes out (the IMGFMT_CUDA hand-off) — no PCIe crossing at all.
  host_ceiling what the PCIe legs alone allow (one frame up, the step's frames down, no kernels), serial and duplex.
  roofline     dominant kernel of the step by device time. The search is integer-ALU work: candidate evaluations/s
               against the packed-SAD issue rate measured live on this GPU (hr_debug_int_peak); the warp and the pack
               are HBM work: algorithmic bytes/s against the measured copy bandwidth. Kernel times come from
               event-timed loops of the serial call sequence (by difference: search only / + pack / + warps); DRAM
               traffic and instruction counts are read from the committed ncu summaries under profiles/.
  configs_measured   the other BASELINE.json configurations on one GPU: 4K P010 24->144 (blend and HSV-flow modes),
               8K P010 24->60 — value, e2e and per-kernel roofline each.
  cpu_baseline the CPU oracle (C restatement of the reference kernels, OpenMP) on a bounded sample, rank 0, N=1.
  reference_gpu  the UNMODIFIED reference (its own opticalFlowCalc.c + .cl kernels, oracle/_ref) on this same GPU through
               the NVIDIA OpenCL ICD, same call sequence, host planes: "the existing GPU kernels to beat".

`--impl reference` times the reference's CPU implementation of the path (the oracle port) on the host cores.
cpu_baseline, reference_gpu and --impl reference are the only places where this file executes anything under oracle/.
"""
import argparse
import ctypes
import json
import os
import re
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (width, height, pixfmt, source fps, display fps, mode)
    "1080p-nv12-24to60": (1920, 1080, 0, 24.0, 60.0, 2),
    "4k-nv12-24to60": (3840, 2160, 0, 24.0, 60.0, 2),
    "4k-p010-24to60": (3840, 2160, 1, 24.0, 60.0, 2),
    "4k-p010-24to144": (3840, 2160, 1, 24.0, 144.0, 2),
    "4k-p010-24to144-hsv": (3840, 2160, 1, 24.0, 144.0, 3),
    "8k-p010-24to60": (7680, 4320, 1, 24.0, 60.0, 2),
}
EXTRA_CONFIGS = ["4k-p010-24to60", "4k-p010-24to144", "4k-p010-24to144-hsv", "8k-p010-24to60"]
L2_BYTES = 126 * 1024 * 1024
MIN_REGION_S = 0.5
KERNEL_OF = {"search": "flow_search", "warp": "warp_fast_kernel", "pack": "pack_frame16_kernel"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="1080p-nv12-24to60", choices=sorted(WORKLOADS))
    ap.add_argument("--radius", type=int, default=5, help="search radius (config.h MIN_SEARCH_RADIUS = 5 is the default)")
    ap.add_argument("--cpu-sample-steps", type=int, default=400)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip configs_measured / reference_gpu / host_ceiling (N = 1 extras)")
    ap.add_argument("--serial", action="store_true", help="device-resident loop without the pipelined mode (one kernel after the other)")
    ap.add_argument("--bands", action="store_true",
                    help="split every frame into spatial bands over the N ranks (8K config, SURVEY.md §8e): strong scaling")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback of B200_PROFILING.md"


def ncu_summaries():
    """{(workload tag, kernel key): {metric: value}} from the committed ncu summaries profiles/r02_ncu_<tag>.txt
    (written by profiles/extract.py): the newest capture of each kernel per tag."""
    out = {}
    pdir = os.path.join(ROOT, "profiles")
    for fn in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        m = re.match(r"r02_ncu_(.+)\.txt$", fn)
        if not m:
            continue
        cur = None
        for line in open(os.path.join(pdir, fn)):
            if not line.startswith(" "):
                name = line.strip()
                cur = next((k for k, pat in KERNEL_OF.items() if pat in name), None)
                if cur:
                    out[(m.group(1), cur)] = {"file": "profiles/" + fn, "kernel": name}
            elif cur:
                parts = line.split()
                if len(parts) >= 2:
                    try:
                        val = float(parts[1])
                    except ValueError:
                        continue
                    unit = parts[2] if len(parts) > 2 else ""
                    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(unit)
                    out[(m.group(1), cur)][parts[0]] = val * scale if scale else val
    return out


def ncu_tag(workload):
    return workload.replace("-24to60", "").replace("-24to144", "").replace("-hsv", "")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            time.sleep(0.05)          # the first sample is on its way when the timed region starts
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 6:
                self.samples.append((time.perf_counter(), parts))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [s for ts, s in self.samples if t0 is None or (t0 <= ts <= t1 + 0.03)]
        for s in inside or [s for _, s in self.samples]:
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
            except ValueError:
                continue
            for n, v in zip(names, s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_inside_timed_region": len(inside)}


def pacing_ts(n_steps, src_fps, disp_fps):
    """Blend scalars for n_steps source frames in steady state (the first source frame of a
    stream produces no warp, so one extra frame primes the pacer)."""
    import hr_pkg

    hr_pkg.load()
    from hopperrender_b200 import pacing

    p = pacing.Pacer(src_fps, disp_fps)
    p.next_source_frame()
    return [p.next_source_frame() for _ in range(n_steps)]


def warp_bytes(w, h, bps, lw, lh, mode=2):
    """Algorithmic bytes of one output frame: both source frames read, one frame written, the flow read. The HSV flow
    mode reads no chroma (its chroma plane is a per-cell constant) but the colour table as well."""
    if mode == 3:
        return 3 * w * h * bps + int(0.5 * w * h * bps) + 8 * lw * lh
    return 3 * int(1.5 * w * h * bps) + 4 * lw * lh


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """CPU arm: the reference's algorithm on the host cores (the oracle port; the reference's own OpenCL host cannot run
    without an OpenCL CPU runtime, which the image does not have)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # all the host threads it can use: torchrun exports OMP_NUM_THREADS=1 to its workers
    os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
    from oracle import hr_oracle_py as O
    import hr_pkg

    hr_pkg.load()
    from hopperrender_b200 import synth

    w, h, pixfmt, sfps, dfps, mode = WORKLOADS[args.workload]
    clip = synth.MovingTextureClip(w, h, pixfmt=pixfmt)
    nfr = 8
    frames = [clip.frame(k) for k in range(nfr)]
    o = O.Oracle(h, w, w, pixfmt)
    o.update_frame(*frames[0])
    o.update_frame(*frames[1])
    # keep the whole run within minutes: the CPU does ~0.1 s per 1080p step
    steps = max(1, min(args.steps, 40))
    warm = max(1, min(args.warmup, 3))
    ts = pacing_ts(warm + steps, sfps, dfps)

    def step(i):
        y, uv = frames[(i + 2) % nfr]
        o.update_frame(y, uv)
        o.calc_flow(args.radius, 8, 6)
        for t in ts[i]:
            o.warp(np.float32(t), mode)
            o.download()
        return len(ts[i])

    for i in range(warm):
        step(i)
    t0 = time.perf_counter()
    outs = sum(step(warm + i) for i in range(steps))
    dt = time.perf_counter() - t0
    val = outs / dt
    line = {
        "impl": "reference", "metric": "interpolated frames/s", "value": val, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8" if pixfmt == 0 else "u16", "data": "synthetic",
        "config": {"workload": args.workload, "search_radius": args.radius, "mode": mode, "device": "host cpu"},
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": O.num_threads(), "kind": "port",
                         "sample": "%d source frames (%d outputs) of %s; C restatement of the reference kernels, OpenMP — no OpenCL runtime in the image" % (steps, outs, args.workload)},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(index):
    """Run this process on the CPU cores next to GPU `index` (NVML's CPU affinity), so that the pinned host
    frames of the end-to-end leg are allocated on the GPU's own NUMA node. Best effort."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


class Env:
    """What every measurement of a run shares: torch, the package, rank / world, the stream."""

    def __init__(self, args):
        import torch
        import hr_pkg

        self.torch = torch
        self.hr = hr_pkg.load()
        from hopperrender_b200 import sharding, synth

        self.synth, self.sharding = synth, sharding
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback")
        self.numa_cpus = bind_to_gpu_numa_node(self.local)
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist_mod

            self.dist = dist_mod
            self.dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.stream = torch.cuda.Stream()
        self.dev = torch.device("cuda", self.local)
        self.peaks, self.peaks_src = peaks()
        self.ncu = ncu_summaries()
        self._int_peak = None

    def barrier(self):
        self.stream.synchronize()
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()

    def reduce(self, count, seconds):
        return self.sharding.reduce_throughput(count, seconds, self.dist, self.dev)

    def max_over_ranks(self, x):
        return self.reduce(0, x)[1]

    def int_peak(self):
        if self._int_peak is None:
            self.hr.debug_int_peak(self.local)
            self._int_peak = self.hr.debug_int_peak(self.local)
        return self._int_peak


def measure(env, workload, K, W, radius, with_e2e=True, serial_only=False, banded=False, full=True):
    """Device-resident value, per-kernel times and the end-to-end legs of one workload on this rank's GPU.
    Returns the rank-local pieces; rank 0 assembles the whole-job numbers from the reductions done here."""
    torch, hr, synth = env.torch, env.hr, env.synth
    stream, local, world, rank, dist = env.stream, env.local, env.world, env.rank, env.dist
    w, h, pixfmt, sfps, dfps, mode = WORKLOADS[workload]
    bps = 2 if pixfmt else 1
    tdtype = torch.uint16 if pixfmt else torch.uint8
    frame_bytes = int(1.5 * w * h * bps)
    # ring of source frames larger than L2, so every step reads its inputs from HBM
    nring = min(96, max(8, (2 * L2_BYTES) // frame_bytes + 2))
    clip = synth.MovingTextureClip(w, h, pixfmt=pixfmt)
    nbase = 8 if frame_bytes < (64 << 20) else 4
    base = [clip.frame(k) for k in range(nbase)]
    g = hr.HrCuda(h, w, w, pixfmt, device=local)
    g.set_stream(stream.cuda_stream)
    lw, lh = g.info.lowWidth, g.info.lowHeight
    r0, r1 = 0, h
    rows = None
    if banded:
        rows = env.sharding.band_rows(h, world, g.info.resScalar)
        hr.connect_bands_distributed(g, dist, rows, max_radius=max(radius, 5))
        r0, r1 = rows[rank]
        nring = max(8, min(96, nring * world))      # a rank keeps only its band of every ring frame
    with torch.cuda.stream(stream):
        ring = []
        for k in range(nring):
            y, uv = base[k % nbase]
            # banded: only this rank's rows live on this GPU
            ring.append((torch.from_numpy(np.ascontiguousarray(y[r0:r1])).to("cuda", non_blocking=False).view(tdtype),
                         torch.from_numpy(np.ascontiguousarray(uv[r0 >> 1:r1 >> 1])).to("cuda", non_blocking=False).view(tdtype)))
        out_ring = [(torch.empty((h, w), dtype=tdtype, device="cuda"), torch.empty((h // 2, w), dtype=tdtype, device="cuda"))
                    for _ in range(max(4, (L2_BYTES // frame_bytes) + 2))]
    stream.synchronize()
    ts = pacing_ts(W + K, sfps, dfps)
    interp_share = sum(1 for i in range(W, W + K) for t in ts[i] if t > 1e-6) / max(1, sum(len(ts[i]) for i in range(W, W + K)))
    oi = [0]

    def feed(i):
        y, uv = ring[i % nring]
        if banded:
            g.band_upload(y, uv, device=True)       # own band (device copy) ...
            g.band_gather(blocking=False)           # ... the others by P2P, then pack
        else:
            g.update_frame_device(y, uv, borrow=True)

    def step_device(i):
        n = len(ts[i])
        feed(i)
        g.calc_flow(radius, 8, 6, blocking=False)
        outs = [out_ring[(oi[0] + j) % len(out_ring)] for j in range(n)]
        oi[0] += n
        if n:
            g.warp_batch(ts[i], outs, mode)
        return n

    CHUNK = 25

    def steps_device(i0, count):
        """`count` source frames from step i0, enqueued CHUNK frames per C call (hr_steps_device)."""
        if banded:
            return sum(step_device(i0 + j) for j in range(count))
        total = 0
        for c0 in range(i0, i0 + count, CHUNK):
            c1 = min(c0 + CHUNK, i0 + count)
            tl = ts[c0:c1]
            n = sum(len(t) for t in tl)
            outs = [out_ring[(oi[0] + j) % len(out_ring)] for j in range(n)]
            oi[0] += n
            g.steps_device([ring[j % nring] for j in range(c0, c1)], tl, outs, radius=radius, mode=mode)
            total += n
        return total

    def timed(fn, reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        n = 0
        for _ in range(reps):
            n += fn()
        g.pipeline_join()              # the main stream now follows every internal stream: e1 closes the whole region
        e1.record(stream)
        g.synchronize()
        return n, e0.elapsed_time(e1) * 1e-3

    res = {"workload": workload, "lw": lw, "lh": lh, "frame_bytes": frame_bytes, "bps": bps, "rows": (r0, r1), "mode": mode,
           "interp_share": interp_share, "nring": nring, "iterations": g.info.iterations, "smCount": g.info.smCount, "chunk": CHUNK}
    # ---- device-resident timed region ---------------------------------------------------------
    pipelined = (not banded) and (not env.args.serial) and not serial_only
    with torch.cuda.stream(stream):
        feed(nring - 1)
        g.set_pipeline(pipelined)
        steps_device(0, W)
        g.synchronize()
        # pilot: how long do K steps take? -> repetitions for a region of MIN_REGION_S (the same on every rank)
        _, pilot = timed(lambda: steps_device(W, K), 1)
        reps = max(1, int(np.ceil(MIN_REGION_S * 1.4 / max(env.max_over_ranks(pilot), 1e-6))))
        env.barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        l0 = g.launch_count()
        t_a = time.perf_counter()
        outs, secs = timed(lambda: steps_device(W, K), reps)
        t_b = time.perf_counter()
        env.barrier()
        res["clocks"] = sampler.stop(t_a, t_b) if rank == 0 else None
        res.update(outs=outs, secs=secs, reps=reps, launches=g.launch_count() - l0, pipelined=pipelined)
        g.set_pipeline(False)
        res["serial_s_per_output"] = None
        if pipelined and full:         # the same steps, one kernel after the other (what the blocking interface sees)
            steps_device(0, W)
            g.synchronize()
            ks = min(K, 100)
            so, st = timed(lambda: steps_device(W, ks), max(1, int(0.1 / max(pilot * ks / K, 1e-6))))
            res["serial_s_per_output"] = st / max(1, so)

    # ---- per-kernel device time, by difference -------------------------------------------------------
    # One CUDA event pair around a single ~5-40 us launch adds several us of front-end latency to it, so
    # the three kernels are timed over whole loops instead (two events per loop, stream kept full):
    #   A: search only            B: pack + search            C: pack + search + warps (one launch per source frame)
    #   search = A / n,  pack = (B - A) / n,  warp = (C - B) / (number of warp launches); outputs per launch recorded
    with torch.cuda.stream(stream):
        nk = min(K, 100)
        # the timed region above ran pipelined: its searches are launches of the third generation (two of them share the
        # SMs); time THAT kernel here, one launch after the other (left to itself a lone launch would pick the first)
        search_gen = 3 if (pipelined and 5 <= radius <= 16) else 1
        g.set_search_generation(search_gen)

        def loop(with_pack, with_warp):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            nrep = max(2, int(np.ceil(0.05 / max(pilot * nk / K, 1e-6))) + 1)
            for rep in range(nrep):                   # the first repetition warms the variant up
                if rep == 1:
                    e0.record(stream)
                    nw = nl = 0
                for i in range(nk):
                    if with_pack:
                        feed(W + i)
                    g.calc_flow(radius, 8, 6, blocking=False)
                    if with_warp and ts[W + i]:
                        outs_ = [out_ring[(oi[0] + j) % len(out_ring)] for j in range(len(ts[W + i]))]
                        oi[0] += len(outs_)
                        g.warp_batch(ts[W + i], outs_, mode)
                        if rep >= 1:
                            nw += len(outs_)
                            nl += 1
            e1.record(stream)
            env.barrier()
            return e0.elapsed_time(e1) / (nrep - 1), nw / (nrep - 1), nl / (nrep - 1)

        tA, _, _ = loop(False, False)
        tB, _, _ = loop(True, False)
        tC, nwarps, nlaunch = loop(True, True)
        res["kernel_ms"] = {"search": tA / nk, "pack": max(tB - tA, 0.0) / nk, "warp": max(tC - tB, 0.0) / max(1.0, nlaunch)}
        res["kernel_count"] = {"search": nk, "pack": nk, "warp": nlaunch}
        res["warp_outputs_per_launch"] = nwarps / max(1.0, nlaunch)
        res["search_generation"] = g.last_search_generation()
        g.set_search_generation(0)
    g.set_output_device(None, None)

    # ---- end-to-end through the reference-facing interface, host buffers -------------------------
    # The compiled C host layer (opticalFlowCalc.c, the drop-in for the reference's file of that name) driven by
    # hrReplay.c, the filter's per-frame call sequence as a C loop: updateFrame (H2D) / calculateOpticalFlow /
    # warpFrames / downloadFrame (D2H) per output, every call blocking like the reference's.
    if with_e2e and not env.args.no_e2e:
        def host_frame(pinned):
            ny, nuv = (r1 - r0) * w, ((r1 >> 1) - (r0 >> 1)) * w
            # host frames as mpv's image pool lays them out: one allocation, the UV plane right behind the Y plane
            buf = torch.empty(ny + nuv, dtype=tdtype)
            if pinned:
                buf = buf.pin_memory()
            return buf[:ny].view(r1 - r0, w), buf[ny:].view((r1 >> 1) - (r0 >> 1), w)

        def host_ring(pinned):
            hring = []
            for k in range(nbase):
                y, uv = base[k]
                ty, tuv = host_frame(pinned)
                ty.copy_(torch.from_numpy(np.ascontiguousarray(y[r0:r1])).view(tdtype))
                tuv.copy_(torch.from_numpy(np.ascontiguousarray(uv[r0 >> 1:r1 >> 1])).view(tdtype))
                hring.append((ty, tuv))
            return hring, host_frame(pinned)

        ets = pacing_ts(5 + 4000, sfps, dfps)
        if banded:
            hring, hout = host_ring(True)
            ofc = hr.OpticalFlowCalc()
            if hr.initOpticalFlowCalc(ofc, h, w, w, pixfmt, device=local):
                raise SystemExit("initOpticalFlowCalc failed")
            ofc.opticalFlowSearchRadius = radius
            hr.connect_bands_distributed(ofc.impl, dist, rows, max_radius=max(radius, 5))

            def step_host(i):                               # a rank moves only its band over PCIe
                ofc.impl.band_upload(*hring[i % nbase])
                ofc.impl.band_gather(blocking=True)
                assert not hr.calculateOpticalFlow(ofc)
                for t in ets[i]:
                    assert not hr.warpFrames(ofc, t, mode)
                    ofc.impl.band_download(hout[0], hout[1])
                return len(ets[i])

            step_host(nbase - 1)
            for i in range(5):
                step_host(i)
            Ke = min(K, 100)
            env.barrier()
            t0 = time.perf_counter()
            eouts = sum(step_host(5 + i) for i in range(Ke))
            torch.cuda.synchronize()
            edt = time.perf_counter() - t0
            hr.freeOFC(ofc)
            res["e2e"] = {"outs": eouts, "secs": edt, "steps": Ke,
                          "api": "band_upload/band_gather/calculateOpticalFlow/warpFrames/band_download per rank, pinned host planes"}
        else:
            lib = hr.load_ofc_library()

            def replay_leg(pinned, pool_out=False):
                hring, hout = host_ring(pinned)
                pool = None
                if pool_out:        # the output image as the filter's pool allocates it with patches/0004 (allocHostPlanes)
                    nbytes = (hout[0].numel() + hout[1].numel()) * hout[0].element_size()
                    pool = lib.allocHostPlanes(nbytes + 64)
                    if not pool:
                        raise SystemExit("allocHostPlanes failed")
                    flat = np.ctypeslib.as_array(ctypes.cast(pool, ctypes.POINTER(ctypes.c_uint8)), (nbytes,)).view(base[0][0].dtype)
                    ny = hout[0].numel()
                    hout = (flat[:ny].reshape(tuple(hout[0].shape)), flat[ny:].reshape(tuple(hout[1].shape)))
                cofc = hr.COpticalFlowCalc()
                cofc.pixelFormat = pixfmt
                cofc.cudaDevice = local + 1
                if lib.initOpticalFlowCalc(ctypes.byref(cofc), h, w, w):
                    raise SystemExit("initOpticalFlowCalc (C host layer) failed")
                cofc.opticalFlowSearchRadius = radius
                hr.replay_stream_c(cofc, hring, nbase - 1, [[]], mode, hout)           # the first frame of the stream
                t0 = time.perf_counter()
                hr.replay_stream_c(cofc, hring, 0, ets[:5], mode, hout)
                per_step = (time.perf_counter() - t0) / 5
                Ke = int(min(4000, max(20, np.ceil(0.3 / max(per_step, 1e-6)))))      # >= 0.3 s of calls
                env.barrier()
                t0 = time.perf_counter()
                eouts = hr.replay_stream_c(cofc, hring, 5, ets[5:5 + Ke], mode, hout)
                edt = time.perf_counter() - t0
                lib.freeOFC(ctypes.byref(cofc))
                if pool:
                    lib.freeHostPlanes(None, pool)
                return {"outs": eouts, "secs": edt, "steps": Ke}

            res["e2e"] = dict(replay_leg(True), api=("libhopperrender_ofc.so: initOpticalFlowCalc, then hrReplayStream = updateFrame/calculateOpticalFlow/"
                                                     "warpFrames/downloadFrame in the filter's order, pinned host planes, every call blocking like the reference's"))
            if full:
                res["e2e_pageable"] = dict(replay_leg(False), api="the same calls with malloc'd (pageable) planes, as mpv's image pool delivers them")
                res["e2e_pinned_pool"] = dict(replay_leg(False, pool_out=True),
                                              api=("the same calls, pageable source frames (a software decoder's) and the output image from allocHostPlanes: "
                                                   "the filter's output pool with patches/0004"))
                res["e2e_zero_copy"] = zero_copy_leg(env, g, ring, out_ring, ets, nring, radius, mode)
    if banded:
        lo, hi, _ = g.band_halo()
        res["band"] = {"rows": [r0, r1], "held_rows": [lo, hi],
                       "nvlink_halo_bytes_per_frame": int(((r0 - lo) + (hi - r1)) * w * bps * 1.5)}
    g.close()
    del ring, out_ring
    torch.cuda.empty_cache()
    return res


def zero_copy_leg(env, g, ring, out_ring, ets, nring, radius, mode):
    """The IMGFMT_CUDA hand-off as the patched filter drives it: source frames arrive as device planes
    (hr_update_frame_device, borrowed), every output is warped into a device image of the pool (hr_set_output_device) and
    handed on without a download; the calls keep the filter's order and block where the filter would read a result
    (calculateOpticalFlow's timing). No PCIe crossing."""
    torch = env.torch
    with torch.cuda.stream(env.stream):
        g.set_pipeline(True)
        g.update_frame_device(*ring[nring - 1], borrow=True)
        oi = 0

        def step(i):
            nonlocal oi
            g.update_frame_device(*ring[i % nring], borrow=True)
            g.calc_flow(radius, 8, 6, blocking=True)
            for t in ets[i]:
                oy, ouv = out_ring[oi % len(out_ring)]
                oi += 1
                g.set_output_device(oy, ouv)
                g.warp(t, mode)
            g.synchronize()        # the frame leaves the filter: its planes must be complete
            return len(ets[i])

        for i in range(5):
            step(i)
        t0 = time.perf_counter()
        for i in range(5, 25):
            step(i)
        per = (time.perf_counter() - t0) / 20
        n = int(min(3000, max(20, np.ceil(0.3 / max(per, 1e-6)))))
        env.barrier()
        t0 = time.perf_counter()
        outs = sum(step(25 + i) for i in range(n))
        dt = time.perf_counter() - t0
        g.set_output_device(None, None)
        g.set_pipeline(False)
    return {"outs": outs, "secs": dt, "steps": n,
            "api": "hr_update_frame_device(borrow) / hr_calc_flow(blocking) / hr_set_output_device + hr_warp per output / hr_synchronize per source frame: device planes in and out"}


def host_ceiling(env, workload, frac=1.0):
    """What the PCIe legs of one step allow on this rank with no kernel at all: one frame host->device, the step's
    outputs device->host, pinned memory. serial = one copy after the other with a wait each (how the blocking calls
    drive them); duplex = upload and downloads on two streams at once (the bound of any asynchronous host interface)."""
    torch = env.torch
    w, h, pixfmt, sfps, dfps, _ = WORKLOADS[workload]
    nbytes = int(1.5 * w * h * (2 if pixfmt else 1) * frac) & ~255     # a band's share of the frame
    per_step = dfps / sfps
    hin = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    hout = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    din = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dout = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    ndown = [3, 2] if abs(per_step - 2.5) < 1e-9 else [int(round(per_step))]
    steps = max(10, int(0.25 / (nbytes * (1 + per_step) / 40e9)))

    def serial():
        for i in range(steps):
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)
            s1.synchronize()
            for _ in range(ndown[i % len(ndown)]):
                with torch.cuda.stream(s1):
                    hout.copy_(dout, non_blocking=True)
                s1.synchronize()

    def duplex():
        for i in range(steps):
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)
            with torch.cuda.stream(s2):
                for _ in range(ndown[i % len(ndown)]):
                    hout.copy_(dout, non_blocking=True)
        s1.synchronize()
        s2.synchronize()

    out = {}
    for name, fn in (("serial", serial), ("duplex", duplex)):
        fn()
        env.barrier()
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        outs = sum(ndown[i % len(ndown)] for i in range(steps))
        tot, mx = env.reduce(outs, dt)
        out[name + "_frames_per_s"] = tot / mx
        out[name + "_gb_per_s_per_gpu"] = nbytes * (steps + outs) / dt / 1e9
    out["note"] = "pinned host memory, %d concurrent processes (one per GPU), %.1f MB per frame; no kernels" % (env.world, nbytes / 1e6)
    return out


def kernel_rooflines(env, res, radius, step_ms=None):
    """Per-kernel roofline blocks from the by-difference kernel times of `res`; step_ms: the pipelined device loop's time per
    source frame (all kernels overlapped), for the issue-slot share of the whole step."""
    pk = env.peaks
    w, h, pixfmt, _, _, _ = WORKLOADS[res["workload"]]
    bps, lw, lh = res["bps"], res["lw"], res["lh"]
    r0, r1 = res["rows"]
    kms, kcount = res["kernel_ms"], res["kernel_count"]
    per_step = {"pack": kms["pack"], "search": kms["search"], "warp": kms["warp"] * (kcount["warp"] / max(1, kcount["search"]))}
    tot = max(1e-12, sum(per_step.values()))
    opl = res["warp_outputs_per_launch"]
    wbytes = int(warp_bytes(w, h, bps, lw, lh, res["mode"]) * ((r1 - r0) / h))   # a band's launch moves the band's rows
    alg = {"warp": wbytes * opl, "pack": int(1.5 * w * h * bps) + 4 * w * h}
    tag = ncu_tag(res["workload"])
    roof = {}
    for k in ("pack", "warp"):
        if kms[k] > 0:
            ach = alg[k] / (kms[k] * 1e-3) / 1e9
            n = env.ncu.get((tag, k), {})
            traffic = (n.get("dram__bytes_read.sum", 0.0) + n.get("dram__bytes_write.sum", 0.0)) if n else None
            roof[k] = {"kernel": KERNEL_OF[k], "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                       "traffic": traffic, "traffic_source": n.get("file"), "avg_us": kms[k] * 1e3, "share_of_step": per_step[k] / tot,
                       "algorithmic_bytes": alg[k], "peak_source": env.peaks_src}
    if roof.get("warp"):
        roof["warp"]["outputs_per_launch"] = opl
        roof["warp"]["traffic_note"] = "ncu traffic is per single-output launch" if roof["warp"]["traffic"] else None
    if kms["search"] > 0:
        evals = 2 * res["iterations"] * radius * lw * lh
        peak_evals, per_clk = env.int_peak()
        ach = evals / (kms["search"] * 1e-3)
        n = env.ncu.get((tag, "search"), {}) if radius == 5 else {}
        hbm_alg = 4 * lw * lh + int(1.5 * w * h) + 4 * lw * lh    # frame-2 lattice words + reachable frame-1 samples + flow out
        gen = res.get("search_generation", 1)
        blk = {"kernel": ("flow_search3_kernel<%d>" if gen == 3 else "flow_search_kernel<%d>") % radius, "bound": "int_alu", "achieved": ach / 1e9, "peak": peak_evals / 1e9, "unit": "G candidate evaluations/s",
               "frac": ach / peak_evals, "avg_us": kms["search"] * 1e3, "share_of_step": per_step["search"] / tot,
               "algorithmic_evaluations": evals,
               "peak_source": "hr_debug_int_peak, this run: a kernel of nothing but independent VABSDIFF4.U8.ACC chains at full occupancy = %.2f warp instructions per SM clock" % per_clk,
               "traffic": (n.get("dram__bytes_read.sum", 0.0) + n.get("dram__bytes_write.sum", 0.0)) if n else None, "traffic_source": n.get("file"),
               "hbm": {"algorithmic_bytes": hbm_alg, "achieved_gbs": hbm_alg / (kms["search"] * 1e-3) / 1e9, "frac": hbm_alg / (kms["search"] * 1e-3) / 1e9 / pk["hbm_gbs"]},
               "note": ("one packed SAD per candidate evaluation is all the reference's arithmetic asks for; the launch is bound by 16 strictly dependent "
                        "steps (tile-to-tile hand-offs, block barriers) and by the instructions around each SAD, not by the SAD pipe; avg_us is one launch "
                        "alone — in the pipelined device loop three launches (consecutive frame pairs) share the SMs, see config.flow_ms_per_pair_pipelined and issue.pipelined_step")}
        if n.get("smsp__inst_executed.sum"):
            slots = pk.get("sm_max_mhz", 1965.0) * 1e6 * 4 * res["smCount"] * (kms["search"] * 1e-3)
            blk["issue"] = {"warp_instructions_per_launch": n["smsp__inst_executed.sum"], "issue_slots_in_launch": slots, "frac": n["smsp__inst_executed.sum"] / slots,
                            "instructions_per_evaluation": n["smsp__inst_executed.sum"] * 32 / evals,
                            "source": "smsp__inst_executed.sum of %s; 4 schedulers x SMs x max clock x measured launch time" % n.get("file")}
            npk, nwp = env.ncu.get((tag, "pack"), {}), env.ncu.get((tag, "warp"), {})
            if step_ms and res.get("pipelined") and npk.get("smsp__inst_executed.sum") and nwp.get("smsp__inst_executed.sum"):
                # every kernel of a source frame: one search, one pack, the warps (the ncu summary holds a single-output launch)
                wps = kcount["warp"] * opl / max(1, kcount["search"])
                inst = n["smsp__inst_executed.sum"] + npk["smsp__inst_executed.sum"] + nwp["smsp__inst_executed.sum"] * wps
                step_slots = pk.get("sm_max_mhz", 1965.0) * 1e6 * 4 * res["smCount"] * (step_ms * 1e-3)
                blk["issue"]["pipelined_step"] = {"warp_instructions_per_source_frame": inst, "issue_slots_in_step": step_slots, "frac": inst / step_slots,
                                                  "outputs_per_source_frame": wps,
                                                  "note": "the searches of three consecutive pairs, the pack and the warps share the SMs in the device loop: instructions of all kernels of one source frame over the issue slots of the measured time per source frame"}
        roof["search"] = blk
    dom = max(per_step, key=per_step.get)
    return roof, dom


def summarize(env, res, radius, K, W):
    """Whole-job figures of one measure() result (collective: every rank calls it)."""
    world, banded = env.world, res.get("banded", False)
    tot_outs, max_s = env.reduce(res["outs"], res["secs"])
    launches, _ = env.reduce(res["launches"], 0.0)
    if banded:          # every rank produced a band of the SAME frames: count each frame once
        tot_outs //= world
    out = {"value": tot_outs / max_s, "ms_per_step": max_s * 1e3 / (K * res["reps"]), "reps": res["reps"], "timed_region_s": max_s,
           "gpu_launches": int(launches), "outs": tot_outs}
    for key in ("e2e", "e2e_pageable", "e2e_pinned_pool", "e2e_zero_copy"):
        if key in res:
            e_outs, e_dt = env.reduce(res[key]["outs"], res[key]["secs"])
            if banded:
                e_outs //= world
            out[key] = {"value": e_outs / e_dt, "steps": res[key]["steps"], "api": res[key]["api"], "outs": e_outs}
    return out


def run_ours(args):
    env = Env(args)
    torch, hr = env.torch, env.hr
    rank, world = env.rank, env.world
    K, W_ = args.steps, max(3, args.warmup)
    radius = args.radius
    banded = bool(args.bands) and world > 1
    w, h, pixfmt, sfps, dfps, mode = WORKLOADS[args.workload]

    res = measure(env, args.workload, K, W_, radius, banded=banded)
    res["banded"] = banded
    summ = summarize(env, res, radius, K, W_)
    frame_bytes = res["frame_bytes"]
    r0, r1 = res["rows"]
    band_frac = (r1 - r0) / h

    extras = []
    ceiling = None
    if not args.no_extra:
        ceiling = host_ceiling(env, args.workload, band_frac)
        if banded:                   # every rank moves a band of the SAME frames
            for k in ("serial_frames_per_s", "duplex_frames_per_s"):
                ceiling[k] /= world
    if world == 1 and not args.no_extra:
        shared = {}
        for name in EXTRA_CONFIGS:
            if name == args.workload:
                continue
            ke = 40 if "8k" in name else 100
            r = measure(env, name, ke, 5, radius, full=True)
            s = summarize(env, r, radius, ke, 5)
            roof, dom = kernel_rooflines(env, r, radius, s["ms_per_step"])
            ew, eh, epf, esf, edf, emode = WORKLOADS[name]
            blk = {"workload": name, "frame": "%dx%d" % (ew, eh), "mode": emode, "search_radius": radius, "steps": ke, "reps": s["reps"], "timed_region_s": s["timed_region_s"],
                   "value": s["value"], "unit": "frames/s", "ms_per_step": s["ms_per_step"],
                   "serial_frames_per_s": (1.0 / r["serial_s_per_output"]) if r["serial_s_per_output"] else None,
                   "interp_only_frames_per_s": s["value"] * r["interp_share"], "kernels": roof, "dominant_kernel": dom}
            for key in ("e2e", "e2e_pageable", "e2e_pinned_pool", "e2e_zero_copy"):
                if key in s:
                    blk[key] = {"value": s[key]["value"], "unit": "frames/s", "steps": s[key]["steps"]}
            if "e2e" in s:
                blk["e2e"].update(h2d_bytes_per_step=r["frame_bytes"], d2h_bytes_per_step=int(r["frame_bytes"] * edf / esf))
            extras.append(blk)

    bands_block = None
    if world > 1 and not banded and not args.no_extra and not os.environ.get("HR_BENCH_NO_BANDS"):
        # the 8K configuration of BASELINE.json on the same N GPUs, every frame split into N spatial bands (strong scaling):
        # a secondary block of this line, so that a scaling run records it next to the independent-streams figure
        bname = "8k-p010-24to60"
        br = measure(env, bname, 40, 5, radius, banded=True, full=False)
        br["banded"] = True
        bs = summarize(env, br, radius, 40, 5)
        halo = env.reduce(br["band"]["nvlink_halo_bytes_per_frame"], 0.0)[0]
        bceil = host_ceiling(env, bname, (br["rows"][1] - br["rows"][0]) / WORKLOADS[bname][1])
        if rank == 0:
            broof, bdom = kernel_rooflines(env, br, radius)
            bands_block = {"workload": bname, "partition": "%d spatial bands of whole lattice tile rows; halo rows by NVLink P2P, band-sharded search with tile totals / edge windows / flow stored into the peers' arenas" % world,
                           "scaling": "strong", "value": bs["value"], "unit": "frames/s", "ms_per_step": bs["ms_per_step"], "reps": bs["reps"], "timed_region_s": bs["timed_region_s"],
                           "e2e": {"value": bs["e2e"]["value"], "unit": "frames/s", "steps": bs["e2e"]["steps"]} if "e2e" in bs else None,
                           "nvlink_halo_bytes_per_frame_all_gpus": int(halo), "rank0_rows": br["band"]["rows"], "rank0_held_rows": br["band"]["held_rows"],
                           "host_ceiling": {k: (v / world if k.endswith("frames_per_s") else v) for k, v in bceil.items()},
                           "kernels_rank0": {k: {"avg_us": v["avg_us"], "frac": v["frac"], "bound": v["bound"]} for k, v in broof.items()}}

    if rank == 0:
        roof, dom = kernel_rooflines(env, res, radius, summ["ms_per_step"])
        line = {
            "metric": "interpolated frames/s", "value": summ["value"], "unit": "frames/s", "n_gpus": world,
            "steps": K, "warmup": W_, "reps": summ["reps"], "timed_region_s": summ["timed_region_s"], "ms_per_step": summ["ms_per_step"],
            "higher_is_better": True, "scaling": "strong" if banded else "weak", "vs_baseline": None,
            "dtype": "u8" if pixfmt == 0 else "u16", "data": "synthetic",
            "config": {"workload": args.workload, "frame": "%dx%d" % (w, h), "search_radius": radius, "mode": mode,
                       "streams_per_gpu": 1, "partition": ("%d spatial bands, NVLink P2P" % world) if banded else ("independent streams" if world > 1 else "none"),
                       "cache": "source ring of %d frames (%d MB) and output ring exceed the 126 MB L2" % (res["nring"], res["nring"] * frame_bytes >> 20),
                       "timed_region": "the %d steps repeated %d times inside one CUDA event pair (>= %.1f s)" % (K, summ["reps"], MIN_REGION_S),
                       "flow_ms_per_pair": res["kernel_ms"]["search"],
                       "flow_ms_per_pair_pipelined": summ["ms_per_step"] if res["pipelined"] else None,   # the search is what bounds the pipelined step
                       "search_kernel": "generation %d (csrc/hr_search%s.cuh)" % (res.get("search_generation", 1), "3" if res.get("search_generation", 1) == 3 else ""),
                       # every delivered frame is a warp output (vf_HopperRender.c:357-375); the ones with t != 0 alone:
                       "interp_only_frames_per_s": summ["value"] * res["interp_share"],
                       "device_loop": ("pipelined: pack || search, three search lanes whose launches share the SMs (three CTAs per SM), the warps of a source frame in one launch, search(k+1) || warps(k); %d source frames per C call" % res["chunk"]
                                       if res["pipelined"] else "serial"),
                       "serial_frames_per_s": (1.0 / res["serial_s_per_output"]) if res["serial_s_per_output"] else None},
            "gpu_launches": summ["gpu_launches"],
            "clocks": res["clocks"],
            "roofline": roof.get(dom),
            "kernels": roof,
            "dominant_kernel": dom,
            "roofline_hbm_kernel": roof.get("warp"),
        }
        if "e2e" in summ:
            line["e2e"] = {"value": summ["e2e"]["value"], "unit": "frames/s", "h2d_bytes_per_step": int(frame_bytes * band_frac),
                           "d2h_bytes_per_step": int(frame_bytes * band_frac * dfps / sfps),
                           "steps": summ["e2e"]["steps"], "api": summ["e2e"]["api"], "host_cpus_near_gpu": env.numa_cpus}
            for key in ("e2e_pageable", "e2e_pinned_pool", "e2e_zero_copy"):
                if key in summ:
                    line[key] = {"value": summ[key]["value"], "unit": "frames/s", "steps": summ[key]["steps"], "api": summ[key]["api"]}
        if ceiling:
            line["host_ceiling"] = ceiling
        if extras:
            line["configs_measured"] = extras
        if bands_block:
            line["bands_8k"] = bands_block
        if world == 1 and not args.no_extra:
            line["reference_gpu"] = reference_gpu(args)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line))
    if env.dist:
        env.dist.barrier()
        env.dist.destroy_process_group()


def reference_gpu(args):
    """The unmodified reference (host + .cl kernels, oracle/_ref) on this GPU through the NVIDIA OpenCL ICD: the same
    call sequence with host planes (1080p NV12: the reference negotiates nothing else). A checker leg, like cpu_baseline."""
    try:
        from oracle import ref_opencl
        import hr_pkg

        hr_pkg.load()
        from hopperrender_b200 import synth

        ok, why = ref_opencl.available()
        if not ok:
            return {"unavailable": why}
        w, h = 1920, 1080
        clip = synth.MovingTextureClip(w, h)
        frames = [clip.frame(k) for k in range(8)]
        r = ref_opencl.Reference(h, w, w)
        r.update_frame(*frames[0])
        ts = pacing_ts(260, 24.0, 60.0)
        y, uv = np.zeros((h, w), np.uint8), np.zeros((h // 2, w), np.uint8)
        planes = (ctypes.c_void_p * 2)(y.ctypes.data, uv.ctypes.data)

        def step(i):
            r.update_frame(*frames[(i + 1) % 8])
            r.calc_flow(args.radius, 8, 6)
            for t in ts[i]:
                assert not r.warp(float(np.float32(t)), 2)
                assert not r.lib.downloadFrame(ctypes.byref(r.s), planes)
            return len(ts[i]), r.s.ofcCalcTime, r.s.warpCalcTime

        for i in range(10):
            step(i)
        flows, warps, outs = [], [], 0
        t0 = time.perf_counter()
        for i in range(10, 260):
            n, f, wv = step(i)
            outs += n
            flows.append(f)
            warps.append(wv)
        dt = time.perf_counter() - t0
        r.close()
        return {"value": outs / dt, "unit": "frames/s", "workload": "1080p-nv12-24to60", "search_radius": args.radius, "steps": 250,
                "ofcCalcTime_ms_median": float(np.median(flows)) * 1e3, "warpCalcTime_ms_median": float(np.median(warps)) * 1e3,
                "kind": "reference: unmodified opticalFlowCalc.c + Kernels/*.cl through the NVIDIA OpenCL ICD on this GPU, host planes, blocking calls"}
    except Exception as e:          # a checker leg must not take the bench line down
        return {"unavailable": "%s: %s" % (type(e).__name__, e)}


def cpu_baseline(args):
    from oracle import hr_oracle_py as O
    import hr_pkg

    hr_pkg.load()
    from hopperrender_b200 import synth

    w, h, pixfmt, sfps, dfps, mode = WORKLOADS[args.workload]
    clip = synth.MovingTextureClip(w, h, pixfmt=pixfmt)
    frames = [clip.frame(k) for k in range(8)]
    o = O.Oracle(h, w, w, pixfmt)
    o.update_frame(*frames[0])
    o.update_frame(*frames[1])
    n = args.cpu_sample_steps
    ts = pacing_ts(n + 1, sfps, dfps)
    outs = 0
    t0 = time.perf_counter()
    done = 0
    for i in range(n):
        o.update_frame(*frames[(i + 2) % 8])
        o.calc_flow(args.radius, 8, 6)
        for t in ts[i]:
            o.warp(np.float32(t), mode)
            o.download()
            outs += 1
        done += 1
        if time.perf_counter() - t0 > 12.0:
            break
    dt = time.perf_counter() - t0
    return {"value": outs / dt, "unit": "frames/s", "cores": O.num_threads(), "kind": "port",
            "sample": "%d source frames (%d outputs) of %s in %.1f s; C restatement of the reference kernels, OpenMP — no OpenCL runtime (PoCL) in the image" % (done, outs, args.workload, dt)}


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
