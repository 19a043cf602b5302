#!/usr/bin/env python
"""bench.py — HopperRender hot path on B200: interpolated frames/s.

One "step" = one source frame of a 24->60 stream through the hot path:
    update (pack) -> calculateOpticalFlow -> the 2 or 3 warps the filter's pacing rule asks for
(vf_HopperRender.c:371-374,481; SURVEY.md Appendix D: 25 outputs per 10 source frames).

  value      interpolated (= delivered, every one is a warp output) frames/s, device-resident:
             the source frames already sit in HBM (a ring larger than L2), output stays in HBM.
  e2e        the same metric through the reference-facing call sequence with HOST buffers:
             updateFrame (H2D) / calculateOpticalFlow / warpFrames / downloadFrame (D2H) per output.
  roofline   dominant kernel of the step by device time, timed with CUDA events on the launch
             stream over whole loops of the same steps (by difference: search only / + pack / + warps).
  cpu_baseline   the CPU oracle (C restatement of the reference kernels, OpenMP) on a bounded
             sample of the same workload, rank 0, N=1 only.

`--impl reference` times the reference's CPU implementation of the path (oracle/_ref when it was
built, else the oracle port) on the host cores — the only place besides cpu_baseline where
oracle/ is executed by this file.
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

WORKLOADS = {
    # name: (width, height, pixfmt, source fps, display fps, mode)
    "1080p-nv12-24to60": (1920, 1080, 0, 24.0, 60.0, 2),
    "4k-nv12-24to60": (3840, 2160, 0, 24.0, 60.0, 2),
    "4k-p010-24to144": (3840, 2160, 1, 24.0, 144.0, 2),
    "4k-p010-24to144-hsv": (3840, 2160, 1, 24.0, 144.0, 3),
    "8k-p010-24to60": (7680, 4320, 1, 24.0, 60.0, 2),
}
L2_BYTES = 126 * 1024 * 1024
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures (profiles/)
NCU_TRAFFIC = {
    ("1080p-nv12-24to60", "search"): 5203968, ("1080p-nv12-24to60", "warp"): 6787328, ("1080p-nv12-24to60", "pack"): 3116288,
    ("4k-p010-24to144", "warp"): 53419008, ("4k-p010-24to144", "pack"): 25016576,
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="1080p-nv12-24to60", choices=sorted(WORKLOADS))
    ap.add_argument("--radius", type=int, default=5, help="search radius (config.h MIN_SEARCH_RADIUS = 5 is the default)")
    ap.add_argument("--cpu-sample-steps", type=int, default=400)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--serial", action="store_true", help="device-resident loop without the pipelined mode (one kernel after the other)")
    ap.add_argument("--bands", action="store_true",
                    help="split every frame into spatial bands over the N ranks (8K config, SURVEY.md §8e): strong scaling, "
                         "bands uploaded/warped/downloaded per GPU, the other bands pulled by NVLink P2P")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


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
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 6:
                self.samples.append(parts)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
            except ValueError:
                continue
            for n, v in zip(names, s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def pacing_ts(n_steps, src_fps, disp_fps):
    """Blend scalars for n_steps source frames in steady state (the first source frame of a
    stream produces no warp, so one extra frame primes the pacer)."""
    import hr_pkg

    hr_pkg.load()
    from hopperrender_b200 import pacing

    p = pacing.Pacer(src_fps, disp_fps)
    p.next_source_frame()
    return [p.next_source_frame() for _ in range(n_steps)]


def warp_bytes(w, h, bps, lw, lh):
    return 3 * int(1.5 * w * h * bps) + 4 * lw * lh


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """CPU arm: the reference's algorithm on the host cores (oracle/_ref if built, else the oracle)."""
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
    kind = "port"
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
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": O.num_threads(), "kind": kind,
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


def run_ours(args):
    import torch
    import hr_pkg

    hr = hr_pkg.load()
    from hopperrender_b200 import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback")
    numa_cpus = bind_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    w, h, pixfmt, sfps, dfps, mode = WORKLOADS[args.workload]
    bps = 2 if pixfmt else 1
    tdtype = torch.uint16 if pixfmt else torch.uint8
    frame_bytes = int(1.5 * w * h * bps)
    # ring of source frames larger than L2, so every step reads its inputs from HBM
    nring = max(8, (2 * L2_BYTES) // frame_bytes + 2)
    nring = min(nring, 96)
    clip = synth.MovingTextureClip(w, h, pixfmt=pixfmt)
    nbase = 8 if frame_bytes < (64 << 20) else 4
    base = [clip.frame(k) for k in range(nbase)]
    stream = torch.cuda.Stream()
    g = hr.HrCuda(h, w, w, pixfmt, device=local)
    g.set_stream(stream.cuda_stream)
    lw, lh = g.info.lowWidth, g.info.lowHeight
    banded = bool(args.bands) and world > 1
    r0, r1 = 0, h
    if banded:
        from hopperrender_b200 import sharding

        rows = sharding.band_rows(h, world, g.info.resScalar)
        hr.connect_bands_distributed(g, dist, rows)
        r0, r1 = rows[rank]
        nring = max(8, min(96, nring * world))      # a rank keeps only its band of every ring frame
    with torch.cuda.stream(stream):
        ring = []
        for k in range(nring):
            y, uv = base[k % nbase]
            # banded: only this rank's rows live on this GPU
            ring.append((torch.from_numpy(np.ascontiguousarray(y[r0:r1])).to("cuda", non_blocking=False).view(tdtype),
                         torch.from_numpy(np.ascontiguousarray(uv[r0 >> 1:r1 >> 1])).to("cuda", non_blocking=False).view(tdtype)))
        out_ring = [(torch.empty((h, w), dtype=tdtype, device="cuda"), torch.empty((h // 2, w), dtype=tdtype, device="cuda")) for _ in range(max(4, (L2_BYTES // frame_bytes) + 2))]
    stream.synchronize()
    K, W_ = args.steps, max(3, args.warmup)
    ts = pacing_ts(W_ + K, sfps, dfps)
    interp_share = sum(1 for i in range(W_, W_ + K) for t in ts[i] if t > 1e-6) / max(1, sum(len(ts[i]) for i in range(W_, W_ + K)))
    radius = args.radius

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
        if banded:
            feed(i)
            g.calc_flow(radius, 8, 6, blocking=False)
            for t in ts[i]:
                oy, ouv = out_ring[oi[0] % len(out_ring)]
                oi[0] += 1
                g.set_output_device(oy, ouv)
                g.warp(t, mode)
            return n
        # one C call per source frame: update (borrowed device planes) + flow + the pacing rule's warps, each into
        # its own output frame; pipelined mode overlaps the independent work of consecutive pairs
        outs = [out_ring[(oi[0] + j) % len(out_ring)] for j in range(n)]
        oi[0] += n
        y, uv = ring[i % nring]
        g.step_device(y, uv, ts[i], outs, radius=radius, mode=mode)
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

    def barrier():
        stream.synchronize()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    # ---- device-resident timed region ---------------------------------------------------------
    pipelined = (not banded) and (not args.serial)
    with torch.cuda.stream(stream):
        feed(nring - 1)
        g.set_pipeline(pipelined)
        steps_device(0, W_)
        g.synchronize()
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        l0 = g.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        outs = steps_device(W_, K)
        g.pipeline_join()              # the main stream now follows every internal stream: e1 closes the whole region
        e1.record(stream)
        g.synchronize()
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        launches = g.launch_count() - l0
        ms = e0.elapsed_time(e1)
        g.set_pipeline(False)
        serial_ms = None
        if pipelined:                  # the same steps, one kernel after the other (what the blocking interface sees)
            steps_device(0, W_)
            g.synchronize()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(stream)
            ks = min(K, 100)
            so = steps_device(W_, ks)
            s1.record(stream)
            g.synchronize()
            serial_ms = s0.elapsed_time(s1) / max(1, so)

    # ---- per-kernel device time, by difference -------------------------------------------------------
    # One CUDA event pair around a single ~5-40 us launch adds several us of front-end latency to it, so
    # the three kernels are timed over whole loops instead (two events per loop, stream kept full):
    #   A: search only            B: pack + search            C: pack + search + warps (= the timed region)
    #   search = A / n,  pack = (B - A) / n,  warp = (C - B) / (number of warps)
    with torch.cuda.stream(stream):
        nk = min(K, 100)

        def loop(with_pack, with_warp):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for rep in range(2):                      # first repetition warms the variant up
                if rep == 1:
                    e0.record(stream)
                nw = 0
                for i in range(nk):
                    if with_pack:
                        feed(W_ + i)
                    g.calc_flow(radius, 8, 6, blocking=False)
                    if with_warp:
                        for t in ts[W_ + i]:
                            oy, ouv = out_ring[oi[0] % len(out_ring)]
                            oi[0] += 1
                            g.set_output_device(oy, ouv)
                            g.warp(t, mode)
                            nw += 1
            e1.record(stream)
            barrier()
            return e0.elapsed_time(e1), nw

        tA, _ = loop(False, False)
        tB, _ = loop(True, False)
        tC, nwarps = loop(True, True)
        kms = {"search": [tA / nk], "pack": [max(tB - tA, 0.0) / nk], "warp": [max(tC - tB, 0.0) / max(1, nwarps)] * max(1, nwarps)}
        kcount = {"search": nk, "pack": nk, "warp": nwarps}
    g.set_output_device(None, None)

    # ---- end-to-end through the reference-facing interface, host buffers -------------------------
    # The compiled C host layer (opticalFlowCalc.c, the drop-in for the reference's file of that name) driven by
    # hrReplay.c, the filter's per-frame call sequence as a C loop: updateFrame (H2D) / calculateOpticalFlow /
    # warpFrames / downloadFrame (D2H) per output, every call blocking like the reference's.
    e2e = None
    if not args.no_e2e:
        # host frames as mpv's image pool lays them out: one pinned allocation, the UV plane right behind the Y plane
        def host_frame():
            ny, nuv = (r1 - r0) * w, ((r1 >> 1) - (r0 >> 1)) * w
            buf = torch.empty(ny + nuv, dtype=tdtype).pin_memory()
            return buf[:ny].view(r1 - r0, w), buf[ny:].view((r1 >> 1) - (r0 >> 1), w)

        hring = []
        for k in range(nbase):
            y, uv = base[k]
            ty, tuv = host_frame()
            ty.copy_(torch.from_numpy(np.ascontiguousarray(y[r0:r1])).view(tdtype))
            tuv.copy_(torch.from_numpy(np.ascontiguousarray(uv[r0 >> 1:r1 >> 1])).view(tdtype))
            hring.append((ty, tuv))
        hout = host_frame()
        Ke = min(K, 100)
        We = min(W_, 5)
        if banded:
            ofc = hr.OpticalFlowCalc()
            if hr.initOpticalFlowCalc(ofc, h, w, w, pixfmt, device=local):
                raise SystemExit("initOpticalFlowCalc failed")
            ofc.opticalFlowSearchRadius = radius
            hr.connect_bands_distributed(ofc.impl, dist, rows)

            def step_host(i):                               # a rank moves only its band over PCIe
                ofc.impl.band_upload(*hring[i % nbase])
                ofc.impl.band_gather(blocking=True)
                assert not hr.calculateOpticalFlow(ofc)
                for t in ts[i]:
                    assert not hr.warpFrames(ofc, t, mode)
                    ofc.impl.band_download(hout[0], hout[1])
                return len(ts[i])

            step_host(nbase - 1)
            for i in range(We):
                step_host(i)
            barrier()
            t0 = time.perf_counter()
            eouts = sum(step_host(We + i) for i in range(Ke))
            torch.cuda.synchronize()
            edt = time.perf_counter() - t0
            hr.freeOFC(ofc)
            e2e_api = "band_upload/band_gather/calculateOpticalFlow/warpFrames/band_download per rank, pinned host planes"
        else:
            import ctypes

            lib = hr.load_ofc_library()
            cofc = hr.COpticalFlowCalc()
            cofc.pixelFormat = pixfmt
            cofc.cudaDevice = local + 1
            if lib.initOpticalFlowCalc(ctypes.byref(cofc), h, w, w):
                raise SystemExit("initOpticalFlowCalc (C host layer) failed")
            cofc.opticalFlowSearchRadius = radius
            hr.replay_stream_c(cofc, hring, nbase - 1, [[]], mode, hout)           # the first frame of the stream
            hr.replay_stream_c(cofc, hring, 0, ts[:We], mode, hout)
            barrier()
            t0 = time.perf_counter()
            eouts = hr.replay_stream_c(cofc, hring, We, ts[We:We + Ke], mode, hout)
            edt = time.perf_counter() - t0
            lib.freeOFC(ctypes.byref(cofc))
            e2e_api = ("libhopperrender_ofc.so: initOpticalFlowCalc, then hrReplayStream = updateFrame/calculateOpticalFlow/warpFrames/downloadFrame "
                       "in the filter's order, pinned host planes, every call blocking like the reference's")
        e2e = (eouts, edt, Ke)

    # ---- reduce over ranks (sharding.py: the N > 1 host logic, covered on CPU by tests/test_sharding_cpu.py) ----
    from hopperrender_b200 import sharding

    dev = torch.device("cuda", local)
    tot_outs, max_s = sharding.reduce_throughput(outs, ms * 1e-3, dist, dev)
    max_ms = max_s * 1e3
    e_outs, e_dt = (e2e[0], e2e[1]) if e2e else (0, 1.0)
    if e2e:
        e_outs, e_dt = sharding.reduce_throughput(e_outs, e_dt, dist, dev)
    if banded:          # every rank produced a band of the SAME frames: count each frame once
        tot_outs //= world
        e_outs //= world
    launches, _ = sharding.reduce_throughput(launches, 0.0, dist, dev)

    if rank == 0:
        pk, pk_src = peaks()
        avg = {k: float(np.mean(v)) if v else 0.0 for k, v in kms.items()}            # ms per launch
        per_step = {"pack": avg["pack"], "search": avg["search"], "warp": avg["warp"] * (kcount["warp"] / max(1, kcount["search"]))}
        dom = max(per_step, key=per_step.get)
        wbytes = int(warp_bytes(w, h, bps, lw, lh) * ((r1 - r0) / h))   # a band's launch moves the band's rows
        # algorithmic bytes per launch
        alg = {
            "warp": wbytes,
            "pack": int(1.5 * w * h * bps) + 4 * w * h,
            "search": 4 * lw * lh + int(1.5 * w * h) + 4 * lw * lh,   # frame2 lattice words + reachable frame1 (<= 1 frame of packed samples' worth) + flow out
        }
        roof = {}
        for k in ("pack", "search", "warp"):
            if avg[k] > 0:
                ach = alg[k] / (avg[k] * 1e-3) / 1e9
                roof[k] = {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                           "traffic": NCU_TRAFFIC.get((args.workload, k)), "avg_us": avg[k] * 1e3, "share_of_step": per_step[k] / max(1e-12, sum(per_step.values())),
                           "algorithmic_bytes": alg[k], "peak_source": pk_src}
        evals = 2 * g.info.iterations * radius * lw * lh
        if avg["search"] > 0 and args.workload == "1080p-nv12-24to60" and radius == 5:
            # the search's own limiter is instruction issue and dependent latency: executed warp instructions per launch
            # from the committed ncu capture (profiles/r01_ncu_1080p_nv12.txt) against the issue slots of the launch
            slots = pk.get("sm_max_mhz", 1965.0) * 1e6 * 4 * g.info.smCount * (avg["search"] * 1e-3)
            roof["search"]["issue"] = {"warp_instructions_per_launch": 14027307, "issue_slots_in_launch": slots, "frac": 14027307 / slots,
                                       "source": "smsp__inst_executed.sum of the ncu capture; 4 schedulers x SMs x max clock x measured launch time"}
        if avg["search"] > 0:
            roof["search"]["candidate_evals_per_s"] = evals / (avg["search"] * 1e-3)
            roof["search"]["note"] = ("latency/issue bound, not HBM: 16 dependent steps with tile-to-tile hand-offs; ncu (profiles/): issue slots 39 % busy, "
                                      "ALU pipe 39 %, 41 instructions per packed SAD; the HBM-bound kernel of the step is the warp, see roofline_hbm_kernel")
        line = {
            "metric": "interpolated frames/s", "value": tot_outs / (max_ms * 1e-3), "unit": "frames/s", "n_gpus": world,
            "steps": K, "warmup": W_, "ms_per_step": max_ms / K, "higher_is_better": True, "scaling": "strong" if banded else "weak", "vs_baseline": None,
            "dtype": "u8" if pixfmt == 0 else "u16", "data": "synthetic",
            "config": {"workload": args.workload, "frame": "%dx%d" % (w, h), "search_radius": radius, "mode": mode,
                       "streams_per_gpu": 1, "partition": ("%d spatial bands, NVLink P2P gather" % world) if banded else ("independent streams" if world > 1 else "none"),
                       "cache": "source ring of %d frames (%d MB) and output ring exceed the 126 MB L2" % (nring, nring * frame_bytes >> 20),
                       "flow_ms_per_pair": avg["search"],
                       # every delivered frame is a warp output (vf_HopperRender.c:357-375); the ones with t != 0 alone:
                       "interp_only_frames_per_s": tot_outs / (max_ms * 1e-3) * interp_share,
                       "device_loop": ("pipelined: pack || search, two search lanes, warps on parallel streams, search(k+1) || warps(k); %d source frames per C call" % CHUNK if pipelined else "serial"),
                       "serial_frames_per_s": (1e3 / serial_ms if serial_ms else None)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof.get(dom),
            "kernels": roof,
            "dominant_kernel": dom,
            "roofline_hbm_kernel": dict(roof.get("warp", {}), kernel="warp_fast_kernel"),
        }
        if e2e:
            band_frac = (r1 - r0) / h
            line["e2e"] = {"value": e_outs / e_dt, "unit": "frames/s", "h2d_bytes_per_step": int(frame_bytes * band_frac),
                           "d2h_bytes_per_step": int(frame_bytes * band_frac * (e2e[0] / e2e[2])), "steps": e2e[2],
                           "api": e2e_api,
                           "host_cpus_near_gpu": numa_cpus}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line))
    if dist:
        dist.barrier()
        dist.destroy_process_group()


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
        if time.perf_counter() - t0 > 30.0:
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
