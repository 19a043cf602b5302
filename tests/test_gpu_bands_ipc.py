"""Spatial bands with one PROCESS per GPU (how bench.py --bands and a torchrun deployment run them): frame slots and
exchange arenas mapped through CUDA IPC handles, set-up over a gloo process group, everything else on the devices.
Two ranks on two GPUs; every rank checks its flow and its band of every output against a single-context run of its own.
Skipped with a reason on a box with one GPU (gpurun --gpus 2)."""
import ctypes
import os
import pathlib
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = pathlib.Path(__file__).resolve().parent.parent

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["HR_ROOT"])
import numpy as np
import torch, torch.distributed as dist
import hr_pkg
hr = hr_pkg.load()
from hopperrender_b200 import sharding, synth
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD"])
torch.cuda.set_device(rank)
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["HR_PORT"], rank=rank, world_size=world)
w, h, pixfmt = int(os.environ["HR_W"]), int(os.environ["HR_H"]), int(os.environ["HR_PF"])
clip = synth.MovingTextureClip(w, h, pixfmt=pixfmt)
single = hr.HrCuda(h, w, w, pixfmt, rank)
band = hr.HrCuda(h, w, w, pixfmt, rank)
rows = sharding.band_rows(h, world, band.info.resScalar)
hr.connect_bands_distributed(band, dist, rows, max_radius=8)
r0, r1 = rows[rank]
for k in range(4):
    y, uv = clip.frame(k)
    single.update_frame(y, uv)
    band.band_upload(np.ascontiguousarray(y[r0:r1]), np.ascontiguousarray(uv[r0 >> 1:r1 >> 1]))
    band.band_gather(blocking=True)
    if k == 0:
        continue
    R = (5, 8, 5)[k - 1]
    single.calc_flow(R)
    band.calc_flow(R, blocking=False)
    band.synchronize()
    sraw, sblur = single.get_offsets()
    raw, blur = band.get_offsets()
    assert np.array_equal(raw, sraw) and np.array_equal(blur, sblur), "flow differs on rank %d frame %d" % (rank, k)
    for t, mode in ((0.4, 2), (0.8, 0), (0.5, 3)):
        single.warp(t, mode)
        sy, suv, _ = single.download()
        band.warp(t, mode)
        by, buv = np.empty((r1 - r0, w), sy.dtype), np.empty(((r1 >> 1) - (r0 >> 1), w), sy.dtype)
        band.band_download(by, buv)
        assert np.array_equal(by, sy[r0:r1]) and np.array_equal(buv, suv[r0 >> 1:r1 >> 1]), "rows differ on rank %d frame %d mode %d" % (rank, k, mode)
lo, hi, nbytes = band.band_halo()
assert nbytes > 0 and (lo, hi) != (0, h) or world == 1
dist.barrier()
band.close()
single.close()
dist.destroy_process_group()
print("rank %d ok: rows [%d, %d) held [%d, %d), %d bytes over NVLink" % (rank, r0, r1, lo, hi, nbytes))
'''


def _ndev():
    cu = ctypes.CDLL("libcuda.so.1")
    cu.cuInit(0)
    n = ctypes.c_int(0)
    cu.cuDeviceGetCount(ctypes.byref(n))
    return n.value


@pytest.mark.parametrize("w,h,pixfmt", [(1920, 1080, 0), (3840, 2160, 1)])
def test_two_processes_two_gpus(tmp_path, w, h, pixfmt):
    if _ndev() < 2:
        pytest.skip("two band processes need two GPUs (run under gpurun --gpus 2)")
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    script = tmp_path / "band_worker.py"
    script.write_text(WORKER)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD="2", HR_PORT=str(port), HR_ROOT=str(ROOT), HR_W=str(w), HR_H=str(h), HR_PF=str(pixfmt))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=300)[0])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            pytest.fail("band worker timed out")
    for rank, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and ("rank %d ok" % rank) in o, o[-3000:]
