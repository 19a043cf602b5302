"""N > 1 host logic on CPU: partitioning, band geometry, and the counter reductions over a world_size-2
gloo group (two real processes, rendezvous on 127.0.0.1)."""
import os
import pathlib
import subprocess
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


def test_items_are_partitioned_exactly_once(hr):
    from hopperrender_b200 import sharding as S
    for n in (0, 1, 7, 64):
        for world in (1, 2, 4, 8):
            got = sorted(i for r in range(world) for i in S.shard_items(n, r, world))
            assert got == list(range(n))
    assert S.shard_items(64, 3, 8) == list(range(3, 64, 8))
    with pytest.raises(ValueError):
        S.shard_items(4, 2, 2)


def test_bands_cover_the_frame_and_are_aligned(hr):
    from hopperrender_b200 import sharding as S
    for h in (1080, 2160, 4320, 4322, 720):
        s = S.res_scalar(h)
        for world in (1, 2, 4, 8):
            if world > -(-h // (32 << s)):
                continue                             # more bands than lattice tile rows
            bands = S.band_rows(h, world)
            assert bands[0][0] == 0 and bands[-1][1] == h
            for (a0, a1), (b0, b1) in zip(bands, bands[1:]):
                assert a1 == b0 and a0 < a1
            for r0, r1 in bands[:-1]:
                assert r1 % (32 << s) == 0          # whole lattice tile rows: the search is split by tiles
    assert S.res_scalar(4320) == 4 and S.res_scalar(2160) == 3 and S.res_scalar(1080) == 2
    assert S.band_rows(4320, 8) == [(512 * k, 512 * (k + 1)) for k in range(7)] + [(3584, 4320)]
    assert S.band_rows(4320, 4) == [(0, 1024), (1024, 2048), (2048, 3072), (3072, 4320)]
    assert S.band_rows(1080, 2) == [(0, 512), (512, 1080)]
    with pytest.raises(ValueError):
        S.band_rows(720, 8)                          # 720 lines = 6 tile rows


def test_halo_follows_the_search_radius(hr):
    from hopperrender_b200 import sharding as S
    assert S.max_offset(5) == (32, 32)
    assert S.max_offset(16) == (512, 392)
    lo, hi = S.band_halo(1088, 1632, 4320, 5)
    assert lo == 1054 and hi == 1666
    assert S.band_halo(0, 544, 4320, 16) == (0, 1058)
    assert S.band_halo(3808, 4320, 4320, 16)[1] == 4320


WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["HR_ROOT"])
import torch, torch.distributed as dist
import hr_pkg
hr_pkg.load()
from hopperrender_b200 import sharding as S
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["HR_PORT"], rank=int(os.environ["RANK"]), world_size=2)
rank = dist.get_rank()
mine = S.shard_items(9, rank, 2)
outs, secs = S.reduce_throughput(len(mine) * 5, 1.0 + rank, dist)
gathered = [None, None]
dist.all_gather_object(gathered, mine)
assert sorted(gathered[0] + gathered[1]) == list(range(9))
assert outs == 45 and secs == 2.0, (outs, secs)
bands = S.band_rows(4320, 2)
assert bands[rank] == ((0, 2048), (2048, 4320))[rank]
dist.barrier()
dist.destroy_process_group()
print("rank %d ok" % rank)
'''


def test_two_process_gloo_reduction(hr, tmp_path):
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), HR_PORT=str(port), HR_ROOT=str(ROOT), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for rank, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert "rank %d ok" % rank in o
