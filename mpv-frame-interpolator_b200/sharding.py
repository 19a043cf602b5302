"""Host-side work partitioning for more than one GPU. No data-path collective:

* independent frame pairs / streams (`shard_items`): item i -> rank i mod N; every rank owns its own
  context, stream and pinned buffers; results meet only as counters and times (`reduce_throughput`).
* 8K spatial bands (`band_rows`, `band_halo`): rows split into whole lattice tile rows (32 << s frame rows); each rank
  uploads, searches, warps and downloads only its band; the rows its search and warp can reach outside the band
  (the halo, fetched from the neighbours by NVLink P2P) follow from the search radius.

Pure Python + torch.distributed for the counters (gloo on CPU in the tests, NCCL under torchrun);
the reference has no counterpart (single device, `HR/opticalFlowCalc.c:279-305`).
"""
import math


def shard_items(n_items, rank, world):
    """Indices of the items (frame pairs / streams) rank `rank` of `world` processes owns."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    return list(range(rank, n_items, world))


def res_scalar(frame_height, max_calc_res=270):
    """opticalFlowResScalar, HR/opticalFlowCalc.c:331-334."""
    s = 0
    while (frame_height >> s) > max_calc_res:
        s += 1
    return s


LATTICE_TILE = 32   # lattice points per tile side of the search kernel (csrc/hr_common.cuh HR_TILE)


def band_rows(frame_height, world, s=None):
    """[(row0, row1)] per rank: contiguous bands of whole lattice tile rows (32 << s frame rows: the search is split
    by tiles, and tile rows are lattice-row and chroma-row aligned), sizes as equal as that allows, the last band ends
    with the frame. Every row belongs to exactly one band."""
    if s is None:
        s = res_scalar(frame_height)
    unit = LATTICE_TILE << s
    units = math.ceil(frame_height / unit)
    if world > units:
        raise ValueError("%d bands do not fit %d rows in units of %d" % (world, frame_height, unit))
    # boundary r as close to r/world of the frame as whole tile rows allow, at least one tile row per band
    cuts = [0]
    for r in range(1, world):
        c = int(round(r * frame_height / world / unit))
        c = max(c, cuts[-1] + 1)
        c = min(c, units - (world - r))
        cuts.append(c)
    cuts.append(units)
    return [(min(a * unit, frame_height), min(b * unit, frame_height)) for a, b in zip(cuts, cuts[1:])]


def max_offset(radius, iterations=8):
    """Largest |accumulated offset| per axis after `iterations` levels, full-resolution pixels
    (candidate shifts (z - R/2)|z - R/2|, HR/Kernels/calcDeltaSumsKernel.cl:68-72): (negative, positive)."""
    neg = (radius // 2) ** 2
    pos = (radius - 1 - radius // 2) ** 2
    return neg * iterations, pos * iterations


def band_halo(row0, row1, frame_height, radius, iterations=8):
    """Source rows [lo, hi) of both frames that the warp of output rows [row0, row1) can read: the blend
    position moves by at most the largest offset (t and 1-t are <= 1), rounded out to even rows for the
    chroma plane, clipped to the frame."""
    neg, pos = max_offset(radius, iterations)
    reach = max(neg, pos) + 1
    lo = max(0, (row0 - reach) & ~1)
    hi = min(frame_height, (row1 + reach + 1) & ~1)
    return lo, hi


def reduce_throughput(outputs, seconds, dist=None, device=None):
    """Whole-job numbers of a sharded run: (sum of outputs over ranks, max of seconds over ranks)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return int(outputs), float(seconds)
    import torch

    t = torch.tensor([float(seconds)], dtype=torch.float64, device=device)
    c = torch.tensor([float(outputs)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return int(round(float(c[0]))), float(t[0])
