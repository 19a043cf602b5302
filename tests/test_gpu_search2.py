"""The three generations of the search kernel (csrc/hr_search.cuh, hr_search2.cuh, hr_search3.cuh) against the CPU oracle and
against one another: raw and blurred offsets, the winning layer of every step, the tables the blur reads — same bits.
Geometries: 16:9 at every resolution scalar, a ragged lattice (points outside the lattice in the last tile column and
row), clips whose motion pushes the candidates over the frame border (reflected sample addressing), every radius the
filter's auto-adjust visits (HR/config.h:6-7), NV12 and P010."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _flow(hr, f1, f2, h, stride, w, R, gen, pixfmt=0, dS=8, nS=6, staged=True):
    g = hr.HrCuda(h, stride, w, pixfmt)
    g.set_search_generation(gen)
    g.set_search_staged(staged)
    g.set_trace(True)
    g.update_frame(*f1)
    g.update_frame(*f2)
    g.calc_flow(R, dS, nS)
    assert g.last_search_generation() == gen
    # the TMA-staged variant serves resolution scalar 2 (1080p, 720p), radius 5..8, and nothing else
    assert g.last_search_staged() == (gen == 2 and staged and g.info.resScalar == 2 and 5 <= R <= 8)
    raw, blur = g.get_offsets()
    layers = [g.get_step_layers(k) for k in range(2 * g.info.iterations)]
    g.close()
    return raw, blur, layers


def _same(a, b, what):
    for name, x, y in (("raw", a[0], b[0]), ("blurred", a[1], b[1])):
        assert np.array_equal(x, y), "%s: %s offsets differ at %d points" % (what, name, int((x != y).sum()))
    for k, (x, y) in enumerate(zip(a[2], b[2])):
        assert np.array_equal(x, y), "%s: winning layers of step %d differ" % (what, k)


@pytest.mark.parametrize("R", list(range(5, 17)))
def test_generations_agree_every_radius(hr, synth, R):
    c = synth.MovingTextureClip(1920, 1080)
    f1, f2 = c.frame(3), c.frame(4)
    first = _flow(hr, f1, f2, 1080, 1920, 1920, R, 1)
    _same(first, _flow(hr, f1, f2, 1080, 1920, 1920, R, 2, staged=False), "1080p R=%d, samples from global memory" % R)
    _same(first, _flow(hr, f1, f2, 1080, 1920, 1920, R, 3), "1080p R=%d, four points per thread" % R)
    if R <= 8:
        _same(first, _flow(hr, f1, f2, 1080, 1920, 1920, R, 2, staged=True), "1080p R=%d, samples staged by TMA" % R)


@pytest.mark.parametrize("velocity", [(7.0, -3.0), (-33.0, 21.0), (70.0, 40.0)])
def test_staged_variant_fast_motion(hr, oracle, synth, velocity):
    """Displacements that stay inside the staged halo (9 lattice cells = 36 samples), that reach its edge, and that leave
    it (the steps that reach further read global memory): R = 5 and R = 8 against the oracle."""
    c = synth.MovingTextureClip(1920, 1080, velocity=velocity)
    f1, f2 = c.frame(1), c.frame(2)
    for R, gen in ((5, 2), (8, 2), (5, 3), (8, 3)):
        got = _flow(hr, f1, f2, 1080, 1920, 1920, R, gen)
        o = oracle.Oracle(1080, 1920, 1920, 0)
        o.update_frame(*f1)
        o.update_frame(*f2)
        o.calc_flow(R, 8, 6)
        oraw, oblur = o.get_offsets()
        assert np.array_equal(got[0], oraw) and np.array_equal(got[1], oblur), "velocity %s R=%d generation %d" % (velocity, R, gen)


@pytest.mark.parametrize("w,h,stride,pixfmt", [(1280, 720, 1280, 0), (854, 480, 896, 0), (480, 270, 480, 0), (1000, 562, 1024, 0),
                                               (3840, 2160, 3840, 1), (2048, 858, 2048, 0), (1918, 1080, 1920, 1)])
@pytest.mark.parametrize("R,gen", [(5, 2), (11, 2), (16, 2), (5, 3), (8, 3), (11, 3), (16, 3)])
def test_generation2_against_oracle(hr, oracle, synth, w, h, stride, pixfmt, R, gen):
    c = synth.MovingTextureClip(w, h, stride=stride, pixfmt=pixfmt)
    f1, f2 = c.frame(1), c.frame(2)
    got = _flow(hr, f1, f2, h, stride, w, R, gen, pixfmt)
    o = oracle.Oracle(h, stride, w, pixfmt)
    o.update_frame(*f1)
    o.update_frame(*f2)
    o.calc_flow(R, 8, 6)
    oraw, oblur = o.get_offsets()
    assert np.array_equal(got[0], oraw), "raw offsets differ at %d points" % int((got[0] != oraw).sum())
    assert np.array_equal(got[1], oblur), "blurred offsets differ at %d points" % int((got[1] != oblur).sum())


def test_generation2_noise_wrap_and_scalars(hr, oracle, synth):
    """Full-range noise: large offsets everywhere (every border warp takes the reflected path), window sums that wrap
    in 32 bits with a large deltaScalar, and the extreme bias scalars."""
    f1, f2 = synth.noise_frame(1080, 1920, 21), synth.noise_frame(1080, 1920, 22)
    for R, dS, nS, gen in ((5, 12, 6, 2), (16, 8, 6, 2), (8, 0, 0, 2), (9, 12, 10, 2), (13, 4, 8, 2), (5, 12, 6, 3), (8, 0, 0, 3), (7, 4, 8, 3), (16, 8, 6, 3), (13, 4, 8, 3)):
        got = _flow(hr, f1, f2, 1080, 1920, 1920, R, gen, 0, dS, nS)
        o = oracle.Oracle(1080, 1920, 1920, 0)
        o.update_frame(*f1)
        o.update_frame(*f2)
        o.calc_flow(R, dS, nS)
        oraw, oblur = o.get_offsets()
        assert np.array_equal(got[0], oraw) and np.array_equal(got[1], oblur), "R=%d dS=%d nS=%d" % (R, dS, nS)


def test_generation2_repeats_and_alternates(hr, synth):
    """Launch after launch on one context, generations alternating: the epoch-tagged tables and totals left by one
    generation never leak into the next launch."""
    c = synth.MovingTextureClip(1920, 1080)
    g = hr.HrCuda(1080, 1920, 1920)
    g.update_frame(*c.frame(0))
    g.update_frame(*c.frame(1))
    ref = None
    for k in range(6):
        g.set_search_generation(1 + k % 3)
        g.calc_flow(5 if k < 4 else 16)
        cur = g.get_offsets()
        if k in (0, 4):
            ref = cur
        else:
            assert np.array_equal(ref[0], cur[0]) and np.array_equal(ref[1], cur[1]), "launch %d" % k
    g.close()


def test_generation_chosen_per_launch(hr, synth):
    """Default policy: a launch that has the GPU to itself (blocking calls) is the first generation; while the previous
    pair's search is still under way (pipelined, enqueued back to back) the third. Same offsets either way."""
    c = synth.MovingTextureClip(1920, 1080)
    g = hr.HrCuda(1080, 1920, 1920)
    g.update_frame(*c.frame(0))
    g.update_frame(*c.frame(1))
    g.calc_flow(5)
    assert g.last_search_generation() == 1
    ref = g.get_offsets()
    g.set_pipeline(True)
    seen = set()
    for _ in range(8):
        g.calc_flow(5, blocking=False)
        seen.add(g.last_search_generation())
    g.synchronize()
    assert seen <= {1, 3} and 3 in seen, seen
    cur = g.get_offsets()
    assert np.array_equal(ref[0], cur[0]) and np.array_equal(ref[1], cur[1])
    g.set_pipeline(False)
    g.close()
