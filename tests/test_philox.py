"""Known-answer tests for the keyed RNG (Random123 kat_vectors, philox4x32 10 rounds)."""
import numpy as np

from oracle import philox

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_kat_scalar():
    for ctr, key, want in KAT:
        assert philox.philox4x32(*ctr, *key) == want


def test_philox_kat_vectorised():
    for ctr, key, want in KAT:
        got = philox.philox4x32_np(*[np.array([c]) for c in ctr], *key)
        assert tuple(int(w[0]) for w in got) == want


def test_vector_matches_scalar_on_detect_counters():
    env = np.arange(0, 50, 7)
    t = np.arange(0, 5)
    got = philox.philox4x32_np(env[:, None], t[None, :], 2, 11, 42, philox.STREAM_DETECT)[0]
    for a, e in enumerate(env):
        for b, tt in enumerate(t):
            assert int(got[a, b]) == philox.philox4x32(int(e), int(tt), 2, 11, 42, philox.STREAM_DETECT)[0]


def test_threshold_equivalence():
    # r / 2**32 <= d  <=>  r <= floor(d * 2**32)   (hard part 1)
    for d in (0.9, 0.5, 0.1, 0.999, 1e-9):
        thr = philox.detect_threshold(d)
        for r in (thr - 1, thr, thr + 1):
            if 0 <= r <= philox.MASK:
                assert (r / 4294967296.0 <= d) == (r <= thr)
    assert philox.detect_threshold(0.9) == 3865470566
    assert philox.detect_threshold(1.0) == philox.MASK
    assert philox.detect_threshold(-0.1) == -1


def test_uniformity_smoke():
    r = philox.philox4x32_np(np.arange(200000), 1, 0, 0, 42, philox.STREAM_DETECT)[0]
    frac = float(np.mean(r <= philox.detect_threshold(0.9)))
    assert abs(frac - 0.9) < 0.004


def test_streams_do_not_repeat_after_65536_episodes():
    """The counter word holds episode & 0xFFFF; the episode's upper bits go into the stream key (cs_stream_key), so
    episode e and e + 65536 of an env draw from different streams."""
    from oracle import philox
    a = [philox.detect_draw(7, 123, 5, t, 1, 2) for t in range(1, 9)]
    b = [philox.detect_draw(7, 123, 5 + 65536, t, 1, 2) for t in range(1, 9)]
    c = [philox.detect_draw(7, 123, 5, t, 1, 2) for t in range(1, 9)]
    assert a == c and a != b
    assert philox.stream_key(philox.STREAM_DETECT, 5) == philox.STREAM_DETECT          # unchanged below 65536 episodes
