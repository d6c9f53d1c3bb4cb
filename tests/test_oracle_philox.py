"""Known-answer tests pinning the oracle's Philox4x32-10 (Random123 kat_vectors,
`philox4x32 10` rows) and the B2N stream helpers."""
import numpy as np
from oracle import philox


def _hex(x):
    return [int(v) for v in x]


def test_philox_kat():
    f = philox.philox4x32_10
    assert _hex(f([0, 0, 0, 0], [0, 0])) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _hex(f([0xffffffff] * 4, [0xffffffff] * 2)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _hex(f([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344],
                  [0xa4093822, 0x299f31d0])) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_u53_open_interval():
    assert philox.u53(0, 0) > 0
    assert philox.u53(0xffffffff, 0xffffffff) < 1


def test_stream_statistics():
    z = np.concatenate([philox.event_normals(1, c, 0, 200) for c in range(200)])
    u = np.concatenate([philox.event_uniforms(1, c, 1, 200) for c in range(200)])
    assert abs(z.mean()) < 0.03 and abs(z.std() - 1) < 0.03
    assert abs(u.mean() - 0.5) < 0.01 and 0 < u.min() and u.max() < 1


def test_scripted_generator_is_generator():
    g = philox.ScriptedGenerator(5, 9)
    assert isinstance(g, np.random.Generator)
    s = philox.ChainStream(5, 9)
    assert g.random() == s.uniform()
    assert np.array_equal(g.standard_normal(size=7), s.normals(7))
    assert np.array_equal(g.random(3), s.uniforms(3))
    assert g.random(0).size == 0 and g.tick == s.tick
    idx = np.arange(6)
    g.shuffle(idx)
    assert np.array_equal(idx, s.permutation(6))
