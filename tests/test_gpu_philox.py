"""Device-side known-answer tests of the counter-based generator (mcp_philox_raw): the Random123 philox4x32-10 vectors and
a million counters against the numpy restatement (which tests/test_oracle.py pins to the C port and to the same vectors)."""
import numpy as np
import pytest

import montecarlooptionspricer_b200 as m
from philox_np import philox4x32_10

pytestmark = pytest.mark.gpu

KATS = [  # Random123 kat_vectors, philox4x32 10 rounds: (ctr, key, out)
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


def test_device_philox_reproduces_the_random123_vectors(engine):
    for ctr, key, want in KATS:
        got = engine.philox_raw(seed=key[0] | (key[1] << 32), first=ctr[0] | (ctr[1] << 32), count=1, c2=ctr[2], c3=ctr[3])
        assert [int(x) for x in got[0]] == want


@pytest.mark.parametrize("seed,first,c2,c3", [(0, 0, 0, 0), (0x0123456789abcdef, (1 << 32) - 500_000, 7, 2), (2**64 - 1, 2**64 - 1_000_000, 0xffffffff, 3)])
def test_device_philox_matches_the_restatement_on_a_million_counters(engine, port, seed, first, c2, c3):
    n = 1_000_000
    got = engine.philox_raw(seed=seed, first=first, count=n, c2=c2, c3=c3)
    idx = (np.uint64(first) + np.arange(n, dtype=np.uint64))  # wraps mod 2^64 like the device counter
    ctr = np.stack([(idx & np.uint64(0xFFFFFFFF)).astype(np.uint32), (idx >> np.uint64(32)).astype(np.uint32),
                    np.full(n, c2, np.uint32), np.full(n, c3, np.uint32)], axis=1)
    want = philox4x32_10(ctr, (seed & 0xFFFFFFFF, seed >> 32))
    assert np.array_equal(got, want)
    for i in (0, 1, n // 2, n - 1):  # and the C port, directly
        assert list(port.philox(ctr[i], [seed & 0xFFFFFFFF, seed >> 32])) == [int(x) for x in got[i]]
