"""Philox4x32-10 known-answer vectors (Random123 kat_vectors) against the oracle's
independent implementation; the CUDA implementation is checked against the same
vectors in tests/test_gpu_parity.py."""

import numpy as np

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def test_philox_known_answers(oracle):
    for ctr, key, want in KAT:
        got = oracle.philox4x32_10(ctr, key)
        assert tuple(int(x) for x in got) == want


def test_uniform53_range_and_pairs(oracle):
    us = np.array([oracle.uniform53(7, i, 3, 1, pair) for i in range(2000) for pair in (0, 1)])
    assert us.min() >= 0.0 and us.max() < 1.0
    assert abs(us.mean() - 0.5) < 0.02
    # pair 0 and pair 1 are different words of the same block
    assert oracle.uniform53(7, 5, 3, 1, 0) != oracle.uniform53(7, 5, 3, 1, 1)
