"""CPU: the job map of the persistent BPTT kernel (csrc/recurrent_bwd.cu: rb_tiling), a host-only function of the dims and
the number of co-resident CTA pairs. The kernel relies on these properties:
  * every output column of the three data-gradient GEMMs, of d z and of the two small GEMMs belongs to exactly one tile;
  * the (tile, K part) jobs of a kind fit the pairs of their class (62 "big" / 12 "small" of the 74 pairs of a B200);
  * the K parts of a tile partition the k-blocks of the 4H gate gradients without an empty part;
  * the split counts stay within the slot counts the workspace is sized for (kernels.cuh: RB_MAX_SPLIT_*).
"""
import ctypes
import itertools

import pytest

from sscvae import _lib

PAIRS_B200 = 74
MAX_SPLIT = dict(A=2, B=4, X=4, Z=16)


def _tiling(dims, batch=256, boxes=36, pairs=PAIRS_B200):
    L = _lib.lib()
    h = ctypes.c_void_p()
    _lib.check(L.sscvae_create(ctypes.byref(dims), ctypes.byref(h)))
    out = (ctypes.c_int32 * 12)()
    rc = L.sscvae_debug_bptt_tiling(h, batch, boxes, pairs, out)
    L.sscvae_destroy(h)
    assert rc in (0, 1)
    if rc == 0:
        return None
    keys = ["nbig", "nsmall", "nA", "splitA", "nB", "splitB", "nX", "splitX", "nZt", "splitZ", "n4", "N4"]
    return dict(zip(keys, list(out)))


def _dims(F, E, H, A, Z, V=10000, sv=1, tied=1):
    return _lib.SscvaeDims(F, E, H, A, Z, V, 20, sv, 0, tied, 0, 1, 1.0, 0.5, 0)


def _pad(x, m=64):
    return (x + m - 1) // m * m


def _check(t, F, H, Z, pairs):
    Fp, Hp, Zp, Gp = _pad(F), _pad(H), _pad(Z), _pad(4 * H)
    KX, kbG = Fp + 2 * Hp, Gp // 64
    assert t["nbig"] + t["nsmall"] == pairs and t["nsmall"] >= 1
    # coverage
    assert (t["nA"] - 1) * 128 < KX <= t["nA"] * 128
    assert t["nB"] * 64 == Hp
    assert t["nX"] * 128 == 2 * Hp
    assert (t["nZt"] - 1) * 96 < Zp <= t["nZt"] * 96
    assert (t["n4"] - 1) * t["N4"] < H <= t["n4"] * t["N4"]
    assert t["N4"] % 16 == 0 and 16 <= t["N4"] <= 128
    # the jobs fit their pairs
    assert t["nA"] * t["splitA"] <= t["nbig"]
    assert t["nB"] * t["splitB"] <= t["nbig"]
    assert t["nX"] * t["splitX"] <= t["nbig"]
    assert t["nZt"] * t["splitZ"] <= t["nsmall"]
    assert t["n4"] <= t["nsmall"]
    # K parts: part i covers k-blocks [i*kbG//S, (i+1)*kbG//S): a partition without an empty part iff S <= kbG
    for k, cap in (("splitA", MAX_SPLIT["A"]), ("splitB", MAX_SPLIT["B"]), ("splitX", MAX_SPLIT["X"]), ("splitZ", MAX_SPLIT["Z"])):
        S = t[k]
        assert 1 <= S <= min(cap, kbG)
        edges = [i * kbG // S for i in range(S + 1)]
        assert edges[0] == 0 and edges[-1] == kbG and all(b > a for a, b in zip(edges, edges[1:]))


def test_dims_y_on_a_b200():
    t = _tiling(_dims(2048, 600, 900, 768, 150))
    assert t == dict(nbig=62, nsmall=12, nA=31, splitA=2, nB=15, splitB=4, nX=15, splitX=4, nZt=2, splitZ=6, n4=12, N4=80)
    _check(t, 2048, 900, 150, PAIRS_B200)


def test_dims_d_on_a_b200():
    t = _tiling(_dims(2048, 1000, 1200, 768, 150, tied=0))
    assert t is not None and t["splitA"] == 1          # 35 tiles of [x_hat | h1 | h_dec] on 62 pairs: no room for a K split
    _check(t, 2048, 1200, 150, PAIRS_B200)


@pytest.mark.parametrize("F,H,Z", list(itertools.product([64, 512, 2048], [32, 40, 64, 300, 900, 1200, 1600], [8, 24, 150])))
@pytest.mark.parametrize("pairs", [8, 33, 66, 74])
def test_properties_over_a_grid_of_shapes(F, H, Z, pairs):
    t = _tiling(_dims(F, 600, H, 48, Z), batch=37, boxes=9, pairs=pairs)
    if t is not None:
        _check(t, F, H, Z, pairs)


def test_refusals():
    assert _tiling(_dims(2048, 600, 900, 768, 151)) is None          # odd Z: the latent stage reads 8-byte pairs
    assert _tiling(_dims(2048, 600, 900, 768, 150), batch=257) is None
    assert _tiling(_dims(2048, 600, 900, 768, 150), pairs=3) is None
    assert _tiling(_dims(2048, 600, 2400, 768, 150)) is None         # [x_hat | h1 | h_dec] needs more tiles than big pairs
