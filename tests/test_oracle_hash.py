"""CPU: the C oracle of Hash3DAnchored against an independent numpy restatement of
reference gfnerf/bindings/field/Hash3DAnchored_cuda.cu:11-155 (indices and weights exact,
blend to one fp16 ulp), plus structural properties."""
import numpy as np

from oracle import oracle as orc
from tests.helpers import hash_inputs


def numpy_hash(feat, prim, bias, pts, anchors, scales):
    n = pts.shape[0]
    local = feat.shape[0] // 16
    n_vol = prim.shape[1]
    out = np.zeros((n, 32), np.float32)
    rows = np.zeros((n, 16, 8), np.int64)
    ws = np.zeros((n, 16, 8), np.float32)
    feat16 = feat.astype(np.float16).astype(np.float32)
    for l in range(16):
        p = (pts * np.float32(scales[l])).astype(np.float32) + bias[l * n_vol + anchors]   # bias == 0: exact
        fl = np.floor(p)
        pos = fl.astype(np.uint32)
        frac = (p - fl).astype(np.float32)
        pr = prim[l, anchors].astype(np.uint32)                                            # [n,3]
        one = np.float32(1)
        for d in range(8):
            dx, dy, dz = (d >> 2) & 1, (d >> 1) & 1, d & 1
            h = ((pos[:, 0] + np.uint32(dx)) * pr[:, 0]) ^ ((pos[:, 1] + np.uint32(dy)) * pr[:, 1]) ^ \
                ((pos[:, 2] + np.uint32(dz)) * pr[:, 2])
            # the level offset l * local is added to a SCALAR pointer in the reference (Hash3DAnchored_cuda.cu:38):
            # in rows that is l * local / 2 -- consecutive levels overlap by half a window
            rows[:, l, d] = (l * local) // 2 + (h % np.uint32(local))
            wa = frac[:, 0] if dx else one - frac[:, 0]
            wb = frac[:, 1] if dy else one - frac[:, 1]
            wc = frac[:, 2] if dz else one - frac[:, 2]
            ws[:, l, d] = ((wa * wb).astype(np.float32) * wc).astype(np.float32)
        for k in range(2):
            acc = np.zeros(n, np.float64)
            for d in range(8):
                acc += ws[:, l, d].astype(np.float64) * feat16[rows[:, l, d], k].astype(np.float64)
            out[:, l * 2 + k] = acc.astype(np.float32).astype(np.float16).astype(np.float32)
    return out, rows, ws


def test_level_scales_match_closed_form():
    s = orc.hash_level_scales()
    assert s[0] == 8.0 and s[15] == 1024.0
    np.testing.assert_allclose(s, 2.0 ** (3 + 7 * np.arange(16) / 15), rtol=2e-6)  # the exponent 7*l/15+3 is itself rounded to fp32


def test_oracle_vs_numpy_restatement():
    with np.errstate(over="ignore"):
        feat, prim, bias, pts, anchors = hash_inputs(2000, 7, 12, seed=3)
        scales = orc.hash_level_scales()
        out, idx = orc.hash_forward(feat, prim, bias, pts, anchors, scales, want_idx=True)
        ref_out, ref_rows, _ = numpy_hash(feat, prim, bias, pts, anchors, scales)
    assert np.array_equal(idx.astype(np.int64), ref_rows)           # bit-exact indices
    # fp16-rounded outputs: at most one fp16 ulp apart (fp64 blend vs FMA chain)
    ulp = np.maximum(np.abs(ref_out), 2.0 ** -14) * 2.0 ** -10
    assert np.all(np.abs(out - ref_out) <= ulp)
    assert (out == ref_out).mean() > 0.99


def test_forward_values_are_fp16_representable():
    feat, prim, bias, pts, anchors = hash_inputs(512, 3, 10, seed=5)
    out = orc.hash_forward(feat, prim, bias, pts, anchors)
    assert np.array_equal(out, out.astype(np.float16).astype(np.float32))


def test_non_pow2_local_size_and_edges():
    # local_size that is a multiple of 16 but not a power of two, points on cell borders / domain corners
    rng = np.random.RandomState(0)
    local = 48 * 16
    feat = rng.uniform(-1, 1, size=(16 * local, 2)).astype(np.float32)
    _, prim, bias, _, _ = hash_inputs(8, 2, 10, seed=1)
    pts = np.array([[0, 0, 0], [1, 1, 1], [0.5, 0.25, 0.125], [0.999999, 0, 1]], np.float32)
    anchors = np.array([0, 1, 1, 0], np.int64)
    out, idx = orc.hash_forward(feat, prim, bias, pts, anchors, want_idx=True)
    # level l reaches rows [l * local / 2, l * local / 2 + local): nothing beyond 8.5 * local
    base = (np.arange(16) * local // 2).reshape(1, 16, 1)
    assert np.all(idx >= base) and np.all(idx < base + local) and idx.max() < 17 * local // 2
    # a point exactly on a lattice node takes the value of corner 000 (weights 1,0,..)
    f16 = feat.astype(np.float16).astype(np.float32)
    np.testing.assert_array_equal(out[0, 0:2], f16[idx[0, 0, 0]])


def test_backward_is_adjoint_of_forward():
    """<J^T g, table> == <g, J table> up to the two fp16 quantisations of the reference backward."""
    feat, prim, bias, pts, anchors = hash_inputs(300, 4, 9, seed=11)
    rng = np.random.RandomState(1)
    # fp16-exact table and gradients that survive x128 -> fp16 exactly: then only w*g rounding remains
    feat = feat.astype(np.float16).astype(np.float32)
    g = (rng.randint(-8, 9, size=(300, 32)) / 64.0).astype(np.float32)
    gt = orc.hash_backward(feat.shape[0] // 16, prim, bias, pts, anchors, g)
    out, idx = orc.hash_forward(feat, prim, bias, pts, anchors, want_idx=True)
    lhs = float((gt * feat).sum())
    rhs = float((g.astype(np.float64) * out).sum())
    assert abs(lhs - rhs) <= 2e-3 * max(1.0, abs(rhs))
    # rows that no sample touches get no gradient
    touched = np.zeros(feat.shape[0], bool)
    touched[idx.reshape(-1)] = True
    assert np.all(gt[~touched] == 0)


def test_backward_skips_zero_gradients_and_empty_input():
    feat, prim, bias, pts, anchors = hash_inputs(64, 2, 8, seed=2)
    gt = orc.hash_backward(1 << 8, prim, bias, pts, anchors, np.zeros((64, 32), np.float32))
    assert not gt.any()
    out = orc.hash_forward(feat, prim, bias, pts[:0], anchors[:0])
    assert out.shape == (0, 32)
