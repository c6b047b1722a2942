"""The width-7 median selection network of align_reduce_kernel (csrc/align.cu median7_x16), restated in numpy and checked
against a sort-based median: sorted column pairs, a 4-merge shared by four outputs, the two middle ranks of the six shared
elements from the k-th-smallest-of-two-sorted-lists identity, the seventh element clamped between them."""
import numpy as np


def median7_x16(v):
    ql = {j: np.minimum(v[2 * j], v[2 * j + 1]) for j in range(1, 11)}
    qh = {j: np.maximum(v[2 * j], v[2 * j + 1]) for j in range(1, 11)}
    out = [None] * 16
    for p0 in range(0, 8, 2):
        a0, a1, b0, b1 = ql[p0 + 2], qh[p0 + 2], ql[p0 + 3], qh[p0 + 3]
        s1, t, s4, u = np.minimum(a0, b0), np.maximum(a0, b0), np.maximum(a1, b1), np.minimum(a1, b1)
        s2, s3 = np.minimum(t, u), np.maximum(t, u)
        for e in range(2):
            c0, c1 = (ql[p0 + 4], qh[p0 + 4]) if e else (ql[p0 + 1], qh[p0 + 1])
            r2 = np.minimum(np.minimum(s3, np.maximum(s2, c0)), np.maximum(s1, c1))
            r3 = np.minimum(np.minimum(s4, np.maximum(s3, c0)), np.maximum(s2, c1))
            k = 2 * (p0 + e)
            out[k] = np.maximum(r2, np.minimum(v[k + 1], r3))
            out[k + 1] = np.maximum(r2, np.minimum(v[k + 8], r3))
    return np.stack(out)


def test_network_equals_sorted_median():
    rs = np.random.RandomState(0)
    for data in (rs.randn(24, 20000).astype(np.float32), rs.randint(0, 4, size=(24, 20000)).astype(np.float32),
                 np.sort(rs.randn(24, 2000).astype(np.float32), axis=0), -np.sort(rs.randn(24, 2000).astype(np.float32), axis=0)):
        got = median7_x16(data)
        want = np.stack([np.median(data[k + 1:k + 8], axis=0) for k in range(16)])
        assert np.array_equal(got, want)
