"""Data-dependent helpers and the remaining table entries on the DEVICE (VERDICT r1:
these used to round-trip through host NumPy): validated integer-array indexing, stream compaction
(argwhere / boolean masks), isin, unravel_index, randint / binomial / permutation / shuffle / choice,
save / load, float64 / integer / stacked matmul, and NumPy's overlap semantics of the in-place family.
Reference behaviour: backend/numpy.py:73-75,84,105-138, tensor.py:503-515,598-659."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    import minidiff_b200.backend as backend

    backend.assert_live()
    return backend


def launches():
    from minidiff_b200.backend._lib import lib

    return int(lib.mdb_launch_count())


# ------------------------------------------------------------------ validated integer-array indexing
def test_out_of_range_index_arrays_raise_like_numpy(B):
    a_np = np.arange(40, dtype=np.float32).reshape(10, 4)
    a = B.asarray(a_np)
    for bad in ([0, 10], [-11, 2], [3, 99999999999]):
        with pytest.raises(IndexError, match="out of bounds"):
            a[B.asarray(np.array(bad))]
        with pytest.raises(IndexError):
            a[B.asarray(np.array(bad))] = 1.0
        with pytest.raises(IndexError):
            B.index_add(a, B.asarray(np.array(bad)), 1.0)
    with pytest.raises(IndexError):
        a[B.asarray(np.array([0, 1])), B.asarray(np.array([0, 4]))]
    with pytest.raises(IndexError):
        B.take_along_axis(a, B.asarray(np.array([[0], [10]] * 5)), axis=0)
    with pytest.raises(IndexError):
        B.put_along_axis(a, B.asarray(np.array([[4]] * 10)), 7.0, axis=1)
    np.testing.assert_array_equal(a.numpy(), a_np)              # nothing was written by the failed calls
    ok = a[B.asarray(np.array([-10, 9, -1, 0]))].numpy()        # the legal extremes still work
    np.testing.assert_array_equal(ok, a_np[[-10, 9, -1, 0]])


@pytest.mark.parametrize("dtype", [np.int8, np.int16, np.float32, np.float64, np.bool_])
def test_fancy_indexing_on_flipped_and_strided_views(B, dtype):
    rng = np.random.default_rng(0)
    a_np = (rng.integers(0, 100, (12, 7)) % (2 if dtype == np.bool_ else 100)).astype(dtype)
    a = B.asarray(a_np)
    idx = np.array([0, 5, -1, 3, 3, -12])
    for view_np, view in ((a_np[::-1], B.flip(a, 0)), (a_np[::-1, ::-1], B.flip(a)), (a_np[::2, ::-1], B.flip(a[::2], 1)),
                          (a_np.T[::-1], B.flip(a.T, 0))):
        k = np.clip(idx, -view_np.shape[0], view_np.shape[0] - 1)
        np.testing.assert_array_equal(view[B.asarray(k)].numpy(), view_np[k])
        cols = (np.arange(view_np.shape[0]) * 3 % view_np.shape[1]).reshape(-1, 1)
        np.testing.assert_array_equal(B.take_along_axis(view, B.asarray(cols), axis=1).numpy(),
                                      np.take_along_axis(view_np, cols, axis=1))
    # scatter through a flipped view lands on the right elements of the base
    b_np = np.zeros(10, dtype=dtype if dtype != np.bool_ else np.int8)
    b = B.asarray(b_np.copy())
    fv = B.flip(b)
    fv[B.asarray(np.array([0, 2, -1]))] = 1
    b_np[::-1][[0, 2, -1]] = 1
    np.testing.assert_array_equal(b.numpy(), b_np)


# ------------------------------------------------------------------ stream compaction
@pytest.mark.parametrize("shape", [(0,), (1,), (7,), (2048,), (2049,), (5000,), (37, 61), (5, 6, 7), (3, 1, 4, 2), (300, 300)])
def test_argwhere_matches_numpy(B, shape):
    rng = np.random.default_rng(sum(shape))
    for density in (0.0, 0.03, 0.5, 1.0):
        a_np = (rng.random(shape) < density) * rng.standard_normal(shape).astype(np.float32)
        got = B.argwhere(B.asarray(a_np))
        want = np.argwhere(a_np)
        assert got.dtype == np.int64 and got.shape == want.shape
        np.testing.assert_array_equal(got.numpy(), want)
    m = rng.random(shape) < 0.4
    np.testing.assert_array_equal(B.argwhere(B.asarray(m)).numpy(), np.argwhere(m))


def test_boolean_mask_getitem_setitem_and_index_add(B):
    rng = np.random.default_rng(5)
    a_np = rng.standard_normal((70, 33)).astype(np.float32)
    m_np = a_np > 0.3
    a, m = B.asarray(a_np.copy()), B.asarray(m_np)
    np.testing.assert_array_equal(a[m].numpy(), a_np[m_np])
    rows = m_np[:, 0]
    np.testing.assert_array_equal(a[B.asarray(rows)].numpy(), a_np[rows])
    a[m] = -1.0
    a_np[m_np] = -1.0
    np.testing.assert_array_equal(a.numpy(), a_np)
    B.index_add(a, m, 2.0)
    np.add.at(a_np, m_np, 2.0)
    np.testing.assert_array_equal(a.numpy(), a_np)
    none = B.asarray(np.zeros((70, 33), bool))
    assert a[none].shape == (0,)
    with pytest.raises(IndexError):
        a[B.asarray(np.zeros((70, 32), bool))]


# ------------------------------------------------------------------ isin / unravel_index
def test_isin_matches_numpy(B):
    rng = np.random.default_rng(1)
    e = rng.integers(-50, 50, (40, 30))
    t = rng.integers(-50, 50, 1500)          # more than one 1024-entry tile
    np.testing.assert_array_equal(B.isin(B.asarray(e), B.asarray(t)).numpy(), np.isin(e, t))
    np.testing.assert_array_equal(B.isin(B.asarray(e), B.asarray(t[:7]), invert=True).numpy(), np.isin(e, t[:7], invert=True))
    ef = rng.integers(0, 9, 100).astype(np.float32) / 2
    np.testing.assert_array_equal(B.isin(B.asarray(ef), [0.5, 2.0, 3]).numpy(), np.isin(ef, [0.5, 2.0, 3]))
    big = np.array([2**62 + 1, 2**62, -2**62 - 1])            # exact above 2^53: integer pairs compare as int64
    np.testing.assert_array_equal(B.isin(B.asarray(big), B.asarray(np.array([2**62 + 1]))).numpy(), [True, False, False])
    assert B.isin(B.asarray(e), B.asarray(np.array([], dtype=np.int64))).numpy().sum() == 0
    assert B.isin(B.asarray(e.T[::2]), B.asarray(t)).shape == e.T[::2].shape


def test_unravel_index_matches_numpy(B):
    rng = np.random.default_rng(2)
    shape = (7, 5, 11)
    flat = rng.integers(0, 7 * 5 * 11, (13, 4))
    got = B.unravel_index(B.asarray(flat), shape)
    want = np.unravel_index(flat, shape)
    assert len(got) == 3
    for g, w in zip(got, want):
        assert g.shape == w.shape and g.dtype == np.int64
        np.testing.assert_array_equal(g.numpy(), w)
    np.testing.assert_array_equal(B.unravel_index(B.asarray(np.array([5])), 9)[0].numpy(), [5])
    with pytest.raises(ValueError, match="out of bounds"):
        B.unravel_index(B.asarray(np.array([0, 385])), shape)
    with pytest.raises(ValueError):
        B.unravel_index(B.asarray(np.array([-1])), shape)


# ------------------------------------------------------------------ random family on the device
def test_randint_binomial_distributions(B):
    B.seed(7)
    l0 = launches()
    r = B.randint(-3, 12, size=(400, 500), dtype=np.int32)
    assert launches() - l0 == 2        # the draw + the one-thread kernel that advances the device-side stream position
    x = r.numpy()
    assert x.dtype == np.int32 and x.min() == -3 and x.max() == 11
    counts = np.bincount(x.ravel() + 3, minlength=15)
    assert np.abs(counts / x.size - 1 / 15).max() < 2e-3
    assert B.randint(5, size=3).dtype == np.int64 and B.randint(5).shape == ()
    with pytest.raises(ValueError):
        B.randint(3, 3)
    l0 = launches()
    b = B.binomial(20, 0.3, size=(200000,))
    assert launches() - l0 == 2
    y = b.numpy()
    assert y.dtype == np.int64 and 0 <= y.min() and y.max() <= 20
    assert abs(y.mean() - 6.0) < 0.03 and abs(y.var() - 4.2) < 0.08
    pv = np.array([0.0, 0.25, 1.0])
    z = B.binomial(8, B.asarray(pv), size=(50000, 3)).numpy()
    assert (z[:, 0] == 0).all() and (z[:, 2] == 8).all() and abs(z[:, 1].mean() - 2.0) < 0.03
    assert B.binomial(0, 0.5, size=4).numpy().sum() == 0
    with pytest.raises(ValueError):
        B.binomial(3, 1.5)


@pytest.mark.parametrize("n", [1, 2, 5, 2047, 2048, 2049, 4097, 100000, 300001])
def test_permutation_is_a_permutation(B, n):
    B.seed(n)
    p = B.permutation(n)
    x = p.numpy()
    assert x.dtype == np.int64 and x.shape == (n,)
    np.testing.assert_array_equal(np.sort(x), np.arange(n))
    if n > 1000:
        assert (x != np.arange(n)).mean() > 0.99                   # really shuffled
        assert abs(np.corrcoef(x, np.arange(n))[0, 1]) < 0.02
        q = B.permutation(n).numpy()
        assert (q != x).mean() > 0.99                               # the stream advances


def test_permutation_uniform_over_small_orders(B):
    B.seed(3)
    seen = {}
    for _ in range(600):
        seen[tuple(B.permutation(3).numpy())] = seen.get(tuple(B.permutation(3).numpy()), 0) + 1
    assert len(seen) == 6 and min(seen.values()) > 50


def test_shuffle_and_choice(B):
    B.seed(11)
    a_np = np.arange(60, dtype=np.float32).reshape(20, 3)
    a = B.asarray(a_np.copy())
    B.shuffle(a)
    s = a.numpy()
    np.testing.assert_array_equal(np.sort(s[:, 0]), a_np[:, 0])
    np.testing.assert_array_equal(s[:, 1] - s[:, 0], np.ones(20))      # rows stay intact
    np.testing.assert_array_equal(B.permutation(B.asarray(a_np)).numpy()[:, 2] % 3, 2 * np.ones(20))
    c = B.choice(10, size=(4, 5)).numpy()
    assert c.shape == (4, 5) and c.min() >= 0 and c.max() < 10
    nr = B.choice(B.asarray(np.arange(100, 130)), size=30, replace=False).numpy()
    np.testing.assert_array_equal(np.sort(nr), np.arange(100, 130))
    with pytest.raises(ValueError):
        B.choice(5, size=6, replace=False)
    w = np.array([0.1, 0.0, 0.6, 0.3])
    pc = B.choice(4, size=200000, p=B.asarray(w)).numpy()
    freq = np.bincount(pc, minlength=4) / pc.size
    assert freq[1] == 0 and np.abs(freq - w).max() < 5e-3
    pc2 = B.choice(B.asarray(np.array([7.0, 8.0, 9.0])), size=1000, p=[0, 0, 1]).numpy()
    assert (pc2 == 9.0).all()


def test_save_load_roundtrip(B, tmp_path):
    rng = np.random.default_rng(0)
    for arr in (rng.standard_normal((17, 5)).astype(np.float32), rng.integers(0, 9, (4, 3, 2)), np.array(3.5)):
        f = tmp_path / "t.npy"
        B.save(str(f), B.asarray(arr))
        back = B.load(str(f))
        assert isinstance(back, B.tensor_class) and back.dtype == arr.dtype and back.shape == arr.shape
        np.testing.assert_array_equal(back.numpy(), arr)
        np.testing.assert_array_equal(np.load(str(f)), arr)            # a plain .npy file NumPy can read
    # a strided view is saved as its logical contents
    v = B.asarray(np.arange(12.0).reshape(3, 4)).T[::2]
    B.save(str(tmp_path / "v.npy"), v)
    np.testing.assert_array_equal(np.load(str(tmp_path / "v.npy")), np.arange(12.0).reshape(3, 4).T[::2])


def test_tensor_level_save_load(tmp_path):
    import minidiff_b200 as md

    t = md.Tensor(np.arange(6, dtype=np.float32).reshape(2, 3), allow_grad=True)
    md.save(str(tmp_path / "x.npy"), t)
    u = md.load(str(tmp_path / "x.npy"))
    assert isinstance(u, md.Tensor) and u.shape == (2, 3)
    np.testing.assert_array_equal(u.as_numpy(), t.as_numpy())


# ------------------------------------------------------------------ float64 / integer / stacked matmul
def test_float64_matmul_needs_no_cubic_temporary(B):
    from minidiff_b200.backend._lib import lib
    import ctypes as C

    rng = np.random.default_rng(0)
    a, b = rng.standard_normal((700, 900)), rng.standard_normal((900, 650))
    da, db = B.asarray(a), B.asarray(b)
    l0 = launches()
    got = B.matmul(da, db)
    assert launches() - l0 == 1
    assert got.dtype == np.float64
    # 3072^3 in float64: an M*K*N temporary would be 232 GB (more than the device has); O(M*N) here
    big_a, big_b = rng.standard_normal((3072, 3072)), rng.standard_normal((3072, 3072))
    big = B.matmul(B.asarray(big_a), B.asarray(big_b))
    np.testing.assert_allclose(big[B.asarray(np.array([0, 1500, 3071]))].numpy(), big_a[[0, 1500, 3071]] @ big_b,
                               rtol=1e-11, atol=1e-10)
    del big
    np.testing.assert_allclose(got.numpy(), a @ b, rtol=1e-12, atol=1e-11)
    np.testing.assert_allclose(B.matmul(da[:300], da[:300].T).numpy(), a[:300] @ a[:300].T, rtol=1e-12, atol=1e-11)
    ai, bi = rng.integers(-9, 9, (33, 40)), rng.integers(-9, 9, (40, 21))
    gi = B.matmul(B.asarray(ai), B.asarray(bi))
    assert gi.dtype == np.int64
    np.testing.assert_array_equal(gi.numpy(), ai @ bi)
    np.testing.assert_allclose(B.matmul(B.asarray(a[0]), db).numpy(), a[0] @ b, rtol=1e-12, atol=1e-11)
    np.testing.assert_allclose(B.dot(da, B.asarray(b[:, 0])).numpy(), a @ b[:, 0], rtol=1e-12, atol=1e-11)
    np.testing.assert_allclose(B.tensordot(B.asarray(a.reshape(700, 30, 30)), B.asarray(b.reshape(30, 30, 650))).numpy(),
                               np.tensordot(a.reshape(700, 30, 30), b.reshape(30, 30, 650)), rtol=1e-11, atol=1e-10)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("sa,sb", [((6, 40, 30), (6, 30, 20)), ((2, 3, 17, 9), (3, 9, 5)), ((5, 1, 8, 12), (1, 4, 12, 7)),
                                   ((8, 33), (4, 33, 6)), ((3, 2, 10, 4), (4,)), ((7,), (2, 7, 5)),
                                   ((4, 130, 33), (4, 33, 257)), ((2, 1, 300, 129), (3, 129, 200))])
def test_stacked_matmul_is_one_launch_and_broadcasts(B, dtype, sa, sb):
    rng = np.random.default_rng(len(sa) + len(sb))
    a, b = rng.standard_normal(sa).astype(dtype), rng.standard_normal(sb).astype(dtype)
    da, db = B.asarray(a), B.asarray(b)
    l0 = launches()
    got = B.matmul(da, db)
    assert launches() - l0 == 1, "every matrix of the batch in ONE launch"
    want = np.matmul(a, b)
    assert got.shape == want.shape and got.dtype == want.dtype
    np.testing.assert_allclose(got.numpy(), want, rtol=2e-5 if dtype == np.float32 else 1e-12, atol=1e-5 if dtype == np.float32 else 1e-12)
    # transposed / strided views of stacked operands need no copies either
    if a.ndim >= 3 and b.ndim >= 3:
        at = B.asarray(np.ascontiguousarray(np.swapaxes(a, -1, -2)))
        l0 = launches()
        got_t = B.matmul(B.swapaxes(at, -1, -2), db)
        assert launches() - l0 == 1
        np.testing.assert_allclose(got_t.numpy(), want, rtol=2e-5 if dtype == np.float32 else 1e-12, atol=1e-5 if dtype == np.float32 else 1e-12)
    with pytest.raises(ValueError):
        B.matmul(da, B.asarray(np.zeros(sb[:-2] + (sb[-2] + 1 if len(sb) > 1 else sb[-1] + 1,) + (sb[-1:] if len(sb) > 1 else ()), dtype)))


# ------------------------------------------------------------------ overlap semantics of the in-place family
def test_inplace_ops_on_overlapping_views_match_numpy(B):
    rng = np.random.default_rng(9)
    a_np = rng.standard_normal((64, 64)).astype(np.float32)
    a = B.asarray(a_np.copy())
    a += a.T
    a_np += a_np.T
    np.testing.assert_array_equal(a.numpy(), a_np)
    x_np = np.arange(1000, dtype=np.float32)
    x = B.asarray(x_np.copy())
    x[1:] = x[:-1]
    x_np[1:] = x_np[:-1].copy()
    np.testing.assert_array_equal(x.numpy(), x_np)
    x -= x[::-1]
    x_np -= x_np[::-1].copy()
    np.testing.assert_array_equal(x.numpy(), x_np)
    r = B.asarray(a_np.copy())
    r *= r[3]                                   # a row of itself, broadcast
    np.testing.assert_array_equal(r.numpy(), a_np * a_np[3].copy())
    s = B.asarray(a_np.copy())
    l0 = launches()
    s += s                                      # exact aliasing stays ONE in-place launch
    assert launches() - l0 == 1
    np.testing.assert_array_equal(s.numpy(), a_np + a_np)


def test_arange_and_stacking_stay_on_the_device(B):
    for args, kw in (((10,), {}), ((2, 11), {}), ((0, 20, 3), {}), ((5, -5, -2), {}), ((0.0, 1.0, 0.125), {}),
                     ((3,), {"dtype": np.float32}), ((7, 2), {}), ((0, 2**40, 2**37), {})):
        got, want = B.arange(*args, **kw), np.arange(*args, **kw)
        assert got.dtype == want.dtype and got.shape == want.shape, (args, got.dtype, want.dtype)
        np.testing.assert_array_equal(got.numpy(), want)
    rows = [B.asarray(np.full((3,), i, np.float32)) for i in range(4)]
    l0 = launches()
    s = B.asarray(rows)
    assert s.shape == (4, 3) and launches() - l0 == 4          # one copy per row, nothing staged through the host
    np.testing.assert_array_equal(s.numpy(), np.stack([r.numpy() for r in rows]))


def test_rng_position_lives_on_the_device(B):
    """seed() makes the stream reproducible; a captured CUDA graph draws FRESH numbers on every replay
    (the stream position is device-resident and advanced by a kernel inside the graph)."""
    import minidiff_b200 as md

    B.seed(123)
    a1, b1 = B.rand(1000).numpy(), B.randn(7, 9).numpy()
    B.seed(123)
    a2, b2 = B.rand(1000).numpy(), B.randn(7, 9).numpy()
    np.testing.assert_array_equal(a1, a2)
    np.testing.assert_array_equal(b1, b2)
    assert not np.array_equal(a1, B.rand(1000).numpy())
    buf = md.Tensor(B.zeros((4096,), dtype=np.float64))

    def step():
        with md.no_grad():
            buf[...] = md.Tensor(B.rand(4096))
        return buf

    g = md.capture_graph(step)
    draws = []
    for _ in range(3):
        g.replay()
        draws.append(buf.as_numpy().copy())
    g.close()
    assert not np.array_equal(draws[0], draws[1]) and not np.array_equal(draws[1], draws[2])
    assert all(0.45 < d.mean() < 0.55 for d in draws)


def test_isin_large_test_set_uses_the_sorted_path(B):
    rng = np.random.default_rng(4)
    e = rng.integers(-10**6, 10**6, 300000)
    t = rng.integers(-10**6, 10**6, 50000)                   # > 4096 test elements: sort + binary search
    np.testing.assert_array_equal(B.isin(B.asarray(e), B.asarray(t)).numpy(), np.isin(e, t))
    np.testing.assert_array_equal(B.isin(B.asarray(e), B.asarray(t), invert=True).numpy(), np.isin(e, t, invert=True))
    ef = np.round(rng.standard_normal(200000), 2)
    tf = np.round(rng.standard_normal(9000), 2)
    tf[:3] = [np.nan, -0.0, np.inf]
    ef[:4] = [np.nan, 0.0, np.inf, -np.inf]
    np.testing.assert_array_equal(B.isin(B.asarray(ef), B.asarray(tf)).numpy(), np.isin(ef, tf))
    big = np.array([2**62 + 1, 2**62, -2**62 - 1, 5])
    tb = np.concatenate([np.arange(5000), [2**62 + 1, -2**62 - 1]])
    np.testing.assert_array_equal(B.isin(B.asarray(big), B.asarray(tb)).numpy(), [True, False, True, True])
