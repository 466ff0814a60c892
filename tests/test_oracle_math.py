"""Oracle math helpers vs (a) the doctest known-answer vectors of the reference's
utils/transformation.py and (b) golden outputs of the reference's utils/math.py (tests/golden/make_golden.py)."""
import ctypes as C

import numpy as np

from oracle import cphys

L = cphys.lib()
dp = C.POINTER(C.c_double)


def _call(fn, *arrs_and_out):
    args = []
    for a in arrs_and_out:
        if isinstance(a, np.ndarray):
            args.append(a.ctypes.data_as(dp))
        else:
            args.append(a)
    fn(*args)


def quat_mul(a, b):
    o = np.zeros(4)
    _call(L.eo_quat_mul, np.ascontiguousarray(a, dtype=float), np.ascontiguousarray(b, dtype=float), o)
    return o


def test_transformation_doctest_vectors():
    # utils/transformation.py:1382-1383 quaternion_multiply([4, 1, -2, 3], [8, -5, 6, 7]) == [28, -44, -14, 48]
    assert np.allclose(quat_mul([4, 1, -2, 3], [8, -5, 6, 7]), [28, -44, -14, 48])
    # utils/transformation.py:1413-1415 q * inverse(q) == identity
    q = np.array([0.3, -0.5, 0.7, 0.2])
    qi = np.zeros(4)
    _call(L.eo_quat_inv, q, qi)
    assert np.allclose(quat_mul(q, qi), [1, 0, 0, 0])
    # utils/transformation.py:1270-1277 quaternion_matrix([0.99810947, 0.06146124, 0, 0]) == rot(0.123, x);
    # [0,1,0,0] -> diag(1,-1,-1): checked through transform_vec 'root' (R^T v)
    v = np.array([0.0, 1.0, 0.0])
    o = np.zeros(3)
    _call(L.eo_transform_vec, v, np.array([0.99810947, 0.06146124, 0, 0]), C.c_int(0), o)
    assert np.allclose(o, [0, np.cos(0.123), -np.sin(0.123)], atol=1e-7)
    _call(L.eo_transform_vec, np.array([1.0, 2.0, 3.0]), np.array([0.0, 1.0, 0, 0]), C.c_int(0), o)
    assert np.allclose(o, [1, -2, -3])
    # quaternion_from_euler 'sxyz' equals the composition qz*qy*qx (static xyz), consistent with :1200 convention
    e = np.array([0.3, -0.8, 1.1])
    qe = np.zeros(4)
    L.eo_quat_from_euler(C.c_double(e[0]), C.c_double(e[1]), C.c_double(e[2]), qe.ctypes.data_as(dp))
    qx = [np.cos(e[0] / 2), np.sin(e[0] / 2), 0, 0]
    qy = [np.cos(e[1] / 2), 0, np.sin(e[1] / 2), 0]
    qz = [np.cos(e[2] / 2), 0, 0, np.sin(e[2] / 2)]
    assert np.allclose(qe, quat_mul(qz, quat_mul(qy, qx)))


def test_math_helpers_golden(golden):
    g = golden('math_helpers')
    n = g['q'].shape[0]
    for i in range(n):
        o4 = np.zeros(4)
        _call(L.eo_quat_mul, g['q'][i].copy(), g['q2'][i].copy(), o4)
        assert np.allclose(o4, g['mul'][i], rtol=0, atol=1e-15)
        _call(L.eo_quat_inv, g['inv_in'][i].copy(), o4)
        assert np.allclose(o4, g['inv'][i], rtol=1e-15, atol=1e-15)
        L.eo_quat_from_euler(C.c_double(g['eul'][i, 0]), C.c_double(g['eul'][i, 1]), C.c_double(g['eul'][i, 2]),
                             o4.ctypes.data_as(dp))
        assert np.allclose(o4, g['from_euler'][i], rtol=0, atol=1e-15)
        _call(L.eo_de_heading, g['q'][i].copy(), o4)
        assert np.allclose(o4, g['de_heading'][i], rtol=0, atol=1e-15)
        o3 = np.zeros(3)
        _call(L.eo_transform_vec, g['v'][i].copy(), g['q'][i].copy(), C.c_int(0), o3)
        assert np.allclose(o3, g['tv_root'][i], rtol=0, atol=1e-14)
        _call(L.eo_transform_vec, g['v'][i].copy(), g['q'][i].copy(), C.c_int(1), o3)
        assert np.allclose(o3, g['tv_heading'][i], rtol=0, atol=1e-14)
        ax, ang = np.zeros(3), C.c_double()
        L.eo_rotation_from_quat(g['q'][i].copy().ctypes.data_as(dp), ax.ctypes.data_as(dp), C.byref(ang))
        assert np.allclose(ax * ang.value, g['rot_from_quat'][i], rtol=0, atol=1e-14)
        for key, tr in (('qvel_fd_none', 0), ('qvel_fd_heading', 1)):
            out = np.zeros(58)
            L.eo_qvel_fd(C.c_int(59), g['qa'][i].copy().ctypes.data_as(dp), g['qb'][i].copy().ctypes.data_as(dp),
                         C.c_double(1 / 30.0), C.c_int(tr), out.ctypes.data_as(dp))
            assert np.allclose(out, g[key][i], rtol=1e-12, atol=1e-11), key
        out = np.zeros(63)
        L.eo_angvel_fd(C.c_int(21), g['bq0'][i].copy().ctypes.data_as(dp), g['bq1'][i].copy().ctypes.data_as(dp),
                       C.c_double(1 / 30.0), out.ctypes.data_as(dp))
        assert np.allclose(out, g['angvel_fd'][i], rtol=1e-12, atol=1e-11)
