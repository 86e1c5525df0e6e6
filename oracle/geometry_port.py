"""numpy restatement of the reference's rotation conversions (TEST INFRASTRUCTURE).

Every function computes in the dtype of its input (feed float32 to mimic the
reference's fp32 arithmetic, float64 for a tighter checker).  Citations are
relative to /root/reference.
"""
import numpy as np


def _f(x):
    x = np.asarray(x)
    return x if x.dtype in (np.float32, np.float64) else x.astype(np.float32)


def rot6d_to_rotmat(x):
    """common/geometry.py:330-344 (and the SPIN twin :308-327, eps 1e-12 there).
    x (...,6) viewed as (M,3,2): a1 = elements 0,2,4; a2 = elements 1,3,5. Returns (M,3,3)
    whose COLUMNS are b1,b2,b3."""
    x = _f(x).reshape(-1, 3, 2)
    dt = x.dtype.type
    a1, a2 = x[:, :, 0], x[:, :, 1]
    b1 = a1 / np.maximum(np.sqrt((a1 * a1).sum(1, keepdims=True)), dt(1e-6))
    d = (b1 * a2).sum(1, keepdims=True)
    u = a2 - d * b1
    b2 = u / np.maximum(np.sqrt((u * u).sum(1, keepdims=True)), dt(1e-6))
    b3 = np.cross(b1, b2)
    return np.stack([b1, b2, b3], axis=-1)


def batch_rodrigues(aa):
    """common/geometry.py:22-65: theta = ||aa + 1e-8||, quaternion (w,x,y,z) = (cos t/2, sin t/2 * aa/theta),
    re-normalised, expanded to a rotation matrix.  Returns flat (M,9), row-major."""
    aa = _f(aa).reshape(-1, 3)
    dt = aa.dtype.type
    ang = np.sqrt(((aa + dt(1e-8)) ** 2).sum(1, keepdims=True))
    axis = aa / ang
    half = ang * dt(0.5)
    q = np.concatenate([np.cos(half), np.sin(half) * axis], axis=1)
    q = q / np.sqrt((q * q).sum(1, keepdims=True))
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    two = dt(2)
    return np.stack([w2 + x2 - y2 - z2, two * xy - two * wz, two * wy + two * xz,
                     two * wz + two * xy, w2 - x2 + y2 - z2, two * yz - two * wx,
                     two * xz - two * wy, two * wx + two * yz, w2 - x2 - y2 + z2], axis=1)


def angle_axis_to_rotation_matrix(aa):
    """common/kornia_geometry_conversion.py:125-201: Rodrigues with axis = aa/(theta+1e-6), blended with the
    first-order Taylor form I+[aa]x where theta^2 <= 1e-6.  Returns (M,3,3)."""
    aa = _f(aa).reshape(-1, 3)
    dt = aa.dtype.type
    t2 = (aa * aa).sum(1)
    th = np.sqrt(t2)
    w = aa / (th + dt(1e-6))[:, None]
    wx, wy, wz = w[:, 0], w[:, 1], w[:, 2]
    c, s = np.cos(th), np.sin(th)
    one = dt(1)
    full = np.stack([c + wx * wx * (one - c), wx * wy * (one - c) - wz * s, wy * s + wx * wz * (one - c),
                     wz * s + wx * wy * (one - c), c + wy * wy * (one - c), -wx * s + wy * wz * (one - c),
                     -wy * s + wx * wz * (one - c), wx * s + wy * wz * (one - c), c + wz * wz * (one - c)], axis=1)
    rx, ry, rz = aa[:, 0], aa[:, 1], aa[:, 2]
    k1 = np.ones_like(rx)
    taylor = np.stack([k1, -rz, ry, rz, k1, -rx, -ry, rx, k1], axis=1)
    big = (t2 > dt(1e-6))[:, None]
    return np.where(big, full, taylor).reshape(-1, 3, 3)


def rotation_matrix_to_quaternion(R, eps=1e-6):
    """common/geometry.py:153-233 on the 3x3 part: four-case selection on the TRANSPOSED matrix,
    quaternion (w,x,y,z)."""
    R = _f(R).reshape(-1, 3, 3)
    dt = R.dtype.type
    m = np.transpose(R, (0, 2, 1))          # rmat_t
    m00, m01, m02 = m[:, 0, 0], m[:, 0, 1], m[:, 0, 2]
    m10, m11, m12 = m[:, 1, 0], m[:, 1, 1], m[:, 1, 2]
    m20, m21, m22 = m[:, 2, 0], m[:, 2, 1], m[:, 2, 2]
    d2 = m22 < dt(eps)
    d01 = m00 > m11
    d0n1 = m00 < -m11
    one = dt(1)
    t0 = one + m00 - m11 - m22
    q0 = np.stack([m12 - m21, t0, m01 + m10, m20 + m02], -1)
    t1 = one - m00 + m11 - m22
    q1 = np.stack([m20 - m02, m01 + m10, t1, m12 + m21], -1)
    t2 = one - m00 - m11 + m22
    q2 = np.stack([m01 - m10, m20 + m02, m12 + m21, t2], -1)
    t3 = one + m00 + m11 + m22
    q3 = np.stack([t3, m12 - m21, m20 - m02, m01 - m10], -1)
    c0, c1, c2, c3 = d2 & d01, d2 & ~d01, ~d2 & d0n1, ~d2 & ~d0n1
    f = lambda c: c.astype(R.dtype)[:, None]
    q = q0 * f(c0) + q1 * f(c1) + q2 * f(c2) + q3 * f(c3)
    t = t0 * c0.astype(R.dtype) + t1 * c1.astype(R.dtype) + t2 * c2.astype(R.dtype) + t3 * c3.astype(R.dtype)
    with np.errstate(invalid="ignore", divide="ignore"):
        q = q / np.sqrt(t)[:, None]
    return q * dt(0.5)


def quaternion_to_angle_axis(q):
    """common/geometry.py:100-150 (w,x,y,z)."""
    q = _f(q).reshape(-1, 4)
    dt = q.dtype.type
    q1, q2, q3 = q[:, 1], q[:, 2], q[:, 3]
    s2 = q1 * q1 + q2 * q2 + q3 * q3
    s = np.sqrt(s2)
    c = q[:, 0]
    two_theta = dt(2) * np.where(c < 0, np.arctan2(-s, -c), np.arctan2(s, c))
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(s2 > 0, two_theta / s, dt(2) * np.ones_like(s))
    return np.stack([q1 * k, q2 * k, q3 * k], axis=1).astype(q.dtype)


def rotation_matrix_to_angle_axis(R):
    """common/geometry.py:68-97 (self-consistent w,x,y,z path; NaN -> 0)."""
    aa = quaternion_to_angle_axis(rotation_matrix_to_quaternion(R))
    aa[np.isnan(aa)] = 0
    return aa


# --- kornia known-answer helpers (common/kornia_geometry_conversion.py docstrings :322-324,:350-354) ----------
def normalize_quaternion(q, eps=1e-12):
    q = _f(q)
    return q / np.maximum(np.sqrt((q * q).sum(-1, keepdims=True)), q.dtype.type(eps))


def quaternion_to_rotation_matrix_xyzw(q):
    """common/kornia_geometry_conversion.py:341-393 (x,y,z,w)."""
    q = normalize_quaternion(_f(q).reshape(-1, 4))
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    return np.stack([1 - (ty * y + tz * z), ty * x - tz * w, tz * x + ty * w,
                     ty * x + tz * w, 1 - (tx * x + tz * z), tz * y - tx * w,
                     tz * x - ty * w, tz * y + tx * w, 1 - (tx * x + ty * y)], -1).reshape(-1, 3, 3)


def rotation_matrix_to_quaternion_kornia_xyzw(R, eps=1e-8):
    """common/kornia_geometry_conversion.py:230-307: trace-based selection, quaternion (x,y,z,w)."""
    R = _f(R).reshape(-1, 3, 3)
    dt = R.dtype.type
    tiny = np.finfo(R.dtype).tiny
    m00, m01, m02 = R[:, 0, 0], R[:, 0, 1], R[:, 0, 2]
    m10, m11, m12 = R[:, 1, 0], R[:, 1, 1], R[:, 1, 2]
    m20, m21, m22 = R[:, 2, 0], R[:, 2, 1], R[:, 2, 2]
    tr = m00 + m11 + m22
    sdiv = lambda a, b: a / np.maximum(b, tiny)
    with np.errstate(invalid="ignore"):
        sq = np.sqrt(tr + dt(1)) * dt(2)
        qp = np.stack([sdiv(m21 - m12, sq), sdiv(m02 - m20, sq), sdiv(m10 - m01, sq), dt(0.25) * sq], -1)
        sq = np.sqrt(dt(1) + m00 - m11 - m22 + dt(eps)) * dt(2)
        q1 = np.stack([dt(0.25) * sq, sdiv(m01 + m10, sq), sdiv(m02 + m20, sq), sdiv(m21 - m12, sq)], -1)
        sq = np.sqrt(dt(1) + m11 - m00 - m22 + dt(eps)) * dt(2)
        q2 = np.stack([sdiv(m01 + m10, sq), dt(0.25) * sq, sdiv(m12 + m21, sq), sdiv(m02 - m20, sq)], -1)
        sq = np.sqrt(dt(1) + m22 - m00 - m11 + dt(eps)) * dt(2)
        q3 = np.stack([sdiv(m02 + m20, sq), sdiv(m12 + m21, sq), dt(0.25) * sq, sdiv(m10 - m01, sq)], -1)
    w2 = np.where((m11 > m22)[:, None], q2, q3)
    w1 = np.where(((m00 > m11) & (m00 > m22))[:, None], q1, w2)
    return np.where((tr > 0)[:, None], qp, w1)


def rotation_matrix_to_angle_axis_kornia_quirk(R):
    """common/kornia_geometry_conversion.py:204-227: feeds the (x,y,z,w) quaternion into a routine that
    reads index 0 as cos(theta/2) -- internally inconsistent.  Reproduced verbatim
    for the explicit compatibility flag only."""
    return quaternion_to_angle_axis(rotation_matrix_to_quaternion_kornia_xyzw(R))
