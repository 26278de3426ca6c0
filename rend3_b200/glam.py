"""f32 restatement of the glam 0.25 formulas the rend3 host side uses.

glam is a third-party dependency of the reference (rend3/Cargo.toml:43, `glam = "0.25"`)
and is NOT vendored under /root/reference, so these are written from glam's published
algorithms (column-major Mat4, SSE2 `mul_vec4` accumulation order x,y,z,w).  They are only
used to *generate scene data* (view / projection / light matrices); both the CPU oracle and
the CUDA path receive the resulting bytes as inputs, so parity never depends on them.  They
are pinned end-to-end by the reference's coordinate-space / shadow goldens
(tests/test_oracle_golden.py).

Call sites in the reference: rend3/src/managers/camera.rs:88-107 (projections),
rend3/src/managers/directional/shadow_camera.rs:6-33 (look_at, transform_point3, inverse),
rend3-routine/src/uniforms.rs:41-43 (inverse), rend3-test/tests/*.rs (scene setup).

All matrices are numpy float32 arrays of shape (4, 4) stored as M[col, row] so that
`M.tobytes()` is exactly glam's / std430's column-major byte layout.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def vec3(x, y, z):
    return np.array([x, y, z], dtype=f32)


def vec4(x, y, z, w):
    return np.array([x, y, z, w], dtype=f32)


def mat4_cols(c0, c1, c2, c3):
    return np.array([c0, c1, c2, c3], dtype=f32)


def identity():
    return np.eye(4, dtype=f32)


def mul_vec4(m, v):
    """glam sse2 Mat4::mul_vec4: ((x_axis*v.x + y_axis*v.y) + z_axis*v.z) + w_axis*v.w, no FMA."""
    m = m.astype(f32, copy=False)
    v = np.asarray(v, dtype=f32)
    res = m[0] * v[0]
    res = res + m[1] * v[1]
    res = res + m[2] * v[2]
    res = res + m[3] * v[3]
    return res.astype(f32)


def mul(a, b):
    """glam Mat4::mul_mat4: column j of the result is a.mul_vec4(b.col(j))."""
    return np.array([mul_vec4(a, b[j]) for j in range(4)], dtype=f32)


def transform_point3(m, p):
    """glam Mat4::transform_point3 (assumes affine): x*px + y*py + z*pz + w."""
    p = np.asarray(p, dtype=f32)
    res = m[0] * p[0]
    res = res + m[1] * p[1]
    res = res + m[2] * p[2]
    res = res + m[3]
    return res[:3].astype(f32)


def dot3(a, b):
    a = np.asarray(a, dtype=f32)
    b = np.asarray(b, dtype=f32)
    return f32(f32(f32(a[0] * b[0]) + f32(a[1] * b[1])) + f32(a[2] * b[2]))


def cross(a, b):
    a = np.asarray(a, dtype=f32)
    b = np.asarray(b, dtype=f32)
    return np.array(
        [a[1] * b[2] - b[1] * a[2], a[2] * b[0] - b[2] * a[0], a[0] * b[1] - b[0] * a[1]],
        dtype=f32,
    )


def length3(a):
    return f32(np.sqrt(dot3(a, a)))


def normalize3(a):
    a = np.asarray(a, dtype=f32)
    # glam: self.mul(self.length_recip()) where length_recip = 1.0 / length
    return (a * f32(f32(1.0) / length3(a))).astype(f32)


def normalize_or_zero3(a):
    a = np.asarray(a, dtype=f32)
    ln = length3(a)
    rcp = f32(1.0) / ln if ln != 0 else f32(np.inf)
    if np.isfinite(rcp) and rcp > 0:
        return (a * rcp).astype(f32)
    return np.zeros(3, dtype=f32)


def from_translation(t):
    m = identity()
    m[3, :3] = np.asarray(t, dtype=f32)
    return m


def from_scale(s):
    m = identity()
    m[0, 0], m[1, 1], m[2, 2] = f32(s[0]), f32(s[1]), f32(s[2])
    return m


def from_rotation_x(angle):
    s, c = f32(np.sin(f32(angle))), f32(np.cos(f32(angle)))
    return mat4_cols([1, 0, 0, 0], [0, c, s, 0], [0, -s, c, 0], [0, 0, 0, 1])


def from_rotation_y(angle):
    s, c = f32(np.sin(f32(angle))), f32(np.cos(f32(angle)))
    return mat4_cols([c, 0, -s, 0], [0, 1, 0, 0], [s, 0, c, 0], [0, 0, 0, 1])


def from_rotation_z(angle):
    s, c = f32(np.sin(f32(angle))), f32(np.cos(f32(angle)))
    return mat4_cols([c, s, 0, 0], [-s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1])


def from_euler_xyz(a, b, c):
    """Mat4::from_euler(EulerRot::XYZ, a, b, c) = Rx(a) * Ry(b) * Rz(c) (examples/src/cube/mod.rs:100)."""
    return mul(mul(from_rotation_x(a), from_rotation_y(b)), from_rotation_z(c))


def quat_to_axes(q):
    """glam Mat3::from_quat axes for q = (x, y, z, w)."""
    x, y, z, w = [f32(v) for v in q]
    x2, y2, z2 = x + x, y + y, z + z
    xx, xy, xz = x * x2, x * y2, x * z2
    yy, yz, zz = y * y2, y * z2, z * z2
    wx, wy, wz = w * x2, w * y2, w * z2
    one = f32(1.0)
    return (
        np.array([one - (yy + zz), xy + wz, xz - wy], dtype=f32),
        np.array([xy - wz, one - (xx + zz), yz + wx], dtype=f32),
        np.array([xz + wy, yz - wx, one - (xx + yy)], dtype=f32),
    )


def from_scale_rotation_translation(scale, quat, translation):
    ax, ay, az = quat_to_axes(quat)
    s = np.asarray(scale, dtype=f32)
    t = np.asarray(translation, dtype=f32)
    return mat4_cols(
        list(ax * s[0]) + [0], list(ay * s[1]) + [0], list(az * s[2]) + [0], list(t) + [1]
    )


QUAT_IDENTITY = (0.0, 0.0, 0.0, 1.0)


def look_to_lh(eye, direction, up):
    f = normalize3(direction)
    s = normalize3(cross(up, f))
    u = cross(f, s)
    return mat4_cols(
        [s[0], u[0], f[0], 0],
        [s[1], u[1], f[1], 0],
        [s[2], u[2], f[2], 0],
        [-dot3(eye, s), -dot3(eye, u), -dot3(eye, f), 1],
    )


def look_at_lh(eye, center, up):
    eye = np.asarray(eye, dtype=f32)
    return look_to_lh(eye, np.asarray(center, dtype=f32) - eye, np.asarray(up, dtype=f32))


def look_at_rh(eye, center, up):
    eye = np.asarray(eye, dtype=f32)
    # look_to_rh(eye, dir, up) == look_to_lh(eye, -dir, up)
    return look_to_lh(eye, eye - np.asarray(center, dtype=f32), np.asarray(up, dtype=f32))


def perspective_infinite_reverse_lh(fovy, aspect, near):
    fovy = f32(fovy)
    half = f32(0.5) * fovy
    sin_fov, cos_fov = f32(np.sin(half)), f32(np.cos(half))
    h = f32(cos_fov / sin_fov)
    w = f32(h / f32(aspect))
    return mat4_cols([w, 0, 0, 0], [0, h, 0, 0], [0, 0, 0, 1], [0, 0, f32(near), 0])


def perspective_infinite_reverse_rh(fovy, aspect, near):
    fovy = f32(fovy)
    half = f32(0.5) * fovy
    f = f32(f32(1.0) / f32(np.tan(half)))
    return mat4_cols([f32(f / f32(aspect)), 0, 0, 0], [0, f, 0, 0], [0, 0, 0, -1], [0, 0, f32(near), 0])


def orthographic_lh(left, right, bottom, top, near, far):
    left, right, bottom, top, near, far = [f32(v) for v in (left, right, bottom, top, near, far)]
    rcp_w = f32(1.0) / (right - left)
    rcp_h = f32(1.0) / (top - bottom)
    r = f32(1.0) / (far - near)
    return mat4_cols(
        [rcp_w + rcp_w, 0, 0, 0],
        [0, rcp_h + rcp_h, 0, 0],
        [0, 0, r, 0],
        [-(left + right) * rcp_w, -(top + bottom) * rcp_h, -r * near, 1],
    )


def orthographic_rh(left, right, bottom, top, near, far):
    left, right, bottom, top, near, far = [f32(v) for v in (left, right, bottom, top, near, far)]
    rcp_w = f32(1.0) / (right - left)
    rcp_h = f32(1.0) / (top - bottom)
    r = f32(1.0) / (near - far)
    return mat4_cols(
        [rcp_w + rcp_w, 0, 0, 0],
        [0, rcp_h + rcp_h, 0, 0],
        [0, 0, r, 0],
        [-(left + right) * rcp_w, -(top + bottom) * rcp_h, r * near, 1],
    )


def inverse(m):
    """Mat4::inverse.  glam uses an f32 cofactor expansion whose exact rounding cannot be
    reproduced from here, so inverses are computed in f64 and rounded to
    f32; they are *inputs* to both the oracle and the CUDA path (uniforms.rs:41-43)."""
    a = m.astype(np.float64).T  # -> row-major math matrix
    inv = np.linalg.inv(a)
    return inv.T.astype(f32)


def to_radians(deg):
    return f32(f32(deg) * f32(np.pi / 180.0))
