"""Camera matrices and push-constant derivation.

Restates reina::graphics::Camera's two matrices (src/graphics/Camera.cpp:10-11):
    inverseProjection = inverse(glm::perspective(fovY, aspect, 0.1, 100))      (right-handed, depth -1..1)
    inverseView       = inverse(glm::lookAt(pos, pos + front, up = +y))
and the RtPushConsts defaults of src/Reina.cpp:142-155 (note defocus_multiplier / 100). Matrices use glm's memory
layout: M[c] is column c, flatten() gives the 16 floats the shaders see. glm is an un-pinned dependency absent from
this image; its formulas are the textbook ones below (parity unpinned, the matrices are inputs to both the kernels
and the oracle so they cannot cause a mismatch).
"""
import numpy as np

from . import abi


def perspective(fovy, aspect, znear, zfar):
    t = np.tan(fovy / 2.0)
    m = np.zeros((4, 4), np.float64)       # m[c][r]
    m[0][0] = 1.0 / (aspect * t)
    m[1][1] = 1.0 / t
    m[2][2] = -(zfar + znear) / (zfar - znear)
    m[2][3] = -1.0
    m[3][2] = -(2.0 * zfar * znear) / (zfar - znear)
    return m


def look_at(eye, center, up=(0.0, 1.0, 0.0)):
    eye, center, up = (np.asarray(v, np.float64) for v in (eye, center, up))
    f = center - eye
    f /= np.linalg.norm(f)
    s = np.cross(f, up)
    s /= np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4, dtype=np.float64)        # m[c][r]
    m[0][0], m[1][0], m[2][0] = s
    m[0][1], m[1][1], m[2][1] = u
    m[0][2], m[1][2], m[2][2] = -f
    m[3][0], m[3][1], m[3][2] = -np.dot(s, eye), -np.dot(u, eye), np.dot(f, eye)
    return m


def inverse_glm(m):
    # m[c][r] is the transpose of the mathematical matrix; inverse commutes with transpose
    return np.linalg.inv(m.T).T


def translate(v):
    m = np.eye(4, dtype=np.float32)
    m[3][:3] = v
    return m


def scale(s):
    m = np.eye(4, dtype=np.float32)
    s = np.broadcast_to(np.asarray(s, np.float32), (3,))
    m[0][0], m[1][1], m[2][2] = s
    return m


def rotate(axis, angle):
    """Rotation by `angle` radians about `axis`, glm layout (m[c][r])."""
    a = np.asarray(axis, np.float64)
    a = a / np.linalg.norm(a)
    c, s_ = np.cos(angle), np.sin(angle)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    R = np.eye(3) * c + s_ * K + (1 - c) * np.outer(a, a)
    m = np.eye(4, dtype=np.float64)
    m[:3, :3] = R
    return m.T.astype(np.float32)


def compose(*ms):
    """Mathematical product M0 * M1 * ... for matrices in glm layout (m[c][r])."""
    out = np.eye(4, dtype=np.float64)
    for m in ms:
        out = out @ np.asarray(m, np.float64).T
    return out.T.astype(np.float32)


# config/config.toml defaults (the reference's only configuration file)
DEFAULTS = dict(focus_dist=2.2, defocus_multiplier=1.5, samples_per_pixel=8, max_bounces=16, direct_clamp=100.0,
                indirect_clamp=10.0, bloom_radius=5.0, bloom_threshold=1.0, bloom_intensity=0.05, exposure=1.0)


def push_constants(width, height, pos, look, fovy_deg, total_emissive_weight=0.0, sample_batch=0, **over):
    cfg = dict(DEFAULTS)
    cfg.update(over)
    pc = abi.RtPushConsts()
    inv_view = inverse_glm(look_at(pos, look)).astype(np.float32)
    inv_proj = inverse_glm(perspective(np.radians(fovy_deg), width / height, 0.1, 100.0)).astype(np.float32)
    pc.invView[:] = [float(x) for x in inv_view.reshape(-1)]
    pc.invProjection[:] = [float(x) for x in inv_proj.reshape(-1)]
    pc.sampleBatch = sample_batch
    pc.totalEmissiveWeight = total_emissive_weight
    pc.focusDist = cfg["focus_dist"]
    pc.defocusMultiplier = cfg["defocus_multiplier"] / 100.0
    pc.directClamp = cfg["direct_clamp"]
    pc.indirectClamp = cfg["indirect_clamp"]
    pc.samplesPerPixel = cfg["samples_per_pixel"]
    pc.maxBounces = cfg["max_bounces"]
    return pc
