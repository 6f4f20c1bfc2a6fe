"""Camera matrices the reference obtains from Taichi GGUI's ti.ui.Camera (scene.py:188-191,
233-237; C++/glm inside the Taichi wheel). Written out explicitly: glm::lookAt (right-handed)
and glm::perspective with GL -1..1 depth, as the comment at scene.py:186-187 describes.
Unpinned by any reference test."""
import math

import numpy as np


def look_at(eye, center, up=(0.0, 1.0, 0.0)):
    """glm::lookAtRH -> row-major 4x4 float64 view matrix."""
    eye = np.asarray(eye, np.float64)
    center = np.asarray(center, np.float64)
    up = np.asarray(up, np.float64)
    f = center - eye
    f = f / np.linalg.norm(f)
    s = np.cross(f, up)
    s = s / np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4)
    m[0, :3] = s
    m[1, :3] = u
    m[2, :3] = -f
    m[0, 3] = -np.dot(s, eye)
    m[1, 3] = -np.dot(u, eye)
    m[2, 3] = np.dot(f, eye)
    return m


def perspective(fovy_rad, aspect, z_near=0.01, z_far=10.0):
    """glm::perspectiveRH_NO -> row-major 4x4 float64 projection matrix."""
    g = 1.0 / math.tan(fovy_rad / 2.0)
    m = np.zeros((4, 4))
    m[0, 0] = g / aspect
    m[1, 1] = g
    m[2, 2] = -(z_far + z_near) / (z_far - z_near)
    m[2, 3] = -(2.0 * z_far * z_near) / (z_far - z_near)
    m[3, 2] = -1.0
    return m


def default_camera_matrices(width, height, pos=(0.4, 0.5, 2.0), target=(0.0, 0.0, 0.0), fov_deg=50.0):
    """Reference defaults: camera (0.4,0.5,2)->origin, up +Y (scene.py:28-30), fov 50 deg
    (pathtracer.py:89), near 0.01 / far 10 (scene.py:190-191). Returns float32 (pos, view, proj)."""
    view = look_at(pos, target)
    proj = perspective(math.radians(fov_deg), width / height)
    return (np.asarray(pos, np.float32), np.ascontiguousarray(view, np.float32), np.ascontiguousarray(proj, np.float32))
