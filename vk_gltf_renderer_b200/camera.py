"""Camera math + per-frame structs for the path tracer boundary.

Mirrors what GltfRenderer::onRender writes into SceneFrameInfo every frame
(reference src/renderer.cpp:675-705) and PathTracer::setupPushConstant into the push constants
(src/renderer_pathtracer.cpp:1496-1574).  nvutils::CameraManipulator is external (nvpro_core2);
its matrices are glm::lookAt and glm::perspectiveRH_ZO with the Vulkan Y flip ([1][1] *= -1).
All arithmetic is done in float64 and rounded once to float32.
"""
import math

import numpy as np

from . import abi


def look_at(eye, center, up):
    eye, center, up = (np.asarray(v, np.float64) for v in (eye, center, up))
    f = center - eye
    f /= np.linalg.norm(f)
    s = np.cross(f, up)
    s /= np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4)
    m[0, :3], m[1, :3], m[2, :3] = s, u, -f
    m[0, 3], m[1, 3], m[2, 3] = -s @ eye, -u @ eye, f @ eye
    return m


def perspective_vk(yfov, aspect, znear, zfar):
    t = math.tan(yfov / 2.0)
    m = np.zeros((4, 4))
    m[0, 0] = 1.0 / (aspect * t)
    m[1, 1] = -1.0 / t  # Vulkan clip space: +Y down
    m[2, 2] = zfar / (znear - zfar)
    m[3, 2] = -1.0
    m[2, 3] = -(zfar * znear) / (zfar - znear)
    return m


def ortho_vk(xmag, ymag, znear, zfar):
    m = np.eye(4)
    m[0, 0] = 1.0 / xmag
    m[1, 1] = -1.0 / ymag
    m[2, 2] = -1.0 / (zfar - znear)
    m[2, 3] = -znear / (zfar - znear)
    return m


def fit_camera(lo, hi, yfov=math.radians(45.0), aspect=1.0):
    """CameraManipulator::fit(bbox) analogue used when the glTF has no camera
    (gltf_camera_utils.hpp:80-86): look at the box centre from +Z far enough to see the bounding sphere."""
    from .scene import Camera
    lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
    c = (lo + hi) * 0.5
    r = float(np.linalg.norm(hi - lo) * 0.5)
    cam = Camera()
    cam.yfov = yfov
    d = r / math.sin(min(yfov, 2 * math.atan(math.tan(yfov / 2) * aspect)) / 2.0)
    cam.eye = (c + np.array([0, 0, d])).astype(np.float32)
    cam.center = c.astype(np.float32)
    cam.znear, cam.zfar = 0.001 * r, 100.0 * r
    return cam


def _glm(m):
    return np.ascontiguousarray(np.asarray(m, np.float64).T.reshape(16), np.float32)


def make_frame_info(cam, width, height, *, use_hdr=True, env_rotation=0.0, env_intensity=1.0, env_blur=0.0,
                    solid_background=False, background=(0, 0, 0), infinite_plane=False, plane_distance=0.0,
                    plane_color=(0.5, 0.5, 0.5), plane_metallic=0.0, plane_roughness=0.5, shadow_catcher=False, catcher_darkness=0.0):
    """SceneFrameInfo as filled at src/renderer.cpp:677-700 (projection uses the *window* aspect)."""
    view = look_at(cam.eye, cam.center, cam.up)
    if cam.type == "orthographic":
        proj = ortho_vk(cam.xmag, cam.ymag, cam.znear, cam.zfar)
    else:
        proj = perspective_vk(cam.yfov, width / float(height), cam.znear, cam.zfar)
    fi = abi.FrameInfo()
    fi.viewMatrix[:] = _glm(view).tolist()
    fi.projInv[:] = _glm(np.linalg.inv(proj)).tolist()
    fi.viewInv[:] = _glm(np.linalg.inv(view)).tolist()
    fi.viewProjMatrix[:] = _glm(proj @ view).tolist()
    fi.prevMVP[:] = _glm(proj @ view).tolist()
    fi.jitter[:] = [0.0, 0.0]
    fi.imageSize[:] = [float(width), float(height)]
    fi.flags = ((abi.SCENE_IS_ORTHOGRAPHIC if cam.type == "orthographic" else 0)
                | (abi.SCENE_USE_SOLID_BACKGROUND if solid_background else 0)
                | (abi.SCENE_USE_HDR_ENVIRONMENT if use_hdr else 0)
                | (abi.SCENE_USE_INFINITE_PLANE if infinite_plane else 0)
                | (abi.SCENE_INFINITE_PLANE_SHADOW_CATCHER if (infinite_plane and shadow_catcher) else 0))
    fi.envRotation = env_rotation
    fi.envBlur = env_blur
    fi.envIntensity = env_intensity
    fi.backgroundColor[:] = list(background)
    fi.visualization = 0
    fi.infinitePlaneDistance = plane_distance
    fi.infinitePlaneBaseColor[:] = list(plane_color)
    fi.infinitePlaneMetallic = plane_metallic
    fi.infinitePlaneRoughness = plane_roughness
    fi.shadowCatcherDarkenAmount = max(catcher_darkness, 0.0)   # src/renderer.cpp:700
    return fi


def make_push_constant(cam, height, *, frame_count, total_samples, num_samples=1, max_depth=5,
                       firefly_clamp=10.0, tex_grad_scale=1.0, aperture=0.0, focal_distance=None):
    """PathtracePushConstant as filled by setupPushConstant (renderer_pathtracer.cpp:1496-1574):
    autofocus focal distance = |eye - center| (:1508-1512); pixelAngle = 2|projInv[1][1]|/H (:1570-1571);
    ePtFirstFrame iff frameCount == 0 (:1542)."""
    pc = abi.PushConstant()
    pc.maxDepth = max_depth
    pc.frameCount = frame_count
    pc.fireflyClampThreshold = firefly_clamp
    pc.texGradScale = tex_grad_scale
    pc.numSamples = num_samples
    pc.totalSamples = total_samples
    pc.focalDistance = (float(np.linalg.norm(np.asarray(cam.eye, np.float64) - np.asarray(cam.center, np.float64)))
                        if focal_distance is None else focal_distance)
    pc.aperture = aperture
    pc.flags = abi.PT_FIRST_FRAME if frame_count == 0 else 0
    if cam.type == "orthographic":
        proj_inv_11 = cam.ymag
    else:
        proj_inv_11 = math.tan(cam.yfov / 2.0)
    pc.pixelAngle = np.float32(2.0 * abs(proj_inv_11) / max(float(height), 1.0))
    pc.mouseCoord[:] = [-1.0, -1.0]
    return pc
