"""ctypes binding of libharpb200.so (the C ABI declared in include/harp_b200.h).

There is deliberately NO fallback: if the shared library is missing or cannot be loaded, importing the CUDA-backed ops
raises.  (The oracle under `oracle/` is test infrastructure and is never imported from here.)
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HARP_LIB") or os.path.join(_HERE, "libharpb200.so")  # HARP_LIB: tuning builds (tools/build_variants.sh)

c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)


class HarpTopology(C.Structure):
    _fields_ = [("V0", C.c_int32), ("F0", C.c_int32), ("faces0", c_int32_p), ("subdivide", C.c_int32),
                ("Vt", C.c_int32), ("faces_uvs", c_int32_p), ("verts_uvs", c_float_p)]


class HarpManoModel(C.Structure):
    _fields_ = [("shapedirs", c_float_p), ("posedirs", c_float_p), ("v_template", c_float_p),
                ("J_regressor", c_float_p), ("weights", c_float_p), ("hands_mean", c_float_p), ("tips", c_int32_p)]


class HarpLbsModel(C.Structure):
    _fields_ = [("J", C.c_int32), ("n_shape", C.c_int32), ("parents", c_int32_p), ("shapedirs", c_float_p),
                ("posedirs", c_float_p), ("v_template", c_float_p), ("J_template", c_float_p), ("J_shapedirs", c_float_p),
                ("weights", c_float_p), ("pose_mean", c_float_p), ("rodrigues_kind", C.c_int32), ("center_joint", C.c_int32),
                ("n_joints_out", C.c_int32), ("joint_out_src", c_int32_p), ("n_tips", C.c_int32), ("tips", c_int32_p)]


class HarpShadeGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("d_verts", "d_normals", "d_cam_R", "d_cam_T", "d_light_R", "d_light_T",
                                          "d_light_pos", "d_amb", "d_texture", "d_normal_map")]


class HarpRenderArgs(C.Structure):
    _fields_ = ([(n, C.c_void_p) for n in ("verts", "normals", "cam_R", "cam_T", "light_R", "light_T", "light_pos", "amb",
                                            "texture", "normal_map")] +
                [("texH", C.c_int32), ("texW", C.c_int32), ("self_shadow", C.c_int32), ("faces_per_pixel", C.c_int32),
                 ("sigma", C.c_float)] +
                [(n, C.c_void_p) for n in ("y_sil", "y_true", "y_mask")] + [("w_sil", C.c_float), ("w_photo", C.c_float)] +
                [(n, C.c_void_p) for n in ("alpha", "rgb", "loss_sil", "loss_photo")] + [("grads", HarpShadeGrads)])


class HarpStepArgs(C.Structure):
    """harp_step_args of include/harp_b200.h (field order matters)."""
    _fields_ = ([(n, C.c_int32) for n in ("B", "S", "stage", "self_shadow", "faces_per_pixel", "texH", "texW")] +
                [(n, C.c_float) for n in ("focal", "sigma", "frame_scale")] + [("weights", C.c_float * 9)] +
                [(n, C.c_void_p) for n in ("fid", "pose", "rot", "trans", "cam", "wrist_pose", "shape", "verts_disps", "texture", "normal_map",
                                           "light_position", "amb_ratio", "y_true", "y_sil", "y_sil_eroded", "init_joints", "ref_verts",
                                           "uv_mask", "off_albedo", "off_normal", "losses", "alpha", "rgb", "g_pose", "g_cam", "g_rot",
                                           "g_wrist_pose", "g_shape", "g_verts_disps", "g_texture", "g_normal_map", "g_light_position",
                                           "g_amb_ratio", "d_rot", "d_trans", "d_wrist", "workspace")])


class HarpError(RuntimeError):
    pass


_lib = None

# every symbol include/harp_b200.h declares (tests/test_abi.py checks the library exports exactly these)
SYMBOLS = ["harp_last_error", "harp_version", "harp_launch_count", "harp_ctx_profile", "harp_ctx_profile_read", "harp_ctx_profile_timeline", "harp_ctx_create", "harp_ctx_create_lbs", "harp_ctx_destroy", "harp_lbs_fwd", "harp_lbs_bwd", "harp_ctx_reserve", "harp_ctx_pin", "harp_ctx_dims",
           "harp_ctx_faces", "harp_ctx_read_counters", "harp_ctx_deferred_pixels", "harp_lbs_mano_fwd", "harp_lbs_mano_bwd", "harp_mesh_build_fwd",
           "harp_mesh_build_bwd", "harp_vertex_normals_fwd", "harp_vertex_normals_bwd", "harp_mesh_losses", "harp_silhouette", "harp_shade", "harp_hit_mask", "harp_render_normals", "harp_render_fused", "harp_step_fused", "harp_step_workspace_floats", "harp_step_workspace_views", "harp_frame_glue_fwd", "harp_frame_glue_bwd", "harp_texture_reg", "harp_mse", "harp_image_metrics", "harp_ingest_u8", "harp_adam", "harp_adam_dev",
           "harp_peer_create", "harp_peer_connect", "harp_peer_allreduce", "harp_peer_destroy"]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HarpError(f"{LIB_PATH} not built: run `python -m harp_b200.build` (needs nvcc). "
                            "harp_b200 has no CPU or PyTorch fallback.")
        L = C.CDLL(LIB_PATH)
        L.harp_last_error.restype = C.c_char_p
        L.harp_ctx_destroy.restype = None
        L.harp_launch_count.restype = C.c_int64
        L.harp_launch_count.argtypes = []
        L.harp_ctx_profile.argtypes = [C.c_void_p, C.c_int]
        L.harp_ctx_profile_read.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        L.harp_ctx_profile_timeline.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int)]
        vp, i, f = C.c_void_p, C.c_int, C.c_float
        L.harp_ctx_create.argtypes = [C.POINTER(vp), C.POINTER(HarpTopology), C.POINTER(HarpManoModel)]
        L.harp_ctx_destroy.argtypes = [vp]
        L.harp_ctx_create_lbs.argtypes = [C.POINTER(vp), C.POINTER(HarpTopology), C.POINTER(HarpLbsModel)]
        L.harp_lbs_fwd.argtypes = [vp, i, vp, vp, i, vp, vp, vp, vp]
        L.harp_lbs_bwd.argtypes = [vp, i, vp, vp, i, vp, vp, vp, vp, vp, vp, vp]
        L.harp_ctx_reserve.argtypes = [vp, i, i]
        L.harp_ctx_pin.argtypes = [vp, i]
        L.harp_ctx_dims.argtypes = [vp, c_int32_p, c_int32_p, c_int32_p, c_int32_p]
        L.harp_ctx_faces.argtypes = [vp, c_int32_p]
        L.harp_ctx_read_counters.argtypes = [vp, C.POINTER(C.c_int64), vp]
        L.harp_ctx_deferred_pixels.argtypes = [vp, C.POINTER(C.c_int64), vp]
        L.harp_lbs_mano_fwd.argtypes = [vp, i, vp, vp, i, vp, vp, vp, vp]
        L.harp_lbs_mano_bwd.argtypes = [vp, i, vp, vp, i, vp, vp, vp, vp, vp, vp, vp]
        L.harp_mesh_build_fwd.argtypes = [vp, i, vp, vp, vp, vp, vp]
        L.harp_mesh_build_bwd.argtypes = [vp, i, vp, vp, vp, vp, vp, vp, vp]
        L.harp_vertex_normals_fwd.argtypes = [vp, i, vp, vp, vp]
        L.harp_vertex_normals_bwd.argtypes = [vp, i, vp, vp, vp, vp]
        L.harp_mesh_losses.argtypes = [vp, i, vp, vp, vp, vp, vp, c_float_p, vp, vp, vp, vp, vp]
        L.harp_silhouette.argtypes = [vp, i, i, i, f, f, i, vp, vp, vp, vp, vp, f, vp, vp, vp, vp, vp]
        L.harp_shade.argtypes = [vp, i, i, i, f, i, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i, i, c_float_p, vp, vp, vp, f, vp,
                                 C.POINTER(HarpShadeGrads), vp]
        L.harp_hit_mask.argtypes = [vp, i, i, vp, vp]
        L.harp_render_normals.argtypes = [vp, i, i, f, i, vp, vp, vp, vp, vp, i, i, vp, vp]
        L.harp_render_fused.argtypes = [vp, i, i, f, C.POINTER(HarpRenderArgs), vp]
        L.harp_step_fused.argtypes = [vp, C.POINTER(HarpStepArgs), vp]
        L.harp_step_workspace_floats.argtypes = [vp, i]
        L.harp_step_workspace_floats.restype = C.c_size_t
        L.harp_step_workspace_views.argtypes = [vp, i, vp] + [C.POINTER(vp)] * 4
        L.harp_frame_glue_fwd.argtypes = [vp, i, i, f] + [vp] * 12
        L.harp_frame_glue_bwd.argtypes = [vp, i, i, f] + [vp] * 15
        L.harp_texture_reg.argtypes = [vp, i, vp, i, i, vp, vp, f, vp, vp, vp]
        L.harp_mse.argtypes = [C.c_int64, vp, vp, f, vp, vp, vp]
        L.harp_image_metrics.argtypes = [i, i, vp, vp, vp, vp, vp, vp, vp, vp]
        L.harp_ingest_u8.argtypes = [i, i, vp, vp, vp, vp, vp, vp]
        L.harp_adam.argtypes = [vp, vp, vp, vp, C.c_int64, f, f, f, f, i, vp]
        L.harp_adam_dev.argtypes = [vp, vp, vp, vp, C.c_int64, f, f, f, f, vp, vp]
        L.harp_peer_create.argtypes = [C.POINTER(vp), i, i, C.c_size_t, C.POINTER(vp), vp]
        L.harp_peer_connect.argtypes = [vp, vp]
        L.harp_peer_allreduce.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, f, vp]
        L.harp_peer_destroy.argtypes = [vp]
        L.harp_peer_destroy.restype = None
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise HarpError(f"libharpb200 error {rc}: {lib().harp_last_error().decode()}")


def ptr(t, dtype=torch.float32):
    """Device pointer of a contiguous CUDA tensor of the dtype the C ABI reads at that argument (None -> NULL).
    The kernels trust the layout they are handed (include/harp_b200.h: fp32 everywhere except the int64 frame ids, the int32
    texel offsets / face tables and the uint8 ingest buffers), so a mismatch is an error here, not a silent misread there."""
    if t is None:
        return None
    if not t.is_cuda:
        raise HarpError("harp_b200 ops need CUDA tensors (there is no CPU path)")
    if not t.is_contiguous():
        raise HarpError("harp_b200 ops need contiguous tensors")
    if t.dtype != dtype:
        raise HarpError(f"harp_b200: expected a {dtype} tensor at this argument, got {t.dtype}")
    return t.data_ptr()


def iptr(t):
    """int32 tensor (texel offsets)."""
    return ptr(t, torch.int32)


def lptr(t):
    """int64 tensor (frame ids)."""
    return ptr(t, torch.int64)


def bptr(t):
    """uint8 tensor (frames / masks as decoded from the image files)."""
    return ptr(t, torch.uint8)


def stream():
    return torch.cuda.current_stream().cuda_stream
