"""CPU: the C-ABI shared library loads without a GPU and exports exactly the symbols include/harp_b200.h declares;
the product has no CPU / oracle fallback."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def libpath():
    from harp_b200 import build
    return build.build()


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "harp_b200.h")).read()
    return sorted(set(re.findall(r"HARP_API[^;(]*?\b(harp_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(libpath):
    L = ctypes.CDLL(libpath)
    declared = _header_symbols()
    assert len(declared) >= 19
    for s in declared:
        assert hasattr(L, s), s
    out = subprocess.run(["nm", "-D", "--defined-only", libpath], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (harp_\w+)", out)))
    assert exported == declared, (set(exported) ^ set(declared))
    from harp_b200._lib import SYMBOLS
    assert sorted(SYMBOLS) == declared


def test_version_and_error_string_without_gpu(libpath):
    L = ctypes.CDLL(libpath)
    L.harp_last_error.restype = ctypes.c_char_p
    assert L.harp_version() == 100
    assert L.harp_ctx_dims(None, None, None, None, None) == -1       # HARP_ERR_INVALID, no CUDA call made
    assert b"null ctx" in L.harp_last_error()


def test_library_is_sm100a_only(libpath):
    out = subprocess.run(["cuobjdump", "--list-elf", libpath], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\w+)\.", out))
    assert archs == {"100a"}, archs


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "harp_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "harp_oracle" not in src and "p3d_shim" not in src and "import pytorch3d" not in src, f


def test_ops_fail_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from harp_b200 import _lib
    with pytest.raises(_lib.HarpError):
        _lib.ptr(torch.zeros(3))


def test_checkpoint_roundtrip_uses_reference_keys(tmp_path):
    """saved_params.pkl is a pickle of {key: ndarray} with the reference's key names (utils/file_utils.py:6-26)."""
    import pickle
    import torch
    from harp_b200 import file_utils
    params = {"pose": torch.randn(4, 45), "cam": torch.randn(4, 3), "shape": torch.randn(10), "amb_ratio": torch.tensor(0.4),
              "texture": torch.rand(1, 8, 8, 3), "verts_disps": torch.zeros(12, 1), "uv_mask": None}
    path = file_utils.save_result(params, str(tmp_path))
    raw = pickle.load(open(path, "rb"))
    assert set(raw) == set(params) and all(k in file_utils.PARAM_KEYS for k in raw)
    back = file_utils.load_result(str(tmp_path))
    for k, v in params.items():
        assert (back[k] is None) if v is None else torch.equal(back[k], v)


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """The ctypes mirrors in harp_b200/_lib.py (what INTEGRATION.md shows a maintainer) have the size and field offsets a C
    compiler gives the structs of include/harp_b200.h."""
    from harp_b200 import _lib
    probes = {"harp_step_args": ("B", "focal", "weights", "fid", "pose", "y_true", "losses", "g_pose", "d_rot", "workspace"),
              "harp_render_args": ("verts", "texH", "sigma", "y_sil", "w_sil", "alpha", "grads"),
              "harp_shade_grads": ("d_verts", "d_normal_map"), "harp_topology": ("V0", "faces0", "verts_uvs"),
              "harp_lbs_model": ("J", "parents", "rodrigues_kind", "tips")}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "harp_b200.h")}"', "int main(void) {"]
    for st, fields in probes.items():
        lines.append(f'  printf("{st} %zu", sizeof({st}));')
        for f in fields:
            lines.append(f'  printf(" %zu", offsetof({st}, {f}));')
        lines.append('  printf("\\n");')
    lines += ["  return 0;", "}"]
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    mirror = {"harp_step_args": _lib.HarpStepArgs, "harp_render_args": _lib.HarpRenderArgs, "harp_shade_grads": _lib.HarpShadeGrads,
              "harp_topology": _lib.HarpTopology, "harp_lbs_model": _lib.HarpLbsModel}
    for line in out.strip().splitlines():
        name, size, *offs = line.split()
        cls = mirror[name]
        assert ctypes.sizeof(cls) == int(size), (name, ctypes.sizeof(cls), size)
        for f, o in zip(probes[name], offs):
            assert getattr(cls, f).offset == int(o), (name, f, getattr(cls, f).offset, o)
