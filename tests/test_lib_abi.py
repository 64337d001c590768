"""The C-ABI library must build for sm_100a, load without a GPU, and export every symbol include/b200q.h declares."""
import os
import re
import subprocess

from convnet_quantization_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "b200q.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200q_[a-z0-9_]+)\s*\(", text)))


def test_exports_match_header(lib):
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200q.h but not exported"
    assert sorted(_lib.EXPORTS) == names
    assert lib.b200q_abi_version() >= 2
    assert lib.b200q_launch_count() >= 0


def test_workspace_size_is_pure_host_math(lib):
    assert lib.b200q_static_workspace_bytes(0) == 1024
    assert lib.b200q_static_workspace_bytes(64) >= 2 * 64 * 65536
    assert lib.b200q_static_workspace_bytes(-1) < 0


def test_sass_is_blackwell_native():
    """tcgen05 / TMA evidence in the shipped binary (B200_PROFILING.md: UTC*MMA, LDTM, UTMALDG)."""
    out = subprocess.run(["cuobjdump", "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "UTCIMMA" in out and "LDTM" in out and "UTMALDG" in out
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
