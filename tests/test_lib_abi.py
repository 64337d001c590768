"""The C-ABI library must build for sm_100a, load without a GPU, and export every symbol include/b200q.h declares."""
import os
import re
import subprocess

from convnet_quantization_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(dev: bool):
    """Functions include/b200q.h declares: outside ``#ifdef B200Q_DEV`` blocks (product) or all of them (dev)."""
    text = open(os.path.join(ROOT, "include", "b200q.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    if not dev:
        text = re.sub(r"#ifdef B200Q_DEV.*?#endif", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200q_[a-z0-9_]+)\s*\(", text)))


def _exported(path):
    out = subprocess.run(["nm", "-D", "--defined-only", str(path)], capture_output=True, text=True).stdout
    return sorted(set(re.findall(r" T (b200q_[a-z0-9_]+)$", out, flags=re.M)))


def test_exports_match_header(lib):
    names = _declared(dev=False)
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200q.h but not exported"
    assert sorted(_lib.EXPORTS) == names
    assert _exported(_lib.LIB_PATH) == names, "the product library must export exactly what the header declares"
    assert lib.b200q_abi_version() == 4
    assert lib.b200q_launch_count() >= 0


def test_dev_library_is_a_superset_and_product_has_no_dev_code():
    """Experiments and role-disabling switches live in libb200q_dev.so only (-DB200Q_DEV)."""
    dev = _lib.load_dev(build_if_missing=True)
    names = _declared(dev=True)
    assert sorted(_lib.DEV_EXPORTS) == names and _exported(_lib.DEV_LIB_PATH) == names
    assert set(_lib.DEV_EXPORTS) - set(_lib.EXPORTS) == {"b200q_conv12_fused", "b200q_conv3x3_simt"}
    assert dev.b200q_abi_version() == 4
    product = open(_lib.LIB_PATH, "rb").read()
    for switch in (b"B200Q_HALO_DEBUG", b"B200Q_PAIR_DEBUG", b"B200Q_NO_HALO", b"B200Q_TC_STREAM_WEIGHTS", b"B200Q_FUSE12",
                   b"B200Q_HALO_EW", b"B200Q_NO_CONV1_TC", b"B200Q_NO_CTA2", b"B200Q_H2_SLOTS", b"B200Q_NO_SMALL", b"B200Q_TINY_MAX_B", b"B200Q_CONV1_TINY_MAX_B",
                   b"B200Q_EAGER_PDL", b"conv12_fused"):
        assert switch not in product, f"{switch.decode()} must not be compiled into the product library"
    assert b"B200Q_HALO_DEBUG" in open(_lib.DEV_LIB_PATH, "rb").read()


def test_header_constants_match_binding():
    text = open(os.path.join(ROOT, "include", "b200q.h")).read()
    assert int(re.search(r"#define B200Q_REDUCE_SCRATCH_BYTES (\d+)", text).group(1)) == _lib.REDUCE_SCRATCH_BYTES
    assert int(re.search(r"#define B200Q_REDUCE_QPARAMS_OFFSET (\d+)", text).group(1)) == _lib.REDUCE_QPARAMS_OFFSET
    assert _lib.REDUCE_QPARAMS_OFFSET + 5 * 4 <= _lib.REDUCE_SCRATCH_BYTES  # the qparams block must fit
    assert _lib.REDUCE_QPARAMS_OFFSET >= 2 * 1024 * 4 + 4


def test_workspace_size_is_pure_host_math(lib):
    assert lib.b200q_static_workspace_bytes(0) == 2048  # alignment slack + ticket word
    assert lib.b200q_static_workspace_bytes(64) >= 2 * 64 * 65536
    assert lib.b200q_static_workspace_bytes(-1) < 0


def test_sass_is_blackwell_native():
    """tcgen05 / TMA evidence in the shipped binary (B200_PROFILING.md: UTC*MMA, LDTM, UTMALDG)."""
    out = subprocess.run(["cuobjdump", "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "UTCIMMA" in out and "LDTM" in out and "UTMALDG" in out
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
