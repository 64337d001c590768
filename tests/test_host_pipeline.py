"""Host-side logic of the chunked host-input pipeline (no GPU needed)."""
import pytest

from convnet_quantization_b200.models._gpu_modules import B200StaticQuantizedNet


class _Cfg:
    HOST_CHUNK = 2048
    TAIL_MIN = 256


@pytest.mark.parametrize("b", [1, 2, 5, 255, 256, 257, 600, 2047, 2048, 2049, 4096, 5000, 16384, 65536 + 3])
def test_chunks_cover_the_batch_in_order(b):
    chunks = list(B200StaticQuantizedNet._chunks(_Cfg, b))
    assert sum(n for _, n in chunks) == b
    assert all(0 < n <= _Cfg.HOST_CHUNK for _, n in chunks)
    lo = 0
    for off, n in chunks:
        assert off == lo
        lo += n


def test_tail_is_split_so_that_little_compute_is_left_uncovered():
    sizes = [n for _, n in B200StaticQuantizedNet._chunks(_Cfg, 16384)]
    assert sizes[:7] == [2048] * 7 and sizes[7:] == [1536, 512]


@pytest.mark.parametrize("b", [1, 2, 1000, 1024, 1025, 3071, 3072, 3073, 16384, 65536 + 3])
def test_ramped_chunks_cover_the_batch_in_order(b):
    chunks = list(B200StaticQuantizedNet._chunks_ramp(_Cfg, b, 1024, 8192))
    assert sum(n for _, n in chunks) == b
    assert all(0 < n <= 8192 for _, n in chunks)
    lo = 0
    for off, n in chunks:
        assert off == lo
        lo += n
    # every chunk's copy must fit inside the kernels of the chunk before it: at most twice its size
    assert all(chunks[i + 1][1] <= 2 * chunks[i][1] for i in range(len(chunks) - 1))


def test_ramp_starts_small_for_a_fast_fill():
    assert [n for _, n in B200StaticQuantizedNet._chunks_ramp(_Cfg, 16384, 1024, 8192)] == [1024, 2048, 4096, 8192, 1024]
