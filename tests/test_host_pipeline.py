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
