"""The composite prefix-cache key (kvcached/integration/vllm/patches.py:260-274, pinned by the reference's
tests/test_make_cache_key.py): hash || 4-byte big-endian group id; a str hash gives the key its bytes would."""
import pytest

from kvcached_b200.block_pool import _make_cache_key, get_max_cached_blocks


def _gid(g):
    return g.to_bytes(4, "big", signed=False)


def test_bytes_and_str_hashes_give_one_key():
    assert _make_cache_key(b"deadbeef", 0) == b"deadbeef" + _gid(0)
    assert _make_cache_key("deadbeef", 3) == b"deadbeef" + _gid(3) == _make_cache_key(b"deadbeef", 3)
    _make_cache_key("0a1b2c3d" * 8, 0)          # a 64-char hex digest must not raise


@pytest.mark.parametrize("group_id", [0, 1, 255, 256, 65535, 2 ** 31 - 1])
def test_group_id_is_a_4_byte_big_endian_suffix(group_id):
    key = _make_cache_key(b"h", group_id)
    assert key == b"h" + _gid(group_id) and len(key) == 5
    assert _make_cache_key(b"samehash", 0) != _make_cache_key(b"samehash", 1)


def test_max_cached_blocks_from_tokens():
    """patches.py:247-257: -1 unlimited, 0 evict on free, else MAX_CACHED_TOKENS // block_size."""
    assert get_max_cached_blocks(16, -1) == -1
    assert get_max_cached_blocks(16, 0) == 0
    assert get_max_cached_blocks(16, 16000) == 1000
    assert get_max_cached_blocks(64, 16000) == 250
