"""The CPU oracle (oracle/kv_oracle.cpp) must reproduce the UNMODIFIED reference on every
committed fixture: return values, available_size, page counters, avail/full page sets in dict
order, reserved blocks, shrink flag and the mapped-page set, after every single op."""
import pytest

import oracle
from golden_util import diff_record, golden_names, load_golden


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_fixture(name):
    fix = load_golden(name)
    out = oracle.replay(fix["trace"])
    assert out["null_block"] == fix["null_block"]
    assert len(out["records"]) == len(fix["records"])
    for i, (got, want) in enumerate(zip(out["records"], fix["records"])):
        bad = diff_record(got, want)
        assert not bad, f"{name}: op {i} {fix['trace']['ops'][i][:2]} differs in {bad}: got { {k: got[k] for k in bad} } want { {k: want[k] for k in bad} }"


def test_group_indices_by_page_known_answers():
    # SURVEY Appendix A.3 probes taken from the reference build
    got = oracle.group_indices_by_page([5, 70, 6, 200, 130, 71, 700, 1, 579], 32768)
    assert got == [(9, [579]), (1, [70, 71]), (0, [5, 6, 1]), (10, [700]), (3, [200]), (2, [130])]
    # 20 pages, one id each: reserve(20//64+1 = 1) -> 2 buckets, then libstdc++'s own rehashes
    order = [p for p, _ in oracle.group_indices_by_page([p * 64 for p in range(20)], 32768)]
    assert order == [19, 18, 17, 16, 15, 14, 13, 12, 11, 4, 3, 2, 0, 1, 5, 6, 7, 8, 9, 10]
    # 20 full pages: reserve(21) -> 23 buckets, no rehash, every key lands at the list head
    ids = [p * 64 + j for p in range(20) for j in range(64)]
    assert [p for p, _ in oracle.group_indices_by_page(ids, 32768)] == list(range(19, -1, -1))


def test_internal_page_block_ranges():
    import ctypes as C
    lib = oracle.lib()
    s, e = C.c_int64(), C.c_int64()
    for pid, bms, want in [(0, 32768, (0, 64)), (3, 32768, (192, 256)), (0, 18432, (0, 113)), (1, 18432, (114, 227))]:
        lib.kvo_block_range(pid, 2 << 20, bms, C.byref(s), C.byref(e))
        assert (s.value, e.value) == want
    assert lib.kvo_num_blocks(2 << 20, 18432) == 113
