"""CPU: the step kernel's launch plan and row ownership (gmrm_debug_step_plan, no device needed).

The dot product of Bayes::dot_product (src/bayes.cpp:709-770) is taken by CTAs that each own rows of 256
individuals in every pass; a row owned twice or not at all would be a silently wrong sum, so the ownership
map is checked over many (N, CTA count) shapes: every row exactly once, at most rows_per_pass per CTA and
pass, balanced to within one row over the whole step, and the plan inside the 227 KB shared-memory limit."""
import numpy as np
import pytest

from gmrm_b200 import api

MAX_DYN_SMEM = 232448


def plan(N, nsm, V, T):
    p = api.step_plan(N, nsm, V, T)
    if p is None:
        return None
    return dict(tc=p["traits_per_launch"], rpp=p["rows_per_pass"], npass=p["npass"], smem=p["smem_bytes"],
                nrows=p["nrows"], ranges=p["ranges"])


def check_ownership(p, nsm):
    owner = np.zeros(p["nrows"], dtype=np.int32)
    per_cta = np.zeros(nsm, dtype=np.int64)
    for q in range(p["npass"]):
        for c in range(nsm):
            s, n = p["ranges"][q, c]
            assert 0 <= n <= p["rpp"], (q, c, s, n)
            if n:
                assert 0 <= s and s + n <= p["nrows"]
                owner[s:s + n] += 1
                per_cta[c] += n
    assert owner.min() == 1 and owner.max() == 1, "every row exactly once"
    assert per_cta.max() - per_cta.min() <= 1, "rows balanced over the CTAs to within one"


def test_ukb_shape_plan():
    # BASELINE configs[1]: N = 458,747 on 148 SMs, 2048 virtual ranks per GPU, one trait
    p = plan(458747, 148, 2048, 1)
    assert p["nrows"] == 1792 and p["tc"] == 1
    assert p["rpp"] == 5 and p["npass"] == 3          # 5 table slots of 41,472 B, 12.1 rows per CTA
    assert p["smem"] <= MAX_DYN_SMEM
    check_ownership(p, 148)


@pytest.mark.parametrize("T", [1, 2, 3, 4, 7])
def test_traits_split_over_launches(T):
    p = plan(458747, 148, 2048, T)
    assert 1 <= p["tc"] <= min(T, 4) and p["tc"] * p["rpp"] <= 5
    assert p["smem"] <= MAX_DYN_SMEM
    check_ownership(p, 148)


def test_ownership_over_random_shapes():
    rng = np.random.default_rng(7)
    shapes = [(1, 1), (1, 148), (255, 148), (256, 3), (257, 2), (1024, 148), (37889, 148), (458747, 132), (2_000_000, 148)]
    shapes += [(int(rng.integers(1, 1_500_000)), int(rng.choice([1, 2, 7, 64, 132, 148, 160]))) for _ in range(300)]
    seen = 0
    for N, nsm in shapes:
        for V in (0, 1, 16, 2048):
            p = plan(N, nsm, V, 1)
            if p is None:                    # does not fit one launch (too many rows per CTA): the engine refuses it too
                continue
            assert p["smem"] <= MAX_DYN_SMEM and p["nrows"] == -(-(-(-N // 4)) // 64)
            check_ownership(p, nsm)
            seen += 1
    assert seen > 600


def test_bad_arguments_are_refused():
    assert plan(0, 148, 16, 1) is None
    assert plan(1000, 0, 16, 1) is None
    assert plan(1000, 148, -1, 1) is None
    assert plan(1000, 148, 16, 0) is None


def test_lane_byte_map_of_a_pass_chunk_is_a_bijection():
    """chunk_offset (kernels.cu): (table slot, lane, byte) <-> byte of the CTA's rows * 64-byte chunk of a column.  Every byte
    exactly once; a lane's words of a 4-slot group are 16 contiguous, 16-byte aligned bytes (one LDG.128), of a 2-slot group
    8 (LDG.64); the table build and the stream use the same map."""
    import ctypes as C
    from gmrm_b200 import api
    f = api.lib().gmrm_debug_chunk_offset
    f.restype = C.c_int
    for rows in range(1, 6):
        seen = {}
        for s_ in range(rows):
            for lane in range(16):
                for k in range(4):
                    o = f(rows, s_, lane, k)
                    assert 0 <= o < rows * 64 and o not in seen
                    seen[o] = (s_, lane, k)
        assert len(seen) == rows * 64
        for lane in range(16):
            if rows >= 4:
                base = f(rows, 0, lane, 0)
                assert base % 16 == 0 and [f(rows, s_, lane, k) for s_ in range(4) for k in range(4)] == list(range(base, base + 16))
            elif rows >= 2:
                base = f(rows, 0, lane, 0)
                assert base % 8 == 0 and [f(rows, s_, lane, k) for s_ in range(2) for k in range(4)] == list(range(base, base + 8))
            last = f(rows, rows - 1, lane, 0)
            assert last % 4 == 0 and [f(rows, rows - 1, lane, k) for k in range(4)] == list(range(last, last + 4))
    assert f(6, 0, 0, 0) < 0 and f(3, 3, 0, 0) < 0 and f(3, 0, 16, 0) < 0
