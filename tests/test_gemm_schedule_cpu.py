"""Stream-K tail schedule of the GEMM (csrc/gemm_tcgen05.cu: sk_plan / sk_share, the code the kernel itself runs, reached through
the host-only diagnostic ub_gemm_sk_schedule): every k-block of every tile is computed exactly once, each split tile has exactly
one finishing piece that waits for exactly the parked pieces, and slots never collide.  No GPU needed."""
import ctypes as C

import pytest

from unite_b200 import _cabi


def share(T, U, KB, u, overhead=4):
    out = (C.c_int32 * 11)()
    split = _cabi.lib.ub_gemm_sk_schedule(T, U, KB, overhead, u, out)
    keys = ("first", "tiles", "units", "n_dp", "part_tile", "part_kb0", "part_kb1", "part_slot", "fin_tile", "fin_kb0", "fin_wait")
    return split, dict(zip(keys, out))


def check_schedule(T, U, KB, overhead=4):
    split, s0 = share(T, U, KB, 0, overhead)
    if not split:
        assert s0["tiles"] == 0
        return False
    first, tiles, units = s0["first"], s0["tiles"], s0["units"]
    assert first + tiles == T and first % U == 0 and 0 < tiles < U and tiles <= units <= U
    cover = {t: [] for t in range(first, T)}          # split tile -> [(kb0, kb1, kind, aux, pair)]
    whole = [0] * first
    load = []
    for u in range(U):
        _, s = share(T, U, KB, u, overhead)
        assert (s["first"], s["tiles"], s["units"]) == (first, tiles, units)
        for i in range(s["n_dp"]):
            whole[u + i * U] += 1
        work = s["n_dp"] * KB
        if s["part_tile"] >= 0:
            assert 0 <= s["part_kb0"] < s["part_kb1"] < KB
            cover[s["part_tile"]].append((s["part_kb0"], s["part_kb1"], "part", s["part_slot"], u))
            work += s["part_kb1"] - s["part_kb0"]
        if s["fin_tile"] >= 0:
            assert 0 <= s["fin_kb0"] < KB
            cover[s["fin_tile"]].append((s["fin_kb0"], KB, "fin", s["fin_wait"], u))
            work += KB - s["fin_kb0"]
        if u >= units:
            assert s["part_tile"] < 0 and s["fin_tile"] < 0
        load.append(work)
    assert all(c == 1 for c in whole), "a whole tile is computed zero or several times"
    for t, pieces in cover.items():
        pieces.sort()
        assert pieces[0][0] == 0 and pieces[-1][1] == KB
        assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:])), f"tile {t}: k-blocks not tiled exactly once: {pieces}"
        assert [p[2] for p in pieces] == ["part"] * (len(pieces) - 1) + ["fin"]
        assert len(pieces) <= 3
        assert [p[3] for p in pieces[:-1]] == list(range(len(pieces) - 1)), f"tile {t}: slots {pieces}"
        assert pieces[-1][3] == len(pieces) - 1, f"tile {t}: the finishing piece waits for {pieces[-1][3]} of {len(pieces) - 1}"
        assert [p[4] for p in pieces] == sorted(p[4] for p in pieces) and len({p[4] for p in pieces}) == len(pieces)
    # the balance the cost model promised
    assert max(load) <= (T // U) * KB + -(-tiles * KB // units)
    assert max(load) < -(-T // U) * KB
    return True


def test_stream_k_schedule_is_an_exact_cover_on_the_step_shapes():
    # (tiles, k-blocks) of the student's M = 10 240 products on 74 CTA pairs of a B200
    assert check_schedule(120, 74, 48)          # fc2 fwd / fc1 dgrad: 10240 x 768 x 3072
    assert check_schedule(120, 74, 36)          # qkv dgrad
    assert not check_schedule(120, 74, 12)      # proj: the fix-up charge eats the gain
    assert check_schedule(480, 74, 12)          # fc1 fwd / fc2 dgrad
    # the teacher's M = 50 432 products fill their last wave: left whole
    for tiles, kb in ((591, 12), (591, 48), (1773, 12), (2364, 12)):
        assert not check_schedule(tiles, 74, kb)


@pytest.mark.parametrize("U", [1, 2, 3, 7, 66, 74])
def test_stream_k_schedule_exhaustive_small(U):
    n_split = 0
    for T in list(range(1, 3 * U + 2)) + [5 * U + U // 2, 11 * U + U - 1]:
        for KB in (1, 3, 4, 5, 8, 12, 13, 36, 48, 64):
            for overhead in (0, 4):
                n_split += bool(check_schedule(T, U, KB, overhead))
    assert U == 1 or n_split > 0


def test_stream_k_workspace_size_covers_a_full_tail():
    nbytes = _cabi.lib.ub_gemm_sk_workspace_bytes()
    assert nbytes >= 8192 + 73 * 2 * 2 * 8 * 4096 * 4
