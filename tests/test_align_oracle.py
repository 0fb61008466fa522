"""The alignment oracle (oracle/sw_align.c): its END CELL against the reference's own visiting order (recorded
output of oracle/_ref/sw_antidiag_ends = the reference source + position bookkeeping beside its running maximum,
tests/golden/*.ref_ends.out), its score against oracle_sw_score, and its START / CIGAR through an independent
re-scoring of the path."""
import numpy as np
import pytest

from conftest import GOLDEN

ENDS_FILES = ["sw_ends_two_letter", "sw_ends_repeats", "sw_ends_ragged", "sw_ends_150", "sw_ends_no_trailing_nl",
              "sw_ends_mid", "sw_ends_alphabet"]


def pairs_of(agx, name, line_buf=20000):
    inp = agx.formats.parse_sw((GOLDEN / f"{name}.in").read_bytes(), line_buf=line_buf)
    data = inp.buf.tobytes()
    return [(data[inp.off[2 * p]:inp.off[2 * p] + inp.len[2 * p]],
             data[inp.off[2 * p + 1]:inp.off[2 * p + 1] + inp.len[2 * p + 1]]) for p in range(inp.n_pairs)]


@pytest.mark.parametrize("name", ENDS_FILES)
def test_end_cell_is_the_cell_of_the_references_running_maximum(agx, oracle_mod, name):
    rows = oracle_mod.parse_ref_sw_ends((GOLDEN / f"{name}.ref_ends.out").read_text())
    pairs = pairs_of(agx, name)
    assert len(rows) == len(pairs) > 0
    ties = 0
    for (a, b), row in zip(pairs, rows):
        s, c, cig = oracle_mod.sw_align(a, b)
        assert s == row[0] == oracle_mod.sw_score(a, b)
        assert (c[1], c[3]) == oracle_mod.ref_ends_to_coords(row)
        assert row[3] == (2 if len(a) > len(b) else 1)
        if s > 0:
            assert oracle_mod.sw_cigar_score(a, b, c, cig) == s
    del ties


def test_end_cell_live_reference(agx, oracle_mod, tmp_path):
    if not oracle_mod.ref_available("sw_antidiag_ends"):
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(77)
    data = agx.synth.sw_random_file(rng, 150, 1, 70, alphabet=b"AC")
    p = tmp_path / "e.in"
    p.write_bytes(data)
    rows = oracle_mod.run_ref_sw_ends(str(p))
    inp = agx.formats.parse_sw(data)
    raw = inp.buf.tobytes()
    assert len(rows) == inp.n_pairs
    for q, row in enumerate(rows):
        a = raw[inp.off[2 * q]:inp.off[2 * q] + inp.len[2 * q]]
        b = raw[inp.off[2 * q + 1]:inp.off[2 * q + 1] + inp.len[2 * q + 1]]
        s, c, _ = oracle_mod.sw_align(a, b)
        assert s == row[0] and (c[1], c[3]) == oracle_mod.ref_ends_to_coords(row)


def test_traceback_properties(oracle_mod):
    """CIGAR spells a path from start to end whose score is the Smith-Waterman score; M/I/D lengths add up to
    the spans; other scoring parameters; zero-score pairs have no alignment."""
    rng = np.random.default_rng(5)
    for t in range(400):
        alpha = np.frombuffer(b"ACGT" if t % 4 else b"AC", np.uint8)
        la, lb = int(rng.integers(1, 120)), int(rng.integers(1, 120))
        a = alpha[rng.integers(0, alpha.size, la)]
        if t % 2:
            b = a.copy()
            m = rng.random(la) < 0.1
            b[m] = alpha[rng.integers(0, alpha.size, int(m.sum()))]
            keep = rng.random(la) >= 0.05
            b = np.concatenate([alpha[rng.integers(0, alpha.size, int(rng.integers(0, 9)))], b[keep]])
        else:
            b = alpha[rng.integers(0, alpha.size, lb)]
        sc = [(1, -1, -3, -1), (3, -2, -5, -2), (2, -3, 0, -2), (5, -4, -10, -1)][t % 4]
        a, b = a.tobytes(), b.tobytes()
        s, c, cig = oracle_mod.sw_align(a, b, sc)
        assert s == oracle_mod.sw_score(a, b, sc)
        assert oracle_mod.sw_ends(a, b, sc) == (s, (c[1], c[3]))         # the rolling-row form used for long pairs
        if s == 0:
            assert c == (-1, -1, -1, -1) and cig == []
            continue
        assert oracle_mod.sw_cigar_score(a, b, c, cig, sc) == s
        m_len = sum(w >> 4 for w in cig if w & 15 == 0)
        i_len = sum(w >> 4 for w in cig if w & 15 == 1)
        d_len = sum(w >> 4 for w in cig if w & 15 == 2)
        assert m_len + i_len == c[1] - c[0] + 1 and m_len + d_len == c[3] - c[2] + 1
        assert cig[0] & 15 == 0 and cig[-1] & 15 == 0          # a local alignment starts and ends on a match column
        assert all((x & 15) != (y & 15) for x, y in zip(cig, cig[1:]))
    assert oracle_mod.sw_align(b"AAAA", b"TTTT") == (0, (-1, -1, -1, -1), [])
    assert oracle_mod.sw_align(b"ACGT\n", b"ACGT\n") == (5, (0, 4, 0, 4), [5 << 4])
    s, c, cig = oracle_mod.sw_align(b"ACGTACGTACGT", b"ACGTACGGGGTACGT", (2, -3, -3, -1))
    assert oracle_mod.cigar_string(cig) == "6M3D6M" and s == 2 * 12 - 3 - 3
