"""The CPU oracle (oracle/oracle.c) against the reference's golden vectors: the one committed golden
(pairHMM/test_set/test.out) and the recorded outputs of the unmodified reference programs
(tests/golden/*.out, written by tests/golden/make_golden.py)."""
import numpy as np
import pytest

from conftest import GOLDEN, read_golden, ref_scores

SW_FILES = ["sw_gen_header", "sw_ragged", "sw_short", "sw_150", "sw_no_trailing_nl", "sw_alphabet",
            "sw_two_letter", "sw_linebuf", "sw_dangling", "sw_header_small", "sw_mid", "sw_1kbp"]


@pytest.mark.parametrize("name", SW_FILES)
def test_sw_oracle_matches_reference_stdout(oracle_mod, name):
    scores, header = oracle_mod.sw_file(str(GOLDEN / f"{name}.in"), 1000)
    text = (GOLDEN / f"{name}.ref.out").read_text()
    assert f"line_num: {header}" in text
    assert scores.tolist() == ref_scores(f"{name}.ref.out")


@pytest.mark.parametrize("name", ["sw_1kbp", "sw_5kbp"])
def test_sw_oracle_long_lines(oracle_mod, name):
    # the MAX_LINE_LENGTH-raised reference build (arithmetic untouched) on un-split long lines
    scores, _ = oracle_mod.sw_file(str(GOLDEN / f"{name}.in"), 4200000)
    assert scores.tolist() == ref_scores(f"{name}.ref_long.out")


def test_sw_newline_symbol_quirk(oracle_mod):
    # the trailing newline is a symbol that matches the other line's newline
    assert oracle_mod.sw_score(b"ACGT\n", b"ACGT\n") == 5
    assert oracle_mod.sw_score(b"AAAA\n", b"TTTT\n") == 1
    assert oracle_mod.sw_score(b"ACGT\n", b"ACGT") == 4
    assert oracle_mod.sw_score(b"AAAA", b"TTTT") == 0


def test_sw_newline_identity(oracle_mod):
    """score(a+'\\n', b+'\\n') == max(best_plain, corner_plain + match): the identity the s16x2
    kernel relies on to keep the newline symbol out of its 2-bit alphabet."""
    rng = np.random.default_rng(3)
    for _ in range(300):
        la, lb = rng.integers(1, 60, size=2)
        a = bytes(rng.choice(list(b"ACGT"), size=la).astype(np.uint8))
        b = a if rng.random() < 0.3 else bytes(rng.choice(list(b"ACGT"), size=lb).astype(np.uint8))
        best, corner = oracle_mod.sw_score(a, b, want_corner=True)
        assert oracle_mod.sw_score(a + b"\n", b + b"\n") == max(best, corner + 1)
        assert oracle_mod.sw_score(a + b"\n", b) == best
        assert oracle_mod.sw_score(a, b + b"\n") == best


def test_pairhmm_committed_golden(oracle_mod, tmp_path):
    # the only golden the reference commits: pairHMM/test_set/test.out
    p = tmp_path / "test.in"
    p.write_bytes(read_golden("pairhmm_test.in"))
    vals, nb = oracle_mod.pairhmm_file(str(p))
    assert nb == 1 and len(vals) == 1
    committed = (GOLDEN / "pairhmm_test.committed.out").read_text().strip()
    assert "%f" % vals[0] == committed == "-4.485565"


@pytest.mark.parametrize("name", ["pairhmm_test", "pairhmm_10s", "pairhmm_synth_small",
                                  "pairhmm_synth_cfg4", "pairhmm_synth_long", "pairhmm_synth_tall"])
def test_pairhmm_oracle_matches_reference_output(oracle_mod, tmp_path, name):
    p = tmp_path / "in.txt"
    p.write_bytes(read_golden(f"{name}.in"))
    vals, _ = oracle_mod.pairhmm_file(str(p))
    for which in ("pairhmm_antidiag", "pairhmm_matrix"):
        ref = np.array([float(x) for x in (GOLDEN / f"{name}.{which}.out").read_text().split()])
        assert len(vals) == len(ref)
        # the reference prints %f (6 decimals); the restatement must round to the same text
        assert np.max(np.abs(vals - ref)) <= 1.0e-6
    mism = sum(("%f" % v) != t for v, t in zip(vals, (GOLDEN / f"{name}.pairhmm_matrix.out").read_text().split()))
    assert mism == 0


def test_pairhmm_10s_survey_anchors(oracle_mod, tmp_path):
    # anchors recorded by the survey for the rebuilt oracle
    p = tmp_path / "10s.in"
    p.write_bytes(read_golden("pairhmm_10s.in"))
    vals, nb = oracle_mod.pairhmm_file(str(p))
    assert nb == 7 and len(vals) == 3550
    assert ["%f" % v for v in vals[:5]] == ["-4.485565", "-1.686275", "-5.842611", "-6.316984", "-2.081956"]
    assert "%f" % vals.min() == "-59.743534" and "%f" % vals.max() == "-1.664610"


def test_oracle_vs_compiled_reference_random(oracle_mod, agx, tmp_path):
    """Property test against the live reference binaries (only where oracle/_ref exists)."""
    if not oracle_mod.ref_available("sw_antidiag"):
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(11)
    data = agx.synth.sw_random_file(rng, 40, 1, 120, alphabet=b"ACGTN")
    p = tmp_path / "r.in"
    p.write_bytes(data)
    ref, header, _ = oracle_mod.run_ref_sw(str(p))
    mine, h2 = oracle_mod.sw_file(str(p))
    assert header == h2 and ref.tolist() == mine.tolist()
    hm = agx.synth.pairhmm_batches(2, 6, 3, seed=int(rng.integers(1 << 30)), read_len=(5, 90), hap_len=(10, 120))
    p2 = tmp_path / "h.in"
    p2.write_bytes(bytes(hm.buf))
    ref_vals, _ = oracle_mod.run_ref_pairhmm(str(p2), "pairhmm_antidiag")
    mine_vals, _ = oracle_mod.pairhmm_file(str(p2))
    assert np.max(np.abs(ref_vals - mine_vals)) <= 1e-6


def test_blocked_oracle_equals_the_plain_one(agx, oracle_mod):
    """oracle_sw_score_blocked (tile by tile on several threads: what computes the 1 Mbp x 1 Mbp expected score)
    is the same function as oracle_sw_score: random and related pairs, alphabets with N and the newline symbol,
    other scoring parameters, tiles that do not divide the lengths, more threads than tiles."""
    rng = np.random.default_rng(20)
    for t in range(160):
        la, lb = int(rng.integers(1, 500)), int(rng.integers(1, 500))
        alpha = np.frombuffer(b"ACGT" if t % 3 else b"ACGTN\n", np.uint8)
        a = alpha[rng.integers(0, alpha.size, la)]
        if t % 2:
            b = a.copy()
            m = rng.random(la) < 0.1
            b[m] = alpha[rng.integers(0, alpha.size, int(m.sum()))]
            b = b[:max(1, lb)]
        else:
            b = alpha[rng.integers(0, alpha.size, lb)]
        sc = (1, -1, -3, -1) if t % 5 else (3, -2, -5, -2)
        want = oracle_mod.sw_score(a.tobytes(), b.tobytes(), sc)
        got = oracle_mod.sw_score_blocked(a.tobytes(), b.tobytes(), sc, tile=int(rng.integers(1, 90)), threads=int(rng.integers(1, 6)))
        assert got == want
    data = agx.synth.sw_long_pair(6000, seed=3, related=True)
    inp = agx.formats.parse_sw(data, line_buf=1 << 30)
    a = inp.buf[inp.off[0]:inp.off[0] + inp.len[0]].tobytes()
    b = inp.buf[inp.off[1]:inp.off[1] + inp.len[1]].tobytes()
    assert oracle_mod.sw_score_blocked(a, b, tile=1024) == oracle_mod.sw_score(a, b)


def test_committed_long_expected_scores_are_well_formed():
    """tests/golden/sw_long_expected.json (made by tests/golden/make_long_expected.py with the blocked oracle)
    holds the pair bench.py's "sw_long" object and the slow GPU test score."""
    import json
    from conftest import GOLDEN
    recs = json.loads((GOLDEN / "sw_long_expected.json").read_text())
    keys = {(r["len"], r["seed"], r["related"]) for r in recs}
    assert (1_000_000, 5, True) in keys
    for r in recs:
        assert r["cells"] == r["line_bytes"][0] * r["line_bytes"][1] and 0 < r["score"] <= min(r["line_bytes"])
