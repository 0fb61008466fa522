"""Pins of the GATK-semantics PairHMM mode (the quirk it corrects:
antidiagsPairHMM.c:111-113 = pairHMMmatrix.c:32-34, mismatch prior Qr instead of Qr / 3).

The reference has no such mode, so the pins are (i) tests/golden/pairhmm_gatk.json -- values of an independent
full-matrix LoglessPairHMM statement (tests/golden/make_gatk_golden.py) on the reference's own test_set inputs,
(ii) closed forms derived by hand for one-row reads, (iii) the exact relation to the reference's value where every
path emits the same number of mismatches.  CPU tests check oracle/oracle.c's gatk branch; the `gpu` tests check
libagx's GATK mode through the C ABI (tolerance 1e-5 relative on log10, as for the reference mode)."""
import importlib.util
import json
import math

import numpy as np
import pytest

from conftest import GOLDEN, read_golden

REL_TOL = 1e-5
LOG10_3 = math.log10(3.0)


def _load_independent():
    spec = importlib.util.spec_from_file_location("make_gatk_golden", GOLDEN / "make_gatk_golden.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _golden_rows(name):
    return json.loads((GOLDEN / "pairhmm_gatk.json").read_text())["files"][name]


def _pairs_of(name, rows):
    """(read 5-tuple, haplotype) of every row of the golden file, from the input itself"""
    ind = _load_independent()
    batches = ind.read_batches(GOLDEN / name)
    return [(batches[r["batch"]][0][r["read"]], batches[r["batch"]][1][r["hap"]]) for r in rows]


# ------------------------------------------------------------------ CPU: the oracle's gatk branch
@pytest.mark.parametrize("name", ["pairhmm_test.in.gz", "pairhmm_10s.in.gz"])
def test_oracle_gatk_branch_equals_independent_statement(oracle_mod, name):
    rows = _golden_rows(name)
    assert len(rows) >= (1 if "test" in name else 200)
    for row, (rd, hp) in zip(rows, _pairs_of(name, rows)):
        for mode, key in ((1, "gatk"), (3, "gatk_floor")):
            got = oracle_mod.pairhmm_forward(rd, hp, gatk=mode)
            assert abs(got - row[key]) <= 1e-11 * abs(row[key]), (name, row, mode, got)


def test_oracle_base_quality_floor_on_low_qualities(oracle_mod):
    """the reference's test_set files hold no base quality below 6, so there the floor changes nothing; seeded
    synthetic pairs with qualities 0 .. 14 (inputs recorded in the golden file) separate the two GATK modes"""
    rows = json.loads((GOLDEN / "pairhmm_gatk.json").read_text())["synthetic_low_quality"]
    assert len(rows) >= 20 and sum(r["gatk"] != r["gatk_floor"] for r in rows) >= 15
    for r in rows:
        rd = tuple(x.encode() for x in r["read"])
        for mode, key in ((1, "gatk"), (3, "gatk_floor")):
            got = oracle_mod.pairhmm_forward(rd, r["hap"].encode(), gatk=mode)
            assert abs(got - r[key]) <= 1e-11 * abs(r[key]), (r, mode, got)


def test_golden_file_is_what_the_script_gives():
    """the committed json is reproducible: recompute a few rows with the independent statement"""
    ind = _load_independent()
    rows = _golden_rows("pairhmm_10s.in.gz")
    pairs = _pairs_of("pairhmm_10s.in.gz", rows)
    for k in (0, 17, 101, len(rows) - 1):
        rd, hp = pairs[k]
        assert ind.gatk_forward(*rd, hp, qual_floor=False) == rows[k]["gatk"]
        assert ind.gatk_forward(*rd, hp, qual_floor=True) == rows[k]["gatk_floor"]


def test_committed_reference_pair_shifts_by_one_mismatch(oracle_mod):
    """pairHMM/test_set/test.in: read and haplotype differ in ONE base (position 21, T vs C), and the all-match path
    dominates, so the corrected prior moves the reference's -4.485565 by log10 3 (to 4 decimals)"""
    row = _golden_rows("pairhmm_test.in.gz")[0]
    ref = float((GOLDEN / "pairhmm_test.committed.out").read_text().split()[0])
    assert abs((ref - row["gatk"]) - LOG10_3) < 1e-4
    rd, hp = _pairs_of("pairhmm_test.in.gz", [row])[0]
    assert sum(a != b for a, b in zip(rd[0], hp)) == 1 and len(rd[0]) == len(hp)
    assert abs(oracle_mod.pairhmm_forward(rd, hp, gatk=0) - ref) < 5.1e-7


ONE_ROW = [  # (base, qual, ins, del, gcp, haplotype)
    (b"A", b"I", b"I", b"I", b"+", b"C"),
    (b"A", b"5", b"?", b"D", b"+", b"CGTTG"),
    (b"G", b"'", b"I", b"I", b"5", b"ACTACT"),
    (b"T", b"#", b"I", b"I", b"+", b"CCCC"),      # quality 2: the floor applies in mode 3
]


@pytest.mark.parametrize("case", ONE_ROW)
def test_one_row_closed_form(oracle_mod, case):
    """R = 1: M[1][j] = prior_j * (1 - Qg) * init, X[1][j] = 0, so the result is log10((1 - Qg) * mean_j prior_j).
    With no matching base every path emits exactly one mismatch: GATK mode = reference mode - log10 3, exactly."""
    b, q, qi, qd, qg, hap = case
    p = lambda c: 10.0 ** (-(c - 33) / 10.0)
    Qr, Qg = p(q[0]), p(qg[0])
    assert b[0] not in hap
    want_ref = math.log10((1 - Qg) * Qr)
    want_gatk = math.log10((1 - Qg) * Qr / 3)
    rd = (b, q, qi, qd, qg)
    assert abs(oracle_mod.pairhmm_forward(rd, hap, gatk=0) - want_ref) < 1e-12
    assert abs(oracle_mod.pairhmm_forward(rd, hap, gatk=1) - want_gatk) < 1e-12
    assert abs(oracle_mod.pairhmm_forward(rd, hap, gatk=0) - oracle_mod.pairhmm_forward(rd, hap, gatk=1) - LOG10_3) < 1e-12
    Qf = max(Qr, 0.0) if q[0] - 33 >= 6 else p(33 + 6)
    assert abs(oracle_mod.pairhmm_forward(rd, hap, gatk=3) - math.log10((1 - Qg) * Qf / 3)) < 1e-12
    # a matching base somewhere: the mean of the priors
    hap2 = hap + b
    pri = [(1 - Qr) if c == b[0] else Qr / 3 for c in hap2]
    assert abs(oracle_mod.pairhmm_forward(rd, hap2, gatk=1) - math.log10((1 - Qg) * sum(pri) / len(hap2))) < 1e-12


def test_matching_pairs_do_not_move(oracle_mod):
    """a read that matches its haplotype base for base, high qualities: mismatching paths are negligible, both modes
    agree to ~1e-4; a read of N matches everything: the modes agree exactly"""
    hap = b"ACGTTGCAAGGCTTAACCGGT"
    rd = (hap, b"I" * len(hap), b"I" * len(hap), b"I" * len(hap), b"+" * len(hap))
    assert abs(oracle_mod.pairhmm_forward(rd, hap, gatk=0) - oracle_mod.pairhmm_forward(rd, hap, gatk=1)) < 1e-3
    rn = (b"N" * 9, b"5" * 9, b"I" * 9, b"I" * 9, b"+" * 9)
    assert oracle_mod.pairhmm_forward(rn, hap, gatk=0) == oracle_mod.pairhmm_forward(rn, hap, gatk=1)


# ------------------------------------------------------------------ GPU: libagx's GATK mode through the C ABI
def _gpu_values(agx, gpu_lib, name, mode):
    inp = agx.formats.parse_pairhmm(read_golden(name[:-3]))
    gpu_lib.set_pairhmm_gatk_mode(mode)
    try:
        got = gpu_lib.pairhmm_forward_flat(inp.buf, inp.read_field_off, inp.read_len, inp.hap_off, inp.hap_len,
                                           inp.batch_read_start, inp.batch_hap_start)
    finally:
        gpu_lib.set_pairhmm_gatk_mode(0)
    return inp, got


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["pairhmm_test.in.gz", "pairhmm_10s.in.gz"])
@pytest.mark.parametrize("mode,key", [(1, "gatk"), (3, "gatk_floor")])
def test_gpu_gatk_mode_equals_independent_statement(agx, gpu_lib, name, mode, key):
    inp, got = _gpu_values(agx, gpu_lib, name, mode)
    nh = np.diff(inp.batch_hap_start)
    nr = np.diff(inp.batch_read_start)
    batch_out = np.concatenate(([0], np.cumsum(nr * nh)))
    rows = _golden_rows(name)
    for row in rows:
        k = int(batch_out[row["batch"]] + row["read"] * nh[row["batch"]] + row["hap"])
        assert abs(got[k] - row[key]) <= REL_TOL * abs(row[key]), (row, got[k])


@pytest.mark.gpu
@pytest.mark.parametrize("fp64", [False, True])
def test_gpu_one_row_closed_form(gpu_lib, fp64):
    """the hand-derived one-row values through pairhmm_forward_batch, FP32 stream kernels and the FP64 kernel"""
    gpu_lib.set_pairhmm_force_fp64(fp64)
    try:
        for b, q, qi, qd, qg, hap in ONE_ROW:
            p = lambda c: 10.0 ** (-(c - 33) / 10.0)
            Qr, Qg = p(q[0]), p(qg[0])
            for mode, want in ((0, math.log10((1 - Qg) * Qr)), (1, math.log10((1 - Qg) * Qr / 3)),
                               (3, math.log10((1 - Qg) * (Qr if q[0] - 33 >= 6 else p(39)) / 3))):
                gpu_lib.set_pairhmm_gatk_mode(mode)
                got = gpu_lib.pairhmm_forward_batch([(b, q, qi, qd, qg)], [hap])
                assert abs(float(np.ravel(got)[0]) - want) <= (1e-12 if fp64 else REL_TOL) * abs(want), (b, hap, mode)
    finally:
        gpu_lib.set_pairhmm_gatk_mode(0)
        gpu_lib.set_pairhmm_force_fp64(False)
