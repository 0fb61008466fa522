"""Alignment end cell / start cell / CIGAR on the GPU (sw_ends_batch_flat, sw_align_batch_flat) through the C ABI:
the END CELL against the recorded output of the instrumented reference build (tests/golden/*.ref_ends.out, the cell
the reference's own running maximum comes from), everything against the CPU oracle (oracle/sw_align.c), bit-exact,
and -- independently of any oracle tie rule -- every CIGAR re-scored to the Smith-Waterman score."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

ENDS_FILES = ["sw_ends_two_letter", "sw_ends_repeats", "sw_ends_ragged", "sw_ends_150", "sw_ends_no_trailing_nl",
              "sw_ends_mid", "sw_ends_alphabet"]


def _flat(a, b):
    parts, off, ln, at = [], [], [], 0
    for x, y in zip(a, b):
        for s in (x, y):
            parts.append(s)
            off.append(at)
            ln.append(len(s))
            at += len(s)
    buf = np.frombuffer(b"".join(parts) + b"\0", np.uint8)
    return buf, np.asarray(off, np.int64), np.asarray(ln, np.int32)


def _pairs(rng, n, lo, hi, alphabet=b"ACGT", nl=True, related=0.6):
    alpha = np.frombuffer(alphabet, np.uint8)
    a, b = [], []
    for _ in range(n):
        la, lb = rng.integers(lo, hi + 1, size=2)
        x = alpha[rng.integers(0, alpha.size, size=la)]
        if rng.random() < related:
            y = x.copy()
            m = rng.random(la) < 0.08
            y[m] = alpha[rng.integers(0, alpha.size, size=int(m.sum()))]
            y = y[rng.random(la) > 0.03]
            if rng.random() < 0.5:
                y = np.concatenate([alpha[rng.integers(0, alpha.size, size=int(rng.integers(0, 12)))], y])
            if rng.random() < 0.3 and y.size > 4:      # a longer gap
                k = int(rng.integers(1, y.size - 1))
                y = np.concatenate([y[:k], y[k + int(rng.integers(1, 9)):]])
        else:
            y = alpha[rng.integers(0, alpha.size, size=lb)]
        tail = b"\n" if nl else b""
        a.append(x.tobytes() + tail)
        b.append(y.tobytes() + tail)
    return a, b


def _check_all(gpu_lib, oracle_mod, a, b, scoring=(1, -1, -3, -1), full=True):
    buf, off, ln = _flat(a, b)
    s1, ends = gpu_lib.sw_ends_flat(buf, off, ln, scoring)
    s2, coords, coff, cig = gpu_lib.sw_align_flat(buf, off, ln, scoring)
    s0 = gpu_lib.sw_score_flat(buf, off, ln, scoring)
    assert s1.tolist() == s0.tolist() == s2.tolist()
    assert coff[0] == 0 and coff[-1] == cig.size
    bad = []
    for p, (x, y) in enumerate(zip(a, b)):
        runs = cig[coff[p]:coff[p + 1]].tolist()
        if full:
            ws, wc, wg = oracle_mod.sw_align(x, y, scoring)
            if (int(s2[p]), tuple(coords[p].tolist()), runs, tuple(ends[p].tolist())) != (ws, wc, wg, (wc[1], wc[3])):
                bad.append((p, int(s2[p]), coords[p].tolist(), oracle_mod.cigar_string(runs), ends[p].tolist(), ws, wc, oracle_mod.cigar_string(wg)))
        else:
            assert (coords[p][1], coords[p][3]) == tuple(ends[p].tolist())
            if s2[p] > 0:
                assert oracle_mod.sw_cigar_score(x, y, coords[p].tolist(), runs, scoring) == s2[p]
            else:
                assert runs == [] and coords[p].tolist() == [-1] * 4
    assert not bad, f"{len(bad)} pairs differ from the oracle, first: {bad[:3]}"


@pytest.mark.parametrize("name", ENDS_FILES)
def test_end_cells_equal_the_reference_scan(agx, gpu_lib, oracle_mod, name):
    inp = agx.formats.parse_sw((GOLDEN / f"{name}.in").read_bytes(), line_buf=20000)
    rows = oracle_mod.parse_ref_sw_ends((GOLDEN / f"{name}.ref_ends.out").read_text())
    scores, ends = gpu_lib.sw_ends_flat(inp.buf, inp.off, inp.len)
    assert scores.tolist() == [r[0] for r in rows]
    assert [tuple(e) for e in ends.tolist()] == [oracle_mod.ref_ends_to_coords(r) for r in rows]
    s2, coords, coff, cig = gpu_lib.sw_align_flat(inp.buf, inp.off, inp.len)
    assert s2.tolist() == scores.tolist()
    assert coords[:, 1].tolist() == ends[:, 0].tolist() and coords[:, 3].tolist() == ends[:, 1].tolist()
    data = inp.buf.tobytes()
    for p in range(inp.n_pairs):
        a = data[inp.off[2 * p]:inp.off[2 * p] + inp.len[2 * p]]
        b = data[inp.off[2 * p + 1]:inp.off[2 * p + 1] + inp.len[2 * p + 1]]
        ws, wc, wg = oracle_mod.sw_align(a, b)
        assert (ws, list(wc), wg) == (int(s2[p]), coords[p].tolist(), cig[coff[p]:coff[p + 1]].tolist())


def test_random_ragged_short(gpu_lib, oracle_mod):
    rng = np.random.default_rng(101)
    a, b = _pairs(rng, 2500, 1, 200)
    _check_all(gpu_lib, oracle_mod, a, b)


def test_tie_rich_two_letter_and_repeats(gpu_lib, oracle_mod):
    rng = np.random.default_rng(102)
    a, b = _pairs(rng, 1500, 1, 120, alphabet=b"AC", related=0.3)
    for _ in range(500):
        unit = bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(1, 5))).astype(np.uint8))
        a.append(unit * int(rng.integers(2, 40)) + b"\n")
        b.append(unit * int(rng.integers(2, 40)) + (b"\n" if rng.random() < 0.8 else b""))
    _check_all(gpu_lib, oracle_mod, a, b)


def test_every_length_class(gpu_lib, oracle_mod):
    rng = np.random.default_rng(103)
    a, b = [], []
    for hi in (30, 64, 96, 128, 152, 192, 256, 384, 512, 768, 1020):
        x, y = _pairs(rng, 24, max(1, hi - 40), hi)
        a += x
        b += y
    _check_all(gpu_lib, oracle_mod, a, b)


def test_wavefront_path_long_and_odd_bytes(gpu_lib, oracle_mod):
    rng = np.random.default_rng(104)
    a, b = _pairs(rng, 10, 1100, 2600)                       # longer than the packed kernel's columns
    x, y = _pairs(rng, 300, 1, 150, alphabet=b"ACGTNacgt")    # bytes outside ACGT
    a += x
    b += y
    x, y = _pairs(rng, 6, 200, 300)
    a += [s[:-1] + b"N" * 3 + b"\n" for s in x]               # N on both sides
    b += [s[:-1] + b"N" * 3 + b"\n" for s in y]
    a += [b"ACGT" * 700 + b"\n"]                              # short columns, many rows
    b += [b"ACGTT" * 9 + b"\n"]
    _check_all(gpu_lib, oracle_mod, a, b)


def test_newline_symbol_and_degenerate_lines(gpu_lib, oracle_mod):
    a = [b"ACGT\n", b"ACGT\n", b"ACGT", b"\n", b"\n", b"", b"A", b"AAAA\n", b"ACGTACGT\n", b"T\n", b"ACGT\n"]
    b = [b"ACGT\n", b"ACGT", b"ACGT\n", b"\n", b"ACGT\n", b"ACGT\n", b"A", b"TTTT\n", b"TTACGTACGTTT\n", b"\n", b"ACGA\n"]
    _check_all(gpu_lib, oracle_mod, a, b)
    buf, off, ln = _flat(a, b)
    s, coords, coff, cig = gpu_lib.sw_align_flat(buf, off, ln)
    assert s.tolist()[:4] == [5, 4, 4, 1]
    assert coords[0].tolist() == [0, 4, 0, 4] and cig[coff[0]:coff[1]].tolist() == [5 << 4]
    assert coords[3].tolist() == [0, 0, 0, 0] and coords[5].tolist() == [-1] * 4


@pytest.mark.parametrize("scoring", [(3, -2, -5, -2), (2, -3, 0, -2), (5, -4, -10, -1), (1, -3, -6, -1), (30, -30, -60, -30)])
def test_other_scoring(gpu_lib, oracle_mod, scoring):
    rng = np.random.default_rng(105 + scoring[0])
    a, b = _pairs(rng, 500, 1, 180)
    x, y = _pairs(rng, 3, 1100, 1500)
    _check_all(gpu_lib, oracle_mod, a + x, b + y, scoring)


def test_unsupported_ranges(gpu_lib, agx):
    buf, off, ln = _flat([b"ACGT"], [b"ACGT"])
    for sc in [(200, -1, -3, -1), (1, -200, -3, -1), (1, -1, -200, -1)]:
        with pytest.raises(agx.capi.AgxError) as e:
            gpu_lib.sw_align_flat(buf, off, ln, sc)
        assert e.value.code == -5


def test_cigar_capacity(gpu_lib, agx, oracle_mod):
    rng = np.random.default_rng(106)
    a, b = _pairs(rng, 200, 50, 150)
    buf, off, ln = _flat(a, b)
    s, coords, coff, cig = gpu_lib.sw_align_flat(buf, off, ln)
    with pytest.raises(agx.capi.AgxError) as e:
        gpu_lib.sw_align_flat(buf, off, ln, cigar_cap=3)
    assert e.value.code == -5 and str(cig.size) in str(e.value)
    s2, c2, o2, g2 = gpu_lib.sw_align_flat(buf, off, ln, cigar_cap=int(cig.size))
    assert g2.tolist() == cig.tolist() and o2.tolist() == coff.tolist()


def test_chunked_matrices_give_the_same_result(gpu_lib, oracle_mod):
    """a small traceback budget cuts the batch into many chunks (AGX_ALIGN_TB_BYTES)"""
    rng = np.random.default_rng(107)
    a, b = _pairs(rng, 3000, 20, 260)
    buf, off, ln = _flat(a, b)
    want = gpu_lib.sw_align_flat(buf, off, ln)
    os.environ["AGX_ALIGN_TB_BYTES"] = str(4 << 20)
    try:
        got = gpu_lib.sw_align_flat(buf, off, ln)
    finally:
        del os.environ["AGX_ALIGN_TB_BYTES"]
    for g, w in zip(got, want):
        assert g.tolist() == w.tolist()


@pytest.mark.parametrize("chunk", [1, 37, 1000])
def test_two_lane_chunk_pipeline_gives_the_same_result(gpu_lib, oracle_mod, chunk):
    """the shard alternates its chunks between two lanes (upload of chunk k+1 under the kernels of chunk k, results of
    chunk k under chunk k+1): any chunk size gives the single-chunk answer (AGX_ALIGN_CHUNK)"""
    rng = np.random.default_rng(108)
    a, b = _pairs(rng, 2500, 1, 200)
    buf, off, ln = _flat(a, b)
    os.environ["AGX_ALIGN_CHUNK"] = "0"
    try:
        want = gpu_lib.sw_align_flat(buf, off, ln)
        want_ends = gpu_lib.sw_ends_flat(buf, off, ln)
        os.environ["AGX_ALIGN_CHUNK"] = str(chunk)
        got = gpu_lib.sw_align_flat(buf, off, ln)
        got_ends = gpu_lib.sw_ends_flat(buf, off, ln)
    finally:
        del os.environ["AGX_ALIGN_CHUNK"]
    for g, w in zip(got + got_ends, want + want_ends):
        assert g.tolist() == w.tolist()
    assert got[2][0] == 0 and got[2][-1] == got[3].size and np.all(np.diff(got[2]) >= 0)
    for p in (0, 36, 37, 38, 999, 1000, 2499):           # chunk seams, against the oracle
        ws, wc, wg = oracle_mod.sw_align(a[p], b[p], (1, -1, -3, -1))
        assert (int(got[0][p]), tuple(got[1][p].tolist()), got[3][got[2][p]:got[2][p + 1]].tolist()) == (ws, wc, wg)


def test_config3_batch_properties(agx, gpu_lib, oracle_mod):
    """BASELINE configs[2] shape (150 x 150): 60 000 pairs; every CIGAR re-scored, 1 000 pairs against the oracle."""
    inp = agx.synth.sw_uniform_pairs(60000, 150, seed=31)
    scores, coords, coff, cig = gpu_lib.sw_align_flat(inp.buf, inp.off, inp.len)
    s0 = gpu_lib.sw_score_flat(inp.buf, inp.off, inp.len)
    s1, ends = gpu_lib.sw_ends_flat(inp.buf, inp.off, inp.len)
    assert scores.tolist() == s0.tolist() == s1.tolist()
    assert coords[:, 1].tolist() == ends[:, 0].tolist() and coords[:, 3].tolist() == ends[:, 1].tolist()
    data = inp.buf.tobytes()
    rng = np.random.default_rng(1)
    pick = set(rng.choice(inp.n_pairs, size=1000, replace=False).tolist())
    for p in range(inp.n_pairs):
        a = data[inp.off[2 * p]:inp.off[2 * p] + inp.len[2 * p]]
        b = data[inp.off[2 * p + 1]:inp.off[2 * p + 1] + inp.len[2 * p + 1]]
        runs = cig[coff[p]:coff[p + 1]].tolist()
        if scores[p] > 0:
            assert oracle_mod.sw_cigar_score(a, b, coords[p].tolist(), runs) == scores[p], p
        if p in pick:
            ws, wc, wg = oracle_mod.sw_align(a, b)
            assert (ws, list(wc), wg) == (int(scores[p]), coords[p].tolist(), runs), p


@pytest.mark.parametrize("n,related,alphabet", [(17000, True, b"ACGT"), (17500, False, b"ACGT"), (18000, True, b"AC")])
def test_whole_gpu_pairs_report_their_end_cell(agx, gpu_lib, oracle_mod, n, related, alphabet):
    """sw_ends_batch_flat on pairs of >= 2^28 cells: the striped long-alignment kernel keeps the end cell too
    (64-bit keys in the reference's visiting order); mixed with ordinary pairs in one call."""
    rng = np.random.default_rng(n)
    alpha = np.frombuffer(alphabet, np.uint8)
    x = alpha[rng.integers(0, alpha.size, size=n)]
    if related:
        y = x.copy()
        m = rng.random(n) < 0.03
        y[m] = alpha[rng.integers(0, alpha.size, size=int(m.sum()))]
        y = y[rng.random(n) > 0.004]
    else:
        y = alpha[rng.integers(0, alpha.size, size=n - 321)]
    big_a, big_b = x.tobytes() + b"\n", y.tobytes() + b"\n"
    a, b = _pairs(rng, 40, 10, 200)
    a.insert(7, big_a)
    b.insert(7, big_b)
    a.append(big_b)                 # the same pair the other way round: line 1 shorter / longer decides ix
    b.append(big_a)
    buf, off, ln = _flat(a, b)
    scores, ends = gpu_lib.sw_ends_flat(buf, off, ln)
    assert scores.tolist() == gpu_lib.sw_score_flat(buf, off, ln).tolist()
    for p in (7, len(a) - 1):
        ws, we = oracle_mod.sw_ends(a[p], b[p])
        assert (int(scores[p]), tuple(ends[p].tolist())) == (ws, we)
    for p in (0, 20, 40):
        ws, wc, _ = oracle_mod.sw_align(a[p], b[p])
        assert (int(scores[p]), tuple(ends[p].tolist())) == (ws, (wc[1], wc[3]))
    with pytest.raises(agx.capi.AgxError) as e:
        gpu_lib.sw_align_flat(buf, off, ln)             # the traceback of such a pair is refused
    assert e.value.code == -5


def test_pointer_array_forms(gpu_lib, oracle_mod):
    """sw_ends_batch / sw_align_batch: the pointer-array signatures (pair p = a[p] vs b[p]), embedded NUL bytes kept"""
    rng = np.random.default_rng(108)
    a, b = _pairs(rng, 300, 1, 160)
    a.append(b"AC\0GTAC\0GT")
    b.append(b"AC\0GTTC\0GT")
    s1, ends = gpu_lib.sw_ends_batch(a, b)
    s2, coords, coff, cig = gpu_lib.sw_align_batch(a, b)
    for p, (x, y) in enumerate(zip(a, b)):
        ws, wc, wg = oracle_mod.sw_align(x, y)
        assert (int(s1[p]), tuple(ends[p].tolist())) == (ws, (wc[1], wc[3]))
        assert (int(s2[p]), tuple(coords[p].tolist()), cig[coff[p]:coff[p + 1]].tolist()) == (ws, wc, wg)


def test_randomised_against_the_oracle(gpu_lib):
    """profiles/align_fuzz.py for a few seconds: lengths around every class boundary, long rows on short columns,
    tandem repeats, alphabets with ties and odd bytes, missing newlines, seven scorings (a 150 s run of the same
    script: 181 376 pairs, 0 mismatches -- profiles/r2an_align_fuzz.jsonl)."""
    import subprocess
    import sys
    from conftest import ROOT
    gpu_lib.shutdown()
    try:
        r = subprocess.run([sys.executable, str(ROOT / "profiles" / "align_fuzz.py"), "8", "7"], capture_output=True, text=True, timeout=300)
    finally:
        gpu_lib.init(1)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"mismatches": 0' in r.stdout


def test_device_resident_entry_points(agx, gpu_lib, oracle_mod):
    """sw_ends_batch_device / sw_align_batch_device on torch tensors: equal to the host entry points"""
    import torch
    rng = np.random.default_rng(109)
    a, b = _pairs(rng, 3000, 1, 300)
    buf, off, ln = _flat(a, b)
    n = len(a)
    dev = torch.device("cuda", 0)
    d_buf, d_off, d_len = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (buf, off, ln))
    d_scores = torch.empty(n, dtype=torch.int32, device=dev)
    d_ends = torch.empty((n, 2), dtype=torch.int32, device=dev)
    d_coords = torch.empty((n, 4), dtype=torch.int32, device=dev)
    d_coff = torch.empty(n + 1, dtype=torch.int64, device=dev)
    d_cig = torch.empty(8 * n, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    gpu_lib.sw_ends_device(0, d_buf.data_ptr(), d_buf.numel(), d_off.data_ptr(), d_len.data_ptr(), n, d_scores.data_ptr(),
                           d_ends.data_ptr(), st)
    torch.cuda.synchronize(dev)
    want_s, want_e = gpu_lib.sw_ends_flat(buf, off, ln)
    assert np.array_equal(d_scores.cpu().numpy(), want_s) and np.array_equal(d_ends.cpu().numpy(), want_e)
    total = gpu_lib.sw_align_device(0, d_buf.data_ptr(), d_buf.numel(), d_off.data_ptr(), d_len.data_ptr(), n, d_scores.data_ptr(),
                                    d_coords.data_ptr(), d_coff.data_ptr(), d_cig.data_ptr(), d_cig.numel(), st)
    torch.cuda.synchronize(dev)
    ws, wc, wo, wg = gpu_lib.sw_align_flat(buf, off, ln)
    assert total == wg.size
    assert np.array_equal(d_scores.cpu().numpy(), ws) and np.array_equal(d_coords.cpu().numpy(), wc)
    assert np.array_equal(d_coff.cpu().numpy(), wo)
    assert np.array_equal(d_cig.cpu().numpy()[:total].view(np.uint32), wg)
    with pytest.raises(agx.capi.AgxError) as e:
        gpu_lib.sw_align_device(0, d_buf.data_ptr(), d_buf.numel(), d_off.data_ptr(), d_len.data_ptr(), n, d_scores.data_ptr(),
                                d_coords.data_ptr(), d_coff.data_ptr(), d_cig.data_ptr(), 5, st)
    assert e.value.code == -5
