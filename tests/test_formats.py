"""Host-side format handling (accelerating-genomics_b200/formats.py, synth.py) against the oracle's
file-level restatement and the recorded reference outputs.  CPU only."""
import numpy as np
import pytest

from conftest import GOLDEN, read_golden, ref_scores


@pytest.mark.parametrize("name", ["sw_gen_header", "sw_ragged", "sw_no_trailing_nl", "sw_linebuf",
                                  "sw_dangling", "sw_header_small", "sw_1kbp", "sw_alphabet"])
def test_parse_sw_reproduces_reference_pairing(agx, oracle_mod, name):
    data = (GOLDEN / f"{name}.in").read_bytes()
    inp = agx.formats.parse_sw(data)
    want = ref_scores(f"{name}.ref.out")
    assert inp.n_pairs == len(want)
    got = oracle_mod.sw_scores_flat(inp.buf, inp.off, inp.len)
    assert got.tolist() == want
    assert f"line_num: {inp.header}" in (GOLDEN / f"{name}.ref.out").read_text()


def test_parse_sw_quirks(agx):
    F = agx.formats
    # SW-Q2: header counts lines; generator.py writes the number of alignments -> half the pairs
    inp = F.parse_sw(b"2\nAAAA\nCCCC\nGGGG\nTTTT\n")
    assert inp.n_pairs == 1 and inp.header == 2
    # SW-Q1: the newline stays in the sequence
    assert inp.len.tolist() == [5, 5]
    # SW-Q3: 1000-byte buffer splits a 1000-character line into 999 + ("X\n")
    long = b"A" * 1000
    inp = F.parse_sw(b"2\n" + long + b"\n" + b"C" * 10 + b"\n")
    assert inp.len.tolist() == [999, 2]
    inp = F.parse_sw(b"2\n" + long + b"\n" + b"C" * 10 + b"\n", line_buf=10000)
    assert inp.len.tolist() == [1001, 11]
    # EOF in the middle of a pair: the dangling line is reported, not scored
    inp = F.parse_sw(b"4\nAA\nCC\nGG\n")
    assert inp.n_pairs == 1 and inp.dangling == b"GG\n"
    # header larger than the file / zero / garbage
    assert F.parse_sw(b"100\nAA\nCC\n").n_pairs == 1
    assert F.parse_sw(b"0\nAA\nCC\n").n_pairs == 0
    assert F.parse_sw(b"x\nAA\nCC\n").n_pairs == 0
    with pytest.raises(ValueError):
        F.parse_sw(b"")


def test_write_sw_roundtrip(agx):
    F = agx.formats
    pairs = [(b"ACGT", b"AGGT"), (b"T", b"TT")]
    inp = F.parse_sw(F.write_sw(pairs))
    assert inp.n_pairs == 2
    seqs = [inp.buf[o:o + l].tobytes() for o, l in zip(inp.off, inp.len)]
    assert seqs == [b"ACGT\n", b"AGGT\n", b"T\n", b"TT\n"]


@pytest.mark.parametrize("name", ["pairhmm_test", "pairhmm_10s", "pairhmm_synth_small", "pairhmm_synth_long"])
def test_parse_pairhmm_against_oracle(agx, oracle_mod, tmp_path, name):
    data = read_golden(f"{name}.in")
    inp = agx.formats.parse_pairhmm(data)
    ref = np.array([float(x) for x in (GOLDEN / f"{name}.pairhmm_matrix.out").read_text().split()])
    assert inp.n_pairs == len(ref)
    limit = 200
    got = oracle_mod.pairhmm_flat(inp, limit=limit)
    assert np.max(np.abs(got - ref[:len(got)])) <= 1e-6


def test_parse_pairhmm_shapes(agx):
    inp = agx.formats.parse_pairhmm(read_golden("pairhmm_10s.in"))
    assert inp.n_batches == 7 and inp.n_pairs == 3550 and inp.cells() == 62380634
    one = agx.formats.parse_pairhmm(read_golden("pairhmm_test.in"))
    assert one.n_batches == 1 and one.read_len.tolist() == [41] and one.hap_len.tolist() == [41]


def test_synth_matches_its_own_index(agx):
    S, F = agx.synth, agx.formats
    hm = S.pairhmm_batches(3, 7, 2, seed=4, read_len=(10, 40), hap_len=(30, 60))
    re = F.parse_pairhmm(bytes(hm.buf))
    for a, b in ((hm.read_field_off, re.read_field_off), (hm.read_len, re.read_len), (hm.hap_off, re.hap_off),
                 (hm.hap_len, re.hap_len), (hm.batch_read_start, re.batch_read_start),
                 (hm.batch_hap_start, re.batch_hap_start)):
        assert np.array_equal(np.asarray(a), np.asarray(b))
    sw = S.sw_uniform_pairs(33, 150, seed=2)
    re = F.parse_sw(bytes(sw.buf))
    assert np.array_equal(sw.off, re.off) and np.array_equal(sw.len, re.len) and re.header == 66
    assert set(np.unique(sw.buf[sw.off[0]:sw.off[0] + 150]).tolist()) <= set(b"ACGT")
    # seeded: same seed, same bytes
    assert np.array_equal(S.sw_uniform_pairs(33, 150, seed=2).buf, sw.buf)


def test_seeded_generator_keeps_the_reference_format(agx, oracle_mod, tmp_path):
    """drivers/generator.py: generator.py's format (line 1 = alignments, 2N lines over ATGC, lengths in
    [MIN, MAX]) with the lengths, the count and the seed as arguments."""
    import subprocess
    import sys
    from conftest import ROOT
    gen = ROOT / "drivers" / "generator.py"
    a, b = tmp_path / "a.txt", tmp_path / "b.txt"
    for out in (a, b):
        subprocess.run([sys.executable, str(gen), "64", "80", "30", "--seed", "7", "--out", str(out)], check=True)
    assert a.read_bytes() == b.read_bytes()                       # seeded
    lines = a.read_text().split("\n")
    assert lines[0] == "30" and lines[-1] == "" and len(lines) == 62
    assert all(64 <= len(l) <= 80 and set(l) <= set("ATGC") for l in lines[1:-1])
    inp = agx.formats.parse_sw(a.read_bytes())
    assert inp.n_pairs == 15                                      # SW-Q2: the programs score half of them
    subprocess.run([sys.executable, str(gen), "64", "80", "30", "--seed", "8", "--header-lines", "--out", str(b)],
                   check=True)
    assert b.read_bytes() != a.read_bytes() and agx.formats.parse_sw(b.read_bytes()).n_pairs == 30


def _irregular_pairhmm_file(agx):
    inp = agx.synth.pairhmm_batches(7, 9, 3, seed=3)
    lines = bytes(inp.buf).split(b"\n")[:-1]
    heads = [i for i, l in enumerate(lines) if len(l.split()) == 2 and l.split()[0].isdigit()]
    a, b = lines[heads[1]].split()
    lines[heads[1]] = b"  " + a + b"\t " + b                     # leading blanks, a tab
    a, b = lines[heads[2]].split()
    lines[heads[2]] = b"+" + a + b" " + b + b" trailing words"    # sign, trailing text
    lines.insert(heads[3], b"0 0")                                # a batch with nothing in it
    return b"\n".join(lines)                                      # and no newline after the last haplotype


@pytest.mark.parametrize("which", ["pairhmm_matrix", "pairhmm_antidiag_noleak"])
def test_pairhmm_header_walk_is_the_references(agx, oracle_mod, tmp_path, which):
    """The mirror of the batch walk (formats.parse_pairhmm, which the GPU parser is tested against) agrees
    with the compiled reference on headers only sscanf("%d %d") would accept."""
    if not oracle_mod.ref_available(which):
        pytest.skip("compiled reference not available")
    data = _irregular_pairhmm_file(agx)
    path = tmp_path / "irregular.in"
    path.write_bytes(data)
    vals, _ = oracle_mod.run_ref_pairhmm(str(path), which)
    host = agx.formats.parse_pairhmm(data)
    assert host.n_batches == 8 and len(vals) == host.n_pairs == 189
    assert np.max(np.abs(vals - oracle_mod.pairhmm_flat(host))) <= 1e-6


@pytest.mark.parametrize("keep,extra,message", [(3 * 13 + 5, b"", b"Error reading haplotypes."),
                                                (3 * 13 + 11, b"", b"Error reading haplotypes."),
                                                (13, b"3 0\nACGT IIII NNNN NNNN ++++\n", b"Error reading reads.")])
def test_pairhmm_truncated_file_message_of_the_reference(agx, oracle_mod, tmp_path, keep, extra, message):
    """What the reference says when the file ends inside a batch: its haplotype cursor runs ahead of its read
    cursor (antidiagsPairHMM.c:388-396), so it is "Error reading haplotypes." unless the batch has none.
    (The reference then crashes in its cleanup; only the message is pinned here.)"""
    import subprocess
    if not oracle_mod.ref_available("pairhmm_antidiag_noleak"):
        pytest.skip("compiled reference not available")
    inp = agx.synth.pairhmm_batches(4, 9, 3, seed=5)
    lines = bytes(inp.buf).split(b"\n")[:-1]                      # a batch = 1 + 9 + 3 lines
    path = tmp_path / "cut.in"
    path.write_bytes(b"\n".join(lines[:keep]) + b"\n" + extra)
    r = subprocess.run([str(oracle_mod.REF / "pairhmm_antidiag_noleak"), str(path), str(tmp_path / "o")], capture_output=True)
    assert r.returncode != 0 and r.stderr.startswith(message)


def _inherited_header_file(agx, trailing_blank=False):
    """Batch 2's header holds one integer, batch 3's is a blank line: sscanf("%d %d") leaves what it cannot parse
    as it was, and the reference declares both counts once (antidiagsPairHMM.c:345-346, :378), so both batches
    inherit from the batch before them."""
    inp = agx.synth.pairhmm_batches(4, 9, 3, seed=11)
    lines = bytes(inp.buf).split(b"\n")[:-1]                      # a batch = 1 + 9 + 3 lines
    assert lines[13] == b"9 3" and lines[26] == b"9 3"
    lines[13] = b"9"
    lines[26] = b""
    return b"\n".join(lines) + b"\n" + (b"\n" if trailing_blank else b"")


@pytest.mark.parametrize("which", ["pairhmm_matrix", "pairhmm_antidiag_noleak"])
def test_pairhmm_header_counts_are_inherited_like_the_references(agx, oracle_mod, tmp_path, which):
    if not oracle_mod.ref_available(which):
        pytest.skip("compiled reference not available")
    data = _inherited_header_file(agx)
    path = tmp_path / "inherit.in"
    path.write_bytes(data)
    vals, _ = oracle_mod.run_ref_pairhmm(str(path), which)
    host = agx.formats.parse_pairhmm(data)
    assert host.n_batches == 4 and len(vals) == host.n_pairs == 4 * 27
    assert np.max(np.abs(vals - oracle_mod.pairhmm_flat(host))) <= 1e-6
    port, nb = oracle_mod.pairhmm_file(str(path))
    assert nb == 4 and np.max(np.abs(vals - port)) <= 1e-6


def test_pairhmm_trailing_blank_line_is_a_truncated_batch_in_the_reference(agx, oracle_mod, tmp_path):
    """A blank line after the last batch is read as one more header that inherits the last counts; the file
    then ends inside that batch: "Error reading haplotypes." and a non-zero exit, the earlier batches kept."""
    import subprocess
    if not oracle_mod.ref_available("pairhmm_antidiag_noleak"):
        pytest.skip("compiled reference not available")
    data = _inherited_header_file(agx, trailing_blank=True)
    path = tmp_path / "blank.in"
    path.write_bytes(data)
    r = subprocess.run([str(oracle_mod.REF / "pairhmm_antidiag_noleak"), str(path), str(tmp_path / "o")], capture_output=True)
    assert r.returncode != 0 and r.stderr.startswith(b"Error reading haplotypes.")
    assert agx.formats.parse_pairhmm(data).n_batches == 4
    _, nb = oracle_mod.pairhmm_file(str(path))
    assert nb == 4
