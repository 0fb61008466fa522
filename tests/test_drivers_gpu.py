"""The C drivers (drivers/smithWaterman.c, drivers/pairHMM.c) are drop-ins for the reference
programs: same command line, same stdout / output-file text on the recorded inputs."""
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, read_golden

pytestmark = pytest.mark.gpu
BIN = ROOT / "drivers" / "bin"


def _strip_elapsed(text):
    return "".join(l + "\n" for l in text.splitlines() if not l.startswith("elapsed"))


@pytest.mark.parametrize("name", ["sw_gen_header", "sw_ragged", "sw_short", "sw_150", "sw_no_trailing_nl",
                                  "sw_alphabet", "sw_linebuf", "sw_dangling", "sw_header_small", "sw_1kbp"])
def test_sw_driver_stdout_is_the_references(name):
    r = subprocess.run([str(BIN / "smithWaterman"), str(GOLDEN / f"{name}.in")], capture_output=True, text=True,
                       env=dict(os.environ, AGX_NUM_GPUS="1"))
    assert r.returncode == 0, r.stderr
    assert r.stdout.splitlines()[-1].startswith("elapsed ")
    assert _strip_elapsed(r.stdout) == (GOLDEN / f"{name}.ref.out").read_text()


def test_sw_driver_long_line_buffer():
    r = subprocess.run([str(BIN / "smithWaterman"), str(GOLDEN / "sw_5kbp.in")], capture_output=True, text=True,
                       env=dict(os.environ, AGX_NUM_GPUS="1", AGX_SW_LINE_BUF="4200000"))
    assert r.returncode == 0, r.stderr
    assert _strip_elapsed(r.stdout) == (GOLDEN / "sw_5kbp.ref_long.out").read_text()


def test_sw_driver_usage_and_errors(tmp_path):
    r = subprocess.run([str(BIN / "smithWaterman")], capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.startswith("Usage: ") and "<file_path>" in r.stderr
    r = subprocess.run([str(BIN / "smithWaterman"), str(tmp_path / "missing")], capture_output=True, text=True)
    assert r.returncode == 1 and "Error opening file" in r.stderr
    empty = tmp_path / "empty"
    empty.write_bytes(b"")
    r = subprocess.run([str(BIN / "smithWaterman"), str(empty)], capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout == "file is empty"


@pytest.mark.parametrize("name", ["pairhmm_test", "pairhmm_10s", "pairhmm_synth_small", "pairhmm_synth_long"])
def test_pairhmm_driver_output(tmp_path, name):
    inp = tmp_path / "in.txt"
    inp.write_bytes(read_golden(f"{name}.in"))
    ref_text = (GOLDEN / f"{name}.pairhmm_antidiag.out").read_text()
    ref = np.array([float(x) for x in ref_text.split()])
    # default (FP32 + rescue) path: within tolerance of the reference's printed values
    out = tmp_path / "out.txt"
    r = subprocess.run([str(BIN / "pairHMM"), str(inp), str(out)], capture_output=True, text=True,
                       env=dict(os.environ, AGX_NUM_GPUS="1"))
    assert r.returncode == 0, r.stderr
    got = np.array([float(x) for x in out.read_text().split()])
    assert got.shape == ref.shape
    assert np.all(np.abs(got - ref) <= 1e-5 * np.abs(ref) + 1.5e-6)
    lines = r.stdout.splitlines()
    n_batches = sum(1 for l in lines if l.startswith("#batch:"))
    assert lines[0] == "#batch: 1" and lines[-1] == f"#batch: {n_batches}"
    assert [l for l in lines if not l.startswith("#")] == out.read_text().split()
    # exact-order FP64 path: the output file is byte-identical to the reference's
    out64 = tmp_path / "out64.txt"
    r = subprocess.run([str(BIN / "pairHMM"), str(inp), str(out64)], capture_output=True, text=True,
                       env=dict(os.environ, AGX_NUM_GPUS="1", AGX_PAIRHMM_FP64="1"))
    assert r.returncode == 0, r.stderr
    assert out64.read_text() == ref_text


def test_pairhmm_driver_usage():
    r = subprocess.run([str(BIN / "pairHMM"), "only-one"], capture_output=True, text=True)
    assert r.returncode == 1 and "<input_file_r> <output_file>" in r.stderr


def test_sw_gpu_cli_of_the_reference(tmp_path, agx, oracle_mod):
    """drivers/smithWatermanGpu.c keeps the command line and the output conventions of hipvers.cpp /
    smithWaterman.cu: <input> <output> <block_size>, results appended, 10000-byte line buffer."""
    import sys
    inp, out = tmp_path / "input.txt", tmp_path / "scores.txt"
    subprocess.run([sys.executable, str(ROOT / "drivers" / "generator.py"), "900", "1400", "40", "--seed", "3",
                    "--out", str(inp)], check=True)
    parsed = agx.formats.parse_sw(inp.read_bytes(), line_buf=10000)
    want = oracle_mod.sw_scores_flat(parsed.buf, parsed.off, parsed.len).tolist()
    assert len(want) == 20                                         # header = alignments: half the pairs (SW-Q2)
    out.write_text("Score: -1\n")                                  # the reference opens the file in append mode
    for block_size in ("64", "256"):
        r = subprocess.run([str(BIN / "smithWatermanGpu"), str(inp), str(out), block_size], capture_output=True,
                           text=True, env=dict(os.environ, AGX_NUM_GPUS="1"))
        assert r.returncode == 0, r.stderr
        lines = r.stdout.splitlines()
        assert lines[0].startswith("[main] Using Device ") and "B200" in lines[0]
        assert lines[1:4] == ["num_of_sequences: 40", f"[main] block_size: {block_size}", "[main] grid_size: 20"]
        assert lines[4].startswith("elapsed ")
    got = [int(l.split()[1]) for l in out.read_text().splitlines()]
    assert got == [-1] + want + want
    r = subprocess.run([str(BIN / "smithWatermanGpu"), str(inp)], capture_output=True, text=True,
                       env=dict(os.environ, AGX_NUM_GPUS="1"))
    assert r.returncode == 1 and "<input_file_path> <output_file_path> <block_size>" in r.stderr


@pytest.mark.parametrize("name", ["sw_ends_ragged", "sw_ends_repeats", "sw_dangling", "sw_linebuf", "sw_ends_no_trailing_nl"])
def test_sw_align_driver(agx, oracle_mod, name):
    """drivers/smithWatermanAlign.c: the reference's line structure with the alignment on every score line; scores
    equal the recorded reference scores, coordinates and CIGAR the oracle's."""
    r = subprocess.run([str(BIN / "smithWatermanAlign"), str(GOLDEN / f"{name}.in")], capture_output=True, text=True,
                       env=dict(os.environ, AGX_NUM_GPUS="1"))
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert lines[0].startswith("line_num: ") and lines[-1].startswith("elapsed ")
    inp = agx.formats.parse_sw((GOLDEN / f"{name}.in").read_bytes())
    rows = [l for l in lines if l.startswith("Score: ")]
    assert len(rows) == inp.n_pairs
    data = inp.buf.tobytes()
    for p, row in enumerate(rows):
        a = data[inp.off[2 * p]:inp.off[2 * p] + inp.len[2 * p]]
        b = data[inp.off[2 * p + 1]:inp.off[2 * p + 1] + inp.len[2 * p + 1]]
        s, c, cig = oracle_mod.sw_align(a, b)
        want = f"Score: {s} a[{c[0]},{c[1]}] b[{c[2]},{c[3]}] {oracle_mod.cigar_string(cig)}" if s > 0 else "Score: 0 - - -"
        assert row == want
    if inp.dangling:
        assert inp.dangling.decode().rstrip("\n") in lines
    ref = GOLDEN / f"{name}.ref.out"
    if ref.exists():
        from conftest import ref_scores
        assert [int(l.split()[1]) for l in rows] == ref_scores(f"{name}.ref.out")
