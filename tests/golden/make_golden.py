#!/usr/bin/env python
"""Regenerates tests/golden/ by RUNNING THE UNMODIFIED REFERENCE PROGRAMS (oracle/_ref, built by
`make -C oracle ref` from the read-only /root/reference mount).  Only runs where /root/reference
exists (the build container); the GPU box uses the committed files.

Inputs are seeded, so the files are reproducible.  Outputs are the reference's own stdout /
output-file text (the SW `elapsed` line, which is a wall-clock time, is dropped).

PairHMM inputs test.in / 10s.in are the reference's own test_set files (data, not source),
stored gzip-compressed so BASELINE config 2 can run on the GPU box where /root/reference is absent.
"""
from __future__ import annotations

import gzip
import shutil
import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
import agxpkg  # noqa: E402

agx = agxpkg.load()
import oracle  # noqa: E402

REFDIR = Path("/root/reference")


def ref_sw(path: Path, long_lines=False) -> str:
    _, _, text = oracle.run_ref_sw(str(path), long_lines=long_lines)
    return "".join(l + "\n" for l in text.splitlines() if not l.startswith("elapsed"))


def make_ends() -> None:
    """END CELLS of the reference's own scan: oracle/_ref/sw_antidiag_ends is the reference source with position
    bookkeeping added beside its running maximum (oracle/Makefile); its "Score: s End: iy ix Sx: line" lines are
    recorded for tie-rich inputs (two-letter alphabets, tandem repeats, equal and unequal lengths, a last line
    without newline)."""
    rng = np.random.default_rng(20261019)
    S = agx.synth

    def repeats(n_pairs):
        out = [str(2 * n_pairs).encode()]
        for _ in range(n_pairs):
            unit = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=int(rng.integers(1, 5)))].tobytes()
            a = unit * int(rng.integers(3, 30))
            b = unit * int(rng.integers(3, 30))
            if rng.random() < 0.5:
                a = a[int(rng.integers(0, 3)):] + np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=5)].tobytes()
            out += [a, b]
        return b"\n".join(out) + b"\n"

    files = {
        "sw_ends_two_letter.in": S.sw_random_file(rng, 120, 1, 90, alphabet=b"AC"),
        "sw_ends_repeats.in": repeats(120),
        "sw_ends_ragged.in": S.sw_random_file(rng, 120, 1, 150, related_frac=0.5),
        "sw_ends_150.in": bytes(S.sw_uniform_pairs(96, 150, seed=11).buf),
        "sw_ends_no_trailing_nl.in": S.sw_random_file(rng, 7, 20, 60, alphabet=b"AC", trailing_newline=False),
        "sw_ends_mid.in": S.sw_random_file(rng, 16, 300, 900, related_frac=0.5),
        "sw_ends_alphabet.in": S.sw_random_file(rng, 60, 5, 120, alphabet=b"ACGTNacgt"),
    }
    for name, data in files.items():
        p = HERE / name
        p.write_bytes(data)
        r = subprocess.run([str(ROOT / "oracle" / "_ref" / "sw_antidiag_ends"), str(p)], capture_output=True, check=True)
        text = "".join(l + "\n" for l in r.stdout.decode().splitlines() if not l.startswith("elapsed"))
        (HERE / (name[:-3] + ".ref_ends.out")).write_text(text)
    print("end-cell golden files written")


def main() -> None:
    subprocess.run(["make", "-C", str(ROOT / "oracle"), "ref", "liboracle.so"], check=True)
    if "--ends-only" in sys.argv:
        make_ends()
        return
    make_ends()
    rng = np.random.default_rng(20261018)
    S = agx.synth

    sw_files = {
        # generator.py's own convention: header = number of ALIGNMENTS, so half are scored (SW-Q2)
        "sw_gen_header.in": S.sw_random_file(rng, 24, 450, 500, header=24),
        "sw_ragged.in": S.sw_random_file(rng, 60, 1, 150),
        "sw_short.in": S.sw_random_file(rng, 80, 1, 40, related_frac=0.7),
        "sw_150.in": bytes(S.sw_uniform_pairs(64, 150, seed=7).buf),
        "sw_no_trailing_nl.in": S.sw_random_file(rng, 9, 5, 90, trailing_newline=False),
        "sw_alphabet.in": S.sw_random_file(rng, 40, 5, 120, alphabet=b"ACGTNacgt"),
        "sw_two_letter.in": S.sw_random_file(rng, 30, 20, 200, alphabet=b"AC"),
        # lines of 998 / 999 / 1000 / 1500 characters: the 1000-byte fgets buffer splits them (SW-Q3)
        "sw_linebuf.in": b"8\n" + b"\n".join(
            np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=n)].tobytes()
            for n in (998, 998, 999, 999, 1000, 1000, 1500, 700)) + b"\n",
        # odd number of lines: EOF in the middle of a pair (the reference echoes the dangling line)
        "sw_dangling.in": b"6\nACGTACGT\nACGTTCGT\nGATTACA\nGATCACA\nTTTT\n",
        "sw_header_small.in": b"3\nACGTACGT\nACGTTCGT\nGATTACA\nGATCACA\nTTTT\nTTTA\n",
        "sw_mid.in": S.sw_random_file(rng, 12, 300, 700),
    }
    for name, data in sw_files.items():
        p = HERE / name
        p.write_bytes(data)
        (HERE / (name[:-3] + ".ref.out")).write_text(ref_sw(p))

    # BASELINE config 1: one 1 kbp x 1 kbp pair.  The unmodified binary mis-parses >= 999-char lines
    # (both outputs are recorded); the MAX_LINE_LENGTH-raised build gives the intended alignment.
    p = HERE / "sw_1kbp.in"
    p.write_bytes(S.sw_long_pair(1000, seed=1))
    (HERE / "sw_1kbp.ref.out").write_text(ref_sw(p))
    (HERE / "sw_1kbp.ref_long.out").write_text(ref_sw(p, long_lines=True))
    p = HERE / "sw_5kbp.in"
    p.write_bytes(S.sw_long_pair(5000, seed=2))
    (HERE / "sw_5kbp.ref_long.out").write_text(ref_sw(p, long_lines=True))

    # PairHMM: the reference's own test_set + seeded synthetic batches
    for name in ("test.in", "10s.in"):
        with open(REFDIR / "pairHMM" / "test_set" / name, "rb") as f, \
                gzip.GzipFile(HERE / f"pairhmm_{name}.gz", "wb", mtime=0) as g:
            shutil.copyfileobj(f, g)
    shutil.copyfile(REFDIR / "pairHMM" / "test_set" / "test.out", HERE / "pairhmm_test.committed.out")
    hmm_inputs = {
        "pairhmm_test": (REFDIR / "pairHMM" / "test_set" / "test.in").read_bytes(),
        "pairhmm_10s": (REFDIR / "pairHMM" / "test_set" / "10s.in").read_bytes(),
        "pairhmm_synth_small": bytes(S.pairhmm_batches(3, 12, 3, seed=5, read_len=(10, 120),
                                                       hap_len=(20, 160), n_frac=0.02).buf),
        "pairhmm_synth_cfg4": bytes(S.pairhmm_batches(2, 20, 5, seed=6).buf),
        # reads longer than the haplotype, and reads longer than 256 rows (striped kernel)
        "pairhmm_synth_long": bytes(S.pairhmm_batches(2, 6, 2, seed=8, read_len=(200, 600),
                                                      hap_len=(700, 900)).buf),
        "pairhmm_synth_tall": bytes(S.pairhmm_batches(2, 5, 2, seed=9, read_len=(60, 140),
                                                      hap_len=(145, 160)).buf),
    }
    for name, data in hmm_inputs.items():
        p = HERE / f"{name}.in"
        if not name.startswith(("pairhmm_test", "pairhmm_10s")):
            p.write_bytes(data)
            src = p
        else:
            src = REFDIR / "pairHMM" / "test_set" / (name.split("_", 1)[1] + ".in")
        for which in ("pairhmm_antidiag", "pairhmm_matrix"):
            _, text = oracle.run_ref_pairhmm(str(src), which)
            (HERE / f"{name}.{which}.out").write_text(text)
    print("golden files written to", HERE)


if __name__ == "__main__":
    main()
