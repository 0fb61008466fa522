"""The drop-in drivers' command lines where no GPU is involved -- wrong argument counts, a file that cannot be opened,
an empty file -- against the compiled reference programs (oracle/_ref, live) and against the reference's own strings
(antidiagonalSmithWaterman.c:190-199, antidiagsPairHMM.c:314-331): same text on the same stream, same exit code."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "drivers" / "bin"
REF = ROOT / "oracle" / "_ref"

PAIRS = [("smithWaterman", "sw_antidiag"), ("pairHMM", "pairhmm_antidiag")]
pytestmark = pytest.mark.skipif(not (BIN / "smithWaterman").exists(), reason="drivers are not built (run build())")


def _run(exe, args):
    r = subprocess.run([str(exe)] + [str(a) for a in args], capture_output=True, text=True, timeout=60)
    norm = lambda t: t.replace(str(exe), "PROG")
    return r.returncode, norm(r.stdout), norm(r.stderr)


@pytest.mark.parametrize("drv,ref", PAIRS)
@pytest.mark.parametrize("args", [[], ["a", "b", "c", "d"], ["/nonexistent/input"], ["/nonexistent/input", "/tmp/agx_cli_out.txt"]])
def test_usage_and_open_errors_equal_the_reference(drv, ref, args):
    got = _run(BIN / drv, args)
    assert got[0] == 1 and got[1] == ""                 # both messages go to stderr (fprintf(stderr, ...) / perror)
    if drv == "smithWaterman":
        want_err = "Usage: PROG <file_path>\n" if len(args) != 1 else "Error opening file: No such file or directory\n"
    else:
        want_err = ("Usage: PROG <input_file_r> <output_file>\n" if len(args) != 2 else
                    "Error opening input file_r: No such file or directory\n")
    assert got[2] == want_err
    if (REF / ref).exists():
        assert got == _run(REF / ref, args)


def test_empty_sw_file(tmp_path):
    f = tmp_path / "empty.txt"
    f.write_bytes(b"")
    got = _run(BIN / "smithWaterman", [f])
    assert got == (1, "file is empty", "")              # antidiagonalSmithWaterman.c:205-208: no newline, exit code 1
    if (REF / "sw_antidiag").exists():
        assert got == _run(REF / "sw_antidiag", [f])


def test_align_driver_usage():
    rc, out, err = _run(BIN / "smithWatermanAlign", [])
    assert rc == 1 and out == "" and err == "Usage: PROG <file_path>\n"
