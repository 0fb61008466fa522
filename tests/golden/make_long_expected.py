"""Expected scores of the long Smith-Waterman pairs bench.py / the tests run (BASELINE configs[4]: 1 Mbp x 1 Mbp),
computed on the CPU by oracle_sw_score_blocked() -- the oracle's recurrence, tile by tile on all host cores
(10^12 cells: ~10-15 minutes on 8 cores).  Writes tests/golden/sw_long_expected.json.

    python tests/golden/make_long_expected.py [LEN:SEED:RELATED ...]      default: 1000000:5:1 1000000:5:0

The pair is accelerating-genomics_b200/synth.py: sw_long_pair(LEN, seed=SEED, related=RELATED), scored with the
reference's constants (antidiagonalSmithWaterman.c:40-43) on the raw lines INCLUDING their newline symbol, i.e.
exactly what `sw_score_batch_flat` is given by bench.py.
"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import agxpkg   # noqa: E402
import oracle   # noqa: E402

OUT = Path(__file__).resolve().parent / "sw_long_expected.json"


def main():
    agx = agxpkg.load()
    specs = sys.argv[1:] or ["1000000:5:1", "1000000:5:0"]
    have = json.loads(OUT.read_text()) if OUT.exists() else []
    for spec in specs:
        n, seed, related = (int(x) for x in spec.split(":"))
        inp = agx.formats.parse_sw(agx.synth.sw_long_pair(n, seed=seed, related=bool(related)), line_buf=1 << 30)
        a = inp.buf[inp.off[0]:inp.off[0] + inp.len[0]].tobytes()
        b = inp.buf[inp.off[1]:inp.off[1] + inp.len[1]].tobytes()
        t0 = time.time()
        score = oracle.sw_score_blocked(a, b, tile=4096)
        dt = time.time() - t0
        rec = {"len": n, "seed": seed, "related": bool(related), "score": score, "line_bytes": [len(a), len(b)],
               "cells": len(a) * len(b), "cpu_seconds": round(dt, 1),
               "how": "oracle.sw_score_blocked (oracle/sw_blocked.c), tile 4096, all host cores"}
        have = [h for h in have if (h["len"], h["seed"], h["related"]) != (n, seed, bool(related))] + [rec]
        OUT.write_text(json.dumps(have, indent=1) + "\n")
        print(rec, flush=True)


if __name__ == "__main__":
    main()
