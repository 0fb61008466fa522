#!/usr/bin/env python3
"""Seeded, argv-aware version of the reference's smithWaterman/generator.py.

    generator.py MIN_LEN MAX_LEN [NUM_OF_ALIGNMENTS] [--seed S] [--out input.txt] [--header-lines]

The reference script ignores its arguments (hiprun.sh passes "$i $i" to it all the same) and is unseeded; this
one keeps its FORMAT -- line 1 = NUM_OF_ALIGNMENTS, then 2 * NUM_OF_ALIGNMENTS lines over "ATGC" with lengths
drawn uniformly from [MIN_LEN, MAX_LEN] (generator.py:4-27) -- and makes the lengths, the count and the seed
arguments.  Note the reference quirk it preserves by default: the programs read line 1 as the number of
SEQUENCE LINES, so they score only the first half of the pairs; --header-lines writes 2 * NUM_OF_ALIGNMENTS
instead so that every pair is scored."""
import argparse
import random


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("min_len", type=int, nargs="?", default=450)
    ap.add_argument("max_len", type=int, nargs="?", default=500)
    ap.add_argument("num_of_alignments", type=int, nargs="?", default=500)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default="input.txt")
    ap.add_argument("--header-lines", action="store_true")
    a = ap.parse_args(argv)
    if a.min_len < 0 or a.max_len < a.min_len:
        ap.error("need 0 <= MIN_LEN <= MAX_LEN")
    rng = random.Random(a.seed)
    with open(a.out, "w") as f:
        f.write(str(2 * a.num_of_alignments if a.header_lines else a.num_of_alignments) + "\n")
        for _ in range(2 * a.num_of_alignments):
            n = rng.randint(a.min_len, a.max_len)
            f.write("".join(rng.choice("ATGC") for _ in range(n)) + "\n")


if __name__ == "__main__":
    main()
