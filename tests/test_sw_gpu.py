"""Smith-Waterman parity on the GPU, through the C ABI (libagx.so), against the recorded reference
outputs and the CPU oracle.  Bit-exact."""
import numpy as np
import pytest

from conftest import GOLDEN, ref_scores

pytestmark = pytest.mark.gpu

SW_FILES = ["sw_gen_header", "sw_ragged", "sw_short", "sw_150", "sw_no_trailing_nl", "sw_alphabet",
            "sw_two_letter", "sw_linebuf", "sw_dangling", "sw_header_small", "sw_mid", "sw_1kbp"]


@pytest.mark.parametrize("name", SW_FILES)
def test_golden_files_bit_exact(agx, gpu_lib, name):
    inp = agx.formats.parse_sw((GOLDEN / f"{name}.in").read_bytes())
    got = gpu_lib.sw_score_flat(inp.buf, inp.off, inp.len)
    assert got.tolist() == ref_scores(f"{name}.ref.out")


@pytest.mark.parametrize("name", ["sw_1kbp", "sw_5kbp"])
def test_golden_long_lines(agx, gpu_lib, name):
    # BASELINE config 1 (1 kbp x 1 kbp) with un-split lines: the MAX_LINE_LENGTH-raised reference
    inp = agx.formats.parse_sw((GOLDEN / f"{name}.in").read_bytes(), line_buf=4200000)
    got = gpu_lib.sw_score_flat(inp.buf, inp.off, inp.len)
    assert got.tolist() == ref_scores(f"{name}.ref_long.out")


def _random_pairs(rng, n, lo, hi, alphabet=b"ACGT", nl=True, related=0.5):
    alpha = np.frombuffer(alphabet, np.uint8)
    a, b = [], []
    for _ in range(n):
        la, lb = rng.integers(lo, hi + 1, size=2)
        x = alpha[rng.integers(0, alpha.size, size=la)]
        if rng.random() < related:
            y = x.copy()
            m = rng.random(la) < 0.1
            y[m] = alpha[rng.integers(0, alpha.size, size=int(m.sum()))]
            y = y[rng.random(la) > 0.04][:lb] if la else y
        else:
            y = alpha[rng.integers(0, alpha.size, size=lb)]
        tail = b"\n" if nl else b""
        a.append(x.tobytes() + tail)
        b.append(y.tobytes() + tail)
    return a, b


def _check(gpu_lib, oracle_mod, a, b, scoring=(1, -1, -3, -1)):
    got = gpu_lib.sw_score_batch(a, b, scoring)
    want = [oracle_mod.sw_score(x, y, scoring) for x, y in zip(a, b)]
    bad = [i for i, (g, w) in enumerate(zip(got.tolist(), want)) if g != w]
    assert not bad, f"{len(bad)} mismatches, first {bad[:5]}: got {got[bad[:5]]}, want {[want[i] for i in bad[:5]]}"


def test_random_ragged_batch(gpu_lib, oracle_mod):
    rng = np.random.default_rng(1)
    a, b = _random_pairs(rng, 3000, 1, 300)
    _check(gpu_lib, oracle_mod, a, b)


def test_every_length_class_boundary(gpu_lib, oracle_mod):
    # column capacities of the duo kernel: 32 64 96 128 152 192 256 384 512 768 1024, then generic
    rng = np.random.default_rng(2)
    a, b = [], []
    for L in (1, 2, 31, 32, 33, 63, 64, 65, 96, 97, 128, 129, 150, 151, 152, 153, 192, 193, 256, 257, 384, 385,
              512, 513, 768, 769, 1023, 1024, 1025, 1500):
        for rows in (L, L + 1, 2 * L + 7, 40):
            x, y = _random_pairs(rng, 1, L, L, related=1.0)
            a.append(x[0])
            yy = (y[0].rstrip(b"\n") * 3)[:rows] + b"\n"
            b.append(yy)
    _check(gpu_lib, oracle_mod, a, b)


def test_long_rows_short_columns(gpu_lib, oracle_mod):
    rng = np.random.default_rng(3)
    alpha = np.frombuffer(b"ACGT", np.uint8)
    a = [alpha[rng.integers(0, 4, size=n)].tobytes() + b"\n" for n in (20, 100, 150, 300)]
    b = [alpha[rng.integers(0, 4, size=n)].tobytes() + b"\n" for n in (5000, 3000, 20000, 9000)]
    # plant the short sequence inside the long one so the best score is large
    b = [y[:1000] + x.rstrip(b"\n") + y[1000:] for x, y in zip(a, b)]
    _check(gpu_lib, oracle_mod, a, b)


@pytest.mark.parametrize("alphabet", [b"ACGTN", b"acgtACGT", b"AC", b"ACGT\r", b"ACGU*-"])
def test_byte_exact_alphabets(gpu_lib, oracle_mod, alphabet):
    rng = np.random.default_rng(4)
    a, b = _random_pairs(rng, 300, 1, 200, alphabet=alphabet)
    _check(gpu_lib, oracle_mod, a, b)


def test_newline_combinations_and_empties(gpu_lib, oracle_mod):
    a = [b"ACGT\n", b"AAAA\n", b"ACGT\n", b"ACGT", b"\n", b"\n", b"", b"ACGT\n", b"A", b"\n\n", b"AC\nGT\n"]
    b = [b"ACGT\n", b"TTTT\n", b"ACGT", b"ACGT", b"\n", b"ACGT\n", b"ACGT\n", b"", b"A", b"\n", b"AC\nGT\n"]
    _check(gpu_lib, oracle_mod, a, b)
    got = gpu_lib.sw_score_batch(a, b)
    assert got[:4].tolist() == [5, 1, 4, 4]       # SURVEY 8b probes


def test_no_newline_batch(gpu_lib, oracle_mod):
    rng = np.random.default_rng(5)
    a, b = _random_pairs(rng, 500, 1, 180, nl=False)
    _check(gpu_lib, oracle_mod, a, b)


@pytest.mark.parametrize("scoring", [(2, -3, -5, -2), (1, -4, 0, -1), (5, -4, -10, -1), (3, -1, -200, -3)])
def test_other_scoring_parameters(gpu_lib, oracle_mod, scoring):
    rng = np.random.default_rng(6)
    a, b = _random_pairs(rng, 400, 1, 220)
    _check(gpu_lib, oracle_mod, a, b, scoring)


def test_scoring_sign_convention_is_enforced(gpu_lib):
    with pytest.raises(gpu_lib.AgxError) as e:
        gpu_lib.sw_score_batch([b"ACGT"], [b"ACGT"], (1, 1, -3, -1))
    assert e.value.code == -5


def test_s16_overflow_is_routed_to_s32(gpu_lib, oracle_mod):
    # match = 40: 1000 matching bases would overflow a signed 16-bit half
    x = (b"ACGT" * 250) + b"\n"
    _check(gpu_lib, oracle_mod, [x], [x], (40, -1, -3, -1))
    assert gpu_lib.sw_score_batch([x], [x], (40, -1, -3, -1))[0] == 40 * 1001


def test_device_resident_entry_point(agx, gpu_lib, oracle_mod):
    import torch
    inp = agx.synth.sw_uniform_pairs(4096, 150, seed=3)
    dev = torch.device("cuda:0")
    d_buf = torch.from_numpy(inp.buf.copy()).to(dev)
    d_off = torch.from_numpy(inp.off).to(dev)
    d_len = torch.from_numpy(inp.len).to(dev)
    d_out = torch.full((inp.n_pairs,), -7, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    gpu_lib.sw_score_device(0, d_buf.data_ptr(), d_buf.numel(), d_off.data_ptr(), d_len.data_ptr(),
                            inp.n_pairs, d_out.data_ptr(), st)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    want = oracle_mod.sw_scores_flat(inp.buf, inp.off[:2 * 512], inp.len[:2 * 512])
    assert got[:512].tolist() == want.tolist()
    assert np.array_equal(got, gpu_lib.sw_score_flat(inp.buf, inp.off, inp.len))


def test_config3_shape_properties(agx, gpu_lib, oracle_mod):
    """BASELINE config 3 shape (150 x 150), 200k pairs: size-independent properties + oracle sample."""
    n = 200_000
    inp = agx.synth.sw_uniform_pairs(n, 150, seed=9)
    got = gpu_lib.sw_score_flat(inp.buf, inp.off, inp.len)
    assert got.min() >= 1 and got.max() <= 151
    # symmetry: swapping a and b never changes the score
    swap = np.arange(2 * n).reshape(n, 2)[:, ::-1].reshape(-1)
    assert np.array_equal(gpu_lib.sw_score_flat(inp.buf, inp.off[swap], inp.len[swap]), got)
    # a sequence against itself scores length + 1 (the newline symbol matches too)
    self_off = np.repeat(inp.off[0::2], 2)
    assert np.all(gpu_lib.sw_score_flat(inp.buf, self_off, inp.len) == 151)
    # order independence: a permuted batch gives the permuted scores
    perm = np.random.default_rng(0).permutation(n)
    idx = np.stack([2 * perm, 2 * perm + 1], axis=1).reshape(-1)
    assert np.array_equal(gpu_lib.sw_score_flat(inp.buf, inp.off[idx], inp.len[idx]), got[perm])
    # oracle on a sample
    pick = np.random.default_rng(1).choice(n, size=1500, replace=False)
    idx = np.stack([2 * pick, 2 * pick + 1], axis=1).reshape(-1)
    want = oracle_mod.sw_scores_flat(inp.buf, inp.off[idx], inp.len[idx])
    assert got[pick].tolist() == want.tolist()


# ---------------------------------------------------------------- long alignments (intra-task wavefront over all SMs)
def _long_pair(agx, n, seed, related=True):
    data = agx.synth.sw_long_pair(n, seed=seed, related=related)
    return agx.formats.parse_sw(data, line_buf=1 << 30)


@pytest.mark.parametrize("n,related", [(20000, True), (18000, False), (33000, True)])
def test_long_alignment_whole_gpu_kernel(agx, gpu_lib, oracle_mod, n, related):
    """>= 2^28 cells: the pair takes the stripe-pipelined whole-GPU kernel (BASELINE config 5 path)."""
    inp = _long_pair(agx, n, seed=n, related=related)
    assert int(inp.len[0]) * int(inp.len[1]) >= 1 << 28
    gpu_lib.set_profiling(True)
    got = gpu_lib.sw_score_flat(inp.buf, inp.off, inp.len)
    want = oracle_mod.sw_scores_flat(inp.buf, inp.off, inp.len)
    assert got.tolist() == want.tolist()
    if related:
        assert got[0] > n // 2


@pytest.mark.parametrize("k,rows", [(2, 2), (4, 4), (7, 4), (7, 2), (8, 2), (14, 4), (16, 2), (27, 2), (27, 4), (32, 2), (32, 4)])
def test_long_alignment_stripe_widths_and_rows_per_step(agx, gpu_lib, oracle_mod, k, rows, monkeypatch):
    """sw_longr_kernel: instantiated stripe widths, two and four rows per step; DNA takes the symbol-coded kernels."""
    monkeypatch.setenv("AGX_LONG_K", str(k))
    monkeypatch.setenv("AGX_LONG_R", str(rows))
    inp = _long_pair(agx, 17000 + 111 * k + rows, seed=k + rows, related=bool(rows & 4))
    got = gpu_lib.sw_score_flat(inp.buf, inp.off, inp.len)
    assert got.tolist() == oracle_mod.sw_scores_flat(inp.buf, inp.off, inp.len).tolist()


@pytest.mark.parametrize("bsteps,rows,n", [(1, 4, 16411), (3, 2, 17003), (8, 4, 16999), (16, 2, 21001), (32, 4, 18001)])
def test_long_alignment_hand_off_block_sizes(agx, gpu_lib, oracle_mod, bsteps, rows, n, monkeypatch):
    """Row steps per hand-off block (AGX_LONG_B) from 1 to the ring's limit, row counts that do not divide R * B."""
    monkeypatch.setenv("AGX_LONG_B", str(bsteps))
    monkeypatch.setenv("AGX_LONG_R", str(rows))
    monkeypatch.setenv("AGX_LONG_K", "7")
    inp = _long_pair(agx, n, seed=n + bsteps, related=True)
    got = gpu_lib.sw_score_flat(inp.buf, inp.off, inp.len)
    assert got.tolist() == oracle_mod.sw_scores_flat(inp.buf, inp.off, inp.len).tolist()


@pytest.mark.parametrize("k,n", [(7, 16777), (8, 17001), (27, 18500)])
def test_long_alignment_software_pipelined_kernel(agx, gpu_lib, oracle_mod, k, n, monkeypatch):
    """sw_longp_kernel (rows of a tile run one column apart and wrap into the next step; AGX_LONG_PIPE=1)."""
    monkeypatch.setenv("AGX_LONG_PIPE", "1")
    monkeypatch.setenv("AGX_LONG_R", "4")
    monkeypatch.setenv("AGX_LONG_K", str(k))
    inp = _long_pair(agx, n, seed=n, related=True)
    got = gpu_lib.sw_score_flat(inp.buf, inp.off, inp.len)
    assert got.tolist() == oracle_mod.sw_scores_flat(inp.buf, inp.off, inp.len).tolist()


@pytest.mark.parametrize("k,chain", [(4, 0), (8, 1), (16, 0), (32, 1)])
def test_long_alignment_raw_byte_kernels(agx, gpu_lib, oracle_mod, k, chain, monkeypatch):
    """More than 7 distinct bytes cannot be symbol-coded: sw_long_kernel, both chain forms (forced here on DNA)."""
    monkeypatch.setenv("AGX_LONG_RAW", "1")
    monkeypatch.setenv("AGX_LONG_K", str(k))
    monkeypatch.setenv("AGX_LONG_CHAIN", str(chain))
    inp = _long_pair(agx, 17000 + 111 * k, seed=k + chain, related=bool(chain))
    got = gpu_lib.sw_score_flat(inp.buf, inp.off, inp.len)
    assert got.tolist() == oracle_mod.sw_scores_flat(inp.buf, inp.off, inp.len).tolist()


@pytest.mark.parametrize("alphabet,k", [(b"ACGTNRYKMSWB", 8), (b"ACGTNRYKMSWB", 32), (b"ACGTNac", 6), (b"AC", 10)])
def test_long_alignment_alphabets(agx, gpu_lib, oracle_mod, alphabet, k, monkeypatch):
    """More than 7 distinct bytes: the raw-byte kernels; up to 7 (here with lower case): the coded ones."""
    monkeypatch.setenv("AGX_LONG_K", str(k))
    rng = np.random.default_rng(k)
    alpha = np.frombuffer(alphabet, np.uint8)
    n = 17500
    a = alpha[rng.integers(0, alpha.size, size=n)]
    b = a.copy()
    mut = rng.random(n) < 0.2
    b[mut] = alpha[rng.integers(0, alpha.size, size=int(mut.sum()))]
    buf = np.concatenate([a, b[: n - 300]])
    off = np.array([0, n], dtype=np.int64)
    ln = np.array([n, n - 300], dtype=np.int32)
    assert int(ln[0]) * int(ln[1]) >= 1 << 28
    got = gpu_lib.sw_score_flat(buf, off, ln)
    assert got.tolist() == oracle_mod.sw_scores_flat(buf, off, ln).tolist()


def test_long_alignment_raw_and_coded_kernels_agree(agx, gpu_lib, monkeypatch):
    inp = _long_pair(agx, 60000, seed=3, related=True)
    coded = gpu_lib.sw_score_flat(inp.buf, inp.off, inp.len)
    monkeypatch.setenv("AGX_LONG_RAW", "1")
    assert gpu_lib.sw_score_flat(inp.buf, inp.off, inp.len).tolist() == coded.tolist()
    monkeypatch.delenv("AGX_LONG_RAW")
    for rows in ("2", "4"):
        monkeypatch.setenv("AGX_LONG_R", rows)
        assert gpu_lib.sw_score_flat(inp.buf, inp.off, inp.len).tolist() == coded.tolist()


def test_long_alignment_device_entry_point_and_mixed_batch(agx, gpu_lib, oracle_mod):
    import torch
    big = agx.synth.sw_long_pair(17000, seed=5)
    small = agx.synth.sw_random_file(np.random.default_rng(7), 50, 1, 300, alphabet=b"ACGTN")
    # one file: 50 short pairs, the long pair, with non-ACGT bytes in the long pair too
    body_big = big.split(b"\n", 1)[1].replace(b"ACGTAC", b"ACNTAC", 3)
    data = b"102\n" + small.split(b"\n", 1)[1] + body_big
    inp = agx.formats.parse_sw(data, line_buf=1 << 30)
    assert inp.n_pairs == 51
    want = oracle_mod.sw_scores_flat(inp.buf, inp.off, inp.len)
    got = gpu_lib.sw_score_flat(inp.buf, inp.off, inp.len)
    assert got.tolist() == want.tolist()
    dev = torch.device("cuda:0")
    d_buf, d_off, d_len = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (inp.buf, inp.off, inp.len))
    d_out = torch.zeros(inp.n_pairs, dtype=torch.int32, device=dev)
    gpu_lib.sw_score_device(0, d_buf.data_ptr(), d_buf.numel(), d_off.data_ptr(), d_len.data_ptr(), inp.n_pairs,
                            d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert d_out.cpu().numpy().tolist() == want.tolist()
    assert gpu_lib.profile_ms(0, gpu_lib.PROF_SW_LONG) > 0


def test_long_alignment_properties_100kbp(agx, gpu_lib):
    """100 kbp x 100 kbp (10^10 cells): too large for the oracle in a test; size-independent properties."""
    inp = _long_pair(agx, 100_000, seed=3, related=True)
    s = int(gpu_lib.sw_score_flat(inp.buf, inp.off, inp.len)[0])
    swap = np.array([1, 0])
    assert int(gpu_lib.sw_score_flat(inp.buf, inp.off[swap], inp.len[swap])[0]) == s      # symmetry
    self_s = int(gpu_lib.sw_score_flat(inp.buf, inp.off[[0, 0]], inp.len[[0, 0]])[0])
    assert self_s == 100_001                                                               # a vs a, newline included
    assert 50_000 < s < self_s
    # a prefix pair can never beat the full pair (local alignment is monotone in its inputs)
    half = inp.len.copy()
    half[:] = 50_000
    assert int(gpu_lib.sw_score_flat(inp.buf, inp.off, half)[0]) <= s


def test_long_alignment_200kbp_against_the_oracle(agx, gpu_lib, oracle_mod):
    """200 kbp x 200 kbp (4 * 10^10 cells) against the oracle's recurrence evaluated tile by tile on the host
    cores (oracle_sw_score_blocked, itself pinned to the plain oracle in tests/test_oracle.py)."""
    inp = _long_pair(agx, 200_000, seed=12, related=True)
    a = inp.buf[inp.off[0]:inp.off[0] + inp.len[0]].tobytes()
    b = inp.buf[inp.off[1]:inp.off[1] + inp.len[1]].tobytes()
    want = oracle_mod.sw_score_blocked(a, b)
    assert int(gpu_lib.sw_score_flat(inp.buf, inp.off, inp.len)[0]) == want > 100_000


@pytest.mark.parametrize("related", [True, False])
def test_long_alignment_1mbp_against_the_committed_expected_score(agx, gpu_lib, related):
    """BASELINE configs[4] at full size: 1 Mbp x 1 Mbp (10^12 cells).  The expected scores were computed once on
    the CPU by tests/golden/make_long_expected.py (blocked oracle, ~11 minutes per pair on 8 cores)."""
    import json
    recs = json.loads((GOLDEN / "sw_long_expected.json").read_text())
    want = [r["score"] for r in recs if (r["len"], r["seed"], r["related"]) == (1_000_000, 5, related)][0]
    inp = _long_pair(agx, 1_000_000, seed=5, related=related)
    assert int(gpu_lib.sw_score_flat(inp.buf, inp.off, inp.len)[0]) == want


# ---------------------------------------------------------------- device-side file parser (sw_score_file_image)
@pytest.mark.parametrize("name", SW_FILES + ["sw_5kbp"])
def test_file_image_entry_point_matches_reference_stdout(agx, gpu_lib, name):
    """The GPU-side fgets() chunker reproduces the reference's pairing, header and dangling-line rules."""
    data = (GOLDEN / f"{name}.in").read_bytes()
    if name == "sw_5kbp":
        scores, header, dangling = gpu_lib.sw_score_file_image(data, line_buf=4200000)
        assert scores.tolist() == ref_scores("sw_5kbp.ref_long.out")
        return
    scores, header, dangling = gpu_lib.sw_score_file_image(data)
    text = (GOLDEN / f"{name}.ref.out").read_text()
    assert scores.tolist() == ref_scores(f"{name}.ref.out")
    assert f"line_num: {header}" in text
    host = agx.formats.parse_sw(data)
    assert dangling == host.dangling


@pytest.mark.parametrize("line_buf", [2, 3, 17, 64, 1000, 5000])
def test_file_image_chunking_equals_host_mirror(agx, gpu_lib, line_buf):
    rng = np.random.default_rng(line_buf)
    # ragged lines, some far longer than the buffer, empty lines, no trailing newline, odd line count
    alpha = np.frombuffer(b"ACGT", np.uint8)
    lines = [alpha[rng.integers(0, 4, size=int(n))].tobytes() for n in rng.integers(0, 3 * line_buf + 40, size=301)]
    data = b"400\n" + b"\n".join(lines)
    host = agx.formats.parse_sw(data, line_buf=line_buf)
    scores, header, dangling = gpu_lib.sw_score_file_image(data, line_buf=line_buf)
    assert header == host.header and dangling == host.dangling      # a tiny buffer splits the header line too
    assert scores.tolist() == gpu_lib.sw_score_flat(host.buf, host.off, host.len).tolist()


def test_file_image_config3_shape(agx, gpu_lib):
    inp = agx.synth.sw_uniform_pairs(100_000, 150, seed=11)
    scores, header, dangling = gpu_lib.sw_score_file_image(inp.buf)
    assert header == 200_000 and dangling == b""
    assert np.array_equal(scores, gpu_lib.sw_score_flat(inp.buf, inp.off, inp.len))


@pytest.mark.parametrize("segment", [64, 1000, 4096, 1 << 20])
def test_file_image_streaming_segments(agx, gpu_lib, segment, monkeypatch):
    """Upload segments of any size (pairs, lines and fgets() chunks straddle them) give the same scores."""
    rng = np.random.default_rng(segment)
    alpha = np.frombuffer(b"ACGT", np.uint8)
    lines = [alpha[rng.integers(0, 4, size=int(n))].tobytes() for n in rng.integers(0, 2600, size=403)]
    data = b"404\n" + b"\n".join(lines) + b"\n"
    host = agx.formats.parse_sw(data)
    want = gpu_lib.sw_score_flat(host.buf, host.off, host.len)
    monkeypatch.setenv("AGX_SW_IMAGE_SEGMENT", str(segment))
    scores, header, dangling = gpu_lib.sw_score_file_image(data)
    assert header == host.header and dangling == host.dangling
    assert scores.tolist() == want.tolist()
    for name in SW_FILES:
        g = (GOLDEN / f"{name}.in").read_bytes()
        scores, header, dangling = gpu_lib.sw_score_file_image(g)
        assert scores.tolist() == ref_scores(f"{name}.ref.out")
        assert dangling == agx.formats.parse_sw(g).dangling


def test_file_image_pinned_image_and_output(agx, gpu_lib):
    import torch
    inp = agx.synth.sw_uniform_pairs(50_000, 150, seed=5)
    h_img = torch.from_numpy(inp.buf).pin_memory()
    h_out = torch.empty(50_000, dtype=torch.int32).pin_memory()
    scores, header, dangling = gpu_lib.sw_score_file_image(h_img.numpy(), out=h_out.numpy())
    assert header == 100_000 and dangling == b"" and scores.size == 50_000
    assert np.array_equal(scores, gpu_lib.sw_score_flat(inp.buf, inp.off, inp.len))


def test_short_pairs_with_n_stay_bit_exact(gpu_lib, oracle_mod):
    """'N' in the shorter / the longer / both sequences, at the ends, unequal lengths around the class
    capacities: the s16x2 kernel swaps roles or hands the pair to the byte-exact kernel; either way raw-byte
    semantics (N == N is a match, N vs ACGT a mismatch) hold."""
    rng = np.random.default_rng(42)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    seqs = []
    for k in range(600):
        la = int(rng.integers(1, 320))
        lb = int(rng.integers(1, 320)) if k % 3 else la
        a = acgt[rng.integers(0, 4, size=la)].copy()
        b = a.copy()[:lb] if (k % 4 == 0 and lb <= la) else acgt[rng.integers(0, 4, size=lb)].copy()
        mode = k % 5
        if mode in (0, 2):
            a[rng.integers(0, la, size=1 + k % 3)] = ord("N")
        if mode in (1, 2):
            b[rng.integers(0, b.size, size=1 + k % 2)] = ord("N")
        if mode == 3:
            a[-1] = ord("N"); b[0] = ord("N")
        nl = b"\n" if k % 7 else b""
        seqs += [a.tobytes() + nl, b.tobytes() + nl]
    buf = np.frombuffer(b"".join(seqs), np.uint8)
    ln = np.array([len(x) for x in seqs], np.int32)
    off = np.concatenate(([0], np.cumsum(ln)[:-1])).astype(np.int64)
    got = gpu_lib.sw_score_flat(buf, off, ln)
    assert got.tolist() == oracle_mod.sw_scores_flat(buf, off, ln).tolist()
