"""PairHMM parity on the GPU, through the C ABI, against the recorded reference outputs and the CPU
oracle.  Tolerance (BASELINE.json north_star): |gpu - ref| <= 1e-5 * |ref| on the log10 likelihood."""
import numpy as np
import pytest

from conftest import GOLDEN, read_golden

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5          # north_star: relative tolerance on log10 likelihoods
PRINT_EPS = 5.1e-7      # the reference prints %f: its recorded outputs carry +-5e-7 rounding


def _run_flat(gpu_lib, inp):
    return gpu_lib.pairhmm_forward_flat(inp.buf, inp.read_field_off, inp.read_len, inp.hap_off, inp.hap_len,
                                        inp.batch_read_start, inp.batch_hap_start)


def _rel_err(got, want):
    return np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300))


GOLDENS = ["pairhmm_test", "pairhmm_10s", "pairhmm_synth_small", "pairhmm_synth_cfg4", "pairhmm_synth_long",
           "pairhmm_synth_tall"]


@pytest.mark.parametrize("name", GOLDENS)
def test_golden_vs_reference_output(agx, gpu_lib, name):
    inp = agx.formats.parse_pairhmm(read_golden(f"{name}.in"))
    got = _run_flat(gpu_lib, inp)
    ref = np.array([float(x) for x in (GOLDEN / f"{name}.pairhmm_antidiag.out").read_text().split()])
    assert got.shape == ref.shape and np.all(np.isfinite(got))
    assert np.all(np.abs(got - ref) <= REL_TOL * np.abs(ref) + PRINT_EPS)


@pytest.mark.parametrize("name", GOLDENS)
def test_golden_vs_oracle_double(agx, gpu_lib, oracle_mod, name):
    inp = agx.formats.parse_pairhmm(read_golden(f"{name}.in"))
    got = _run_flat(gpu_lib, inp)
    want = oracle_mod.pairhmm_flat(inp)
    assert _rel_err(got, want) <= REL_TOL


def test_committed_golden_value(agx, gpu_lib):
    inp = agx.formats.parse_pairhmm(read_golden("pairhmm_test.in"))
    got = _run_flat(gpu_lib, inp)
    assert "%f" % got[0] == "-4.485565"     # pairHMM/test_set/test.out


def test_fp64_kernel_keeps_reference_operation_order(agx, gpu_lib, oracle_mod):
    inp = agx.formats.parse_pairhmm(read_golden("pairhmm_10s.in"))
    gpu_lib.set_pairhmm_force_fp64(True)
    try:
        got = _run_flat(gpu_lib, inp)
    finally:
        gpu_lib.set_pairhmm_force_fp64(False)
    want = oracle_mod.pairhmm_flat(inp)
    assert _rel_err(got, want) <= 1e-13
    ref_text = (GOLDEN / "pairhmm_10s.pairhmm_antidiag.out").read_text().split()
    assert ["%f" % v for v in got] == ref_text


def test_pointer_array_entry_point(gpu_lib, oracle_mod):
    reads = [(b"ACGTACGTAC", b"IIIIIIIIII", b"IIIIIIIIII", b"IIIIIIIIII", b"++++++++++"),
             (b"ACGTNCGTAC", b"5555555555", b"IIIIDIIIII", b"HHHHCHHHHH", b"++++++++++"),
             (b"T", b"#", b"I", b"I", b"+")]
    haps = [b"ACGTACGTAC", b"TTACGTACGTACGG", b"ACGTANGTAC", b"G"]
    got = gpu_lib.pairhmm_forward_batch(reads, haps)
    assert got.shape == (3, 4)
    for r, rd in enumerate(reads):
        for h, hp in enumerate(haps):
            want = oracle_mod.pairhmm_forward(rd, hp)
            assert abs(got[r, h] - want) <= REL_TOL * abs(want)


def test_unrelated_pairs_take_the_fp64_rescue(agx, gpu_lib, oracle_mod):
    # uniform random reads against unrelated haplotypes: likelihoods far below FP32 range
    inp = agx.synth.pairhmm_batches(2, 16, 3, seed=21, unrelated_frac=1.0, read_len=(150, 250))
    want = oracle_mod.pairhmm_flat(inp)
    assert want.min() < -70
    got = _run_flat(gpu_lib, inp)
    assert np.all(np.isfinite(got))
    assert _rel_err(got, want) <= REL_TOL


def test_extreme_qualities(gpu_lib, oracle_mod):
    # Phred 0 ('!') through 93 ('~') in every quality track, N in read and haplotype
    rng = np.random.default_rng(8)
    L = 120
    bases = bytes(np.frombuffer(b"ACGTN", np.uint8)[rng.integers(0, 5, size=L)])
    hap = bytes(np.frombuffer(b"ACGTN", np.uint8)[rng.integers(0, 5, size=200)])
    reads = []
    for _ in range(12):
        q = bytes(rng.integers(33, 127, size=L).astype(np.uint8))
        qi = bytes(rng.integers(43, 127, size=L).astype(np.uint8))
        qd = bytes(rng.integers(43, 127, size=L).astype(np.uint8))
        qg = bytes(rng.integers(36, 127, size=L).astype(np.uint8))
        reads.append((bases, q, qi, qd, qg))
    got = gpu_lib.pairhmm_forward_batch(reads, [hap, hap[:50]])
    for r, rd in enumerate(reads):
        for h, hp in enumerate([hap, hap[:50]]):
            want = oracle_mod.pairhmm_forward(rd, hp)
            assert np.isfinite(got[r, h]) == np.isfinite(want)
            if np.isfinite(want):
                assert abs(got[r, h] - want) <= REL_TOL * abs(want)


def test_gatk_mode_is_separate(agx, gpu_lib, oracle_mod):
    inp = agx.synth.pairhmm_batches(2, 10, 3, seed=5)
    gpu_lib.set_pairhmm_gatk_mode(True)
    try:
        got = _run_flat(gpu_lib, inp)
    finally:
        gpu_lib.set_pairhmm_gatk_mode(False)
    assert _rel_err(got, oracle_mod.pairhmm_flat(inp, gatk=True)) <= REL_TOL
    assert _rel_err(_run_flat(gpu_lib, inp), oracle_mod.pairhmm_flat(inp)) <= REL_TOL


def test_argument_ranges(gpu_lib):
    with pytest.raises(gpu_lib.AgxError) as e:
        gpu_lib.pairhmm_forward_batch([(b"", b"", b"", b"", b"")], [b"ACGT"])
    assert e.value.code == -5


def test_device_resident_entry_point(agx, gpu_lib, oracle_mod):
    import torch
    inp = agx.synth.pairhmm_batches(6, 40, 5, seed=12)
    dev = torch.device("cuda:0")
    nb = inp.n_batches
    nh_b = np.diff(inp.batch_hap_start)
    read_batch = np.repeat(np.arange(nb, dtype=np.int32), np.diff(inp.batch_read_start))
    out_off = np.concatenate(([0], np.cumsum(nh_b[read_batch])))[:-1].astype(np.int64)
    n_pairs = inp.n_pairs
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_buf, d_rfo, d_rl = t(inp.buf.copy()), t(inp.read_field_off.reshape(-1)), t(inp.read_len)
    d_rb, d_roo, d_ho, d_hl, d_bhs = t(read_batch), t(out_off), t(inp.hap_off), t(inp.hap_len), t(inp.batch_hap_start)
    d_out = torch.zeros(n_pairs, dtype=torch.float64, device=dev)
    gpu_lib.pairhmm_forward_device(0, d_buf.data_ptr(), d_buf.numel(), d_rfo.data_ptr(), d_rl.data_ptr(),
                                   d_rb.data_ptr(), d_roo.data_ptr(), inp.read_len.size, d_ho.data_ptr(),
                                   d_hl.data_ptr(), inp.hap_len.size, d_bhs.data_ptr(), nb, n_pairs,
                                   d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    assert np.allclose(got, _run_flat(gpu_lib, inp), rtol=0, atol=0)
    want = oracle_mod.pairhmm_flat(inp, limit=300)
    assert _rel_err(got[:300], want) <= REL_TOL


# ---------------------------------------------------------------- device-side file parser (pairhmm_forward_file_image)
@pytest.mark.parametrize("name", GOLDENS)
def test_file_image_entry_point_equals_host_parsed_call(agx, gpu_lib, name):
    """Batch walk, field splitting and index arrays built on the GPU give the values of the host-parsed call."""
    data = read_golden(f"{name}.in")
    inp = agx.formats.parse_pairhmm(data)
    vals, batch_pairs, incomplete = gpu_lib.pairhmm_forward_file_image(data)
    assert incomplete == 0
    assert batch_pairs.tolist() == (np.diff(inp.batch_read_start) * np.diff(inp.batch_hap_start)).tolist()
    assert np.array_equal(vals, _run_flat(gpu_lib, inp))
    ref = np.array([float(x) for x in (GOLDEN / f"{name}.pairhmm_antidiag.out").read_text().split()])
    assert np.all(np.abs(vals - ref) <= REL_TOL * np.abs(ref) + PRINT_EPS)


def _small_batches(agx, seed, n_batches=7):
    inp = agx.synth.pairhmm_batches(n_batches, 9, 3, seed=seed)
    return inp, bytes(inp.buf)


def test_file_image_irregular_headers_follow_the_reference_walk(agx, gpu_lib):
    """Headers that sscanf("%d %d") accepts but that are not plain 'digits blank digits' lines (leading
    blanks, '+', trailing text), an empty batch and a last line without newline."""
    inp, data = _small_batches(agx, 3)
    lines = data.split(b"\n")
    assert lines[-1] == b""
    lines = lines[:-1]
    heads = [i for i, l in enumerate(lines) if len(l.split()) == 2 and l.split()[0].isdigit()]
    assert len(heads) == 7
    a, b = lines[heads[1]].split()
    lines[heads[1]] = b"  " + a + b"\t " + b
    a, b = lines[heads[2]].split()
    lines[heads[2]] = b"+" + a + b" " + b + b" trailing words"
    lines.insert(heads[3], b"0 0")                      # a batch with nothing in it
    mod = b"\n".join(lines)                             # and no newline after the last haplotype
    host = agx.formats.parse_pairhmm(mod)
    assert host.n_batches == 8
    vals, batch_pairs, incomplete = gpu_lib.pairhmm_forward_file_image(mod)
    assert incomplete == 0
    assert batch_pairs.tolist() == (np.diff(host.batch_read_start) * np.diff(host.batch_hap_start)).tolist()
    assert np.array_equal(vals, _run_flat(gpu_lib, host))
    assert np.array_equal(vals, _run_flat(gpu_lib, inp))


@pytest.mark.parametrize("cut,code", [("reads", 2), ("haps", 2)])
def test_file_image_truncated_last_batch(agx, gpu_lib, cut, code):
    """The reference's haplotype cursor runs ahead of its read cursor (antidiagsPairHMM.c:388-396), so a file
    that ends anywhere inside a batch with haplotypes makes it say "Error reading haplotypes." (code 2)."""
    inp, data = _small_batches(agx, 5, n_batches=4)
    lines = data.split(b"\n")[:-1]
    # batch = 1 header + 9 reads + 3 haplotypes = 13 lines; cut inside the last batch
    keep = 3 * 13 + (1 + 4 if cut == "reads" else 1 + 9 + 1)
    mod = b"\n".join(lines[:keep]) + b"\n"
    vals, batch_pairs, incomplete = gpu_lib.pairhmm_forward_file_image(mod)
    assert incomplete == code and batch_pairs.tolist() == [27, 27, 27]
    assert np.array_equal(vals, _run_flat(gpu_lib, inp)[:81])


def test_file_image_rejects_lines_beyond_the_reference_line_buffer(agx, gpu_lib):
    hap = b"ACGT" * 1300                                 # 5200 > 5000
    read = (b"ACGTACGTAC", b"IIIIIIIIII", b"NNNNNNNNNN", b"NNNNNNNNNN", b"++++++++++")
    data = agx.formats.write_pairhmm([([read], [hap])])
    with pytest.raises(gpu_lib.AgxError) as e:
        gpu_lib.pairhmm_forward_file_image(data)
    assert "line buffer" in str(e.value)
    with pytest.raises(gpu_lib.AgxError):
        gpu_lib.pairhmm_forward_file_image(b"1 1\nAC II\nACGT\n")      # fewer than five fields


def test_file_image_config4_shape(agx, gpu_lib):
    inp = agx.synth.pairhmm_batches(40, 200, 5, seed=17)
    vals, batch_pairs, incomplete = gpu_lib.pairhmm_forward_file_image(inp.buf)
    assert incomplete == 0 and batch_pairs.tolist() == [1000] * 40
    assert np.array_equal(vals, _run_flat(gpu_lib, inp))


@pytest.mark.parametrize("segment", [700, 5000, 1 << 16])
def test_file_image_streaming_segments(agx, gpu_lib, segment, monkeypatch):
    """Upload segments far smaller than a batch, batches straddling them, a truncated last batch."""
    inp = agx.synth.pairhmm_batches(12, 14, 3, seed=segment)
    data = bytes(inp.buf)
    want = _run_flat(gpu_lib, inp)
    monkeypatch.setenv("AGX_HMM_IMAGE_SEGMENT", str(segment))
    vals, batch_pairs, incomplete = gpu_lib.pairhmm_forward_file_image(data)
    assert incomplete == 0 and batch_pairs.tolist() == [42] * 12
    assert np.array_equal(vals, want)
    lines = data.split(b"\n")[:-1]
    cut = b"\n".join(lines[: 11 * 18 + 6]) + b"\n"                  # batch = 1 + 14 + 3 lines; EOF inside the reads
    vals, batch_pairs, incomplete = gpu_lib.pairhmm_forward_file_image(cut)
    assert incomplete == 2 and batch_pairs.tolist() == [42] * 11
    assert np.array_equal(vals, want[: 11 * 42])
    # a batch without haplotypes that runs out of reads is the one case of "Error reading reads." (code 1)
    vals, batch_pairs, incomplete = gpu_lib.pairhmm_forward_file_image(b"\n".join(lines[:18]) + b"\n3 0\n" + lines[1] + b"\n")
    assert incomplete == 1 and batch_pairs.tolist() == [42]


def test_config4_shape_properties(agx, gpu_lib, oracle_mod, monkeypatch):
    """BASELINE config 4 shape (200 reads x 5 haplotypes per batch, reads 100-250, haplotypes 200-500),
    100 000 pairs: size-independent properties + oracle sample."""
    inp = agx.synth.pairhmm_batches(100, 200, 5, seed=21, unrelated_frac=0.001)
    got = _run_flat(gpu_lib, inp)
    assert got.shape == (100_000,) and np.all(np.isfinite(got))
    # two reads per warp (packed f32x2) and one read per warp give bit-identical values
    monkeypatch.setenv("AGX_PAIRHMM_NO_DUO", "1")
    assert np.array_equal(_run_flat(gpu_lib, inp), got)
    monkeypatch.delenv("AGX_PAIRHMM_NO_DUO")
    # the FP64 kernel (reference operation order) agrees within the stated tolerance everywhere
    gpu_lib.set_pairhmm_force_fp64(True)
    try:
        f64 = _run_flat(gpu_lib, inp)
    finally:
        gpu_lib.set_pairhmm_force_fp64(False)
    assert _rel_err(got, f64) <= REL_TOL
    # order independence: reversing the reads of every batch (other partners in the two-read kernel) permutes the output
    rs = inp.batch_read_start
    perm = np.concatenate([np.arange(rs[b], rs[b + 1])[::-1] for b in range(inp.n_batches)])
    rev = _run_flat(gpu_lib, type(inp)(inp.buf, inp.read_field_off[perm], inp.read_len[perm], inp.hap_off, inp.hap_len,
                                        inp.batch_read_start, inp.batch_hap_start))
    nh = 5
    assert np.array_equal(rev.reshape(-1, nh), got.reshape(-1, nh)[perm])
    # a read listed twice scores the same in both slots
    dup = np.repeat(np.arange(200), 2)[:200]
    one = type(inp)(inp.buf, inp.read_field_off[dup], inp.read_len[dup], inp.hap_off[:5], inp.hap_len[:5],
                    np.array([0, 200], np.int64), np.array([0, 5], np.int64))
    d = _run_flat(gpu_lib, one).reshape(200, nh)
    assert np.array_equal(d[0::2], d[1::2])
    # oracle on a sample
    want = oracle_mod.pairhmm_flat(inp, limit=400)
    assert _rel_err(got[:400], want) <= REL_TOL


def test_device_entry_point_with_reads_in_any_order(agx, gpu_lib):
    """d_read_batch need not be sorted: the read pairing then steps aside (one read per warp) and the values,
    written through d_read_out_off, do not change."""
    import torch
    inp = agx.synth.pairhmm_batches(8, 30, 4, seed=31)
    want = _run_flat(gpu_lib, inp)
    dev = torch.device("cuda:0")
    nb = inp.n_batches
    nh_b = np.diff(inp.batch_hap_start)
    read_batch = np.repeat(np.arange(nb, dtype=np.int32), np.diff(inp.batch_read_start))
    out_off = np.concatenate(([0], np.cumsum(nh_b[read_batch])))[:-1].astype(np.int64)
    perm = np.random.default_rng(1).permutation(read_batch.size)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_buf, d_rfo, d_rl = t(inp.buf.copy()), t(inp.read_field_off[perm].reshape(-1)), t(inp.read_len[perm])
    d_rb, d_roo, d_ho, d_hl, d_bhs = t(read_batch[perm]), t(out_off[perm]), t(inp.hap_off), t(inp.hap_len), t(inp.batch_hap_start)
    d_out = torch.zeros(inp.n_pairs, dtype=torch.float64, device=dev)
    gpu_lib.pairhmm_forward_device(0, d_buf.data_ptr(), d_buf.numel(), d_rfo.data_ptr(), d_rl.data_ptr(),
                                   d_rb.data_ptr(), d_roo.data_ptr(), inp.read_len.size, d_ho.data_ptr(),
                                   d_hl.data_ptr(), inp.hap_len.size, d_bhs.data_ptr(), nb, inp.n_pairs,
                                   d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), want)


def test_file_image_header_counts_are_inherited(agx, gpu_lib, oracle_mod):
    """A header with fewer than two integers keeps the previous batch's count(s) (antidiagsPairHMM.c:345-346, :378);
    a blank line after the last batch is then one more, truncated, batch: "Error reading haplotypes."."""
    from test_formats import _inherited_header_file
    data = _inherited_header_file(agx)
    host = agx.formats.parse_pairhmm(data)
    vals, batch_pairs, incomplete = gpu_lib.pairhmm_forward_file_image(data)
    assert incomplete == 0 and batch_pairs.tolist() == [27] * 4
    want = oracle_mod.pairhmm_flat(host)
    assert np.max(np.abs(vals - want) / np.abs(want)) <= 1e-5
    vals2, batch_pairs2, incomplete2 = gpu_lib.pairhmm_forward_file_image(_inherited_header_file(agx, trailing_blank=True))
    assert incomplete2 == 2 and batch_pairs2.tolist() == [27] * 4 and np.array_equal(vals2, vals)


@pytest.mark.parametrize("parts", [2, 3, 7])
def test_shard_parts_on_two_lanes_give_the_same_result(agx, gpu_lib, oracle_mod, parts, monkeypatch):
    """a large shard is cut into parts that alternate between two lanes (index arrays of part k+1 built and uploaded
    under the kernels of part k, FP64 rescue without a host round trip): any number of parts -- cuts inside batches
    included -- gives the one-part answer bit for bit, for pageable and for pinned result arrays"""
    import torch
    inp = agx.synth.pairhmm_batches(9, 31, 4, seed=77, unrelated_frac=0.05)
    args = (inp.buf, inp.read_field_off, inp.read_len, inp.hap_off, inp.hap_len, inp.batch_read_start, inp.batch_hap_start)
    monkeypatch.setenv("AGX_HMM_PARTS", "1")
    want = gpu_lib.pairhmm_forward_flat(*args)
    monkeypatch.setenv("AGX_HMM_PARTS", str(parts))
    got = gpu_lib.pairhmm_forward_flat(*args)
    pinned = torch.empty(want.size, dtype=torch.float64).pin_memory().numpy()
    got_pinned = gpu_lib.pairhmm_forward_flat(*args, out=pinned)
    assert np.array_equal(got, want, equal_nan=True) and np.array_equal(got_pinned, want, equal_nan=True)
    ref = oracle_mod.pairhmm_flat(inp)
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), fin) and _rel_err(got[fin], ref[fin]) <= REL_TOL
