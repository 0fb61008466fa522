"""Multi-GPU paths of libagx inside ONE process (the library's own dispatcher: one host thread + one
stream per GPU).  Needs >= 2 GPUs; skipped on a single-GPU box."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def multi(agx):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    cap = agx.capi
    cap.shutdown()
    n = cap.init(0)              # every visible GPU
    assert n == torch.cuda.device_count()
    yield cap, n
    cap.shutdown()


def test_sw_batch_is_sharded_over_all_gpus(agx, multi, oracle_mod):
    cap, n = multi
    rng = np.random.default_rng(0)
    data = agx.synth.sw_random_file(rng, 4000, 1, 260, alphabet=b"ACGT")
    inp = agx.formats.parse_sw(data)
    got = cap.sw_score_flat(inp.buf, inp.off, inp.len)
    cap.shutdown()
    cap.init(1)
    one = cap.sw_score_flat(inp.buf, inp.off, inp.len)
    cap.shutdown()
    cap.init(0)
    assert np.array_equal(got, one)
    pick = rng.choice(inp.n_pairs, size=300, replace=False)
    idx = np.stack([2 * pick, 2 * pick + 1], axis=1).reshape(-1)
    assert got[pick].tolist() == oracle_mod.sw_scores_flat(inp.buf, inp.off[idx], inp.len[idx]).tolist()


def test_pairhmm_reads_are_sharded_over_all_gpus(agx, multi, oracle_mod):
    cap, n = multi
    inp = agx.synth.pairhmm_batches(7, 30, 4, seed=33)
    args = (inp.buf, inp.read_field_off, inp.read_len, inp.hap_off, inp.hap_len, inp.batch_read_start, inp.batch_hap_start)
    got = cap.pairhmm_forward_flat(*args)
    want = oracle_mod.pairhmm_flat(inp)
    assert np.max(np.abs(got - want) / np.abs(want)) <= 1e-5
    # a single batch is split by reads too
    one_batch = agx.synth.pairhmm_batches(1, 64, 3, seed=34)
    args = (one_batch.buf, one_batch.read_field_off, one_batch.read_len, one_batch.hap_off, one_batch.hap_len,
            one_batch.batch_read_start, one_batch.batch_hap_start)
    got = cap.pairhmm_forward_flat(*args)
    want = oracle_mod.pairhmm_flat(one_batch)
    assert np.max(np.abs(got - want) / np.abs(want)) <= 1e-5


@pytest.mark.parametrize("n,related", [(20000, True), (24000, False)])
def test_long_alignment_column_stripes_across_gpus(agx, multi, oracle_mod, n, related):
    """One pair, columns striped over the GPUs, boundary columns exchanged as NVLink peer stores."""
    cap, ngpu = multi
    data = agx.synth.sw_long_pair(n, seed=n + 1, related=related)
    inp = agx.formats.parse_sw(data, line_buf=1 << 30)
    got = cap.sw_score_flat(inp.buf, inp.off, inp.len)
    want = oracle_mod.sw_scores_flat(inp.buf, inp.off, inp.len)
    assert got.tolist() == want.tolist()


@pytest.mark.parametrize("k", [4, 8, 16, 32])
def test_long_alignment_every_stripe_width_across_gpus(agx, multi, oracle_mod, k, monkeypatch):
    """Stripe widths 128..1024 columns: the GPU cut must sit on a stripe boundary for every one."""
    cap, ngpu = multi
    monkeypatch.setenv("AGX_LONG_K", str(k))
    data = agx.synth.sw_long_pair(21000 + 37 * k, seed=k, related=True)
    inp = agx.formats.parse_sw(data, line_buf=1 << 30)
    got = cap.sw_score_flat(inp.buf, inp.off, inp.len)
    want = oracle_mod.sw_scores_flat(inp.buf, inp.off, inp.len)
    assert got.tolist() == want.tolist()


def test_long_alignment_multi_gpu_equals_single_gpu_200kbp(agx, multi):
    cap, ngpu = multi
    data = agx.synth.sw_long_pair(200_000, seed=9, related=True)
    inp = agx.formats.parse_sw(data, line_buf=1 << 30)
    many = int(cap.sw_score_flat(inp.buf, inp.off, inp.len)[0])
    cap.shutdown()
    cap.init(1)
    one = int(cap.sw_score_flat(inp.buf, inp.off, inp.len)[0])
    cap.shutdown()
    cap.init(0)
    assert many == one and many > 100_000


@pytest.mark.parametrize("seed,header_frac,trail", [(1, 1.0, True), (2, 1.0, False), (3, 0.7, True), (4, 1.3, True)])
def test_sw_file_image_is_cut_into_one_range_per_gpu(agx, multi, seed, header_frac, trail):
    """sw_score_file_image on several GPUs: ranges that start at an odd chunk index (their first chunk pairs
    with the previous range's last one), lines split by the fgets() buffer, a header that asks for fewer or
    more lines than the file holds, a dangling last line."""
    cap, n = multi
    rng = np.random.default_rng(seed)
    alpha = np.frombuffer(b"ACGT", np.uint8)
    n_lines = 14000 * n + 1                                      # >= 4 MiB per GPU; odd: the last line dangles when all are asked for
    lens = rng.integers(1, 700, size=n_lines)
    lens[rng.integers(0, n_lines, size=40)] = rng.integers(1000, 2600, size=40)   # split by the 1000-byte buffer
    lines = [alpha[rng.integers(0, 4, size=int(l))].tobytes() for l in lens]
    header = int(header_frac * n_lines)
    data = str(header).encode() + b"\n" + b"\n".join(lines) + (b"\n" if trail else b"")
    assert len(data) >= n * (4 << 20)
    host = agx.formats.parse_sw(data)
    want = cap.sw_score_flat(host.buf, host.off, host.len)
    scores, hdr, dangling = cap.sw_score_file_image(data)
    assert hdr == header and dangling == host.dangling
    assert scores.tolist() == want.tolist()


def test_pairhmm_file_image_is_cut_into_one_range_per_gpu(agx, multi, oracle_mod):
    """pairhmm_forward_file_image on several GPUs: ranges of whole batches, cut at header lines and verified to be
    batch boundaries; results in file order, a truncated last batch reported like on one GPU."""
    cap, n = multi
    inp = agx.synth.pairhmm_batches(60 * n, 200, 5, seed=77, unrelated_frac=0.002)
    assert inp.buf.size >= n * (8 << 20)
    vals, batch_pairs, incomplete = cap.pairhmm_forward_file_image(inp.buf)
    assert incomplete == 0 and batch_pairs.tolist() == [1000] * (60 * n)
    flat = cap.pairhmm_forward_flat(inp.buf, inp.read_field_off, inp.read_len, inp.hap_off, inp.hap_len,
                                    inp.batch_read_start, inp.batch_hap_start)
    assert np.array_equal(vals, flat, equal_nan=True)
    want = oracle_mod.pairhmm_flat(inp, limit=300)
    fin = np.isfinite(want)
    assert np.max(np.abs(vals[:300][fin] - want[fin]) / np.abs(want[fin])) <= 1e-5
    # the file ends inside its last batch
    lines = bytes(inp.buf).split(b"\n")[:-1]
    cut = b"\n".join(lines[:-3]) + b"\n"
    vals2, batch_pairs2, incomplete2 = cap.pairhmm_forward_file_image(cut)
    assert incomplete2 == 2 and batch_pairs2.size == 60 * n - 1 and np.array_equal(vals2, vals[:vals2.size], equal_nan=True)


def test_pairhmm_file_image_with_a_header_shaped_line_inside_a_batch_falls_back(agx, multi):
    """A haplotype line that looks like a header near a cut point: the range check fails and the first GPU takes the
    whole image (the reference reads that line as a haplotype)."""
    cap, n = multi
    inp = agx.synth.pairhmm_batches(60 * n, 200, 5, seed=78)
    data = bytearray(bytes(inp.buf))
    # overwrite the first haplotype line after the middle of the file with "12 34" padded by spaces... it must
    # keep its length; a haplotype of digits and blanks is legal input for the reference (bytes are bytes)
    mid = len(data) // 2
    h = int(inp.hap_off[np.searchsorted(inp.hap_off, mid)])
    L = int(inp.hap_len[np.searchsorted(inp.hap_off, mid)])
    data[h:h + L] = b"7" * (L - 3) + b" 55"
    one_gpu, bp1, inc1 = cap.pairhmm_forward_file_image(bytes(data))
    cap.shutdown()
    cap.init(1)
    ref, bp0, inc0 = cap.pairhmm_forward_file_image(bytes(data))
    cap.shutdown()
    cap.init(0)
    assert inc1 == inc0 == 0 and bp1.tolist() == bp0.tolist() and np.array_equal(one_gpu, ref, equal_nan=True)


def test_sharded_device_entry_points(agx, multi):
    """sw_score_shards_device / pairhmm_forward_shards_device: one device-resident shard per GPU in one call."""
    import torch
    cap, n = multi
    sw = agx.synth.sw_uniform_pairs(20000, 150, seed=5)
    want = cap.sw_score_flat(sw.buf, sw.off, sw.len)
    shards, keep = [], []
    for k in range(n):
        p0, p1 = 20000 * k // n, 20000 * (k + 1) // n
        b0, b1 = int(sw.off[2 * p0]), int(sw.off[2 * p1 - 1] + sw.len[2 * p1 - 1])
        dev = torch.device("cuda", k)
        t = [torch.from_numpy(x).to(dev) for x in (sw.buf[b0:b1], sw.off[2 * p0:2 * p1] - b0, sw.len[2 * p0:2 * p1])]
        out = torch.zeros(p1 - p0, dtype=torch.int32, device=dev)
        keep.append((t, out))
        shards.append(cap.SwShard(k, t[0].data_ptr(), t[0].numel(), t[1].data_ptr(), t[2].data_ptr(), p1 - p0, out.data_ptr()))
    cap.sw_score_shards_device(shards)
    assert np.array_equal(np.concatenate([o.cpu().numpy() for _, o in keep]), want)
    with pytest.raises(agx.capi.AgxError):
        cap.sw_score_shards_device([shards[0], shards[0]])          # one shard per device


def test_alignments_are_sharded_over_all_gpus(agx, multi, oracle_mod):
    """sw_ends_batch_flat / sw_align_batch_flat with several GPUs bound: ranges of pairs per GPU, CIGAR runs and
    offsets concatenated in pair order; equal to the one-GPU result and (a sample) to the oracle."""
    cap, n = multi
    rng = np.random.default_rng(5)
    data = agx.synth.sw_random_file(rng, 5000, 1, 300, alphabet=b"ACGT", related_frac=0.7)
    inp = agx.formats.parse_sw(data)
    got = cap.sw_align_flat(inp.buf, inp.off, inp.len)
    got_e = cap.sw_ends_flat(inp.buf, inp.off, inp.len)
    cap.shutdown()
    cap.init(1)
    one = cap.sw_align_flat(inp.buf, inp.off, inp.len)
    one_e = cap.sw_ends_flat(inp.buf, inp.off, inp.len)
    cap.shutdown()
    cap.init(0)
    for g, w in zip(got + got_e, one + one_e):
        assert np.array_equal(g, w)
    scores, coords, coff, cig = got
    raw = inp.buf.tobytes()
    for p in rng.choice(inp.n_pairs, size=300, replace=False).tolist():
        a = raw[inp.off[2 * p]:inp.off[2 * p] + inp.len[2 * p]]
        b = raw[inp.off[2 * p + 1]:inp.off[2 * p + 1] + inp.len[2 * p + 1]]
        ws, wc, wg = oracle_mod.sw_align(a, b)
        assert (ws, list(wc), wg) == (int(scores[p]), coords[p].tolist(), cig[coff[p]:coff[p + 1]].tolist())


@pytest.mark.parametrize("n,related", [(20000, True), (22000, False)])
def test_long_alignment_end_cell_across_gpus(agx, multi, oracle_mod, n, related):
    """sw_ends_batch_flat on a whole-GPU pair with several GPUs bound: column stripes over the GPUs, the END CELL is
    the maximum of the per-GPU 64-bit keys (global columns); both orientations of the pair."""
    cap, ngpu = multi
    data = agx.synth.sw_long_pair(n, seed=n + 3, related=related)
    inp = agx.formats.parse_sw(data, line_buf=1 << 30)
    raw = inp.buf.tobytes()
    a = raw[inp.off[0]:inp.off[0] + inp.len[0]]
    b = raw[inp.off[1]:inp.off[1] + inp.len[1]][:-1 - 7]          # a little shorter, no newline: line 2 becomes sx
    for x, y in ((a, b), (b, a)):
        buf = np.frombuffer(x + y, np.uint8)
        off = np.array([0, len(x)], np.int64)
        ln = np.array([len(x), len(y)], np.int32)
        scores, ends = cap.sw_ends_flat(buf, off, ln)
        ws, we = oracle_mod.sw_ends(x, y)
        assert (int(scores[0]), tuple(ends[0].tolist())) == (ws, we)
