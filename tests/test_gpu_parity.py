"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes), against the CPU oracle,
bit-exact on every FeatureMatch.  Reference call being replaced:
/root/reference/integration/op_cpp/sequential_matching.cc:154 (colmap::MatchSiftFeaturesCPU)."""
import os

import numpy as np
import pytest

from scanner_colmap_b200 import SiftMatcher, synth, sequential_pairs

pytestmark = pytest.mark.gpu

ENGINES = ["tcgen05", "dp4a"]


@pytest.fixture(scope="module")
def oracle(built):
    from oracle import oracle as o
    return o


def _check_pairs(oracle, m, imgs, ids, pairs, **opts):
    got = m.match_pairs(pairs)
    id2k = {i: k for k, i in enumerate(ids)}
    kp = [(id2k[int(a)], id2k[int(b)]) for a, b in pairs]
    want, _ = oracle.match_many(imgs, kp, **opts)
    assert len(got) == len(want)
    for (a, b), g, w in zip(pairs, got, want):
        assert g.dtype == np.uint32 and g.shape[1] == 2
        assert np.array_equal(g, w), f"pair ({a},{b}): {len(g)} vs oracle {len(w)} matches"
    return sum(len(g) for g in got)


@pytest.mark.parametrize("engine", ENGINES)
def test_small_sequential_window(oracle, engine):
    ids = list(range(100, 108))
    sizes = [300, 256, 257, 1, 129, 1000, 127, 640]
    imgs = [synth.make_image(i, n, track_step=32) for i, n in zip(ids, sizes)]
    pairs = sequential_pairs(ids, 4)
    with SiftMatcher(engine=engine) as m:
        m.put_images(ids, imgs)
        total = _check_pairs(oracle, m, imgs, ids, pairs)
    assert total > 0


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("n1,n2", [(0, 5), (5, 0), (0, 0), (1, 1), (127, 129), (128, 256), (129, 127), (255, 257),
                                   (4095, 4097), (4097, 4095)])
def test_ragged_edge_shapes(oracle, engine, n1, n2):
    a = synth.make_image(1, n1, track_step=8)
    b = synth.make_image(2, n2, track_step=8)
    with SiftMatcher(engine=engine) as m:
        got = m.match(a, b)
    want = oracle.match(a, b)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("cross_check", [True, False])
@pytest.mark.parametrize("max_ratio,max_distance", [(0.8, 0.7), (0.95, 1.2), (0.6, 0.5), (1.0, 1.5707964), (0.0, 0.7),
                                                    (0.8, 0.0), (2.0, 3.0)])
def test_option_sweep(oracle, engine, cross_check, max_ratio, max_distance):
    """Thresholds move the integer pre-filter (down to 'every positive score is a candidate'): results stay exact."""
    a = synth.make_image(10, 600, track_step=16, noise=0.25)
    b = synth.make_image(11, 520, track_step=16, noise=0.25)
    with SiftMatcher(engine=engine, max_ratio=max_ratio, max_distance=max_distance, cross_check=cross_check) as m:
        got = m.match(a, b)
    want = oracle.match(a, b, max_ratio=max_ratio, max_distance=max_distance, cross_check=cross_check)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("engine", ENGINES)
def test_adversarial_inputs(oracle, engine):
    rng = np.random.default_rng(7)
    base = synth.make_image(3, 300, track_step=8)
    cases = {
        "all_identical": (np.tile(base[:1], (260, 1)), np.tile(base[:1], (300, 1))),      # every score ties
        "all_zero": (np.zeros((130, 128), np.uint8), np.zeros((140, 128), np.uint8)),     # no positive score
        "zero_vs_real": (np.zeros((130, 128), np.uint8), base),
        "duplicated_rows": (base, np.concatenate([base[:150], base[:150]])),              # tie -> lowest index... and best==second
        "saturating": (rng.integers(200, 256, (257, 128), dtype=np.uint8),                # dot >> 512^2: acos saturates
                       rng.integers(200, 256, (255, 128), dtype=np.uint8)),
        "uniform_random": (rng.integers(0, 256, (300, 128), dtype=np.uint8),
                           rng.integers(0, 256, (280, 128), dtype=np.uint8)),
        "max_value": (np.full((129, 128), 255, np.uint8), np.full((131, 128), 255, np.uint8)),
        "self_match": (base, base.copy()),
    }
    with SiftMatcher(engine=engine) as m:
        for name, (a, b) in cases.items():
            for cc in (True, False):
                m.set_options(cross_check=cc)
                got = m.match(a, b)
                want = oracle.match(a, b, cross_check=cc)
                assert np.array_equal(got, want), f"{name} cross_check={cc}: {len(got)} vs {len(want)}"
    # the self-match case must be non-trivial
    assert len(oracle.match(base, base.copy())) > 100


@pytest.mark.parametrize("engine", ENGINES)
def test_saturated_scores_keep_argmax(oracle, engine):
    """Scores above 512^2 all map to distance 0, but the best index must still be the true arg-max
    (lowest index among equal maxima), so saturated scores may not be clamped before the top-2."""
    a = np.zeros((3, 128), np.uint8)
    a[:, :8] = 250
    b = np.zeros((300, 128), np.uint8)
    b[:, :8] = 100
    b[137, :8] = 255
    b[20, :8] = 254
    with SiftMatcher(engine=engine, cross_check=False, max_ratio=1.5, max_distance=2.0) as m:
        got = m.match(a, b)
    want = oracle.match(a, b, cross_check=False, max_ratio=1.5, max_distance=2.0)
    assert np.array_equal(got, want)


def test_engines_agree_and_cache_semantics(oracle):
    ids = [5, 9, 4000000000]
    imgs = [synth.make_image(i % 1000, n, track_step=16) for i, n in zip(ids, (512, 400, 300))]
    pairs = np.array([[5, 9], [9, 5], [5, 4000000000], [5, 5]], dtype=np.uint32)
    with SiftMatcher() as m:
        m.put_images(ids, imgs)
        assert m.has_image(9) and not m.has_image(10)
        _check_pairs(oracle, m, imgs, ids, pairs)
        # replace an image: later pairs must see the new descriptors
        imgs[1] = synth.make_image(77, 650, track_step=16)
        m.put_image(9, imgs[1])
        _check_pairs(oracle, m, imgs, ids, pairs)
        m.evict_image(9)
        with pytest.raises(Exception):
            m.match_pairs(pairs)
        ptr, n = m.image_device_ptr(5)
        assert ptr != 0 and n == 512
        m.put_image_device(6, ptr, n)  # device-to-device put (the halo path)
        ids2, imgs2 = [5, 6], [imgs[0], imgs[0]]
        _check_pairs(oracle, m, imgs2, ids2, np.array([[5, 6]], dtype=np.uint32))


def test_config1_20x4096_overlap10(oracle):
    """BASELINE.json configs[0]: 20 images x 4096, overlap=10 -> 135 pairs, every FeatureMatch bit-exact."""
    ids = list(range(20))
    imgs = synth.make_images(20, 4096)
    pairs = sequential_pairs(ids, 10)
    assert len(pairs) == 135
    with SiftMatcher(profile=True) as m:
        m.put_images(ids, imgs)
        total = _check_pairs(oracle, m, imgs, ids, pairs)
        t = m.timing()
    assert total > 1000
    assert t["score_launches"] >= 1 and t["ops"] == 135 * 2 * 4096 * 4096 * 128
    assert 1.0 <= t["cta_busy_max_over_mean"] < 1.6        # per-CTA busy time of the persistent score kernel (load balance)


def test_pool_growth_and_many_images(oracle):
    """More descriptor bytes than the initial 64 MiB pool: the pool grows, rows stay valid."""
    ids = list(range(70))
    imgs = [synth.make_image(i, 8192 if i % 2 else 8000, track_step=512) for i in ids]
    with SiftMatcher() as m:
        for i, d in zip(ids, imgs):
            m.put_image(i, d)
        pairs = np.array([[0, 1], [33, 34], [68, 69], [1, 69]], dtype=np.uint32)
        _check_pairs(oracle, m, imgs, ids, pairs)


def test_full_size_properties_8192():
    """BASELINE full size (8192 x 8192) through size-independent properties: symmetry of cross-checked
    matching under swapping the images, idempotence, and self-matching = identity on distinct rows."""
    a = synth.make_image(0, 8192)
    b = synth.make_image(1, 8192)
    with SiftMatcher() as m:
        m.put_images([0, 1], [a, b])
        ab, ba, ab2, aa = m.match_pairs(np.array([[0, 1], [1, 0], [0, 1], [0, 0]], dtype=np.uint32))
    assert len(ab) > 1000
    assert np.array_equal(ab, ab2)
    swapped = ba[:, ::-1]
    swapped = swapped[np.argsort(swapped[:, 0], kind="stable")]
    assert np.array_equal(ab, swapped)            # cross-check makes the relation symmetric
    assert np.all(np.diff(ab[:, 0].astype(np.int64)) > 0)  # ascending idx1, unique
    assert len(np.unique(ab[:, 1])) == len(ab)    # one-to-one
    assert np.array_equal(aa[:, 0], aa[:, 1])     # an image matched with itself pairs each row with itself


def test_survivor_log_overflow_falls_back_exactly(oracle, monkeypatch):
    """With a tiny survivor log the fast (fire-and-forget) insertion overflows and the batch is repeated with
    returning atomics; results must not change.  Also exercises log chunks straddling the capacity."""
    a = synth.make_image(20, 1500, track_step=32)
    b = synth.make_image(21, 1400, track_step=32)
    want = oracle.match(a, b)
    assert len(want) > 300
    for cap in ("0", "256", "700"):
        monkeypatch.setenv("SMB_LOG_CAP", cap)
        with SiftMatcher() as m:
            assert np.array_equal(m.match(a, b), want), f"SMB_LOG_CAP={cap}"
    monkeypatch.delenv("SMB_LOG_CAP")
    dense_a = np.tile(a[:1], (300, 1))
    with SiftMatcher(max_ratio=1.1, max_distance=2.0, cross_check=False) as m:   # every score survives: 90k log entries
        got = m.match(dense_a, b)
    assert np.array_equal(got, oracle.match(dense_a, b, max_ratio=1.1, max_distance=2.0, cross_check=False))


def test_ragged_video_set_reduced(oracle):
    """BASELINE.json configs[3] (ragged 1k-16k descriptors per image, overlap 10) at 1/8 of the sizes so the
    CPU oracle finishes in seconds: every pair of a 24-image window set, bit-exact."""
    sizes = (synth.ragged_sizes(24, lo=128, hi=2048, seed=99)).tolist()
    ids = list(range(500, 524))
    imgs = [synth.make_image(i, n, track_step=24) for i, n in zip(ids, sizes)]
    pairs = sequential_pairs(ids, 10)
    with SiftMatcher() as m:
        m.put_images(ids, imgs)
        total = _check_pairs(oracle, m, imgs, ids, pairs)
    assert total > 500 and min(sizes) < 300 and max(sizes) > 1500


def test_exhaustive_pairs_reduced(oracle):
    """BASELINE.json configs[4] (exhaustive matching) reduced to 12 images x 640 descriptors: all 66 pairs."""
    ids = list(range(12))
    imgs = [synth.make_image(i, 640, track_step=40) for i in ids]
    pairs = np.array([(a, b) for a in ids for b in ids if a < b], dtype=np.uint32)
    with SiftMatcher(max_num_matches=32768) as m:
        m.put_images(ids, imgs)
        _check_pairs(oracle, m, imgs, ids, pairs)


def test_full_size_ragged_properties():
    """Ragged full-size images (up to 16384 descriptors) through size-independent properties."""
    sizes = [16384, 1024, 5000, 9999]
    imgs = [synth.make_image(i, n) for i, n in enumerate(sizes)]
    prs = np.array([[0, 1], [1, 0], [0, 2], [2, 0], [2, 3], [3, 2], [0, 3], [0, 0]], dtype=np.uint32)
    with SiftMatcher() as m:
        m.put_images(range(4), imgs)
        res = m.match_pairs(prs)
        res2 = m.match_pairs(prs)
    for a, b in zip(res, res2):
        assert np.array_equal(a, b)
    for k in (0, 2, 4):
        ab, ba = res[k], res[k + 1][:, ::-1]
        ba = ba[np.argsort(ba[:, 0], kind="stable")]
        assert np.array_equal(ab, ba)
        n1, n2 = sizes[prs[k][0]], sizes[prs[k][1]]
        assert (ab[:, 0] < n1).all() and (ab[:, 1] < n2).all()
    assert len(res[0]) > 50 and np.array_equal(res[7][:, 0], res[7][:, 1])


def test_accumulator_budget_forces_sub_batches(oracle, monkeypatch):
    """A small accumulator budget cuts one call into many sub-batches (score / runner-up / decide per sub-batch,
    all writing into the one pinned result); the concatenated results must be unchanged."""
    ids = list(range(12))
    imgs = [synth.make_image(i, 500 + 37 * i, track_step=24) for i in ids]
    pairs = sequential_pairs(ids, 5)
    monkeypatch.setenv("SMB_ACC_BUDGET", "3000")       # ~2 pairs per sub-batch
    with SiftMatcher() as m:
        m.put_images(ids, imgs)
        total = _check_pairs(oracle, m, imgs, ids, pairs)
    assert total > 300


@pytest.mark.parametrize("waves", ["4", "7"])
def test_waves_pipelined_over_two_streams_are_invisible(oracle, monkeypatch, waves):
    """A call of >= 256 pairs cut into waves: runner-up / decide of wave k run on a second stream underneath the score
    kernel of wave k+1, on alternating accumulator regions and survivor-log halves (chosen adaptively when the
    delivery tail is long; SMB_WAVES forces it).  Per-pair results must not change."""
    ids = list(range(30))
    imgs = [synth.make_image(i, 600 + 31 * (i % 9), track_step=32) for i in ids]
    pairs = sequential_pairs(ids, 12)
    assert len(pairs) >= 256
    monkeypatch.setenv("SMB_WAVES", waves)
    with SiftMatcher(profile=True) as m:
        m.put_images(ids, imgs)
        total = _check_pairs(oracle, m, imgs, ids, pairs)
        t = m.timing()
        assert t["sub_batches"] == int(waves) and t["score_launches"] == int(waves)
        total2 = _check_pairs(oracle, m, imgs, ids, pairs)          # plan reused, regions clean again
        assert m.timing()["plan_uploaded"] == 0
    assert total == total2 > 2000


def test_result_buffer_overflow_is_repeated_exactly(oracle, monkeypatch):
    """decide_kernel writes matches straight into pinned host memory sized from earlier calls; if they do not fit,
    the call is repeated with a worst-case sized buffer (SMB_RESULT_CAP forces a tiny first attempt)."""
    ids = list(range(24))
    imgs = [synth.make_image(i, 700 + 29 * (i % 7), track_step=32) for i in ids]
    pairs = sequential_pairs(ids, 5)                    # 86 pairs
    monkeypatch.setenv("SMB_RESULT_CAP", "100")
    with SiftMatcher(profile=True) as m:
        m.put_images(ids, imgs)
        total = _check_pairs(oracle, m, imgs, ids, pairs)
        total2 = _check_pairs(oracle, m, imgs, ids, pairs)   # the pooled buffer is large enough now, the limit is not
    assert total == total2 > 300


def test_product_library_rejects_the_test_engine(monkeypatch):
    from scanner_colmap_b200 import matcher
    monkeypatch.setenv("SMB_LIB", matcher.LIB_PATH)     # engine="dp4a" would otherwise pick libsmb_test.so
    with pytest.raises(matcher.SmbError):
        SiftMatcher(engine="dp4a")


def test_begin_wait_and_plan_reuse(oracle):
    """smb_match_pairs_begin / smb_result_wait: uploads and evictions may be issued while a call is in flight; an
    identical follow-up call reuses the device-side plan (no plan fetch) and returns identical matches."""
    import torch
    ids = list(range(10))
    imgs = [synth.make_image(i, 1500 + 100 * i, track_step=64) for i in ids]
    extra = [torch.from_numpy(synth.make_image(50 + i, 2000, track_step=64)).pin_memory().numpy() for i in range(2)]
    pairs = sequential_pairs(ids, 4)
    want, _ = oracle.match_many(imgs, [(int(a), int(b)) for a, b in pairs])
    with SiftMatcher(profile=True) as m:
        m.put_images(ids, imgs)
        r = m.match_pairs_begin(pairs)
        m.put_images_async([50, 51], extra)           # queued under the running call
        m.evict_image(50)                             # rows parked until the call has been waited for
        with pytest.raises(Exception):
            m.match_pairs_begin(pairs)                # one call in flight per handle
        r.wait()
        assert m.timing()["plan_uploaded"] == 1
        for k in range(len(pairs)):
            assert np.array_equal(r.matches(k), want[k])
        r.release()
        with m.match_pairs_result(pairs) as r2:
            assert m.timing()["plan_uploaded"] == 0 and m.timing()["total_launches"] == 3   # score, runner-up, decide
            assert all(np.array_equal(r2.matches(k), want[k]) for k in range(len(pairs)))
        # an image of the same size is refreshed IN PLACE (the halo of every multi-GPU step): the stored plan stays valid,
        # the new bytes must be what the next call sees
        imgs2 = list(imgs)
        imgs2[3] = synth.make_image(333, imgs[3].shape[0], track_step=64)
        m.put_image(3, imgs2[3])
        want2, _ = oracle.match_many(imgs2, [(int(a), int(b)) for a, b in pairs])
        with m.match_pairs_result(pairs) as r3:
            assert m.timing()["plan_uploaded"] == 0
            assert all(np.array_equal(r3.matches(k), want2[k]) for k in range(len(pairs)))
        m.put_image(3, imgs[3][:-1])                  # another size: rows move, the plan is rebuilt
        imgs2[3] = imgs[3][:-1]
        want3, _ = oracle.match_many(imgs2, [(int(a), int(b)) for a, b in pairs])
        with m.match_pairs_result(pairs) as r4:
            assert m.timing()["plan_uploaded"] == 1
            assert all(np.array_equal(r4.matches(k), want3[k]) for k in range(len(pairs)))
        m.put_image(3, imgs[3])
        got = m.match_pairs(np.array([[51, 0]], dtype=np.uint32))[0]
        assert np.array_equal(got, oracle.match(extra[1], imgs[0]))
        m.stream_wait_uploads(m.stream)
        m.synchronize()


def test_full_size_pairs_against_the_oracle(oracle):
    """The sizes the benchmark quotes, bit-exact against the CPU oracle (not only through properties): one
    16384 x 16384 pair (configs[4]), eight full-size ragged pairs spanning 1k-16k descriptors (configs[3]) and an
    8192 x 8192 pair (configs[1]/[2]).  ~10 s of oracle time on the box's cores."""
    sizes = [16384, 16384, 8192, 8192] + [1024, 16000, 1100, 13000, 2047, 9000, 4097, 5555, 16384, 1025, 3000, 12288]
    ids = list(range(900, 900 + len(sizes)))
    imgs = [synth.make_image(i, n) for i, n in zip(ids, sizes)]
    pairs = np.array([[ids[0], ids[1]], [ids[2], ids[3]]] + [[ids[4 + 2 * k], ids[5 + 2 * k]] for k in range(6)] +
                     [[ids[5], ids[4]], [ids[12], ids[9]]], dtype=np.uint32)
    with SiftMatcher() as m:
        m.put_images(ids, imgs)
        total = _check_pairs(oracle, m, imgs, ids, pairs, num_threads=os.cpu_count())
    assert total > 3000


def test_halo_style_step_at_8192(oracle):
    """The multi-GPU step at full size on one GPU: own images resident, 'halo' images adopted from device buffers a
    side stream is still filling (smb_put_images_device_async, the NVLink path), own images of a second window still
    crossing PCIe (smb_put_images_async -> waited for inside the score kernel).  A sample of every pair class meets the
    oracle."""
    import torch
    ids = list(range(12))
    imgs = [synth.make_image(i, 8192) for i in ids]
    pinned = [torch.from_numpy(im).pin_memory() for im in imgs]
    side = torch.cuda.Stream()
    pairs = sequential_pairs(ids, 4)                       # 30 pairs: own-own, own-halo, late-upload pairs
    with SiftMatcher(profile=True) as m:
        m.put_images(ids[:6], imgs[:6])                     # resident window
        bufs = [torch.zeros(8192 * 128, dtype=torch.uint8, device="cuda") for _ in range(3)]   # halo 6..8
        torch.cuda.synchronize()
        with torch.cuda.stream(side):
            torch.cuda._sleep(4_000_000)
            for b, hb in zip(bufs, pinned[6:9]):
                b.copy_(hb.reshape(-1), non_blocking=True)
        m.put_images_device_async(ids[6:9], [b.data_ptr() for b in bufs], [8192] * 3, side.cuda_stream)
        m.put_images_async(ids[9:], [p.numpy() for p in pinned[9:]])   # host ticket: its pairs form a later sub-batch
        with m.match_pairs_result(pairs) as r:
            t = m.timing()
            sample = [0, 1, 2, 9, 12, 14, 17, 20, 23, 26, 28, 29]
            want, _ = oracle.match_many(imgs, [(int(pairs[k][0]), int(pairs[k][1])) for k in sample],
                                        num_threads=os.cpu_count())
            for k, w in zip(sample, want):
                assert np.array_equal(r.matches(k), w), f"pair {tuple(pairs[k])}"
        assert t["sub_batches"] == 1 and t["score_launches"] == 1     # one launch: no sub-batch per upload
        m.synchronize()


def test_pairs_are_planned_in_upload_landing_order(oracle):
    """With host uploads in flight the pairs are taken in the order their images land and the score kernel waits, item
    by item, for the upload ticket an item depends on (one launch); the results must still come back per pair in
    the CALLER's order."""
    import torch
    ids = list(range(12))
    imgs = [torch.from_numpy(synth.make_image(i, 2048 + 64 * (i % 5), track_step=128)).pin_memory().numpy() for i in ids]
    pairs = sequential_pairs(ids, 4)
    rng = np.random.default_rng(5)
    for trial in range(3):
        order = rng.permutation(len(pairs)) if trial else np.arange(len(pairs))[::-1]
        scrambled = pairs[order]
        with SiftMatcher() as m:
            for k in range(0, 12, 3):                       # four tickets, none synchronised
                m.put_images_async(ids[k:k + 3], imgs[k:k + 3])
            _check_pairs(oracle, m, imgs, ids, scrambled)
            _check_pairs(oracle, m, imgs, ids, scrambled)   # everything landed (or lands): same answer
            m.synchronize()


def test_device_buffers_adopted_behind_their_producer_stream(oracle):
    """smb_put_images_device_async: the buffers are still being written by work queued on another stream (the
    halo recv in the multi-GPU path); pairs that name them must wait for it, the others need not."""
    import torch
    ids = list(range(6))
    imgs = [synth.make_image(i, 2304 + 128 * i, track_step=128) for i in ids]
    side = torch.cuda.Stream()
    with SiftMatcher() as m:
        m.put_images(ids[:4], imgs[:4])
        bufs = [torch.zeros(im.size, dtype=torch.uint8, device="cuda") for im in imgs[4:]]
        host = [torch.from_numpy(im.reshape(-1)).pin_memory() for im in imgs[4:]]
        torch.cuda.synchronize()
        with torch.cuda.stream(side):
            torch.cuda._sleep(20_000_000)                      # ~10 ms: the buffers are still zero when adopted
            for b, hbuf in zip(bufs, host):
                b.copy_(hbuf, non_blocking=True)
        m.put_images_device_async(ids[4:], [b.data_ptr() for b in bufs], [im.shape[0] for im in imgs[4:]],
                                  side.cuda_stream)
        _check_pairs(oracle, m, imgs, ids, np.array([[4, 5], [0, 1], [3, 4], [1, 2], [2, 5]], dtype=np.uint32))
        m.synchronize()


def test_async_uploads_are_ordered_before_their_pairs(oracle):
    """smb_put_images_async returns at once; a match call must still see every image it names (device-side
    wait on the upload ticket), also across re-puts, evictions and pool growth."""
    import torch
    ids = list(range(8))
    imgs = [torch.from_numpy(synth.make_image(i, 3000 + 100 * i, track_step=96)).pin_memory().numpy() for i in ids]
    with SiftMatcher() as m:
        for k in range(0, 8, 2):
            m.put_images_async(ids[k:k + 2], imgs[k:k + 2])
        _check_pairs(oracle, m, imgs, ids, sequential_pairs(ids, 4))
        # replace two images while nothing is synchronised, then match again
        imgs[3] = torch.from_numpy(synth.make_image(33, 2500, track_step=96)).pin_memory().numpy()
        imgs[4] = torch.from_numpy(synth.make_image(44, 3500, track_step=96)).pin_memory().numpy()
        m.put_images_async([3, 4], [imgs[3], imgs[4]])
        _check_pairs(oracle, m, imgs, ids, np.array([[2, 3], [3, 4], [4, 5]], dtype=np.uint32))
        m.evict_image(0)
        big = [torch.from_numpy(synth.make_image(100 + i, 8192, track_step=512)).pin_memory().numpy() for i in range(70)]
        m.put_images_async(range(100, 170), big)        # > 64 MiB: the pool grows underneath pending uploads
        _check_pairs(oracle, m, [big[0], big[1], big[69], imgs[5]], [100, 101, 169, 5],
                     np.array([[100, 101], [169, 5]], dtype=np.uint32))
        m.synchronize()
