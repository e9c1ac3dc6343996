"""GPU tests at (or near) BASELINE.json's full sizes.  The oracle cannot follow at these sizes; results are pinned by
size-independent properties instead: a batch signed by the engine's synthetic-data helper verifies item by item, seeded
corruptions are found exactly, aggregates verify under aggregate keys, wrong aggregates do not."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def eng():
    import blsful_b200 as B
    e = B.Engine([0])
    yield e
    e.close()


@pytest.fixture(scope="module")
def batch_1m(eng):
    import bench
    n = 1_000_000
    return (n,) + bench.synth_batch(eng, n, seed=77)


def test_cfg2_one_million_signatures_accept_and_exact_rejections(eng, batch_1m):
    n, pks, sigs, msgs, off = batch_1m
    st = eng.verify_batch_packed(2, 0, pks, sigs, msgs, off)
    assert st.shape == (n,) and int(st.max()) == 0
    rng = np.random.default_rng(3)
    bad = np.sort(rng.choice(n, 37, replace=False))
    s2 = sigs.copy().reshape(n, 96)
    s2[bad] = s2[(bad + 1) % n]          # another item's valid signature: a subgroup point, but the wrong one
    p2 = pks.copy().reshape(n, 48)
    p2[12345] = 0                        # undecodable public key
    s2[54321] = 0
    s2[54321, 0] = 0xC0                  # identity signature
    st = eng.verify_batch_packed(2, 0, p2.reshape(-1), s2.reshape(-1), msgs, off)
    want = np.zeros(n, dtype=np.uint8)
    want[bad] = 1
    want[12345] = 4
    want[54321] = 2
    assert np.array_equal(st, want)
    # idempotence: the same call again gives the same vector (fresh RLC salt, same statuses)
    assert np.array_equal(eng.verify_batch_packed(2, 0, p2.reshape(-1), s2.reshape(-1), msgs, off), want)
    # a sample of the million against the CPU oracle (oracle/c64, the reference's per-signature path): 96 random
    # positions of the clean batch, and every tampered position plus 27 clean ones of the tampered batch
    from oracle import c64_oracle as C
    if not C.available():
        C.build()

    def oracle_statuses(pk_rows, sig_rows, idx):
        idx = np.asarray(idx)
        o = np.arange(idx.size + 1, dtype=np.uint64) * 32
        m = np.ascontiguousarray(msgs.reshape(n, 32)[idx]).reshape(-1)
        return C.verify_many(2, 0, 1, np.ascontiguousarray(pk_rows[idx]).reshape(-1), np.ascontiguousarray(sig_rows[idx]).reshape(-1), m, o)

    sample = np.sort(rng.choice(n, 96, replace=False))
    assert int(oracle_statuses(pks.reshape(n, 48), sigs.reshape(n, 96), sample).max()) == 0
    idx2 = np.unique(np.concatenate([bad, [12345, 54321], rng.choice(n, 27, replace=False)]))
    assert np.array_equal(oracle_statuses(p2, s2, idx2), want[idx2])


def test_cfg3_same_message_aggregation_of_one_million_signers(eng):
    n = 1_000_000
    rng = np.random.default_rng(5)
    scal = np.zeros((n, 32), dtype=np.uint8)
    scal[:, 8:] = rng.integers(0, 256, size=(n, 24), dtype=np.uint8)
    scal[:, 31] |= 1
    msg = np.frombuffer(b"one message for every signer....", dtype=np.uint8)
    o_all = np.arange(n + 1, dtype=np.uint64) * 32
    for impl, pkg, sgg in ((2, 1, 2), (1, 2, 1)):
        cnt = n
        pk, sg = [], []
        for lo in range(0, cnt, 1 << 18):
            hi = min(cnt, lo + (1 << 18))
            p, s = eng.testdata_sign(impl, 2, scal[lo:hi].reshape(-1), np.tile(msg, hi - lo), np.ascontiguousarray(o_all[lo:hi + 1] - o_all[lo]))
            pk.append(p)
            sg.append(s)
        pk, sg = np.concatenate(pk), np.concatenate(sg)
        apk, asg = eng.sum_points(pkg, pk), eng.sum_points(sgg, sg)
        assert eng.verify_batch(impl, 2, [apk], [asg], [msg.tobytes()]).tolist() == [0]
        # linearity: the sum over two halves adds up to the whole
        half = (cnt // 2) * (48 if pkg == 1 else 96)
        assert eng.sum_points(pkg, [eng.sum_points(pkg, pk[:half]), eng.sum_points(pkg, pk[half:])]) == apk
        # dropping one signer breaks the aggregate
        sl = 96 if sgg == 2 else 48
        assert eng.verify_batch(impl, 2, [apk], [eng.sum_points(sgg, sg[sl:])], [msg.tobytes()]).tolist() == [1]


def test_cfg4_aggregate_verify_over_100k_distinct_messages(eng, batch_1m):
    import blsful_b200 as B
    n, pks, sigs, msgs, off = batch_1m
    m = 100_000
    agg = eng.sum_points(2, sigs[:m * 96])
    msgs_list = [msgs[i * 32:(i + 1) * 32].tobytes() for i in range(m)]
    eng.aggregate_verify(2, 0, pks[:m * 48], msgs_list, agg)
    with pytest.raises(B.BlsError) as e:
        eng.aggregate_verify(2, 0, pks[:m * 48], msgs_list, eng.sum_points(2, sigs[96:(m + 1) * 96]))
    assert e.value.status == B.ST_INVALID_SIGNATURE
    with pytest.raises(B.BlsError) as e:  # duplicate messages are rejected before any curve work (sig_basic.rs:46-58)
        eng.aggregate_verify(2, 0, pks[:m * 48], msgs_list[:-1] + [msgs_list[7]], agg)
    assert e.value.status == B.ST_DUPLICATE_MESSAGES


def test_cfg5_full_size_ten_thousand_quorums_of_400(eng):
    """BASELINE configs[4] at its full size: 10,000 quorums x 400 members through aggregate_secure and verify_secure
    (Modern), flat buffers.  Pinned by properties: every aggregate made from the members' signatures verifies under its
    own key set, the aggregate of another quorum does not, a quorum with a repeated key is accepted exactly like the
    reference accepts it (coefficients are per position), and 3 sampled quorums agree with the CPU oracle."""
    from oracle import c64_oracle as C
    if not C.available():
        C.build()
    q, mem = 10_000, 400
    tot = q * mem
    rng = np.random.default_rng(19)
    scal = np.zeros((tot, 32), dtype=np.uint8)
    scal[:, 8:] = rng.integers(0, 256, size=(tot, 24), dtype=np.uint8)
    scal[:, 31] |= 1
    qm = rng.integers(0, 256, size=(q, 32), dtype=np.uint8)
    qm[:, :8] = np.arange(q, dtype=np.uint64).view(np.uint8).reshape(q, 8)
    pk5, sg5 = np.empty(tot * 48, dtype=np.uint8), np.empty(tot * 96, dtype=np.uint8)
    chunk = 1 << 18
    for lo in range(0, tot, chunk):
        hi = min(tot, lo + chunk)
        o = np.arange(hi - lo + 1, dtype=np.uint64) * 32
        p, g = eng.testdata_sign(2, 0, scal[lo:hi].reshape(-1), np.ascontiguousarray(qm[np.arange(lo, hi) // mem]).reshape(-1), o)
        pk5[lo * 48:hi * 48], sg5[lo * 96:hi * 96] = p, g
    koff = np.arange(q + 1, dtype=np.uint64) * mem
    qoff = np.arange(q + 1, dtype=np.uint64) * 32
    stq, aggs = eng.aggregate_secure_batch_packed(2, koff, pk5, sg5, 1)
    assert stq.shape == (q,) and int(stq.max()) == 0
    st = eng.verify_secure_batch_packed(2, 0, koff, pk5, aggs, qm.reshape(-1), qoff, 1)
    assert st.shape == (q,) and int(st.max()) == 0
    wrong = aggs.copy().reshape(q, 96)
    swap = np.sort(rng.choice(q, 20, replace=False))
    wrong[swap] = wrong[np.roll(swap, 1)]
    st = eng.verify_secure_batch_packed(2, 0, koff, pk5, wrong.reshape(-1), qm.reshape(-1), qoff, 1)
    assert np.nonzero(st)[0].tolist() == swap.tolist() and set(st[swap].tolist()) == {1}
    for j in (0, 4321, q - 1):
        keys = [pk5[(j * mem + i) * 48:(j * mem + i + 1) * 48].tobytes() for i in range(mem)]
        members = [sg5[(j * mem + i) * 96:(j * mem + i + 1) * 96].tobytes() for i in range(mem)]
        st_o, agg_o = C.aggregate_secure(2, 1, keys, members)
        assert st_o == 0 and agg_o == aggs[j * 96:(j + 1) * 96].tobytes()
        assert C.verify_secure(2, 0, 1, keys, agg_o, qm[j].tobytes()) == 0
        assert C.verify_secure(2, 0, 1, keys, aggs[((j + 1) % q) * 96:((j + 1) % q + 1) * 96].tobytes(), qm[j].tobytes()) == 1


def test_cfg5_verify_secure_quorums_of_400(eng):
    q, mem = 60, 400
    rng = np.random.default_rng(9)
    tot = q * mem
    scal = np.zeros((tot, 32), dtype=np.uint8)
    scal[:, 8:] = rng.integers(0, 256, size=(tot, 24), dtype=np.uint8)
    scal[:, 31] |= 1
    qmsgs = [(b"quorum %06d message" % j).ljust(32, b".") for j in range(q)]
    m5 = np.frombuffer(b"".join(m for m in qmsgs for _ in range(mem)), dtype=np.uint8)
    o5 = np.arange(tot + 1, dtype=np.uint64) * 32
    pk5, sg5 = eng.testdata_sign(2, 0, scal.reshape(-1), m5, o5)
    key_sets = [[pk5[(j * mem + i) * 48:(j * mem + i + 1) * 48].tobytes() for i in range(mem)] for j in range(q)]
    sig_sets = [[sg5[(j * mem + i) * 96:(j * mem + i + 1) * 96].tobytes() for i in range(mem)] for j in range(q)]
    for fmt in (1, 0):  # Modern, Legacy
        if fmt == 0:
            st, ks = eng.recode_points(1, [k for s in key_sets for k in s], 1, 0)
            assert int(np.max(st)) == 0
            st, ss = eng.recode_points(2, [x for s in sig_sets for x in s], 1, 0)
            assert int(np.max(st)) == 0
            key_sets = [ks[j * mem:(j + 1) * mem] for j in range(q)]
            sig_sets = [ss[j * mem:(j + 1) * mem] for j in range(q)]
        stq, aggs = eng.aggregate_secure_batch(2, key_sets, sig_sets, fmt)
        assert int(np.max(stq)) == 0
        assert eng.verify_secure_batch(2, 0, key_sets, aggs, qmsgs, fmt).tolist() == [0] * q
        shuffled = [ks[::-1] for ks in key_sets]   # the key order does not matter (sorted inside)
        assert eng.verify_secure_batch(2, 0, shuffled, aggs, qmsgs, fmt).tolist() == [0] * q
        swapped = list(aggs)
        swapped[3], swapped[4] = aggs[4], aggs[3]
        st = eng.verify_secure_batch(2, 0, key_sets, swapped, qmsgs, fmt)
        assert [i for i in range(q) if st[i]] == [3, 4]
