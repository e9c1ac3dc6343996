"""Drives every C-ABI entry point once on small inputs - the workload for compute-sanitizer runs
(memcheck / initcheck / racecheck / synccheck; logs are committed under profiles/).

    compute-sanitizer --tool memcheck --log-file gpurun_out/san_memcheck.log python tools/sanitize_all.py 300

Results are asserted against expectations known by construction (valid batches verify, planted failures are found), so
a run that corrupts data under the sanitizer's serialisation also fails here."""
import hashlib
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "agora-blsful_b200"))
import numpy as np

import blsful_b200 as B

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
big = int(sys.argv[2]) if len(sys.argv) > 2 else 0      # an extra batch on the bucket path (>= 4096) if given
eng = B.Engine([0])
rnd = random.Random(5)
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001


def scalars(m):
    return np.frombuffer(b"".join(rnd.randrange(1, 2 ** 250).to_bytes(32, "big") for _ in range(m)), dtype=np.uint8)


for impl in (2, 1):
    pl, sl = B.pk_len(impl), B.sig_len(impl)
    for scheme in (0, 1):
        k = scalars(n)
        msgs = [hashlib.sha256(b"m%d-%d" % (scheme, i)).digest()[: 1 + i % 32] for i in range(n)]
        data, off = B.pack_messages(msgs)
        pks, sigs = eng.testdata_sign(impl, scheme, k, data, off)
        assert int(eng.verify_batch_packed(impl, scheme, pks, sigs, data, off).max()) == 0
        s2 = sigs.copy().reshape(n, sl)
        s2[[3, n // 2]] = s2[[4, n // 2 + 1]]
        s2[7] = 0
        s2[7, 0] = 0xC0
        p2 = pks.copy().reshape(n, pl)
        p2[9] = 0xFF
        st = eng.verify_batch_packed(impl, scheme, p2.reshape(-1), s2.reshape(-1), data, off)
        assert np.nonzero(st)[0].tolist() == sorted([3, 7, 9, n // 2]), np.nonzero(st)[0].tolist()
        if scheme == 0:
            # Legacy round trip, recode, sums, aggregate verify
            stl, pk_l = eng.recode_points(1 if impl == 2 else 2, pks, 1, 0)
            stl2, sg_l = eng.recode_points(2 if impl == 2 else 1, sigs, 1, 0)
            assert int(stl.max()) == 0 and int(stl2.max()) == 0
            assert int(eng.verify_batch(impl, 0, pk_l, sg_l, msgs, 0).max()) == 0
            m = min(n, 40)
            agg = eng.sum_points(2 if impl == 2 else 1, sigs[: m * sl])
            pk_list = [pks[i * pl:(i + 1) * pl].tobytes() for i in range(m)]
            assert eng.aggregate_verify_status(impl, 0, pk_list, msgs[:m], agg)[0] == 0
            assert eng.aggregate_verify_status(impl, 0, pk_list, msgs[:m - 1] + [b"zz"], agg)[0] == 1
            # wire format
            tagged = [bytes([0]) + sigs[i * sl:(i + 1) * sl].tobytes() for i in range(m)]
            assert int(eng.verify_batch_wire(impl, pk_list, tagged, msgs[:m]).max()) == 0
    # proofs of possession
    k = scalars(n)
    pks, pops = eng.testdata_sign(impl, 3, k, np.zeros(0, dtype=np.uint8), np.zeros(n + 1, dtype=np.uint64))
    assert int(eng.pop_verify_batch(impl, pks, pops).max()) == 0
    # secure aggregation: a few quorums, one with duplicate keys
    q, mem = 5, 17
    k = scalars(q * mem)
    qmsgs = [b"quorum-%d" % j for j in range(q)]
    m5, o5 = B.pack_messages([qmsgs[j] for j in range(q) for _ in range(mem)])
    pk5, sg5 = eng.testdata_sign(impl, 0, k, m5, o5)
    key_sets = [[pk5[(j * mem + i) * pl:(j * mem + i + 1) * pl].tobytes() for i in range(mem)] for j in range(q)]
    sig_sets = [[sg5[(j * mem + i) * sl:(j * mem + i + 1) * sl].tobytes() for i in range(mem)] for j in range(q)]
    key_sets[2][5], sig_sets[2][5] = key_sets[2][1], sig_sets[2][1]
    stq, aggs = eng.aggregate_secure_batch(impl, key_sets, sig_sets)
    assert int(stq.max()) == 0
    st = eng.verify_secure_batch(impl, 0, key_sets, aggs, qmsgs)
    assert st.tolist() == [0] * q, st.tolist()
    aggs[1], aggs[3] = aggs[3], aggs[1]
    assert eng.verify_secure_batch(impl, 0, key_sets, aggs, qmsgs).tolist() == [0, 1, 0, 1, 0]
    # threshold shares of public keys: f(x) = a + b x over Fr, shares [f(x)]G at x = 1..4 combine to [a]G
    a, b = rnd.randrange(1, R), rnd.randrange(1, R)
    none = np.zeros(0, dtype=np.uint8)
    kk = np.frombuffer(b"".join(((a + b * x) % R).to_bytes(32, "big") for x in range(1, 5)), dtype=np.uint8)
    pks4, _ = eng.testdata_sign(impl, 0, kk, none, np.zeros(5, dtype=np.uint64))
    shares = [x.to_bytes(32, "big") + pks4[(x - 1) * pl: x * pl].tobytes() for x in range(1, 5)]
    st, outs = eng.combine_shares_batch(1 if impl == 2 else 2, [shares[:2], shares[1:4], shares[:1]])
    want, _ = eng.testdata_sign(impl, 0, np.frombuffer(a.to_bytes(32, "big"), dtype=np.uint8), none, np.zeros(2, dtype=np.uint64))
    assert st.tolist() == [0, 0, 11] and outs[0] == want.tobytes() and outs[1] == want.tobytes()

# pairing checks
g1 = eng.sum_points(1, [])
ok, st = eng.pairing_check_batch([[], [(g1, bytes([0xC0]) + bytes(95))]])
assert ok.tolist() == [1, 1] and st.tolist() == [0, 0]

if big:
    k = scalars(big)
    data, off = B.pack_messages([hashlib.sha256(b"big%d" % i).digest() for i in range(big)])
    pks, sigs = eng.testdata_sign(2, 0, k, data, off)
    assert int(eng.verify_batch_packed(2, 0, pks, sigs, data, off).max()) == 0
    s2 = sigs.copy().reshape(big, 96)
    s2[[5, big - 2]] = s2[[6, big - 1]]
    st = eng.verify_batch_packed(2, 0, pks, s2.reshape(-1), data, off)
    assert np.nonzero(st)[0].tolist() == [5, big - 2]
    eng.set_rlc_bits(128)
    assert np.nonzero(eng.verify_batch_packed(2, 0, pks, s2.reshape(-1), data, off))[0].tolist() == [5, big - 2]
print("sanitize_all: every entry point ran and every expectation held; kernel launches:", eng.launch_count())
eng.close()
