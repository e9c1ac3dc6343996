"""Scale sanity of the other BASELINE.json configs on one GPU (times are wall clock of the host call, inputs in host memory):
cfg 3 point sums over 1M signers + one verify of the aggregate, cfg 4 aggregate verify over 100k distinct messages,
cfg 5 verify_secure over quorums of 400 members, and cfg 2 with corrupted signatures (bisection cost).
Every result is cross-checked by an algebraic identity (no oracle at these sizes)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "agora-blsful_b200"))
import numpy as np
import blsful_b200 as B
import bench

eng = B.Engine([0])
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
out = {}

def tm(f):
    t = time.time(); r = f(); return r, (time.time() - t) * 1e3

# cfg 2 robustness: 1M batch with 1 and with 1000 corrupted signatures
n = int(os.environ.get("SCALE_N", 1_000_000))
pks, sigs, msgs, off = bench.synth_batch(eng, n, seed=5)
eng.verify_batch_packed(2, 0, pks, sigs, msgs, off)
st, t0 = tm(lambda: eng.verify_batch_packed(2, 0, pks, sigs, msgs, off))
assert int(st.max()) == 0
rng = np.random.default_rng(1)
for nbad in (1, 1000):
    bad = np.sort(rng.choice(n, nbad, replace=False))
    s2 = sigs.copy().reshape(n, 96)
    s2[bad] = s2[(bad + 1) % n]
    st, t = tm(lambda: eng.verify_batch_packed(2, 0, pks, s2.reshape(-1), msgs, off))
    assert np.array_equal(np.nonzero(st)[0], bad) and set(st[bad].tolist()) == {1}
    out[f"cfg2_{nbad}_bad_ms"] = round(t, 1)
out["cfg2_all_valid_ms"] = round(t0, 1)

# cfg 3: same-message multi-signature over n signers (PoP scheme): sum pks, sum sigs, one verify
scal = np.zeros((n, 32), dtype=np.uint8); scal[:, 8:] = rng.integers(0, 256, size=(n, 24), dtype=np.uint8); scal[:, 31] |= 1
msg = np.frombuffer(b"one message for every signer....", dtype=np.uint8)
m_all = np.tile(msg, n); o_all = (np.arange(n + 1, dtype=np.uint64) * 32)
pk3, sg3 = [], []
for lo in range(0, n, 1 << 18):
    hi = min(n, lo + (1 << 18))
    p, s = eng.testdata_sign(2, 2, scal[lo:hi].reshape(-1), m_all[lo * 32:hi * 32], np.ascontiguousarray(o_all[lo:hi + 1] - o_all[lo]))
    pk3.append(p); sg3.append(s)
pk3 = np.concatenate(pk3); sg3 = np.concatenate(sg3)
apk, t1 = tm(lambda: eng.sum_points(1, pk3))
asg, t2 = tm(lambda: eng.sum_points(2, sg3))
st, t3 = tm(lambda: eng.verify_batch(2, 2, [apk], [asg], [msg.tobytes()]))
assert st.tolist() == [0]
out["cfg3_sum_pk_ms"], out["cfg3_sum_sig_ms"], out["cfg3_verify_ms"] = round(t1, 1), round(t2, 1), round(t3, 1)

# cfg 4: aggregate verify over 100k distinct messages
m4 = 100_000
agg = eng.sum_points(2, sigs[:m4 * 96])
pk_list = pks[:m4 * 48]
msgs_list = [msgs[i * 32:(i + 1) * 32].tobytes() for i in range(m4)]
(_, t4) = tm(lambda: eng.aggregate_verify(2, 0, pk_list, msgs_list, agg))
out["cfg4_aggregate_verify_100k_ms"] = round(t4, 1)
try:
    eng.aggregate_verify(2, 0, pk_list, msgs_list, eng.sum_points(2, sigs[96:(m4 + 1) * 96]))
    raise SystemExit("cfg4: a wrong aggregate verified")
except B.BlsError:
    pass
print(out, flush=True)

# cfg 5: verify_secure over q quorums of 400 members (Modern and Legacy), signatures built with aggregate_secure
q, mem = int(os.environ.get("SCALE_Q", 1000)), 400
tot = q * mem
sc5 = np.zeros((tot, 32), dtype=np.uint8); sc5[:, 8:] = rng.integers(0, 256, size=(tot, 24), dtype=np.uint8); sc5[:, 31] |= 1
qmsgs = [(b"quorum %06d message" % j).ljust(32, b".") for j in range(q)]
m5 = np.frombuffer(b"".join(m for m in qmsgs for _ in range(mem)), dtype=np.uint8)
o5 = (np.arange(tot + 1, dtype=np.uint64) * 32)
pk5, sg5 = [], []
for lo in range(0, tot, 1 << 18):
    hi = min(tot, lo + (1 << 18))
    p, s = eng.testdata_sign(2, 0, sc5[lo:hi].reshape(-1), m5[lo * 32:hi * 32], np.ascontiguousarray(o5[lo:hi + 1] - o5[lo]))
    pk5.append(p); sg5.append(s)
pk5 = np.concatenate(pk5); sg5 = np.concatenate(sg5)
key_sets = [[pk5[(j * mem + i) * 48:(j * mem + i + 1) * 48].tobytes() for i in range(mem)] for j in range(q)]
sig_sets = [[sg5[(j * mem + i) * 96:(j * mem + i + 1) * 96].tobytes() for i in range(mem)] for j in range(q)]
(stq, aggs), t5a = tm(lambda: eng.aggregate_secure_batch(2, key_sets, sig_sets))
assert int(np.max(stq)) == 0
st, t5 = tm(lambda: eng.verify_secure_batch(2, 0, key_sets, aggs, qmsgs))
assert st.tolist() == [0] * q
aggs_bad = list(aggs); aggs_bad[3], aggs_bad[4] = aggs[4], aggs[3]
st = eng.verify_secure_batch(2, 0, key_sets, aggs_bad, qmsgs)
assert [i for i in range(q) if st[i]] == [3, 4]
out = {"cfg5_quorums": q, "cfg5_aggregate_secure_ms": round(t5a, 1), "cfg5_verify_secure_ms": round(t5, 1)}
print(out, flush=True)
