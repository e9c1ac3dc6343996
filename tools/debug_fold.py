import hashlib, os, random, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "agora-blsful_b200"))
import numpy as np
import blsful_b200 as B
from oracle import bls_oracle as O
os.environ["BLSGPU_DEBUG_FOLD"] = "1"
rnd = random.Random(1)
n = 7
k = np.frombuffer(b"".join(rnd.randrange(1, O.R).to_bytes(32, "big") for _ in range(n)), dtype=np.uint8)
data, off = B.pack_messages([hashlib.sha256(b"dbg%d" % i).digest() for i in range(n)])
e = B.Engine([0])
pks, sigs = e.testdata_sign(2, 0, k, data, off)
gt, sm = e.miller_partial(2, 0, pks, sigs, data, off)
print("GT", gt.hex()); print("SM", sm.hex())
sys.stdout.flush()
print("fold(k=1):", e.final_exp_is_one(2, [gt], [sm]))
one = (1).to_bytes(48, "big") + bytes(528)
print("fold (1,O):", e.final_exp_is_one(2, [one], [bytes([0xC0]) + bytes(95)]))
e.close()
