"""Extracts the golden vectors the reference's own tests hold for the verification hot path into JSON.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Sources: tests/cpp_integration_test.rs:19-82,171-179 (C++ bls-signatures sk/pk/sig x3, message "hello",
plain aggregate bytes) and tests/secure_aggregation_test.rs:143-235 (57 production keys, aggregate
signature, message). Only data (hex constants) is extracted; no code.
"""
import json, re, os

REF = "/root/reference/tests"
here = os.path.dirname(os.path.abspath(__file__))


def byte_arrays(src):
    out = {}
    for m in re.finditer(r"const (\w+): \[u8; \d+\] = \[(.*?)\];", src, re.S):
        out[m.group(1)] = bytes(int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{2})", m.group(2))).hex()
    return out


cpp = open(f"{REF}/cpp_integration_test.rs").read()
arr = byte_arrays(cpp)
m = re.search(r"let normal_agg_bytes = \[(.*?)\];", cpp, re.S)
normal = bytes(int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{2})", m.group(1))).hex()
golden_cpp = {
    "source": "reference tests/cpp_integration_test.rs:19-82,171-179",
    "impl": "Bls12381G2Impl", "scheme": "Basic", "message": arr["MESSAGE_HELLO"],
    "signers": [{"sk": arr[f"CPP_SK{i}_BYTES"], "pk": arr[f"CPP_PK{i}_BYTES"], "sig": arr[f"CPP_SIG{i}_BYTES"]}
                for i in (1, 2, 3)],
    "normal_agg_sig12": normal,
}
json.dump(golden_cpp, open(f"{here}/cpp_integration.json", "w"), indent=1)

sec = open(f"{REF}/secure_aggregation_test.rs").read()
body = sec[sec.index("fn test_large_scale_aggregate_signature_verification"):]
sig_hex = re.search(r'let sig_hex = "([0-9a-f]+)"', body).group(1)
keys = re.findall(r'^\s+"([0-9a-f]{96})",?$', body, re.M)
msg = re.search(r'let message_hex = "([0-9a-f]+)"', body).group(1)
assert len(keys) == 57, len(keys)
json.dump({"source": "reference tests/secure_aggregation_test.rs:143-235", "impl": "Bls12381G2Impl",
           "scheme": "Basic", "format": "Modern", "sig": sig_hex, "keys": keys, "message": msg},
          open(f"{here}/secure_57.json", "w"), indent=1)
print("wrote", len(keys), "keys")
