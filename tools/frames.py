"""Stack frame, spill and register table of every kernel (ptxas -v), largest frames first.
    python tools/frames.py [substring ...]      # builds build/blsgpu.cubin with the flags of __graft_entry__.py"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as G
os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
log = os.path.join(ROOT, "build", "ptxas_v.txt")
with open(log, "w") as f:
    subprocess.run(["/usr/local/cuda/bin/nvcc"] + G.NVCC_FLAGS + ["-Xptxas", "-v", "-cubin", "-o", os.path.join(ROOT, "build", "blsgpu.cubin"),
                    os.path.join(G.CSRC, "blsgpu.cu")], stderr=f, check=True)
rows = []
for b in re.split(r"ptxas info\s+: Compiling entry function '", open(log).read())[1:]:
    name = subprocess.run(["c++filt", b.split("'")[0]], capture_output=True, text=True).stdout.strip()
    st = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", b)
    rg = re.search(r"Used (\d+) registers", b)
    rows.append((int(st.group(1)), int(st.group(2)), int(st.group(3)), int(rg.group(1)), name[:110]))
for r in sorted(rows, reverse=True):
    if len(sys.argv) < 2 or any(a in r[4] for a in sys.argv[1:]):
        print("frame %6d  spill st/ld %4d/%4d  regs %3d  %s" % r)
