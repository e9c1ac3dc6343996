"""Reduce gpurun_out/<TAG>_* (tools/gpu_evidence.sh) into profiles/: per-kernel ncu summaries, the traffic table, the launch
list and the bench lines.   python tools/collect_evidence.py r2"""
import csv, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
G = os.path.join(ROOT, "gpurun_out")
P = os.environ.get("PROFILES_OUT", os.path.join(ROOT, "profiles"))  # on the GPU box: a directory under gpurun_out/
os.makedirs(P, exist_ok=True)

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct", "l1tex__t_sector_pipe_lsu_mem_local_op_st_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
ALGO = {  # algorithmic bytes per signature: what the kernel has to read + write once
    "k_decode_g1": 48 + 100, "k_decode_g2": 96 + 196, "k_subgroup_check_g1": 100 + 1, "k_subgroup_check_g2": 196 + 1, "k_hash": 32 + 8 + 288,
    "k_clear_cofactor": 2 * 288, "k_to_affine_batch": 288 + 196, "k_m6_prep": 100 + 196 + 8 + 680, "k_m6_lines": 40000, "k_m6_accum": 40000,
}


def short(name):
    base = re.sub(r"^void ", "", name).split("(")[0].split("<")[0].replace("bls::", "")
    if base in ("k_decode", "k_subgroup_check"):
        base += "_g2" if "Aff<Fp2>" in name.split("(")[0] or "Aff<bls::Fp2>" in name.split("(")[0] else "_g1"
    return base


def bytes_of(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


rep = os.path.join(G, f"{tag}_hot.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    best = {}
    for k, vals in enumerate(rows[2:]):
        d = dict(zip(hdr, vals))
        nm = short(d["Kernel Name"])
        ms = float(d["gpu__time_duration.sum"]) * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}[units[hdr.index("gpu__time_duration.sum")]]
        if nm not in best or ms > best[nm][0]:
            best[nm] = (ms, k, d)
    n = 245760
    traffic = {}
    for nm, (ms, k, d) in sorted(best.items()):
        out = os.path.join(P, f"ncu_{tag}_{nm}_summary.csv")
        with open(out, "w") as f:
            f.write(f"# ncu --set full --clock-control none --import-source on (tools/gpu_evidence.sh), n = {n:,} items, the longest launch of this kernel in one pass\n")
            f.write(f"# {d['Kernel Name'][:150]}\n# metric,unit,value\n")
            for h, u in zip(hdr, units):
                if h in WANT or ("issue_stalled" in h and "per_issue_active" in h and float(d[h] or 0) >= 0.05):
                    f.write(f"{h},{u},{d[h]}\n")
            src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(k), "--launch-count", "1"],
                                 capture_output=True, text=True).stdout
            tmp = os.path.join(G, f"{tag}_src_{nm}.csv")
            open(tmp, "w").write(src)
            mix = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_opmix.py"), tmp, "14"], capture_output=True, text=True).stdout
            f.write("# --- executed opcode mix (tools/ncu_opmix.py; IMAD.WIDE occupies the multiplier pipe 4 cycles per warp, other IMAD forms 2) ---\n")
            f.write("".join("# " + l + "\n" for l in mix.splitlines()[1:]))
        rd = bytes_of(d["dram__bytes_read.sum"], units[hdr.index("dram__bytes_read.sum")])
        wr = bytes_of(d["dram__bytes_write.sum"], units[hdr.index("dram__bytes_write.sum")])
        traffic[nm] = {"kernel": d["Kernel Name"][:70], "ms_under_ncu": ms, "dram_bytes_per_sig": (rd + wr) / n,
                       "algorithmic_bytes_per_sig": ALGO.get(nm), "fmaheavy_pct": float(d["sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"])}
        print(f"{nm:24s} {ms:8.2f} ms  dram/sig {(rd + wr) / n:10.0f} B  fmaheavy {traffic[nm]['fmaheavy_pct']:.1f}%")
    # aliases the bench line looks up by the engine's kernel-timer names
    for a, b in (("k_decode", "k_decode_g1"), ("k_subgroup_check", "k_subgroup_check_g1")):
        if b in traffic:
            traffic[a] = dict(traffic[b], note="first launch = public keys (G1)")
    json.dump(traffic, open(os.path.join(P, f"traffic_{tag}.json"), "w"), indent=1)
for f in sorted(os.listdir(os.path.join(G, f"{tag}_profiles"))) if os.path.isdir(os.path.join(G, f"{tag}_profiles")) and "PROFILES_OUT" not in os.environ else []:
    shutil.copy(os.path.join(G, f"{tag}_profiles", f), os.path.join(P, f))  # reduced on the box (the report itself is too large to bring back)
    print("copied", f)
for src, dst in ((f"{tag}_launches.csv", f"launches_{tag}_bench_1M.csv"), (f"{tag}_bench.json", f"bench_{tag}_1M.json"),
                 (f"{tag}_bench_reference.json", f"bench_{tag}_reference_arm.json"), (f"{tag}_pytest.log", f"gputest_{tag}_pytest.log"),
                 (f"{tag}_smoke.log", f"gputest_{tag}_smoke.log")):
    if os.path.exists(os.path.join(G, src)) and "PROFILES_OUT" not in os.environ:
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))
        print("copied", dst)
