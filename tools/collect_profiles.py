"""Copies the outputs of tools/gpu_round_end.sh (gpurun_out/<prefix>_*) into profiles/ under the round's names, exports the
ncu raw / details pages, writes the launch-share table and profiles/traffic.json (keyed by the hash of the kernel sources).
    python tools/collect_profiles.py r02f r02"""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
src, dst = sys.argv[1], sys.argv[2]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
shutil.copy(f"{G}/{src}_bench_2p20.json", f"{P}/{dst}_tensor_bench_1gpu_lqr_d20_2p20.json")
for w in ("lqr_d20", "lqr_d5", "vdp_d10", "ekn_d20", "lqr_var_d20"):
    shutil.copy(f"{G}/{src}_bench_cfg_{w}.json", f"{P}/{dst}_tensor_bench_cfg_{w}.json")
shutil.copy(f"{G}/{src}_bench_reference.json", f"{P}/{dst}_reference_arm_cpu.json")
shutil.copy(f"{G}/{src}_launches.csv", f"{P}/{dst}_tensor_bench_launches_lqr_d20_2p20.csv")
for k in ("critic", "actor"):
    rep = f"{G}/{src}_{k}_full.ncu-rep"
    open(f"{P}/{dst}_tensor_{k}_kernel_bench_raw.csv", "w").write(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)
    open(f"{P}/{dst}_tensor_{k}_kernel_bench_details.txt", "w").write(subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout)
rows = list(csv.reader(open(f"{P}/{dst}_tensor_bench_launches_lqr_d20_2p20.csv")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.defaultdict(float), collections.Counter()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    v, u = float(r[vi].replace(",", "")), r[ui]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v if u in ("ms", "msecond") else v * 1e3
    name = r[ki].split("(")[0]
    tot[name] += ms; cnt[name] += 1
T = sum(tot.values())
lines = ["ncu launch list of `python bench.py --steps 1 --warmup 3 --no-cpu-baseline` (lqr_d20, 2^20 paths, 1 GPU; first 400 launches; per-launch times are",
         "cold-cache and serialised: compare SHARES): kernel, launches, total ms, ms per launch, share\n"]
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    lines.append(f"{k:60s} {cnt[k]:4d} {v:12.3f} {v / cnt[k]:10.3f} {v / T:7.2%}")
open(f"{P}/{dst}_tensor_bench_launch_shares.txt", "w").write("\n".join(lines) + "\n")
h = bench.kernel_source_hash()
def grab(p, names):
    rows = list(csv.reader(open(p))); hdr, units, vals = rows[0], rows[1], rows[2]
    out = {}
    for n in names:
        i = hdr.index(n)
        out[n] = (float(vals[i].replace(",", "")), units[i])
    return out
t = {}
summary = []
for k in ("critic", "actor"):
    m = grab(f"{P}/{dst}_tensor_{k}_kernel_bench_raw.csv", ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
             "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
             "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum", "launch__registers_per_thread"])
    sc = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
    r, w = m["dram__bytes_read.sum"][0] * sc[m["dram__bytes_read.sum"][1]], m["dram__bytes_write.sum"][0] * sc[m["dram__bytes_write.sum"][1]]
    t[k] = {"kernel": f"{k}_tc_kernel", "dram_bytes_read": r, "dram_bytes_write": w, "source_hash": h,
            "capture": f"profiles/{dst}_tensor_{k}_kernel_bench_raw.csv (ncu --set full --clock-control none, one launch of `bench.py --steps 1 --warmup 3`, kernel sources {h})"}
    summary.append(f"{k}: " + ", ".join(f"{n.split('.')[0]}={v[0]:g} {v[1]}" for n, v in m.items()))
json.dump({"lqr_d20_2p20": t}, open(f"{P}/traffic.json", "w"), indent=1)
print("\n".join(lines[2:5])); print("\n".join(summary)); print("source hash", h)
