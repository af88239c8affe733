#!/usr/bin/env python
"""Turn the raw profiler outputs in gpurun_out/ into the tracked summaries under profiles/.
  python scripts/profile_summaries.py sass                      -> profiles/r2_sass_census.md (cuobjdump of the built library)
  python scripts/profile_summaries.py full <raw.csv> <label>    -> profiles/r2_ncu_full_<label>.md + r2_dominant_kernel.json
  python scripts/profile_summaries.py launches <list.csv> <label> -> profiles/r2_launches_<label>.md
(raw.csv = `ncu -i X.ncu-rep --page raw --csv`; list.csv = the --metrics gpu__time_duration.sum launch list)"""
import collections, csv, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")


def sass():
    lib = os.path.join(ROOT, "xingtian_b200", "lib", "libxtb200.so")
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    pat = re.compile(r"\b(UBLKCP|UTMALDG|UTMASTG|UTCHMMA|UTCQMMA|LDTM|STTM|UTCBAR|SYNCS|ELECT|ACQBULK|HMMA|FFMA)\b")
    cur, cnt = None, collections.OrderedDict()
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); cnt[cur] = collections.Counter(); continue
        if cur:
            for t in pat.findall(line):
                cnt[cur][t] += 1
    cols = ["UBLKCP", "UTMALDG", "UTCHMMA", "LDTM", "UTCBAR", "SYNCS", "ELECT", "HMMA", "FFMA"]
    out = ["# SASS census of libxtb200.so (sm_100a), round 2", "",
           "`cuobjdump -sass xingtian_b200/lib/libxtb200.so`, mnemonic counts per kernel.  UBLKCP = `cp.async.bulk` (the TMA engine's",
           "linear bulk copy: operands are batch-planar, i.e. already tiles, so no tensor map and hence no UTMALDG); UTCHMMA =",
           "`tcgen05.mma` (kind::f16, bf16 inputs); LDTM = `tcgen05.ld`; UTCBAR = `tcgen05.commit`; SYNCS = mbarrier ops.  No HMMA",
           "(legacy mma.sync) anywhere.", "", "| kernel | " + " | ".join(cols) + " |", "|---|" + "---|" * len(cols)]
    for k, v in cnt.items():
        if not (v["UBLKCP"] or v["UTCHMMA"] or v["LDTM"] or v["HMMA"]):
            continue
        name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        out.append("| `%s` | " % name + " | ".join(str(v[c]) for c in cols) + " |")
    others = sum(1 for v in cnt.values() if not (v["UBLKCP"] or v["UTCHMMA"] or v["LDTM"] or v["HMMA"]))
    out += ["", "%d further kernels (loss / GAE / V-trace / optimiser / layout / fp32 fallback GEMMs) use no tensor-core or bulk-copy instructions." % others]
    open(os.path.join(PROF, "r2_sass_census.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


def full(raw, label, names):
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    g = lambda r, k: r[ix[k]] if k in ix else ""
    cols = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
            ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
            ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("launch__registers_per_thread", "regs"),
            ("smsp__inst_executed.sum", "warp insts")]
    out = ["# ncu --set full, %s" % label, "",
           "`ncu --set full --clock-control none --import-source on -k regex:'bp_wgrad_kernel|bp_rows_kernel' python scripts/one_step.py 1`",
           "(one eager PPO SGD step, B=320; cold caches, serialised launches -- durations here are NOT bench numbers).", "",
           "| # | op | kernel | grid | " + " | ".join(c[1] for c in cols) + " |", "|---|---|---|---|" + "---|" * len(cols)]
    dom = {}
    for n, r in enumerate(data):
        kn = re.sub(r"\(.*", "", g(r, "Kernel Name")).replace("void ", "")
        vals = []
        for k, _ in cols:
            v, u = g(r, k), units[ix[k]] if k in ix else ""
            try:
                v = "%.4g" % float(v)
            except ValueError:
                pass
            vals.append((v + " " + u).strip())
        op = names[n] if n < len(names) else ""
        out.append("| %d | %s | `%s` | %s | " % (n, op, kn, g(r, "Grid Size")) + " | ".join(vals) + " |")
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = sum(float(g(r, k)) * mult.get(units[ix[k]], 1) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        m = re.match(r"L(\d+) (fwd|wgrad|dgrad)", op)
        if m:
            dom["ppo:L%s/%d" % (m.group(1), {"fwd": 0, "wgrad": 1, "dgrad": 2}[m.group(2)])] = int(tot)
    open(os.path.join(PROF, "r2_ncu_full_%s.md" % label), "w").write("\n".join(out) + "\n")
    json.dump({"source": "profiles/r2_ncu_full_%s.md (ncu --set full, cold cache, B=320)" % label, "dram_bytes_per_launch": dom},
              open(os.path.join(PROF, "r2_dominant_kernel.json"), "w"), indent=1, sort_keys=True)
    print("\n".join(out)); print(dom)


def launches(path, label):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        n = re.sub(r"\(.*", "", r[4]).replace("void ", "")[:70]
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += float(r[-1]) / 1e3
    tot = sum(v[1] for v in agg.values())
    out = ["# ncu launch list, %s" % label, "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none ... python bench.py --steps 1 --warmup 1 --no-cpu`: %d launches, %.1f ms of kernel time."
           % (len(rows), tot / 1e3), "Per-launch times under ncu are cold-cache and serialised: read the SHARE column, not the absolute.", "",
           "| kernel | launches | total us | share | avg us |", "|---|---|---|---|---|"]
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.1f | %.1f %% | %.2f |" % (n, c, t, 100 * t / tot, t / c))
    open(os.path.join(PROF, "r2_launches_%s.md" % label), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    if sys.argv[1] == "sass":
        sass()
    elif sys.argv[1] == "full":
        names = ["L0 fwd", "L1 fwd", "L2 fwd", "L3 fwd (split-K partial)", "L3 wgrad", "L3 dgrad", "L2 wgrad", "L2 dgrad", "L1 wgrad", "L1 dgrad", "L0 wgrad"]
        full(sys.argv[2], sys.argv[3], names)
    else:
        launches(sys.argv[2], sys.argv[3])
