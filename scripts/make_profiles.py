"""Turn the ncu outputs a gpurun call brought back (gpurun_out/) into the committed summaries under profiles/:
    python scripts/make_profiles.py launches gpurun_out/launches.csv profiles/r2_launches_step.txt "<command line>"
    python scripts/make_profiles.py full gpurun_out/prof_r2_head.ncu-rep profiles/r2_ncu_full_summary.txt [profiles/r2_traffic.json B P]
"""
import collections, csv, json, subprocess, sys


def launches(src, dst, cmd):
    rows, hdr = [], None
    for line in csv.reader(open(src)):
        if hdr is None:
            if line and line[0] == "ID":
                hdr = line
            continue
        if len(line) >= len(hdr):
            rows.append(line)
    idx = {h: i for i, h in enumerate(hdr)}
    names = [r[idx["Kernel Name"]] for r in rows]
    vals = [float(r[idx["Metric Value"]].replace(",", "")) for r in rows]
    grids = [r[idx["Grid Size"]] + " " + r[idx["Block Size"]] for r in rows]
    marks = [i for i, n in enumerate(names) if "gather_embed" in n]
    if len(marks) >= 3:
        a, b = marks[-3], marks[-2]  # one whole step between two loader gathers, late in the run
    else:
        a, b = 0, len(names)  # scripts/ncu_step.py: the capture IS one step (cudaProfilerStart/Stop around it)
    agg = collections.OrderedDict()
    for n, v, g in zip(names[a:b], vals[a:b], grids[a:b]):
        k = (n[:118], g)
        c = agg.setdefault(k, [0, 0.0])
        c[0] += 1
        c[1] += v
    tot = sum(v for _, v in agg.values())
    ours = sum(v for (n, _), (_, v) in agg.items() if "mde::" in n or "tc::" in n or n.startswith("void tc::") or "mde" in n.split("(")[0])
    with open(dst, "w") as f:
        f.write(f"# {cmd}\n# one inference step of BASELINE config 2 (B=16, 416x544): launches {a}..{b} of the capture "
                f"\n# {b - a} launches, {tot / 1000:.1f} us total (cold-cache, serialised: compare shares); "
                f"hand-written kernels (mde::*, tc::*): {ours / 1000:.1f} us = {100 * ours / tot:.1f} %\n")
        for (n, g), (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v / 1000:9.1f} us {100 * v / tot:5.1f}%  x{c:<3d} {n}  grid/block {g}\n")
    print(f"{dst}: {b - a} launches, {tot / 1000:.1f} us, ours {100 * ours / tot:.1f} %")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def full(src, dst, traffic=None, B=None, P=None):
    """src: one .ncu-rep, or a comma-separated list of raw-page csv exports (ncu -i rep --page raw --csv), concatenated"""
    seen, tr = set(), {}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none ... ({src}); one entry per distinct (kernel, grid)\n")
        for part in src.split(","):
            if part.endswith(".csv"):
                raw = open(part).read()
            else:
                raw = subprocess.run(["ncu", "-i", part, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
            rows = list(csv.reader(raw.splitlines()))
            if len(rows) < 3:
                continue
            hdr, units = rows[0], rows[1]
            idx = {h: i for i, h in enumerate(hdr)}
            _full_rows(f, rows[2:], idx, units, seen, tr)
    if traffic:
        for v in tr.values():
            v["batch"], v["P"] = int(B), int(P)
        json.dump(tr, open(traffic, "w"), indent=1)
    print(dst, "kernels:", len(seen))


def _full_rows(f, data, idx, units, seen, tr):
    if True:
        for r in data:
            name = r[idx["Kernel Name"]]
            key = name[:90] + (r[idx["Grid Size"]] if "Grid Size" in idx else "") + r[idx["gpu__time_duration.sum"]][:2]
            if key in seen:
                continue
            seen.add(key)
            f.write(f"\n{name[:160]}\n")
            for c in WANT:
                if c in idx:
                    f.write(f"    {c:75s} {r[idx[c]]} {units[idx[c]]}\n")
            short = name.split("(")[0].split("::")[-1].split("<")[0].replace("void ", "").strip()
            def num(c):
                v, u = float(r[idx[c]].replace(",", "")), units[idx[c]]
                return v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(u, 1)
            tr.setdefault(short, {"dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
                                  "duration_us_under_ncu": float(r[idx["gpu__time_duration.sum"]].replace(",", "")), "kernel": name[:120]})


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        full(*sys.argv[2:])
