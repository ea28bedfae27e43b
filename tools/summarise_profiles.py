"""Turn gpurun_out/final/ (tools/collect_profiles.sh) into the tracked artefacts under profiles/:
   <tag>_bench.json, <tag>_launches_{cold,warm}.csv (+ .md summary), <tag>_ncu_kernels.md / .json (headline ncu metrics per kernel)."""
import collections
import csv
import json
import os
import re
import shutil
import sys

SRC = "gpurun_out/final"
WANT = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs/thread"), ("smsp__inst_executed.sum", "warp instructions"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("lts__t_bytes.sum", "L2 bytes"), ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 read sectors (tex)"),
        ("l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "global load bytes (L1)"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor instr"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (hmma subpipe)"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %")]


def to_bytes(v, u):
    v = float(v.replace(",", "")) if v else 0.0
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def launches(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for x in csv.DictReader(lines):
        if x["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", x["Kernel Name"]))[:64]
        v = float(x["Metric Value"].replace(",", ""))
        v = v / 1000 if x["Metric Unit"] == "ns" else v * 1000 if x["Metric Unit"] == "ms" else v
        k = (name, x["Grid Size"], x["Block Size"])
        agg[k][0] += 1
        agg[k][1] += v
    return agg


def main(tag):
    os.makedirs("profiles", exist_ok=True)
    shutil.copy(f"{SRC}/bench.json", f"profiles/{tag}_bench.json")
    md = [f"# {tag}: ncu launch lists of one training step (`bench.py --steps 1 --warmup 3 --no-graph --cpu-iters 0 --no-extras`, 1 x B200)\n",
          "`ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c 150` (tools/collect_profiles_r2.sh): a window of 150 launches = one full step",
          "(129 launches of this library + a few torch glue kernels) and the head of the next, serialised by the profiler -- the concurrency of the",
          "captured graph (side stream, background lane: profiles/r2_j_step_timeline.md) is NOT visible here.",
          "cold = ncu's default cache control (L2 flushed before every launch); warm = `--cache-control none`",
          "(L2 as the previous kernel left it -- the steady state of the captured graph).  Per-launch durations carry ~2.5 us of fixed",
          "profiler overhead (a 1-CTA elementwise kernel reads 2.5 us): compare SHARES.\n"]
    for mode in ("cold", "warm"):
        shutil.copy(f"{SRC}/launches_{mode}.csv", f"profiles/{tag}_launches_{mode}.csv")
        agg = launches(f"{SRC}/launches_{mode}.csv")
        tot = sum(v[1] for v in agg.values())
        md.append(f"\n## {mode}: {tot:.0f} us over {sum(v[0] for v in agg.values())} launches\n")
        md.append("| kernel | grid | block | launches | mean us | total us | share |\n|---|---|---|---|---|---|---|")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
            md.append(f"| `{k[0]}` | {k[1]} | {k[2]} | {v[0]} | {v[1] / v[0]:.2f} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |")
    open(f"profiles/{tag}_launches.md", "w").write("\n".join(md) + "\n")
    # per-kernel ncu --set full headline metrics
    out, md = {}, [f"# {tag}: `ncu --set full --clock-control none` headline metrics, one launch per kernel (steady state unless marked cold)\n"]
    for f in sorted(os.listdir(SRC)):
        if not f.endswith("_raw.csv"):
            continue
        rows = list(csv.reader(open(f"{SRC}/{f}")))
        if len(rows) < 3:
            continue
        hdr, units, r = rows[0], rows[1], rows[2]
        idx = {h: i for i, h in enumerate(hdr)}
        name = f[:-8]
        d = {"kernel": r[idx["Kernel Name"]]}
        md.append(f"\n## {name}: `{d['kernel'][:110]}`\n\n| metric | value |\n|---|---|")
        for key, label in WANT:
            if key in idx:
                d[key] = [r[idx[key]], units[idx[key]]]
                md.append(f"| {label} (`{key}`) | {r[idx[key]]} {units[idx[key]]} |")
        if "dram__bytes_read.sum" in idx:
            d["traffic_bytes"] = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
                to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            md.append(f"| **traffic = dram read + write** | {d['traffic_bytes'] / 1e6:.3f} MB |")
        stalls = []
        for h in hdr:
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                v = float(r[idx[h]] or 0)
                if v >= 0.3:
                    stalls.append((v, h.split("issue_stalled_")[1].split("_per_issue")[0]))
        md.append("| warp stalls per issued instruction (>= 0.3) | " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)) + " |")
        out[name] = d
    open(f"profiles/{tag}_ncu_kernels.md", "w").write("\n".join(md) + "\n")
    json.dump(out, open(f"profiles/{tag}_ncu_kernels.json", "w"), indent=1)
    shutil.copy(f"{SRC}/clocks.csv", f"profiles/{tag}_clocks.csv")
    print("wrote profiles/" + tag + "_*")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r1_f")
