"""Turn ncu output (gpurun_out/) into the small text summaries committed under profiles/.

    python scripts/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
    python scripts/summarize_ncu.py full gpurun_out/prof_force_r1d.ncu-rep profiles/r1_force_kernel.md [atoms]

`full` also writes the numbers bench.py quotes (DRAM traffic per launch, FP64 flops executed per launch) as JSON next to
the summary (same name, .json); bench.py reads profiles/force_kernel_ncu.json, a copy of the capture at its workload.
"""
import json
import csv
import io
import subprocess
import sys
from collections import Counter, defaultdict


def short(name):
    name = name.replace("void ", "").replace("<unnamed>::", "")
    return name.split("(")[0][:90]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    first = next((i for i, r in enumerate(rows) if "annp_force_kernel" in r[4]), 0)
    # one MD step = everything from one FIRST-PASS annp_force_kernel launch up to the next (the overflow pass is a second,
    # near-empty launch of the same kernel right behind it: told apart by its duration)
    tmax = max([float(r[14]) for r in rows if "annp_force_kernel" in r[4]] + [0.0])
    idx = [i for i, r in enumerate(rows) if "annp_force_kernel" in r[4] and float(r[14]) >= 0.5 * tmax]
    lines = ["# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)", "",
             f"source: {path}; {len(rows)} launches captured, {len(idx)} force-kernel launches", ""]
    if len(idx) >= 2:
        a, b = idx[-2], idx[-1]
        # a step spans from just after the previous force kernel to (and including) this one, shifted so that it
        # starts at k_nve_initial
        step = rows[a + 1:b + 1]
        tot = sum(float(r[14]) for r in step)
        lines += ["## one MD step (launches between two consecutive force kernels)", "", "| kernel | grid | block | ns | share |", "|---|---|---|---|---|"]
        for r in step:
            lines.append(f"| {short(r[4])} | {r[8]} | {r[7]} | {float(r[14]):.0f} | {100 * float(r[14]) / tot:.2f} % |")
        fk = sum(float(r[14]) for r in step if "annp_force_kernel" in r[4])
        lines += ["", f"step total {tot / 1e6:.3f} ms; annp_force_kernel share {100 * fk / tot:.2f} % (first pass + overflow pass)", ""]
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[first:]:
        agg[short(r[4])][0] += 1
        agg[short(r[4])][1] += float(r[14])
    tot = sum(v[1] for v in agg.values())
    lines += ["## all launches from the first force kernel on", "", "| kernel | launches | total ns | share |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| {k} | {v[0]} | {v[1]:.0f} | {100 * v[1] / tot:.2f} % |")
    open(out, "w").write("\n".join(lines) + "\n")


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def full(rep, out, atoms=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    lines = [f"# ncu --set full --clock-control none ({rep})", "", f"kernel: {data[0][hdr.index('Kernel Name')]}", "",
             "| metric | unit | value |", "|---|---|---|"]
    for i, h in enumerate(hdr):
        if h in WANT or "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h.lower():
            lines.append(f"| {h} | {units[i]} | {data[0][i]} |")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    if hi:
        h = rows[hi[0]]
        body = rows[hi[0] + 1:]
        S, IE = h.index("# Samples"), h.index("Instructions Executed")

        def op(r):
            t = r[1].split()
            return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        c, cs = Counter(), Counter()
        for r in body:
            c[op(r)] += int(r[IE])
            cs[op(r)] += int(r[S])
        te, ts = sum(c.values()), sum(cs.values())
        lines += ["", "## SASS opcode mix (warp-level instructions executed / stall samples)", "", "| opcode | executed | share | sample share |", "|---|---|---|---|"]
        for k, v in c.most_common(16):
            lines.append(f"| {k} | {v} | {100 * v / te:.2f} % | {100 * cs[k] / ts:.2f} % |")
        fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DADD", "DMUL", "DSETP", "MUFU"))
        tens = sum(v for k, v in c.items() if k.startswith(("HMMA", "DMMA", "UTC", "IMMA")))
        lines += ["", f"FP64-pipe instructions: {fp64} ({100 * fp64 / te:.1f} % of all); tensor-pipe instructions: {tens}"]
        val = lambda name: float(data[0][hdr.index(name)].replace(",", ""))
        unit = lambda name: units[hdr.index(name)]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tpi = val("smsp__thread_inst_executed_per_inst_executed.ratio")
        flops = (2 * c["DFMA"] + c["DMUL"] + c["DADD"]) * tpi
        js = {"source": rep, "kernel": data[0][hdr.index("Kernel Name")],
              "gpu_time_ms": val("gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[unit("gpu__time_duration.sum")],
              "dram_bytes_read": val("dram__bytes_read.sum") * scale[unit("dram__bytes_read.sum")],
              "dram_bytes_write": val("dram__bytes_write.sum") * scale[unit("dram__bytes_write.sum")],
              "warp_instructions": te, "threads_per_instruction": tpi,
              "opcodes": {k: v for k, v in c.most_common(24)},
              "fp64_flops_executed": flops, "atoms": int(atoms) if atoms else None,
              "fp64_flops_executed_per_atom": flops / int(atoms) if atoms else None,
              "fp64_pipe_active_pct": val("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
              "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active")}
        json.dump(js, open(out.rsplit(".", 1)[0] + ".json", "w"), indent=1)
    open(out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:])
