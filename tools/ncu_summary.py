#!/usr/bin/env python
"""Summarise an .ncu-rep (one `ncu --set full` capture) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [top_n] [--json profiles/dram_bytes_per_point.json --points N] > profiles/rNN_kernel.txt

``--json`` also writes the measured DRAM bytes per point of each kernel class (what bench.py reports as
``roofline.traffic``), so that the number in the bench line is read from the committed capture, not typed in.

Reads the report with `ncu -i ... --page raw --csv` and `--page source --csv` (no GPU needed).
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def kernel_class(name):
    """bench.py's kernel classes from the (demangled) kernel name: chain_kernel<DT, PAR, PAIR, SW> with SW = 1 forward /
    reverse programs, 2 tangent / adjoint programs."""
    if "chain_kernel" in name:
        args = name[name.index("<") + 1:name.index(">")].replace(" ", "").split(",") if "<" in name else []
        sw = args[3] if len(args) > 3 else "3"
        return "chain_tan_adj" if sw.startswith("2") else "chain_fwd_rev"
    if "dw2_kernel" in name or "dw_kernel" in name:
        return "dw_gemm"
    return "other"


def main():
    argv = [a for a in sys.argv[1:]]
    json_out, points = None, None
    if "--json" in argv:
        i = argv.index("--json")
        json_out = argv[i + 1]
        del argv[i:i + 2]
    if "--points" in argv:
        i = argv.index("--points")
        points = int(argv[i + 1])
        del argv[i:i + 2]
    rep = argv[0]
    top_n = int(argv[1]) if len(argv) > 1 else 25
    rows = ncu(rep, "raw")
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}")
    traffic = {}
    for k, r in enumerate(rows[2:]):
        print(f"\n## launch {k}: {r[ix['Kernel Name']]}")
        for key in KEYS:
            if key in ix:
                print(f"{key:72s} {r[ix[key]]} {units[ix[key]]}")
        if "dram__bytes_read.sum" in ix:
            def gb(key):
                v, u = float(r[ix[key]]), units[ix[key]]
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
            tb = gb('dram__bytes_read.sum') + gb('dram__bytes_write.sum')
            print(f"{'traffic (dram read + write) bytes':72s} {tb:.4g}")
            cls = kernel_class(r[ix['Kernel Name']])
            traffic.setdefault(cls, []).append(tb)
    if json_out and points:
        import json
        # one launch of each hot class per step at this size: the largest launch of the class is the step's kernel
        out = {"source": f"ncu --set full capture {rep.split('/')[-1]} (tools/ncu_summary.py), bytes per point = dram read + write of the "
                         f"class's launch / {points} points", "points": points,
               "bytes_per_point": {c: max(v) / points for c, v in traffic.items() if c != "other"}}
        with open(json_out, "w") as f:
            json.dump(out, f, indent=1)
    src = ncu(rep, "source")
    blocks, hdr2, cur = [], None, None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = []
            blocks.append((r[1], cur))
        elif r and r[0] == "Address":
            hdr2 = r
        elif cur is not None and len(r) > 10:
            cur.append(r)
    if not blocks:
        return
    jx = {h: i for i, h in enumerate(hdr2)}
    stalls = [h for h in hdr2 if h.startswith("stall_") and "Not Issued" not in h]
    # the source page holds one block per profiled launch IN THE ORDER OF THE RAW PAGE, each headed by its kernel name:
    # label the tables with that name (round 1 numbered them and mis-assigned two) and drop byte-identical repeats
    seen = set()
    for k, (name, b) in enumerate(blocks):
        samp = sum(int(r[jx["# Samples"]]) for r in b) or 1
        sig = (name, samp, len(b))
        if sig in seen:
            continue
        seen.add(sig)
        print(f"\n## warp-stall samples of {name} ({samp} total, {len(b)} SASS instructions)")
        tot = {s: sum(int(r[jx[s]]) for r in b) for s in stalls}
        for s, v in sorted(tot.items(), key=lambda t: -t[1])[:8]:
            print(f"  {s:28s} {100.0 * v / samp:5.1f} %")
        print(f"  top {top_n} instructions by samples: samples / executed / SASS / dominant stall")
        for r in sorted(b, key=lambda r: -int(r[jx["# Samples"]]))[:top_n]:
            st = max(((int(r[jx[s]]), s) for s in stalls))
            print(f"  {int(r[jx['# Samples']]):7d} {int(r[jx['Instructions Executed']]):10d}  {r[jx['Source']].strip()[:64]:64s} {st[1]}")


if __name__ == "__main__":
    main()
