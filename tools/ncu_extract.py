"""Extract the per-launch evidence columns from an `ncu --set full` report (run HERE, no GPU needed) and, for the three roofline
kernels of bench.py, the DRAM bytes per launch.

    python tools/ncu_extract.py gpurun_out/prof_r02.ncu-rep profiles/r02_ncu_full_top_kernels.csv profiles/r02_ncu_traffic.json
    python tools/ncu_extract.py profiles/r02_ncu_full_raw.csv profiles/r02_ncu_full_top_kernels.csv profiles/r02_ncu_traffic.json
(second form: the raw page was already exported ON the GPU box with `ncu -i rep --page raw --csv`, because a full report exceeds the
64 MB that travel back from it)
"""
import csv, io, json, subprocess, sys

rep, out_csv, out_json = sys.argv[1], sys.argv[2], sys.argv[3]
METRICS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"]
if rep.endswith(".csv"):
    rows = list(csv.reader(open(rep)))
    keep = [i for i, n in enumerate(rows[0]) if n in METRICS or n in ("ID", "Kernel Name", "Grid Size", "Block Size")]
    rows = [[r[i] for i in keep] for r in rows if len(r) == len(rows[0])]
    w = csv.writer(open(out_csv, "w", newline=""))
    w.writerows(rows)
else:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)], capture_output=True, text=True, check=True).stdout
    open(out_csv, "w").write(raw)
    rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(hdr)}


def mb(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[unit]


def pick(pattern, grid=None):
    sel = [r for r in data if pattern in r[col["Kernel Name"]] and (grid is None or r[col["Grid Size"]].replace(" ", "") == grid)]
    return sel[-1] if sel else None


out = {"source": f"{out_csv} (ncu --set full --clock-control none, tools/ncu_target.py, one B200); bytes per launch = dram__bytes_read.sum + dram__bytes_write.sum"}
for key, pat, note in (("roofline", "gemm_tap2_kernel<160>", "3x3 conv 320->320 @25x72x128; algorithmic A + out = 294.9 MB"),
                       ("roofline_attention", "flash_attn_d64_kernel", "5 heads N=9216 x 25; algorithmic q+k+v read once 442.4 MB + out 147.5 MB"),
                       ("roofline_groupnorm", "gn_apply_kernel", "C=320 @25x72x128, statistics from the producing conv's epilogue: one op = gn_part_finalize_kernel "
                                                                  "+ gn_apply_kernel (per-frame GroupNorm, grid 12 x 25); algorithmic read+write once 294.9 MB + 2.3 MB of partial sums"),
                       ("groupnorm_fused_statistics_pass", "gn_fused_kernel", "the same tensor without producer sums (statistics pass + normalise pass in one cooperative launch)")):
    sel = [r for r in data if pat in r[col["Kernel Name"]]]
    if not sel:
        continue
    if key == "roofline_groupnorm":
        # the per-frame (4-D) op of tools/ncu_target.py: the gn_apply launch with a 2-D grid and the gn_part_finalize launch right before it
        ap = [r for r in sel if "," in r[col["Grid Size"]] and not r[col["Grid Size"]].replace(" ", "").endswith(",1,1)")] or sel
        a = ap[-1]
        fin = [r for r in data if "gn_part_finalize" in r[col["Kernel Name"]] and int(r[col["ID"]]) < int(a[col["ID"]])]
        grp = ([fin[-1]] if fin else []) + [a]
    elif key == "roofline":
        grp = [sel[0]]                               # tools/ncu_target.py launches the plain 3x3 conv (bench.py's `roofline` launch) first
    else:
        grp = [sel[-1]]
    rd = sum(mb(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) for r in grp)
    wr = sum(mb(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]]) for r in grp)
    out[key] = {"kernel": pat, "launches_per_op": len(grp), "duration_us": round(sum(float(r[col["gpu__time_duration.sum"]]) * {"ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(units[col["gpu__time_duration.sum"]], 1.0) for r in grp), 1), "dram_read_mb": round(rd, 1), "dram_write_mb": round(wr, 1),
                "traffic_bytes": (rd + wr) * 1e6, "note": note,
                "tensor_pipe_pct": float(grp[-1][col["sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"]]),
                "dram_pct": float(grp[-1][col["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]) if "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed" in col else None}
json.dump(out, open(out_json, "w"), indent=1)
print(json.dumps(out, indent=1))
