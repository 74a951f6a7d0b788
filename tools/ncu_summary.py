#!/usr/bin/env python
"""Summarise an .ncu-rep (read offline with `ncu -i`): per-launch key metrics, and the hottest source lines.
usage: tools/ncu_summary.py <rep> [--source KERNEL_REGEX] [--top N]"""
import csv, io, re, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = re.sub(r"\(.*", "", d.get("Kernel Name", ""))
        print(f"== {d.get('ID')} {name}")
        for k in hdr:
            if k in KEYS or k in ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
                                  "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
                                  "smsp__inst_executed.avg.per_cycle_active", "sm__inst_executed.avg.per_cycle_elapsed",
                                  "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
                                  "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"):
                print(f"   {k:90s} {d[k]} {units[hdr.index(k)]}")


def source(rep, kregex, top):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kregex, "-c", "1"],
                         capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"') or l.startswith('"#"'))
    rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
    hdr = rows[0]
    si = hdr.index("# Samples") if "# Samples" in hdr else hdr.index("Warp Stall Sampling (All Samples)")
    srci = hdr.index("Source")
    ii = hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    body = []
    for idx, r in enumerate(rows[1:]):
        try:
            body.append((int(r[si] or 0), idx, r))
        except (ValueError, IndexError):
            pass
    tot = sum(b[0] for b in body)
    print(f"total samples {tot}, instructions {len(body)}")
    for s, idx, r in sorted(body, reverse=True)[:top]:
        stalls = sorted(((int(r[c] or 0), hdr[c][6:]) for c in stall_cols), reverse=True)[:3]
        print(f"{100.0 * s / max(tot, 1):5.1f}%  #{idx:5d} ex={r[ii]:>9s}  {r[srci].strip()[:70]:70s} {stalls}")


if __name__ == "__main__":
    rep = sys.argv[1]
    if "--source" in sys.argv:
        k = sys.argv[sys.argv.index("--source") + 1]
        top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
        source(rep, k, top)
    else:
        raw(rep)
