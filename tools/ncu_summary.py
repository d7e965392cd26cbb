#!/usr/bin/env python
"""Turns an .ncu-rep (ncu --set full --import-source on) into the JSON summary kept under profiles/:
selected raw metrics per captured kernel and, with --lines, the share of executed instructions,
shared-memory wavefronts and stall samples per source line.  Runs where ncu is installed (no GPU needed).
usage: ncu_summary.py REPORT.ncu-rep OUT.json [--lines 0.01]"""
import argparse, collections, csv, io, json, subprocess

METRICS = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed dram__throughput.avg.pct_of_peak_sustained_elapsed
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed
l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
l1tex__t_sector_hit_rate.pct lts__t_sector_hit_rate.pct lts__t_sector_op_read_hit_rate.pct lts__t_sectors_op_read.sum
lts__t_sectors_srcunit_tex_op_read.sum l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum
launch__block_size launch__grid_size launch__registers_per_thread launch__shared_mem_per_block_dynamic
sm__cycles_elapsed.avg sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__throughput.avg.pct_of_peak_sustained_elapsed sm__warps_active.avg.pct_of_peak_sustained_active
smsp__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active sm__inst_issued.avg.pct_of_peak_sustained_active
smsp__thread_inst_executed_per_inst_executed.ratio smsp__thread_inst_executed.sum""".split()


def run(args):
    return subprocess.run(["ncu", "-i"] + args, check=True, capture_output=True, text=True).stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out")
    ap.add_argument("--lines", type=float, default=0.0, help="also list source lines above this share")
    a = ap.parse_args()
    rows = list(csv.reader(io.StringIO(run([a.report, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    kernels = []
    for vals in rows[2:]:
        k = {"Kernel Name": vals[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in METRICS:
                k[h] = (vals[i] + " " + units[i]).strip()
        kernels.append(k)
    out = {"kernels": kernels}
    if a.lines > 0:
        src = list(csv.reader(io.StringIO(run([a.report, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
        cur, h = None, None
        agg = collections.defaultdict(lambda: [0, 0, 0])
        text = {}
        for r in src:
            if r and r[0] == "File Path":
                cur = r[1].split("/")[-1]
            elif r and r[0] == "Line No":
                h = r
            elif h is not None and len(r) >= len(h) - 5 and r[0].isdigit():
                d = dict(zip(h, r))
                key = (cur, int(r[0]))
                for j, name in enumerate(("Instructions Executed", "L1 Wavefronts Shared", "# Samples")):
                    try:
                        agg[key][j] += int(d.get(name) or 0)
                    except ValueError:
                        pass
                text[key] = r[1].strip()[:100]
        tot = [max(1, sum(v[j] for v in agg.values())) for j in range(3)]
        out["totals"] = {"instructions_executed": tot[0], "shared_wavefronts": tot[1], "stall_samples": tot[2]}
        out["lines"] = [
            {"file": k[0], "line": k[1], "inst_pct": round(100 * v[0] / tot[0], 2), "smem_wavefront_pct": round(100 * v[1] / tot[1], 2),
             "sample_pct": round(100 * v[2] / tot[2], 2), "source": text[k]}
            for k, v in sorted(agg.items()) if any(v[j] > a.lines * tot[j] for j in range(3))]
    json.dump(out, open(a.out, "w"), indent=1)
    for k in kernels:
        print(k["Kernel Name"], k.get("gpu__time_duration.sum"))


if __name__ == "__main__":
    main()
