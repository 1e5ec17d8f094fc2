"""Text summary of one .ncu-rep (ncu --set full --import-source on): key counters + hottest SASS lines.
usage: python scripts/ncu_summary.py gpurun_out/r2_x.ncu-rep [title] > profiles/r2_x_ncu_summary.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[-1]
m = {h: (v, u) for h, v, u in zip(hdr, vals, units)}
KEYS = ["Kernel Name", "Grid Size", "Block Size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum",
        "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "smsp__mem_tensor_reads_op_utcmma_matrix_c.sum.pct_of_peak_sustained_elapsed",
        "smsp__mem_tensor_writes_op_utcmma.sum.pct_of_peak_sustained_elapsed", "smsp__sass_inst_executed_op_utcmma.sum", "smsp__sass_inst_executed_op_tmem_ldt.sum",
        "smsp__sass_inst_executed_op_tmem_stt.sum"]
print(f"# {title}")
print(f"# source: {rep} (ncu --set full --clock-control none --import-source on; one launch; values as ncu reports them)")
for k in KEYS:
    if k in m and m[k][0] not in ("", "n/a"):
        print(f"{k:95s} {m[k][0]} {m[k][1]}")
st = [(float(v[0]), k) for k, v in m.items() if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k and v[0] not in ("", "n/a")]
tot = sum(s for s, _ in st) or 1.0
print("\nwarp stall samples (all warps, incl. warps parked in wait loops):")
for s_, k in sorted(st, reverse=True)[:8]:
    print(f"  {k.replace('smsp__pcsamp_warps_issue_stalled_', ''):28s} {s_:9.0f}  {100 * s_ / tot:5.1f}%")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(src)))
if len(r) > 2:
    h = r[1]
    ix = {c: i for i, c in enumerate(h)}
    data = r[2:]

    def f(d, c):
        try:
            return float(d[ix[c]])
        except Exception:
            return 0.0
    tot = sum(f(d, "# Samples") for d in data) or 1.0
    stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    print("\nhottest SASS instructions (share of stall samples, executions, two largest stall reasons):")
    for d in sorted(data, key=lambda d: -f(d, "# Samples"))[:14]:
        top = sorted(((f(d, c), c) for c in stall_cols), reverse=True)[:2]
        print(f"  {100 * f(d, '# Samples') / tot:5.1f}%  {d[ix['Source']].strip()[:64]:64s} exec={f(d, 'Instructions Executed'):11.0f}  " + " ".join(f"{c}={v:.0f}" for v, c in top if v))
