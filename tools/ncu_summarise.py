"""Summarise an Nsight Compute report of the window-BA kernel into a small markdown file for profiles/ (run here, on the
.ncu-rep brought back in gpurun_out/):  python tools/ncu_summarise.py gpurun_out/x.ncu-rep profiles/x_summary.md"""
import collections, csv, io, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "launch__registers_per_thread",
        "launch__cluster_size", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.max.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__pcsamp_sample_buffer_full", "lts__t_sector_hit_rate.pct"]
lines = [f"# ncu summary of `{rep}`", "", "| launch | metric | unit | value |", "|---|---|---|---|"]
for li, r in enumerate(rows[2:]):
    for i, h in enumerate(hdr):
        if h in want:
            lines.append(f"| {li} | {h} | {units[i]} | {r[i]} |")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; H = None
tot = collections.Counter(); inst = collections.Counter(); stalls = collections.defaultdict(collections.Counter); text = {}
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 2 and r[0] == "Line No": H = r; sc = [i for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]; continue
    if H is None or len(r) < len(H) or r[0] == "": continue
    k = (cur, int(r[0])); text[k] = r[1]
    try: tot[k] += int(r[H.index("# Samples")] or 0); inst[k] += int(r[H.index("Instructions Executed")] or 0)
    except ValueError: pass
    for i in sc:
        try: stalls[k][H[i]] += int(r[i] or 0)
        except ValueError: pass
dup = {"helpers.h"}  # inlined wrappers repeat the samples of cooperative_groups.h lines
T = sum(n for k, n in tot.items() if k[0] not in dup)
lines += ["", f"Warp-state samples: {T} (cluster-barrier wrappers counted once), executed warp instructions: {sum(inst.values())}", "",
          "| samples | % | warp insts | file:line | top stall reasons | source |", "|---|---|---|---|---|---|"]
for k, n in [kv for kv in tot.most_common(60) if kv[0][0] not in dup][:40]:
    st = ", ".join(f"{a[6:]}={b}" for a, b in stalls[k].most_common(3))
    lines.append(f"| {n} | {100 * n / max(T, 1):.1f} | {inst[k]} | {k[0]}:{k[1]} | {st} | `{text[k].strip()[:90].replace('|', '/')}` |")
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out, "samples", T)
