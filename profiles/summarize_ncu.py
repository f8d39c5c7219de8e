#!/usr/bin/env python
"""Turn an `ncu --set full` report of bench.py's timed region into profiles/ncu_summary.json (+ a readable table).
    python profiles/summarize_ncu.py gpurun_out/<rep>.ncu-rep <mode> [out_table.md]
Launch order inside one forward chunk is fixed (conv_tc.cu:tc_forward_chunk), so the i-th conv_tc_kernel launch maps to a unit."""
import csv, io, json, os, subprocess, sys
ORDER = ["conv1_1", "conv1_2", "conv1_3", "side_op1", "conv2_1", "conv2_2", "conv2_3", "side_op2", "conv3_1", "conv3_2", "conv3_3",
         "side_op3", "conv4_1", "conv4_2", "conv4_3", "side_op4", "merge_conv", "merge_conv2"]
rep, mode = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, body = rows[0], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try:
        return float(r[col[k]].replace(",", ""))
    except Exception:
        return None
unit_scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}
units = rows[1]
def scaled(r, k):
    v = f(r, k)
    return None if v is None else v * unit_scale.get(units[col[k]], 1.0)
out = {}
convs = [r for r in body if "conv_tc_kernel" in r[col["Kernel Name"]]]
for i, r in enumerate(convs[:len(ORDER)]):
    out[ORDER[i]] = {
        "kernel": r[col["Kernel Name"]][:40], "grid": r[col["launch__grid_size"]],
        "duration_s_under_ncu": scaled(r, "gpu__time_duration.sum"),
        "dram_bytes_per_launch": (scaled(r, "dram__bytes_read.sum") or 0) + (scaled(r, "dram__bytes_write.sum") or 0),
        "tensor_pipe_active_pct": f(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        "tc_smem_wavefronts_pct": f(r, "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        "l2_throughput_pct": f(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        "dram_throughput_pct": f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "registers": f(r, "launch__registers_per_thread"), "smem_per_block_kb": f(r, "launch__shared_mem_per_block"),
    }
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_summary.json")
allm = json.load(open(path)) if os.path.exists(path) else {}
allm[mode] = out
json.dump(allm, open(path, "w"), indent=1)
lines = ["| unit | grid | ms (ncu) | DRAM MB/launch | tensor pipe % | tc smem wavefront % | L2 % | regs | smem KB |", "|---|---|---|---|---|---|---|---|---|"]
for k, v in out.items():
    lines.append("| %s | %s | %.3f | %.1f | %.1f | %.1f | %.1f | %d | %.1f |" % (k, v["grid"], (v["duration_s_under_ncu"] or 0) * 1e3, v["dram_bytes_per_launch"] / 1e6,
                 v["tensor_pipe_active_pct"] or 0, v["tc_smem_wavefronts_pct"] or 0, v["l2_throughput_pct"] or 0, v["registers"] or 0, v["smem_per_block_kb"] or 0))
txt = "\n".join(lines)
print(txt)
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write("# ncu --set full, %s mode, timed region of bench.py (one step)\n\n" % mode + txt + "\n")
