"""Per-source-line table of one kernel from an ncu report taken with --set full --import-source on (run here, no GPU):
the SASS page of the report (samples, stall reasons, executed instructions per SASS instruction) joined, instruction by
instruction, with the line annotations nvdisasm -g prints for the same kernel of the library in the tree.
usage: python tools/source_line_table.py REPORT.ncu-rep KERNEL_REGEX MANGLED_NAME OUT.csv
  e.g. python tools/source_line_table.py gpurun_out/r02z_wave.ncu-rep k_extend _ZN5rb2008k_extendILb0EEEvNS_10WaveParamsEi profiles/r02z_k_extend_source_lines.csv
The report and the library must be the same code: the opcode sequences are compared and the script stops on a mismatch."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, kregex, mangled, out = sys.argv[1:5]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "reina-vk_b200", "csrc", "librb200.so")], cwd=tmp, check=True, capture_output=True)
listing = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, "wavefront.sm_100a.cubin")], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(listing) if l.startswith(".text." + mangled + ":")][0]
ins, cur = [], ("?", 0)
for l in listing[start + 1:]:
    if l.startswith("//---------------------"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m:
        ins.append((cur, m.group(2).strip()))
page = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kregex], capture_output=True, text=True).stdout
rows = list(csv.reader(page.split("\n")))
hdr, sec = rows[1], rows[2:2 + len(ins)]          # the first captured launch of the kernel
ci = {n: i for i, n in enumerate(hdr)}
op = lambda t: t.strip().split()[0].lstrip("@!P0123456789U ")[:4]
assert all(r and r[0].startswith("0x") for r in sec) and all(op(t) == op(r[ci["Source"]]) for (_, t), r in zip(ins, sec)), "report and library differ"
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
for (k, _), r in zip(ins, sec):
    a = agg[k]
    a[0] += int(r[ci["# Samples"]]); a[1] += int(r[ci["stall_long_sb"]])
    a[2] += int(r[ci["Instructions Executed"]]); a[3] += int(r[ci["Thread Instructions Executed"]])
tot = [sum(a[j] for a in agg.values()) for j in range(4)]
with open(out, "w") as f:
    w = csv.writer(f)
    w.writerow(["file", "line", "stall_samples_pct", "long_scoreboard_pct", "warp_instructions_pct", "avg_active_lanes"])
    for (fn, ln), a in sorted(agg.items()):
        w.writerow([fn, ln, round(100 * a[0] / tot[0], 2), round(100 * a[1] / tot[0], 2), round(100 * a[2] / tot[2], 2), round(a[3] / a[2], 1) if a[2] else 0])
print(len(ins), "instructions,", tot[2], "warp instructions,", round(tot[3] / tot[2], 2), "active lanes on average ->", out)
