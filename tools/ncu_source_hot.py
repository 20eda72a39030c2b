"""Per-instruction view of an ncu --set full --import-source on capture: the hottest SASS instructions by
stall samples and by executed count for one kernel launch.
usage: python tools/ncu_source_hot.py x.ncu-rep LAUNCH_INDEX [TOP]"""
import csv
import io
import subprocess
import sys

rep, idx = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]   # one section per launch
h = heads[2 * idx]   # two sections (SASS view twice) per launch
end = heads[2 * idx + 1] - 1 if 2 * idx + 1 < len(heads) else len(rows)
hdr = rows[h]
data = [r for r in rows[h + 1:end] if len(r) == len(hdr)]
iS, iN, iX = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
tot_s = sum(int(r[iN] or 0) for r in data)
tot_x = sum(int(r[iX] or 0) for r in data)
print(f"# {rep} launch {idx}: {len(data)} SASS instructions, {tot_s} stall samples, {tot_x} warp instructions executed")
print("# hottest by samples: line, samples (%), executed, top stall reasons, SASS")
order = sorted(range(len(data)), key=lambda i: -int(data[i][iN] or 0))[:top]
for i in sorted(order):
    r = data[i]
    st = sorted(((int(r[c] or 0), hdr[c][6:]) for c in stall_cols), reverse=True)[:3]
    st = " ".join(f"{n}:{v}" for v, n in st if v)
    print(f"{i:5d} {int(r[iN]):7d} ({100.0 * int(r[iN]) / max(tot_s, 1):4.1f}%) x{int(r[iX]):9d}  [{st}]  {r[iS].strip()[:90]}")
# executed-instruction histogram by region (sum of executed per 50-instruction bucket)
print("# executed warp instructions per 64-instruction bucket")
for b in range(0, len(data), 64):
    ex = sum(int(r[iX] or 0) for r in data[b:b + 64])
    sm = sum(int(r[iN] or 0) for r in data[b:b + 64])
    print(f"  [{b:5d}, {b + 64:5d})  executed {ex:11d} ({100.0 * ex / max(tot_x, 1):4.1f}%)  samples {sm:7d} ({100.0 * sm / max(tot_s, 1):4.1f}%)")
