"""Summarise `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` per CUDA source line:
share of stall samples, share of warp instructions, average active lanes."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr, data = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif "Instructions Executed" in r:
        hdr = {h: i for i, h in enumerate(r)}
    elif hdr and len(r) > 10 and r[0].isdigit():
        try:
            data.append((float(r[hdr["# Samples"]]), float(r[hdr["Instructions Executed"]]), float(r[hdr["Thread Instructions Executed"]]),
                         cur_file, int(r[0]), r[1].strip()[:100]))
        except ValueError:
            pass
ts, ti = sum(d[0] for d in data), sum(d[1] for d in data)
print(f"total stall samples {ts:.0f}, warp instructions {ti:.3e}, avg lanes {sum(d[2] for d in data)/ti:.2f}")
for d in sorted(data, reverse=True)[:top]:
    print(f"{100*d[0]/ts:5.1f}% smp {100*d[1]/ti:5.1f}% inst lanes {d[2]/max(d[1],1):4.1f} | {d[3]}:{d[4]} {d[5]}")
