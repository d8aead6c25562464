"""Per-source-line attribution of an ncu capture (needs -lineinfo + --import-source on):
python tools/ncu_lines.py file.ncu-rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
hdr = rows[hi]
isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
inot = hdr.index("Warp Stall Sampling (Not-issued Samples)")
data = []
for r in rows[hi + 1:]:
    if r and r[0] and r[0].isdigit():
        try:
            data.append((int(r[isamp]), int(r[inot]), int(r[iex]), int(r[0]), r[1].strip()[:100]))
        except ValueError:
            pass
S = sum(d[0] for d in data); N = sum(d[1] for d in data); E = sum(d[2] for d in data)
print(f"samples {S} (not issued {N}), warp instructions {E}")
print("  %smp  %notis  %instr  line  source")
for d in sorted(data, reverse=True)[:top]:
    print(f"{100*d[0]/S:6.1f} {100*d[1]/max(N,1):6.1f} {100*d[2]/E:6.1f}  L{d[3]:<4} {d[4]}")
