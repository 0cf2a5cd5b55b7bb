"""Instructions executed and stall samples of one kernel per SOURCE LINE (ncu report with --import-source on, code built
with -lineinfo).   python tools/ncu_lines.py <report.ncu-rep> [top N]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
agg = collections.OrderedDict()
cur = None; first_kernel = None
lines = txt.splitlines()
i = 0
while i < len(lines):
    ln = lines[i]
    if ln.startswith('"File Path"'):
        cur = next(csv.reader([ln]))[1].split("/")[-1]
        fn = next(csv.reader([lines[i + 1]]))[1]
        if first_kernel is None: first_kernel = fn
        if fn != first_kernel: break
        hdr = next(csv.reader([lines[i + 2]]))
        i += 3
        while i < len(lines) and not lines[i].startswith('"File Path"'):
            r = next(csv.reader([lines[i]]))
            if len(r) == len(hdr) and r[0].isdigit():
                d = dict(zip(hdr, r))
                k = (cur, int(d["Line No"]))
                a = agg.setdefault(k, [0, 0, 0, d["Source"].strip()[:90]])
                a[0] += int(d["Instructions Executed"]); a[1] += int(d["# Samples"]); a[2] += int(d["Thread Instructions Executed"])
            i += 1
        continue
    i += 1
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print(first_kernel[:100]); print(f"{tot_i} warp instructions, {tot_s} samples")
byfile = collections.Counter()
for (f, l), a in agg.items(): byfile[f] += a[0]
print("by file:", ", ".join(f"{f} {100*v/tot_i:.1f}%" for f, v in byfile.most_common()))
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*a[0]/tot_i:5.2f}% inst {100*a[1]/tot_s:5.2f}% smp  lanes {a[2]/max(a[0],1):4.1f}  {f}:{l}  {a[3]}")
