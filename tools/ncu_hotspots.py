"""Per-instruction picture of one kernel from an ncu report (source page, SASS view): where the stall samples sit and
which stall reasons dominate.   python tools/ncu_hotspots.py <report.ncu-rep> [top N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
# several kernels (launches) follow one another, each introduced by a "Kernel Name" line: take the first
blocks = txt.split('"Kernel Name",')
blk = blocks[1]
name, rest = blk.split("\n", 1)
rows = list(csv.DictReader(io.StringIO(rest)))
rows = [r for r in rows if r.get("Address", "").startswith("0x")]
tot = sum(int(r["# Samples"]) for r in rows)
inst = sum(int(r["Instructions Executed"]) for r in rows)
print(name.strip()[:120])
print(f"{len(rows)} SASS instructions, {tot} samples, {inst} warp instructions executed")
reasons = [k for k in rows[0].keys() if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: sum(int(r[k]) for r in rows) for k in reasons}
print("stall reasons (all samples):", ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.005 * tot))
ops = {}
for r in rows:
    op = r["Source"].split()[0] if not r["Source"].strip().startswith("@") else r["Source"].split()[1]
    op = op.split(".")[0]
    o = ops.setdefault(op, [0, 0])
    o[0] += int(r["Instructions Executed"]); o[1] += int(r["# Samples"])
print("by opcode (executed %, samples %):", ", ".join(f"{k} {100 * v[0] / inst:.1f}/{100 * v[1] / tot:.1f}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1][0])[:24]))
print("top instructions by samples:")
for i, r in sorted(enumerate(rows), key=lambda ir: -int(ir[1]["# Samples"]))[:top]:
    why = sorted(((int(r[k]), k[6:]) for k in reasons), reverse=True)[:2]
    print(f"  #{i:5d} {100 * int(r['# Samples']) / tot:5.2f}%  lanes {r['Avg. Threads Executed']:>5}  {r['Source'].strip()[:70]:70s} {why[0][1]} {why[0][0]}, {why[1][1]} {why[1][0]}")
