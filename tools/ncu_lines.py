"""Aggregate an ncu report's source page by CUDA line / opcode / stall reason for one kernel.
usage: python tools/ncu_lines.py report.ncu-rep kernel_regex [launch_index_from_end]"""
import collections, csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else -1
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# kernels are separated by 'Function Name' rows preceded by 'File Path'
kern = []
cur = None
for r in rows:
    if r and r[0] == "File Path":
        cur = {"file": r[1], "rows": []}
        kern.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
# group consecutive file blocks into launches: a launch starts at a block whose file is the .cu of interest
launches = []
for b in kern:
    if b["file"].endswith(".cu") and (not launches or launches[-1][-1]["file"].endswith(".hpp") or True):
        pass
cu = [b for b in kern if b["file"].endswith(".cu")]
blk = cu[which]
# include the header-file blocks that follow this .cu block until the next .cu block
i0 = kern.index(blk)
group = [blk]
for b in kern[i0 + 1:]:
    if b["file"].endswith(".cu"):
        break
    group.append(b)
lines = collections.defaultdict(lambda: [0, 0, ""])
stall = collections.Counter()
ops = collections.Counter(); opsamp = collections.Counter()
for b in group:
    hdr = b["rows"][1]
    ci = {h: i for i, h in enumerate(hdr)}
    si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
    key = None
    for r in b["rows"][2:]:
        if not r:
            continue
        if r[0] != "":
            key = (b["file"].split("/")[-1], int(r[0]))
            lines[key][2] = r[1][:100]
        elif len(r) > ii:
            try:
                s, n = int(r[si]), int(r[ii])
            except ValueError:
                continue
            lines[key][0] += s
            lines[key][1] += n
            t = r[3].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            ops[op] += n; opsamp[op] += s
            for h, j in ci.items():
                if h.startswith("stall_") and "Not Issued" not in h and j < len(r):
                    try:
                        stall[h] += int(r[j])
                    except ValueError:
                        pass
ts = sum(v[0] for v in lines.values()); ti = sum(v[1] for v in lines.values())
print("samples", ts, "warp-instructions", ti)
print("stalls:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / max(ts, 1)) for k, v in stall.most_common(9)))
print("opcodes:", ", ".join("%s %.1f%%(%d smp)" % (k, 100 * v / max(ti, 1), opsamp[k]) for k, v in ops.most_common(14)))
print("-- top lines by samples")
for k, (s, n, src) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:22]:
    print("%-10s %4d smp %5.1f%% inst %5.1f%%  %s" % (k[0][:10], k[1], 100 * s / max(ts, 1), 100 * n / max(ti, 1), src.strip()))
print("-- top lines by instructions")
for k, (s, n, src) in sorted(lines.items(), key=lambda kv: -kv[1][1])[:14]:
    print("%-10s %4d smp %5.1f%% inst %5.1f%%  %s" % (k[0][:10], k[1], 100 * s / max(ts, 1), 100 * n / max(ti, 1), src.strip()))
if len(sys.argv) > 4:      # optional: "file:lo-hi=name,..." groups -> share of samples / instructions per group
    groups = []
    for g in sys.argv[4].split(","):
        rng_, name = g.split("=")
        f, lh = rng_.split(":")
        lo, hi = lh.split("-")
        groups.append((f, int(lo), int(hi), name))
    agg = collections.defaultdict(lambda: [0, 0])
    for k, (s, n, src) in lines.items():
        name = "other"
        for f, lo, hi, nm in groups:
            if k[0].startswith(f) and lo <= k[1] <= hi:
                name = nm
                break
        agg[name][0] += s; agg[name][1] += n
    print("-- groups")
    for nm, (s, n) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-14s smp %5.1f%% inst %5.1f%%" % (nm, 100 * s / max(ts, 1), 100 * n / max(ti, 1)))
