"""Static SASS instruction count per source region of one kernel (code-footprint check).
usage: python tools/sass_lines.py file.cubin kernel_substring [file:lo-hi=name,...]"""
import sys, re, collections, subprocess
cubin, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
cur = None; cnt = collections.Counter(); active = False
for l in out:
    if l.startswith(".text."):
        active = kern in l
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+[A-Z@]', l):
        cnt[cur] += 1
print("total", sum(cnt.values()))
groups = []
if len(sys.argv) > 3:
    for g in sys.argv[3].split(","):
        r, name = g.split("="); f, lh = r.split(":"); lo, hi = lh.split("-"); groups.append((f, int(lo), int(hi), name))
agg = collections.Counter()
for k, n in cnt.items():
    name = "other:" + (k[0] if k else "none")
    if k:
        for f, lo, hi, nm in groups:
            if k[0].startswith(f) and lo <= k[1] <= hi:
                name = nm; break
    agg[name] += n
for k, v in agg.most_common():
    print("%-24s %d" % (k, v))
if not groups:
    for k, v in cnt.most_common(25):
        print(k, v)
