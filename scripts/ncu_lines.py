"""Executed warp instructions and stall samples per CUDA source line of one kernel.
  python scripts/ncu_lines.py report.ncu-rep <kernel regex> <mangled-name substring> [launch-skip]
Joins `ncu --page source --csv` (SASS view: per-instruction counters) with `nvdisasm -g` line markers of the
in-tree library (must be the build that was profiled)."""
import csv, os, re, subprocess, sys, tempfile, collections
rep, kre, mangled = sys.argv[1:4]
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
k = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[k]
iA, iS, iI, iT, iSm = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
inst = [(int(r[iA], 16), r[iS].strip(), int(r[iI]), int(r[iT]), int(r[iSm])) for r in rows[k + 1:] if len(r) > iSm and r[iA].startswith("0x")]
base = inst[0][0]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "realtimeparticles_b200", "lib", "librtp_cuda.so")], cwd=tmp, capture_output=True)
line_of = {}
for f in os.listdir(tmp):
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if mangled not in dis:
        continue
    on, cur = False, ("?", 0)
    for ln in dis.splitlines():
        if ln.startswith("\t.section") or ln.startswith("//---"):
            on = (mangled in ln) if ".text." in ln else on and not ln.startswith("//---")
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur
agg = collections.OrderedDict()
tot = sum(i[2] for i in inst)
for a, s, n, t, sm in inst:
    key = line_of.get(a - base, ("?", 0))
    v = agg.setdefault(key, [0, 0, 0])
    v[0] += n; v[1] += t; v[2] += sm
print("total warp instructions %d, samples %d" % (tot, sum(i[4] for i in inst)))
src = {}
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    if f not in src:
        p = os.path.join(ROOT, "realtimeparticles_b200", "csrc", f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ""
    print("%5.1f%% inst %9d lanes %4.1f samples %5d  %s:%d  %s" % (100.0 * v[0] / tot, v[0], v[1] / max(v[0], 1), v[2], f, l, text))
