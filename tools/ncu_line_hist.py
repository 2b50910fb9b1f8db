"""Per-source-line executed-instruction / stall histogram of one kernel: joins `ncu --page source --csv` (per SASS instruction, in
address order) with `nvdisasm -g -gi` of the same build (source line + inlining chain per SASS instruction).

usage: python tools/ncu_line_hist.py REPORT.ncu-rep MANGLED_KERNEL_SUBSTRING [lo hi] [--so path/to/libnfc_b200.so] [--file nfc_tc.cuh]
  lo hi: attribute every instruction to the OUTERMOST line inside [lo, hi] of --file (i.e. the statement of the kernel body that the
         inlined helpers were called from); without them the innermost line is used.
Prints lines sorted by executed warp-instructions with their share and stall samples."""
import collections, csv, io, os, re, subprocess, sys, tempfile

args = [a for a in sys.argv[1:] if not a.startswith("--")]
rep, kname = args[0], args[1]
lo, hi = (int(args[2]), int(args[3])) if len(args) >= 4 else (None, None)
so = "neuralfreeconvection.jl_b200/libnfc_b200.so"
srcfile = "nfc_tc.cuh"
for i, a in enumerate(sys.argv):
    if a == "--so":
        so = sys.argv[i + 1]
    if a == "--file":
        srcfile = sys.argv[i + 1]
args = [a for a in args if a not in (so, srcfile) or a == args[0] or a == args[1]]
per_n = float(os.environ.get("PER", "0")) or None          # divide counts by this (e.g. warp-steps) for per-unit numbers

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
dis = None
for f in os.listdir(tmp):
    if f.endswith(".cubin"):
        t = subprocess.run(["nvdisasm", "-g", "-gi", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kname in t:
            dis = t
            break
assert dis, "kernel not found in any cubin"
# instructions of the kernel, in order, with their line chains
lines = dis.split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l and l.rstrip().endswith(":"))
insts = []            # (offset, text, chain)  chain: list of (file, line) innermost first
chain = []
pending = []
for l in lines[start + 1:]:
    if l.startswith("\t.section") or l.startswith("//-----"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        pending.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        if pending:
            chain, pending = pending, []
        insts.append((int(m.group(1), 16), m.group(2).strip(), chain))

# REPORT may also be the text of `ncu -i X.ncu-rep --page source --csv` (exported on the GPU box: reports are ~28 MB each)
txt = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi_ = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi_]
ci = {h: i for i, h in enumerate(hdr)}
prof = [r for r in rows[hi_ + 1:] if len(r) >= len(hdr)]
assert len(prof) == len(insts), (len(prof), len(insts), "report and library are different builds")

def key(ch):
    if lo is None:
        return ch[0] if ch else ("?", 0)
    for f, n in reversed(ch):       # outermost first
        if f == srcfile and lo <= n <= hi:
            return (f, n)
    return ch[-1] if ch else ("?", 0)

ex, st = collections.Counter(), collections.Counter()
tot = 0
for (off, text, ch), r in zip(insts, prof):
    n = int(r[ci["Instructions Executed"]]); s = int(r[ci["Warp Stall Sampling (All Samples)"]])
    k = key(ch)
    ex[k] += n; st[k] += s; tot += n
src = {}
def source_line(f, n):
    if f not in src:
        p = os.path.join("neuralfreeconvection.jl_b200/csrc", f)
        src[f] = open(p).read().split("\n") if os.path.exists(p) else []
    return src[f][n - 1].strip()[:110] if 0 < n <= len(src[f]) else ""
stot = sum(st.values())
print(f"total warp-instructions {tot}  stall samples {stot}" + (f"  per unit {tot / per_n:.0f}" if per_n else ""))
for k, n in ex.most_common(int(os.environ.get("TOP", "60"))):
    per = f" {n / per_n:8.1f}" if per_n else ""
    print(f"{k[0]}:{k[1]:<5d} {n:>13d} {100 * n / tot:5.1f}%{per}  stalls {100 * st[k] / max(1, stot):5.1f}%   {source_line(*k)}")
