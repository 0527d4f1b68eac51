#!/usr/bin/env python
"""Attribute ncu warp-stall samples / executed instructions of one kernel to source lines.
usage: tools/ncu_lines.py report.ncu-rep library.so kernel_substring [top]
(ncu's CSV source page is SASS-level; line numbers come from nvdisasm -g of the same library, so the library
must be the build that was profiled.)"""
import csv, io, os, re, subprocess, sys, tempfile, glob

rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
line_of = {}
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    txt = subprocess.run(["nvdisasm", "-c", "-g", cub], capture_output=True, text=True).stdout
    on, cur = False, "?"
    for l in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            on = kern in m.group(1)
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            cur = os.path.basename(m.group(1)) + ":" + m.group(2)
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/", l)
        if m:
            line_of[int(m.group(1), 16)] = cur
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
sec, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}; sec.append(cur); continue
    if cur is None: continue
    if cur["hdr"] is None: cur["hdr"] = r; continue
    cur["rows"].append(r)
for s in sec:
    if kern not in s["name"].replace("::", ""): 
        if kern not in s["name"]: continue
    h = s["hdr"]; ia, isamp, iinst, ithr = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    base = int(s["rows"][0][ia], 16)
    agg = {}
    tot_s = tot_i = 0
    stall_tot = {h[i]: 0 for i in stall_cols}
    for r in s["rows"]:
        off = int(r[ia], 16) - base
        ln = line_of.get(off, "?")
        a = agg.setdefault(ln, [0, 0, 0, {}])
        sm, ni, nt = int(r[isamp]), int(r[iinst]), int(r[ithr])
        a[0] += sm; a[1] += ni; a[2] += nt
        for i in stall_cols:
            v = int(r[i]) if r[i] else 0
            if v:
                a[3][h[i]] = a[3].get(h[i], 0) + v; stall_tot[h[i]] += v
        tot_s += sm; tot_i += ni
    print(f"kernel {s['name']}: {tot_i} warp instructions, {tot_s} samples")
    print("stall mix:", {k: f"{100*v/max(1,sum(stall_tot.values())):.1f}%" for k, v in sorted(stall_tot.items(), key=lambda x: -x[1])[:8]})
    for ln, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        ts = sorted(a[3].items(), key=lambda x: -x[1])[:2]
        print(f"{100*a[0]/tot_s:5.1f}% samples {100*a[1]/tot_i:5.1f}% inst  lanes {a[2]/max(1,a[1]):4.1f}  {ln:28s} {ts}")
