#!/usr/bin/env python3
"""Per-source-line attribution of an ncu report (--set full --import-source on): instructions, active threads per
instruction, stall samples.  Usage: tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None
hdr = None
out = []
for r in csv.reader(txt.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) >= 2 and r[0] == "Line No":
        hdr = r
        continue
    if not hdr or len(r) != len(hdr) or not r[0]:
        continue
    i_ie = hdr.index("Instructions Executed"); i_te = hdr.index("Thread Instructions Executed"); i_s = hdr.index("# Samples")
    i_lsb = hdr.index("stall_long_sb"); i_ni = hdr.index("stall_no_inst"); i_w = hdr.index("stall_wait")
    try:
        ie, te, smp = int(r[i_ie]), int(r[i_te]), int(r[i_s])
        lsb, ni, w = int(r[i_lsb]), int(r[i_ni]), int(r[i_w])
    except ValueError:
        continue
    if ie > 0 or smp > 0:
        out.append((ie, te, smp, lsb, ni, w, cur, r[0], r[1][:100]))
tot = sum(o[0] for o in out) or 1
tots = sum(o[2] for o in out) or 1
print(f"total warp instructions {tot}, thread instructions {sum(o[1] for o in out)}, threads/instr {sum(o[1] for o in out)/tot:.2f}, samples {tots}")
print(f"stall samples: long_sb {100*sum(o[3] for o in out)/tots:.1f}%  no_inst {100*sum(o[4] for o in out)/tots:.1f}%  wait {100*sum(o[5] for o in out)/tots:.1f}%")
out.sort(reverse=True)
for ie, te, smp, lsb, ni, w, f, l, s in out[:top]:
    print(f"{100*ie/tot:5.1f}% instr  {te/max(ie,1):5.1f} thr/instr  {100*smp/tots:5.1f}% samples (lsb {100*lsb/tots:4.1f})  {f}:{l}  {s}")
