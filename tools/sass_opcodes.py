#!/usr/bin/env python
"""Per-kernel counts of the SASS opcodes that prove the Blackwell paths (cuobjdump -sass of libawqk.so):
UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA tensor load), UBLKCP (1-D bulk TMA), UTCBAR (tcgen05.commit),
SYNCS (mbarrier), FFMA2/FMUL2/FADD2 (packed fp32x2), REDUX, and the register count per kernel.
    python tools/sass_opcodes.py > profiles/sass_opcodes.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "awq-converter_b200", "libawqk.so")
KEYS = ["UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "FFMA2", "FMUL2", "FADD2", "REDUX", "VIADDMNMX",
        "VHMNMX", "FMNMX3", "MUFU.RCP", "STG", "LDG", "ATOM", "RED"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
regs = {}
for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+)", res):
    regs[m.group(1)] = int(m.group(2))
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
counts, order, cur = {}, [], None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
print(f"# {os.path.relpath(lib, ROOT)}: {len(order)} kernels, cubin architectures: {', '.join(arch)}")
print("# kernel | registers | SASS instructions | opcode counts (only the listed families)")
for fn in sorted(order, key=lambda f: (-counts[f]["UTCHMMA"], -counts[f]["UTMALDG"], -counts[f]["UBLKCP"], -counts[f]["_total"])):
    c = counts[fn]
    d = demangle(fn)
    name = (d.split(">(")[0] + ">") if ">(" in d else d.split("(")[0]
    fam = " ".join(f"{k}={c[k]}" for k in KEYS if c[k])
    print(f"{name} | {regs.get(fn, '?')} | {c['_total']} | {fam}")
