#!/usr/bin/env python
"""SASS excerpts of the hot loops (cuobjdump -sass on the built objects; no GPU needed):
    python tools/sass_excerpts.py > profiles/r02_sass_excerpts.txt
For every kernel: the instruction mix of the whole function and the listing around the loop that matters."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = os.path.join(ROOT, "gpu-gmres_b200", "build")
WANT = [
    ("orth.o", r"orth_kernelIddE", "UBLKCP", 6, 10, "fused CGS step, FP64: bulk-async (TMA unit) tile loads"),
    ("orth.o", r"orth_kernelIddE", "DFMA", 8, 24, "fused CGS step, FP64: consumer inner loop (LDS.64 + DFMA per element)"),
    ("orth.o", r"orth_kernelIdfE", "FFMA", 8, 24, "fused CGS step, FP32 basis under FP64 vectors: projections as FFMA, no F2F per element"),
    ("orth.o", r"mgs_kernelIdE", "UBLKCP", 4, 8, "fused strict MGS: basis slices by bulk copy"),
    ("trsv_block.o", r"trsv_block_fusedIdddLb0ELb1E", "SHFL.UP", 14, 60, "block-Jacobi lattice step (L): 2 SHFL + 2 DFMA + LDS.64/LDS.128/STS.64 + selects per step"),
    ("trsv.o", r"trsv_lattice_globalIddddLb1ELi128E", "SHFL.UP", 10, 40, "global wavefront strip, chain warp (L)"),
    ("trsv.o", r"trsv_syncfreeIdddLb0E", "VOTE", 30, 12, "sync-free slices: warp-uniform poll of 16-byte packets (LD.E.128.STRONG + VOTE.ALL)"),
    ("spmv.o", r"spmv_sell_wIdLi5ELb0ELb0E", "LDG.E.NA", 2, 16, "fixed-width sliced-ELL SpMV (FP64: int2 + double2 per thread and entry): every matrix load issued before the first gather"),
]
def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(B, obj)], capture_output=True, text=True).stdout
    cur, name = [], None
    res = {}
    for ln in out.split("\n"):
        m = re.search(r"Function : (\S+)", ln)
        if m:
            if name: res[name] = cur
            name, cur = m.group(1), []
        elif re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
            cur.append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", ln).rstrip())
    if name: res[name] = cur
    return res
cache = {}
for obj, fre, pat, before, after, what in WANT:
    if obj not in cache: cache[obj] = functions(obj)
    hit = [k for k in cache[obj] if re.search(fre, k)]
    if not hit:
        print(f"== {obj}: no function matching {fre}\n"); continue
    name = hit[0]
    lines = cache[obj][name]
    ops = collections.Counter(re.sub(r"@!?U?P\d+\s+", "", l.split("*/", 1)[1].strip()).split(" ")[0].rstrip(";") for l in lines)
    idx = [i for i, l in enumerate(lines) if pat in l]
    print(f"== {what}\n   {obj}: {name}\n   {len(lines)} instructions; " + ", ".join(f"{k} x{v}" for k, v in ops.most_common(14)))
    if not idx:
        print(f"   (no {pat} in this function)\n"); continue
    i0 = idx[0]
    for l in lines[max(0, i0 - before): i0 + after]:
        print(l)
    print()
