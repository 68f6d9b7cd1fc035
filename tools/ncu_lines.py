#!/usr/bin/env python
"""Join an ncu SASS source page (ncu -i X.ncu-rep --page source --csv) with nvdisasm -g line info of the
same cubin and aggregate executed warp instructions / stall samples per source line.

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep k_rollout [top_n] [kernel_index]
"""
import csv
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gym_pcgrl_b200", "csrc", "libpcgrl_b200.so")


def sass_lines(kernel_substr):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, stdout=subprocess.DEVNULL)
    txt = ""   # one cubin per translation unit: scan them all
    for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
        txt += subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout + "\n"
    funcs, cur, line = {}, None, None
    for ln in txt.split("\n"):
        m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = {}
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and cur is not None:
            funcs[cur][int(m.group(1), 16)] = (line, m.group(2).strip())
    return funcs


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    kidx = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.split("\n")))
    blocks, cur = [], None
    for r in rows:
        if len(r) >= 2 and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif r and r[0] == "Address" and cur is not None:
            cur["hdr"] = r
        elif cur is not None and r and r[0].startswith("0x"):
            cur["rows"].append(r)
    blocks = [b for b in blocks if ksub in b["name"]]
    b = blocks[kidx]
    hdr = b["hdr"]
    ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = int(b["rows"][0][0], 16)
    funcs = sass_lines(ksub)
    # pick the SASS function whose mangled name contains the kernel substring and instantiation
    # template arguments "(int)0, (int)-1" -> mangled "ILi0ELin1EE"
    targs = re.findall(r"\(int\)(-?\d+)", b["name"].split("(pcgrl_config")[0])
    mang = ("I" + "".join("Li%sE" % (a.replace("-", "n")) for a in targs) + "E") if targs else None
    cands = [f for f in funcs if (ksub + (mang or "")) in f]
    if not cands:   # namespaced kernels (pcgrl_smb::...)
        cands = [f for f in funcs if ksub in f]
    fn = funcs[cands[0]]
    stall_cols = [c for c in ("stall_no_inst", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_branch_resolving", "stall_lg")
                  if c in hdr] if os.environ.get("NCU_LINES_STALLS") else []
    stall_idx = [hdr.index(c) for c in stall_cols]
    per_line = collections.defaultdict(lambda: [0, 0] + [0] * len(stall_cols))
    total_i = total_s = 0
    for r in b["rows"]:
        off = int(r[0], 16) - base
        n, s = int(r[ie] or 0), int(r[isamp] or 0)
        line = fn.get(off, (("?", 0), ""))[0] or ("?", 0)
        per_line[line][0] += n
        per_line[line][1] += s
        for q, ci in enumerate(stall_idx):
            per_line[line][2 + q] += int(r[ci] or 0)
        total_i += n
        total_s += s
    # NCU_LINES_FILE=<substring>: keep only source files matching (percentages become relative to the kept lines);
    # NCU_LINES_SORT=samp: rank by stall samples instead of executed instructions
    flt, by = os.environ.get("NCU_LINES_FILE"), (1 if os.environ.get("NCU_LINES_SORT") == "samp" else 0)
    if flt:
        per_line = {k: v for k, v in per_line.items() if flt in k[0]}
        total_i, total_s = sum(v[0] for v in per_line.values()), sum(v[1] for v in per_line.values())
    print("kernel:", b["name"][:80], "| SASS:", cands[0][:60])
    print("total warp-instructions %d, stall samples %d" % (total_i, total_s))
    src_cache = {}
    if stall_cols:   # NCU_LINES_STALLS=1: per-line stall reasons (sample counts) after the percentages
        print("stall columns:", " ".join(c.replace("stall_", "") for c in stall_cols))
    for (f, l), vals in sorted(per_line.items(), key=lambda kv: -kv[1][by])[:topn]:
        n, s = vals[0], vals[1]
        if f not in src_cache:
            p = os.path.join(ROOT, "gym_pcgrl_b200", "csrc", f)
            src_cache[f] = open(p).read().split("\n") if os.path.exists(p) else []
        text = src_cache[f][l - 1].strip()[:90] if 0 < l <= len(src_cache[f]) else ""
        extra = (" [" + " ".join("%d" % v for v in vals[2:]) + "]") if stall_cols else ""
        print("%6.2f%% inst %6.2f%% samp%s  %-20s %s" % (100.0 * n / max(total_i, 1), 100.0 * s / max(total_s, 1), extra, "%s:%d" % (f, l), text))


if __name__ == "__main__":
    main()
