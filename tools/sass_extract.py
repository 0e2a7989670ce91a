#!/usr/bin/env python
"""SASS evidence for profiles/: for every hot kernel, the static instruction mix of its SASS (cuobjdump -sass of the
in-tree objects) and the instructions that show which hardware paths it uses (UBLKCP / SYNCS = bulk async copy +
mbarrier, VOTE / R2P = ballot ranking, ATOMS = shared-memory atomics, DADD / DFMA = fp64 vote, IMAD.WIDE / LOP3 = Philox).
usage: tools/sass_extract.py OUT.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
BUILD = ROOT / "precise-mrd-mini_b200" / "build"
HOT = {
    "radix_sort.o": ["rs_downsweep_direct_kernel", "rs_downsweep_staged_kernel", "rs_upsweep_tiles_kernel", "rs_cellscan_kernel", "rs_downsweep_kernel"],
    "collapse.o": ["cell_consensus_peel_kernel", "cell_consensus_bulk_kernel", "cell_consensus_kernel", "small_cell_kernel", "group_rows_kernel", "family_kernel"],
    "reads.o": ["reads_gen_kernel", "reads_size_kernel"],
    "fused.o": ["fused_cell_kernel", "fused_replay_kernel"],
    "pvalues.o": ["pvalues_kernel", "bh_group_kernel"],
    "lod.o": ["lod_fit_warp_kernel"],
}
MARK = ["UBLKCP", "SYNCS", "VOTE", "R2P", "ATOMS", "ATOMG", "RED", "DADD", "DFMA", "DMUL", "IMAD.WIDE", "LOP3", "SHFL", "LDS", "STS", "LDG", "STG",
        "MATCH", "BAR", "POPC", "STL", "LDL"]


def main():
    out = []
    for obj, names in HOT.items():
        sass = subprocess.run(["cuobjdump", "-sass", str(BUILD / obj)], capture_output=True, text=True).stdout
        res = subprocess.run(["cuobjdump", "-res-usage", str(BUILD / obj)], capture_output=True, text=True).stdout
        usage = {}
        fn = None
        for line in res.splitlines():
            m = re.search(r"Function (\S+):", line)
            if m:
                fn = m.group(1)
            elif fn and "REG:" in line:
                usage[fn] = line.strip()
                fn = None
        cur, body = None, collections.defaultdict(list)
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = m.group(1)
                continue
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", line)
            if cur and m:
                body[cur].append(m.group(1).strip())
        for mangled, ins in body.items():
            demangled = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
            base = demangled.split("(")[0].replace("void ", "")
            if not any(base.startswith(n) or ("<" in base and base.split("<")[0].endswith(n)) for n in names):
                continue
            ops = collections.Counter()
            for i in ins:
                t = i.split()
                op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
                ops[op.split(".")[0] if not op.startswith("IMAD.WIDE") else "IMAD.WIDE"] += 1
            marks = {k: sum(v for o, v in collections.Counter(
                (x.split()[1] if x.split()[0].startswith("@") and len(x.split()) > 1 else x.split()[0]) for x in ins).items() if o.startswith(k)) for k in MARK}
            out.append(f"== {base}")
            out.append(f"   {usage.get(mangled, '')}")
            out.append(f"   static SASS instructions: {len(ins)}")
            out.append("   marks: " + "  ".join(f"{k}={v}" for k, v in marks.items() if v))
            out.append("   top opcodes: " + "  ".join(f"{k}={v}" for k, v in ops.most_common(14)))
            if any(k in ("UBLKCP",) and v for k, v in marks.items()):
                out.append("   bulk-copy lines:")
                out += ["      " + x for x in ins if x.split()[-1 if False else 0].startswith(("UBLKCP", "SYNCS")) or " UBLKCP" in x or " SYNCS" in x][:8]
    Path(sys.argv[1]).write_text("\n".join(out) + "\n")
    print(f"wrote {sys.argv[1]} ({len(out)} lines)")


if __name__ == "__main__":
    main()
