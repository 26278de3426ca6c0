"""SASS evidence for the design claims (cuobjdump -sass of the objects build() compiled): per kernel, how often the mnemonics that
matter occur, plus the instructions themselves.  python tools/sass_evidence.py > profiles/r2_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "rend3_b200", "csrc")
CLAIMS = [
    ("r3_cull_bake.o", "cull_bake_kernel", ["LDG.E.128", "STG.E.128", "STG.E.EF.128", "LDG.E.EF.128", "LDS", "STS", "SHFL", "BAR.SYNC", "VOTE"],
     "stream kernel: 128-bit coalesced loads / stores, operands from the constant bank, no shuffles; shared memory only for the per-CTA survivor count"),
    ("r3_cull_bake.o", "compact_visible_kernel", ["ST.E.STRONG.SYS", "STG.E.STRONG.SYS", "MEMBAR.SC.SYS", "MEMBAR.ALL.SYS", "ATOMG", "RED"],
     "fused exchange: plain peer stores, system-scope fence, release store of the epoch flag by the last CTA"),
    ("r3_cull_bake.o", "exchange_wait_kernel", ["LD.E.STRONG.SYS", "LDG.E.STRONG.SYS", "NANOSLEEP"], "consumer: acquire load of the epoch flag at system scope"),
    ("r3_tri_cull.o", "triangle_test_kernel", ["UBLKCP", "SYNCS.ARRIVE.TRANS64", "SYNCS.PHASECHK", "BAR.SYNC", "LDS", "LDG.E.128", "MUFU.RCP", "FCHK", "ATOMG", "RED"],
     "index runs staged by 1-D bulk async copies (TMA unit) on per-warp mbarriers; no block barrier"),
    ("r3_raster.o", "raster_setup_kernel", ["RED.E.MAX.64", "REDG.E.MAX.64", "RED.E.MAX", "ATOMG"], "visibility buffer: fire-and-forget 64-bit RED.MAX per covered sample"),
    ("r3_raster.o", "raster_band_kernel", ["RED.E.MAX.64", "REDG.E.MAX.64", "RED.E.MAX", "ATOMG"], "same for the band kernel"),
    ("r3_peer.o", "peer_copy_kernel", ["LDG.E.128", "STG.E.128", "ST.E.128"], "peer copies: 128-bit loads, 128-bit stores into the mapped peer buffers"),
    ("r3_peer.o", "peer_signal_kernel", ["STRONG.SYS", "MEMBAR"], "epoch flag: system-scope fence + release store"),
    ("r3_peer.o", "peer_wait_kernel", ["STRONG.SYS", "NANOSLEEP"], "flag wait: acquire loads at system scope"),
]


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(CSRC, obj)], capture_output=True, text=True).stdout
    fns, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            fns[cur] = []
        elif cur and "/*" in line and ";" in line:
            fns[cur].append(line.split("*/", 1)[1].split(";")[0].strip())
    return fns


def main():
    cache = {}
    print("SASS evidence (cuobjdump -sass, sm_100a objects of this tree; `tools/sass_evidence.py`)\n")
    for obj, kernel, mnemonics, claim in CLAIMS:
        if not os.path.exists(os.path.join(CSRC, obj)):
            print(f"{obj}: not built\n")
            continue
        fns = cache.setdefault(obj, functions(obj))
        for name, ins in fns.items():
            if kernel not in name:
                continue
            demangled = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
            print(f"== {demangled[:150]}\n   claim: {claim}\n   {len(ins)} SASS instructions")
            for m in mnemonics:
                hits = [i for i in ins if m in i]
                print(f"   {m:24s} {len(hits):5d}" + (f"   e.g. {hits[0][:90]}" if hits else ""))
            print()


if __name__ == "__main__":
    main()
