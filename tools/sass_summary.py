#!/usr/bin/env python3
"""Mnemonic counts per kernel of the shipped library (cuobjdump -sass), the SASS evidence kept under profiles/.

  python tools/sass_summary.py [radiodsp_sdr_rx_b200/librdsp_gpu.so] > profiles/rNNx_sass.txt
"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "radiodsp_sdr_rx_b200/librdsp_gpu.so"
TAGS = ["UTCIMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "LDGSTS", "SYNCS", "UTCATOMSWS", "FFMA2", "FADD2", "FMUL2", "F2I.S16", "FCHK", "ACQBULK"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
fn, counts, total = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        total[fn] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        total[fn] += 1
        op = m.group(1)
        for t in TAGS:
            if op == t or op.startswith(t + "."):
                counts[fn][t] += 1
print(f"# SASS evidence: cuobjdump -sass {LIB}, mnemonic counts per kernel (sm_100a); tools/sass_summary.py")
print("# UTCIMMA = tcgen05.mma kind::i8, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTMALDG = cp.async.bulk.tensor (TMA), LDGSTS = cp.async,")
print("# FFMA2/FADD2/FMUL2 = packed f32x2, F2I.S16 = cvt.rzi.s16.f32 (arm_float_to_q15 in one instruction), FCHK = IEEE-division slow-path check, ACQBULK = griddepcontrol.wait\n")
for fn in counts:
    short = re.sub(r"^_ZN\d+_GLOBAL__N__[0-9a-f]+_\d+_", "", fn)
    print(f"{short:<92} instr {total[fn]:6d}  " + "  ".join(f"{t} {n}" for t, n in counts[fn].items()))
