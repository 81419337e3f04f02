"""SASS opcode histogram of the product library (cuobjdump -sass, runs without a GPU): the tensor-core / TMA / TMEM /
multimem opcodes per kernel, and the check that no legacy HMMA (mma.sync) exists.  python tools/sass_opcodes.py > profiles/..."""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pesr_b200", "libpesr_b200.so")
KEEP = re.compile(r"^(UTCHMMA|UTCBAR|UTMALDG|UTMASTG|UTMACCTL|LDTM|STTM|UTCATOMSWS|SYNCS|LDGMC|STGMC|REDGMC|HMMA|UCGABAR|MEMBAR\.ALL\.SYS|"
                  r"LDG\.E\.128\.STRONG\.SYS|STG\.E\.128\.STRONG\.SYS|LD\.E\.STRONG\.SYS|ST\.E\.STRONG\.SYS)")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kernels, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur is not None:
        cur["__total__"] += 1
        op = m.group(1)
        if KEEP.match(op):
            cur[op] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
print("# SASS opcode evidence, pesr_b200/libpesr_b200.so (cuobjdump -sass, sm_100a), tools/sass_opcodes.py")
print("# tensor-core / TMA / TMEM / multimem opcodes per kernel; HMMA (legacy mma.sync) must be absent")
hmma = 0
for name, cnt in zip(demangle, kernels.values()):
    ops = " ".join(f"{k}={v}" for k, v in sorted(cnt.items()) if k != "__total__")
    hmma += sum(v for k, v in cnt.items() if k.startswith("HMMA"))
    if ops:
        print(f"{name}\n   total instructions {cnt['__total__']} {ops}")
print(f"# kernels in the library: {len(kernels)}; HMMA instructions anywhere: {hmma}")
sys.exit(1 if hmma else 0)
