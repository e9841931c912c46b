"""Count the SASS mnemonics that prove which hardware paths the shipped kernels use.

  python profiles/sass_mnemonics.py [donk_b200/csrc/libdonk_b200.so] > profiles/rNN/sass_mnemonics.txt

UTCHMMA = tcgen05.mma (kind::tf32 / f16), LDTM / STTM = tcgen05.ld / .st (tensor memory), UTMALDG = cp.async.bulk.tensor
(TMA tile load), UBLKCP = cp.async.bulk (1-D TMA copy), SYNCS = mbarrier, UTCBAR = tcgen05.commit, DMMA = mma.sync f64,
LDGSTS = cp.async.  (Mnemonics per /opt/skills/guides/B200_PROFILING.md.)
"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "donk_b200/csrc/libdonk_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
KEYS = ("UTCHMMA", "UTCMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS", "DMMA", "DFMA",
        "LDGSTS")
cur, cnt = None, collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        for key in KEYS:
            if m.group(1).startswith(key):
                cnt[cur][key] += 1
names = subprocess.run(["c++filt"], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
total = collections.Counter()
for (f, c), name in sorted(zip(cnt.items(), names), key=lambda x: x[1]):
    total.update(c)
    print(f"{name[:120]:120s} " + " ".join(f"{k}={v}" for k, v in sorted(c.items())))
print("TOTAL " + " ".join(f"{k}={v}" for k, v in sorted(total.items())))
