"""Count the Blackwell-specific SASS mnemonics per kernel of csrc/libcomemb_b200.so (cuobjdump -sass): tcgen05.mma ->
UTCHMMA, tcgen05.ld -> LDTM, TMEM alloc -> UTCATOMSWS, tcgen05.commit -> UTCBAR, cp.async.bulk (TMA engine) -> UBLKCP,
mbarrier -> SYNCS; legacy mma.sync -> HMMA.  Writes profiles/r2_sass_evidence.txt."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "nodeembedding-to-communityembedding_b200", "csrc", "libcomemb_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCHMMA|UBLKCP\.S\.G|LDTM[.\w]*|UTCBAR|UTCATOMSWS[.\w]*|SYNCS[.\w]*|HMMA[.\w]*|RED\.E[.\w]*|REDG[.\w]*)")
counts, fn = collections.defaultdict(collections.Counter), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = pat.search(line)
    if m and fn:
        counts[fn][m.group(1)] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
out = ["# SASS mnemonics per kernel (cuobjdump -sass libcomemb_b200.so); kernels without any of them are omitted",
       "# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCATOMSWS = tcgen05.alloc/dealloc, UTCBAR = tcgen05.commit,",
       "# UBLKCP.S.G = cp.async.bulk global->shared (TMA engine), SYNCS = mbarrier, HMMA = legacy mma.sync", ""]
for mangled, name in sorted(zip(counts, names), key=lambda t: t[1]):
    c = counts[mangled]
    if not any(k.startswith(("UTC", "UBLKCP", "LDTM", "HMMA")) for k in c):
        continue
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(.*", "", name)
    out.append("%-70s %s" % (name, "  ".join("%s x%d" % kv for kv in sorted(c.items()))))
open(os.path.join(ROOT, "profiles", "r2_sass_evidence.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:40]))
