"""SASS mnemonic evidence: counts of the Blackwell-specific instructions per kernel of libmsw_b200.so.
    python tools/sass_mnemonics.py > profiles/<name>_sass_mnemonics.txt   (no GPU needed: cuobjdump -sass)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "minesweeper_ppo_b200", "libmsw_b200.so")
WANT = {"UTCHMMA", "UTCQMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "LDGSTS", "UTCBAR", "ELECT", "UTMAPF"}
out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
cur, cnt = None, collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur and m.group(1) in WANT:
        cnt[(cur, m.group(1))] += 1
mangled = sorted({k[0] for k in cnt})
dem = subprocess.run(["c++filt"] + mangled, stdout=subprocess.PIPE, text=True).stdout.splitlines()
names = dict(zip(mangled, [re.sub(r"\(.*", "", d) for d in dem]))
print("# SASS mnemonics per kernel of libmsw_b200.so (cuobjdump -sass, sm_100a; counts of static instructions)")
print("# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTMALDG / UTMASTG = cp.async.bulk.tensor load / store,")
print("# UTMAPF = prefetch.tensormap, UBLKCP = cp.async.bulk, SYNCS = mbarrier ops, ELECT = elect.sync, HMMA = mma.sync, LDGSTS = cp.async")
for (k, op), c in sorted(cnt.items(), key=lambda kv: (names[kv[0][0]], kv[0][1])):
    print(names[k], op, c)
