"""Developer script: SASS mnemonic counts of the built library (cuobjdump -sass) -> profiles/rNN_sass_counts.txt
    python tools/sass_counts.py > profiles/r02_sass_counts.txt"""
import collections, os, re, subprocess, sys
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hermipy_b200", "libhermite_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = {"UTMALDG": r"\bUTMALDG", "UBLKCP": r"\bUBLKCP", "UTMASTG": r"\bUTMASTG", "SYNCS": r"\bSYNCS", "DMMA": r"\bDMMA",
       "UTCxMMA": r"\bUTC[A-Z]*MMA", "LDTM": r"\bLDTM", "BAR": r"\bBAR\.", "MEMBAR/FENCE": r"\bMEMBAR|\bFENCE", "ATOMG/RED": r"\bATOMG|\bRED\."}
tot, rows = collections.Counter(), []
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    c = {k: len(re.findall(v, f)) for k, v in pat.items()}
    tot.update(c)
    if c["DMMA"] or c["UTMALDG"] or c["UBLKCP"]:
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        rows.append((re.sub(r"\(CUtensorMap_st.*", "", dem), c))
print("SASS mnemonic counts of hermipy_b200/libhermite_b200.so (cuobjdump -sass, sm_100a)")
print("FP64 tensor cores on sm_100a are reached through mma.sync m8n8k4 f64 -> DMMA.8x8x4 (tcgen05.mma has no f64 kind: no")
print("UTCxMMA / LDTM expected); operands arrive by TMA (UTMALDG) behind mbarriers (SYNCS); the row-block scatter of the")
print("sharded transform leaves by bulk asynchronous copies shared -> global (UBLKCP).\n")
print("whole library: " + ", ".join("%s %d" % kv for kv in tot.items()) + "\n")
print("per kernel (those with DMMA / TMA / bulk-copy instructions):")
for dem, c in rows:
    print("  %-95s DMMA %4d  UTMALDG %3d  UBLKCP %3d  SYNCS %3d" % (dem[:95], c["DMMA"], c["UTMALDG"], c["UBLKCP"], c["SYNCS"]))
