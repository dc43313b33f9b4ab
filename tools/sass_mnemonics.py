"""SASS mnemonic counts of the hot kernels in the built library (cuobjdump -sass) -> profiles/r02_sass_mnemonics.txt.
Runs without a GPU.  The mnemonics that prove the Blackwell paths: UTCHMMA (tcgen05.mma), UTMALDG / UTMASTG / UTMAREDG (TMA),
LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bsc-nav_b200", "libbscnav_b200.so")
WATCH = ("UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "HMMA", "LDSM", "MUFU", "DFMA", "DMUL",
         "FMNMX3", "ATOMG", "REDG", "SYNCS")
HOT = re.compile(r"k_gemm_f16|k_attention|k_layernorm|k_preprocess_tiled|k_index_points|k_select|k_commit|k_count_new|k_assign_slots|"
                 r"k_query_slots|k_blk_select|k_batch_rescore|k_build_shadow|k_publish_counts")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
names = {}
for m in re.finditer(r"Function : (\S+)", sass):
    names[m.group(1)] = None
dem = subprocess.run(["cu++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
names = dict(zip(names, dem))
out = ["SASS mnemonic counts of the hot kernels in bsc-nav_b200/libbscnav_b200.so (cuobjdump -sass; sm_100a; tools/sass_mnemonics.py).",
       "UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,",
       "HMMA / LDSM = legacy warp-level mma.sync / ldmatrix (attention tail rows, agent towers, one-frame attention), SYNCS = mbarrier ops.", ""]
cur, counts, total = None, None, 0
def flush():
    if cur and HOT.search(cur):
        hits = ", ".join(f"{k} {counts[k]}" for k in WATCH if counts.get(k))
        out.append(re.sub(r"\(.*", "", re.sub(r"\((?:int|bool)\)", "", cur))); out.append(f"    total {total}: {hits}")
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush(); cur, counts, total = names[m.group(1)], collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        total += 1; counts[m.group(1)] += 1
flush()
open(os.path.join(ROOT, "profiles", "r02_sass_mnemonics.txt"), "w").write("\n".join(out) + "\n")
print(len(out) // 2 - 2, "kernels listed")
