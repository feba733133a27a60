"""Per kernel of libspb200.so: how many SASS instruction sites use the Blackwell tensor / TMA / TMEM paths.
   python profiles/sass_table.py > profiles/r02_sass_mnemonics.txt"""
import collections
import re
import subprocess

SO = "scoreperformer_b200/csrc/libspb200.so"
COLS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "HMMA", "FFMA2", "FADD2", "FMUL2", "MUFU", "SYNCS"]
sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
cur, table = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        cur = table.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        for c in COLS:
            if op.startswith(c):
                cur[c] += 1
print(f"{'kernel':58s} " + " ".join(f"{c:>8s}" for c in COLS))
for name, cnt in table.items():
    if any(cnt[c] for c in COLS[:7]):
        print(f"{name[:58]:58s} " + " ".join(f"{cnt[c]:8d}" for c in COLS))
