"""Development tool: SASS instruction count per kernel of the built library (and TMA/bulk-copy mnemonics).
usage: python tools/sass_sizes.py [path/to/libygl_b200.so]"""
import collections
import re
import subprocess
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "yocto-gl_b200/lib/libygl_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
cnt, special, name = collections.Counter(), collections.defaultdict(collections.Counter), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(\S+)", line)
    if name and m:
        cnt[name] += 1
        op = m.group(1).split(".")[0]
        if op in ("UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS", "SYNCS", "LDTM", "STTM") or op.startswith("UTC"):
            special[name][op] += 1
names = list(cnt)
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.split("\n")
for n, d in sorted(zip(names, dem), key=lambda x: cnt[x[0]]):
    print(f"{cnt[n]:7d}  {d.split('(')[0][:70]:70s} {dict(special[n]) if special[n] else ''}")
