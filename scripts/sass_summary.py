"""profiles/sass_summary.txt: per kernel of libdfvit.so, the counts of the SASS mnemonics that prove which hardware
path it uses (B200_PROFILING.md): UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG / UTMASTG (TMA tensor load / store),
UBLKCP (bulk copy), UTCBAR (tcgen05.commit), SYNCS (mbarrier), FHFMA (mixed bf16 x bf16 + fp32 FMA), FFMA2 / HMUL2.BF16."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "deepfake_vit_b200", "libdfvit.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
WANT = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "FHFMA", "FFMA2", "HMUL2", "HFMA2", "FFMA", "MUFU", "LDS", "STS", "LDG", "STG", "RED", "ATOM"]
rows, cur, cnt = [], None, None
it = iter(names)
for line in sass.split("\n"):
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        if cur:
            rows.append((cur, cnt))
        cur, cnt = next(it), collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        for w in WANT:
            if op == w or (w in ("HMUL2", "HFMA2") and op.startswith(w)):
                cnt[w] += 1
if cur:
    rows.append((cur, cnt))


def short(n):
    n = re.sub(r"\(.*", "", n).replace("dfv::", "").replace("(anonymous namespace)::", "")
    return n[:100]


agg = collections.OrderedDict()
for n, c in rows:
    base = re.sub(r"<.*", "", short(n))
    a = agg.setdefault(base, [0, collections.Counter()])
    a[0] += 1
    for k, v in c.items():
        a[1][k] = max(a[1][k], v)
out = ["libdfvit.so (sm_100a) -- SASS evidence per kernel family: max count over the family's template instances", ""]
out.append(f"{'kernel family':38s} {'inst':>4s} " + " ".join(f"{w:>7s}" for w in WANT))
for base, (n, c) in agg.items():
    out.append(f"{base[:38]:38s} {n:4d} " + " ".join(f"{c[w]:7d}" for w in WANT))
tc = [b for b, (n, c) in agg.items() if c["UTCHMMA"]]
tma = [b for b, (n, c) in agg.items() if c["UTMALDG"] or c["UTMASTG"] or c["UBLKCP"]]
out += ["", "tcgen05 tensor-core kernels (UTCHMMA + LDTM): " + ", ".join(tc), "TMA / bulk-copy kernels: " + ", ".join(tma)]
text = "\n".join(out) + "\n"
dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_summary.txt")
open(dst, "w").write(text)
print(text)
