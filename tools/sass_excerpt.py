"""Write profiles/<out>.md: per-kernel counts of the Blackwell-native SASS instructions (tcgen05.mma = UTCHMMA, TMA = UTMALDG /
UTMASTG, tensor-memory loads / stores = LDTM / STTM, tcgen05.commit = UTCBAR) in the built library.
  python tools/sass_excerpt.py profiles/r2_sass_excerpt.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "drl-energy-optimal-routing-for-electric-vehicles_b200", "libevrp_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, cnt, ex = None, collections.OrderedDict(), {}
pat = re.compile(r'\b(UTCHMMA|UTCQMMA|UTCIMMA|UTCMMA|UTMALDG|UTMASTG|UTMAPF|LDTM|STTM|UTCBAR|UTCATOMSWS|UTCCP)\b[^\s;]*')
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        fn = m.group(1); continue
    m = pat.search(line)
    if m and fn:
        cnt.setdefault(fn, collections.Counter())[m.group(0)] += 1
        ex.setdefault((fn, m.group(1)), re.sub(r'/\*[0-9a-f]{4}\*/', '', line).strip().split(';')[0])
dem = lambda n: subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip().split('(')[0]
out = ["# SASS evidence of the Blackwell-native instructions in libevrp_b200.so (cuobjdump -sass, sm_100a; counts per kernel)\n",
       "`UTCHMMA` = tcgen05.mma (kind::f16 / kind::tf32), `UTMALDG` / `UTMASTG` = TMA tensor load / store, `LDTM` / `STTM` = tcgen05.ld / "
       "tcgen05.st (tensor memory), `UTCBAR` = tcgen05.commit -> mbarrier, `UTCATOMSWS` = tcgen05.alloc / dealloc.\n",
       "| kernel | instruction counts | one example line |", "|---|---|---|"]
for f, c in cnt.items():
    key = [k for k in ex if k[0] == f and 'MMA' in k[1]] or [k for k in ex if k[0] == f]
    out.append(f"| `{dem(f)}` | " + ", ".join(f"{k} × {v}" for k, v in sorted(c.items(), key=lambda kv: -kv[1])) + f" | `{ex[key[0]][:110]}` |")
open(sys.argv[1], "w").write("\n".join(out) + "\n")
print(f"{len(cnt)} kernels -> {sys.argv[1]}")
