"""SASS opcode histogram of the built library (CPU only: cuobjdump).
  python tools/sass_histogram.py > profiles/r02/sass_opcode_histogram.txt
Counts instruction mnemonics over every kernel of big_vision_b200/libbv_b200.so; the Blackwell-native ones are
listed first (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMA* = TMA, UTCBAR = tcgen05.commit), then the
legacy tensor paths (must be absent), then per-kernel local-memory (spill) instruction counts."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "big_vision_b200", "libbv_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = lambda names: subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.split("\n")

ops, full = collections.Counter(), collections.Counter()
per_kernel_local = collections.Counter()
kernels, cur = [], None
ins = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[0-9T]\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)")
for line in sass.split("\n"):
  m = re.match(r"\s+Function : (\S+)", line)
  if m:
    cur = m.group(1)
    kernels.append(cur)
    continue
  m = ins.match(line)
  if m:
    ops[m.group(1)] += 1
    full[m.group(1) + m.group(2)] += 1
    if m.group(1) in ("LDL", "STL"):
      per_kernel_local[cur] += 1

print(f"# SASS opcode histogram of big_vision_b200/libbv_b200.so (cuobjdump -sass, sm_100a), {len(kernels)} kernels,")
print(f"# {sum(ops.values())} instructions; tools/sass_histogram.py")
print("# -- Blackwell-native paths")
rows = [("UTCHMMA", "tcgen05.mma kind::f16"), ("UTCHMMA.2CTA", "  of which cta_group::2"), ("LDTM", "tcgen05.ld"),
        ("STTM", "tcgen05.st (P into tensor memory)"), ("UTCBAR", "tcgen05.commit"), ("UTMALDG", "TMA load"),
        ("UTMASTG", "TMA store"), ("UTMAREDG", "TMA reduce-add"), ("UTMAPF", "TMA descriptor prefetch"),
        ("SYNCS", "mbarrier"), ("ELECT", "elect.sync"), ("USETMAXREG", "setmaxnreg"), ("UCGABAR_ARV", "cluster barrier"),
        ("LDGSTS", "cp.async"), ("MUFU.EX2", "ex2.approx"), ("MUFU.TANH", "tanh.approx"), ("REDG", "global reduction (atomicAdd, no return)"),
        ("ATOMG", "global atomic with return"), ("ATOMS", "shared-memory atomic")]
for key, what in rows:
  if "." in key:
    n = sum(v for k, v in full.items() if k.startswith(key.split(".")[0]) and "." + key.split(".", 1)[1] in k)
  else:
    n = ops.get(key, 0)
  print(f"{key:16s} {n:7d}   {what}")
print("# -- legacy tensor-core paths (must be zero)")
for key, what in [("HMMA", "mma.sync / wmma"), ("IMMA", "integer mma.sync"), ("HGMMA", "Hopper wgmma"), ("QGMMA", "Hopper wgmma fp8")]:
  print(f"{key:16s} {ops.get(key, 0):7d}   {what}")
print("# -- local memory (LDL + STL) by kernel; absent kernels have none")
def short(n):
  n = re.sub(r"(\(anonymous namespace\)|<unnamed>)::", "", n)
  n = re.sub(r"\((?:int|bool)\)", "", n)
  return re.sub(r"[(].*$", "", n).replace("void ", "")


names = list(per_kernel_local)
for mangled, nice in sorted(zip(names, demangle(names)), key=lambda kv: -per_kernel_local[kv[0]]):
  print(f"{per_kernel_local[mangled]:6d}  {short(nice)[:110]}")
print("# -- kernels in the library (template instances per name)")
cnt = collections.Counter(re.sub(r"<.*$", "", short(n)) for n in demangle(kernels) if n)
for k, v in sorted(cnt.items(), key=lambda kv: (-kv[1], kv[0])):
  print(f"{v:6d}  {k}")
