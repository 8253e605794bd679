"""Inner loop of direct_sum_kernel<4,false,false,false> as ptxas emitted it: FP64 instruction mix and how many of the accumulate
DFMAs (three distinct register operands = 3 FP64-pipe cycles) can take an operand from the reuse cache,
i.e. directly follow an FP64 instruction that flagged the same register `.reuse`.
    python scripts/sass_reuse_report.py > profiles/r2_direct_sum_inner_loop_sass.txt"""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
obj = ROOT / "gravity-simulator_b200" / "build" / "direct_sum.o"
fun = "_ZN2gb17direct_sum_kernelILi4ELb0ELb0ELb0EEEvNS_6DSArgsE"
out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, str(obj)], capture_output=True, text=True, check=True).stdout
ins = [re.sub(r"/\* 0x[0-9a-f]+ \*/", "", l).rstrip() for l in out.splitlines() if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
# the loop: from the first LDS.128 after the second BAR.SYNC to the backward BRA.U
bars = [i for i, l in enumerate(ins) if "BAR.SYNC" in l]
start = next(i for i in range(bars[1], len(ins)) if "LDS.128" in ins[i])
def is_back_branch(i):
    m = re.search(r"/\*([0-9a-f]{4})\*/.*\bBRA(?:\.U)?\b.*0x([0-9a-f]+)", ins[i])
    return bool(m) and int(m.group(2), 16) <= int(m.group(1), 16) and "MUFU" in " ".join(ins[start:i])
end = next(i for i in range(start, len(ins)) if is_back_branch(i))
start = max(k for k in range(start, end) if re.search(r"/\*([0-9a-f]{4})\*/", ins[k]) and
            int(re.search(r"/\*([0-9a-f]{4})\*/", ins[k]).group(1), 16) <= int(re.search(r"0x([0-9a-f]+)", ins[end].split("BRA")[1]).group(1), 16))
loop = ins[start:end + 1]
print(f"# {fun}\n# inner loop: {len(loop)} instructions for 2 sources x 4 targets = 8 interactions\n")
for l in loop:
    print(l)
ops = [re.sub(r"^\s*/\*[0-9a-f]+\*/\s*", "", l) for l in loop]
fp64 = [o for o in ops if re.match(r"(@!?P\d\s+)?D(FMA|MUL|ADD)\b", o)]
mufu = [o for o in ops if "MUFU" in o]
print(f"\n# FP64-pipe instructions: {len(fp64)} (= {len(fp64) / 8:.2f} per interaction) + {len(mufu)} MUFU.RSQ64H; other: {len(ops) - len(fp64) - len(mufu)}")
# accumulate DFMAs: destination == third source (d = a*b + d) with a, b distinct registers
acc, hits = 0, 0
prev_reuse = set()
for o in ops:
    m = re.match(r"(?:@!?P\d\s+)?(DFMA|DMUL|DADD)\s+(R\d+),\s*(-?R\d+(?:\.reuse)?|[^,]+),\s*(-?R\d+(?:\.reuse)?|[^,;]+)(?:,\s*(-?R\d+(?:\.reuse)?|[^,;]+))?", o)
    if not m:
        if "MUFU" not in o and not o.startswith(("IMAD", "MOV", "LDS", "UIADD", "UISETP", "BRA")):
            prev_reuse = set()
        continue
    kind, d, *src = m.groups()
    src = [s for s in src if s]
    regs = [s.replace("-", "").replace(".reuse", "") for s in src if s.lstrip("-").startswith("R")]
    if kind == "DFMA" and len(src) == 3 and regs[-1:] == [d] and len(set(regs)) == 3:
        acc += 1
        if any(r in prev_reuse for r in regs[:2]):
            hits += 1
    prev_reuse = {s.replace("-", "").replace(".reuse", "") for s in src if s.endswith(".reuse")}
print(f"# accumulate DFMAs (3 distinct register operands): {acc}; directly preceded by an FP64 instruction that flagged one of their "
      f"operands .reuse: {hits}")
cyc = 2 * len(fp64) + (acc - hits)
print(f"# FP64-pipe cycles per loop trip by the measured rule (2 per instruction, +1 per three-operand DFMA without a reuse hit): {cyc} "
      f"-> ceiling of the instruction rate {2 * len(fp64) / cyc:.3f}; measured in bench.py: fp64_pipe_util ~0.906")
