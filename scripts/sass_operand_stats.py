"""Static look at the FP64 instructions of the force kernel's inner loops: how many 64-bit register operands per
DFMA/DADD/DMUL are NOT served by the operand-reuse cache (development aid for DESIGN.md 4.1 "what limits it").

    python scripts/sass_operand_stats.py [object-or-library] [kernel-substring]

Model (from scripts/microbench/bwd_pat.cu: 2.63 cycles per DFMA with three fresh register operands in every second
instruction, 2.2 with one operand from the constant bank): an FP64 instruction costs about max(2, fresh operands) issue
cycles of its SM sub-partition, where an operand is "fresh" unless the same register was read in the same operand slot
by an earlier instruction that carried the .reuse flag on that slot and no other .reuse-flagged read replaced it.
Prints, per inner loop (back edge with >= 40 FP64 instructions in the body): instruction counts and the model's cycles
per FP64 instruction.  It is a proxy, not a measurement."""
import re
import subprocess
import sys

obj = sys.argv[1] if len(sys.argv) > 1 else "meng_zhang_b200/lib/annp_force.o"
want = sys.argv[2] if len(sys.argv) > 2 else "annp_force_kernelILi9ELi19ELi0ELb1"
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
ins, on = [], False
for line in txt.split("\n"):
    if "Function :" in line:
        on = want in line
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if on and m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
FP = ("DFMA", "DADD", "DMUL")


def srcs(text):
    t = text.split()
    if t[0].startswith("@"):
        t = t[1:]
    op = t[0].split(".")[0]
    ops = " ".join(t[1:]).split(",")
    ops = [o.strip() for o in ops]
    return op, ops[1:]          # drop the destination


for i, (a, text) in enumerate(ins):
    m = re.search(r"BRA 0x([0-9a-f]+)", text)
    if not m or "BRA.DIV" in text:
        continue
    tgt = int(m.group(1), 16)
    if tgt not in addr or addr[tgt] >= i:
        continue
    body = ins[addr[tgt]:i + 1]
    if any(re.search(r"BRA 0x", t) and "BRA.DIV" not in t for _, t in body[:-1]):
        continue                # not an innermost loop
    nfp = sum(srcs(t)[0] in FP for _, t in body)
    if nfp < 40:
        continue
    cache = [None, None, None]
    fresh_hist, cyc = {}, 0.0
    for rep in range(2):        # second trip sees the cache state left by the first
        fresh_hist, cyc = {}, 0.0
        for _, t in body:
            op, ss = srcs(t)
            fresh = 0
            for slot, o in enumerate(ss[:3]):
                neg = o.lstrip("-|")
                reg = re.match(r"(R\d+)(\.reuse)?", neg)
                if not reg or neg.startswith("RZ"):
                    continue
                hit = cache[slot] == reg.group(1)
                if op in FP and not hit:
                    fresh += 1
                if reg.group(2):
                    cache[slot] = reg.group(1)
            if op in FP:
                fresh_hist[fresh] = fresh_hist.get(fresh, 0) + 1
                cyc += max(2, fresh)
    print(f"loop {tgt:#x}..{a:#x}: {len(body)} instructions, {nfp} FP64, fresh-operand histogram {dict(sorted(fresh_hist.items()))}, "
          f"model {cyc / nfp:.2f} cycles per FP64 instruction = {cyc:.0f} per trip")
