"""Developer helper (no GPU needed): static opcode histogram of the time loop of one kernel.

    python tools/sass_loop.py <object-or-.so> <kernel-name-substring> [--all]

Disassembles the kernel with cuobjdump, takes the LARGEST backward branch as the time loop (the filter kernels have exactly
one loop that big) and counts the opcodes between its target and the branch.  Rarely-taken side blocks that sit inside that
address range (the trunc(R^T) slow path) are included, so the figures are an upper bound of what one trip executes; the
ncu source-page count (tools/ncu_hot.py) is the executed figure.  Used by tests/test_host_cpu.py to keep the flop constants of
bench.py honest."""
import collections
import re
import subprocess
import sys

FP64 = ("DFMA", "DMUL", "DADD")
FP32 = ("FFMA", "FMUL", "FADD", "FFMA2", "FMUL2", "FADD2")


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, body = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
        elif cur is not None:
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                body[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return body


def opcode(text):
    t = re.sub(r"^@!?U?P[0-9T]+\s+", "", text)
    return t.split()[0].split(".")[0]


def loop_of(insts):
    best = None
    for addr, text in insts:
        if opcode(text) == "BRA":
            m = re.search(r"0x([0-9a-f]+)", text)
            if m:
                tgt = int(m.group(1), 16)
                if tgt < addr and (best is None or addr - tgt > best[1] - best[0]):
                    best = (tgt, addr)
    return best


def loop_histogram(path, name_sub):
    res = {}
    for name, insts in kernels(path).items():
        if name_sub not in name:
            continue
        lp = loop_of(insts)
        if lp is None:
            continue
        h = collections.Counter(opcode(t) for a, t in insts if lp[0] <= a <= lp[1])
        res[name] = (h, lp)
    return res


def flops(h):
    """(flops, fp instructions) of a histogram: FMA = 2 flops, packed FP32 instructions carry two lanes."""
    f = 2 * h["DFMA"] + h["DMUL"] + h["DADD"] + 2 * h["FFMA"] + h["FMUL"] + h["FADD"] + 4 * h["FFMA2"] + 2 * h["FMUL2"] + 2 * h["FADD2"]
    n = sum(h[k] for k in FP64 + FP32)
    return f, n


if __name__ == "__main__":
    for name, (h, lp) in loop_histogram(sys.argv[1], sys.argv[2]).items():
        tot = sum(h.values())
        f, n = flops(h)
        print(f"{name}\n  loop 0x{lp[0]:x}..0x{lp[1]:x}: {tot} instructions, {n} floating-point ({f} flops)")
        show = h.most_common() if "--all" in sys.argv else h.most_common(14)
        print("  " + "  ".join(f"{k}:{v}" for k, v in show))
