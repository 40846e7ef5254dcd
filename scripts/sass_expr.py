#!/usr/bin/env python
"""TEST INFRASTRUCTURE (parity work): symbolic read-out of floating-point dataflow from an `nvdisasm -gi` listing.

    python scripts/sass_expr.py LISTING FIRST_LINE LAST_LINE [--depth N] [--show REGS]

Walks the SASS instructions of a straight-line region in program order and keeps, per register, the expression
that produced it (FFMA/FMUL/FADD/MUFU/FMNMX/FSEL and loads as leaves named by their address offset), so the exact
FMA CONTRACTION the compiler chose for a source expression of the reference build can be read off and pinned in the
product's arithmetic.  Values that enter the region are leaves named after their register."""
import re
import sys

INS = re.compile(r"^\s*/\*([0-9a-f]+)\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)\s+(.*?)\s*;")


def clean(op):
    return op.replace(".reuse", "").strip()


def main():
    path, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    depth = 6
    show = None
    if "--depth" in sys.argv:
        depth = int(sys.argv[sys.argv.index("--depth") + 1])
    if "--show" in sys.argv:
        show = set(sys.argv[sys.argv.index("--show") + 1].split(","))
    lines = open(path).read().splitlines()[lo - 1:hi]
    env = {}

    def val(op, d=0):
        op = clean(op)
        neg = op.startswith("-")
        if neg:
            op = op[1:]
        ab = op.startswith("|") and op.endswith("|")
        if ab:
            op = op[1:-1]
        if re.fullmatch(r"R\d+", op):
            e = env.get(op, op)
        else:
            e = op
        if ab:
            e = f"|{e}|"
        if neg:
            e = f"-{e}"
        return e

    src_line = ""
    for ln in lines:
        if "//## File" in ln:
            m = re.search(r'"([^"]+)", line (\d+)', ln)
            if m and "inlined at" not in ln.split(m.group(0))[0]:
                src_line = m.group(1).split("/")[-1] + ":" + m.group(2)
            continue
        m = INS.match(ln)
        if not m:
            continue
        addr, pred, opc, rest = m.groups()
        ops = [o.strip() for o in rest.split(",")]
        base = opc.split(".")[0]
        if pred and base in ("FMUL",) and ops[-1].strip() in ("16777216", "4096", "1.84467440737095516160e+19", "5.42101086242752217004e-20"):
            continue                      # denormal-range rescaling guards around MUFU
        dst = clean(ops[0])
        e = None
        if base == "FFMA":
            e = f"fma({val(ops[1])}, {val(ops[2])}, {val(ops[3])})"
        elif base == "FMUL":
            e = f"mul({val(ops[1])}, {val(ops[2])})"
        elif base == "FADD":
            e = f"add({val(ops[1])}, {val(ops[2])})"
        elif base == "MUFU":
            e = f"{opc.split('.')[1].lower()}({val(ops[1])})"
        elif base in ("FMNMX", "FMNMX3"):
            e = f"{opc.lower()}({', '.join(val(o) for o in ops[1:])})"
        elif base == "FSEL":
            e = f"sel({val(ops[1])}, {val(ops[2])}, {ops[3]})"
        elif base in ("LDG", "LD", "LDS", "LDL", "LDC"):
            mm = re.search(r"\[(.*)\]", rest)
            e = f"{base.lower()}[{mm.group(1) if mm else rest}]"
            nreg = 1
            if ".128" in opc:
                nreg = 4
            elif ".64" in opc:
                nreg = 2
            r0 = int(dst[1:]) if re.fullmatch(r"R\d+", dst) else None
            if r0 is not None:
                for k in range(nreg):
                    env[f"R{r0 + k}"] = e + (f".{k}" if nreg > 1 else "")
                continue
        elif base in ("MOV", "IMAD") and opc.startswith("IMAD.MOV") or base == "MOV":
            e = val(ops[-1])
        elif base == "FSETP":
            print(f"{addr} [{src_line}] {pred or ''}{opc} {ops[0]},{ops[1]} <- {val(ops[2])} ? {val(ops[3])}")
            continue
        else:
            if re.fullmatch(r"R\d+", dst):
                env.pop(dst, None)       # produced by something we do not model: becomes a leaf again
            continue
        if pred:
            e = f"({pred.strip()}? {e} : {env.get(dst, dst)})"
        if show is None or dst in show:
            print(f"{addr} [{src_line}] {dst} = {e}")
        if len(e) > depth * 12:
            e = f"{dst}@{addr}"          # long expression: later uses refer to it by its definition site
        env[dst] = e


if __name__ == "__main__":
    main()
