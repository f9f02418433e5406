#!/usr/bin/env python
"""Generates csrc/ali_glibcmath.cuh: sin / cos / tan / atan with EXACTLY the results of the host's glibc.

Why.  The reference (numba) calls libm's double sin / cos / tan / atan; the ALI update chooses between
stencils by min |dT|, which is discontinuous, so one last-ulp difference in atan / sin / cos can move a
whole field by 1e-4 (DESIGN.md "Parity").  Round 1 used correctly rounded device functions, which agree
with glibc only where glibc itself is correctly rounded (99.9 % of the calls).  This tool restates the
very code glibc runs: it reads the x86-64 machine code of the variants the dynamic loader selects on
FMA-capable CPUs (__sin_fma, __cos_fma, __tan_fma, __atan_fma: glibc's IBM Accurate Mathematical
Library routines, sysdeps/ieee754/dbl-64/s_sin.c, s_tan.c, s_atan.c, compiled with -mfma -mavx2) from
the host's libm.so.6 and translates it instruction by instruction into C: every add / mul / div / fma,
every comparison and every table look-up in the same order, constants and tables taken from .rodata.
The result compiles for the device (fma -> __fma_rn, the translation unit is built with -fmad=false)
and for the host replay (tests/emu; g++ -ffp-contract=off), and tests/test_kernel_replay.py checks
bit equality with the running libm on >= 1e8 arguments.

Scope: round-to-nearest mode; |x| < 105414350 for sin / cos and |x| <= 1e8 for tan (beyond that glibc
calls its big-argument reduction __branred, which is not translated: the generated code falls back to the
platform's own function there -- the hot path never leaves [-2 pi, 2 pi]).

    python tools/gen_glibc_math.py [--libm /lib/x86_64-linux-gnu/libm.so.6]

Needs objdump (binutils) and an x86-64 glibc 2.28+ whose sin resolves through an IFUNC with an FMA variant.
"""
import argparse
import os
import re
import struct
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "ali_fmm_and_ray_tracing_b200", "csrc", "ali_glibcmath.cuh")

REG64 = ["rax", "rbx", "rcx", "rdx", "rsi", "rdi", "r8", "r9", "r10", "r11", "r12", "r13", "r14", "r15"]
REG32 = {"eax": "rax", "ebx": "rbx", "ecx": "rcx", "edx": "rdx", "esi": "rsi", "edi": "rdi"}
REG32.update({"r%dd" % k: "r%d" % k for k in range(8, 16)})
REG8 = {"al": "rax", "bl": "rbx", "cl": "rcx", "dl": "rdx", "sil": "rsi", "dil": "rdi"}
REG8.update({"r%db" % k: "r%d" % k for k in range(8, 16)})
REG8H = {"ah": "rax", "bh": "rbx", "ch": "rcx", "dh": "rdx"}


def sh(cmd):
    return subprocess.run(cmd, stdout=subprocess.PIPE, check=True, text=True).stdout


class Lib:
    def __init__(self, path):
        self.path = path
        self.data = open(path, "rb").read()
        self.sections = []
        for ln in sh(["readelf", "-S", "-W", path]).splitlines():
            m = re.match(r"\s*\[\s*\d+\]\s+(\S+)\s+\S+\s+([0-9a-f]+)\s+([0-9a-f]+)\s+([0-9a-f]+)", ln)
            if m:
                self.sections.append((m.group(1), int(m.group(2), 16), int(m.group(3), 16), int(m.group(4), 16)))

    def read(self, vaddr, n):
        for name, va, off, size in self.sections:
            if va <= vaddr and vaddr + n <= va + size and name in (".rodata", ".data.rel.ro", ".data"):
                return self.data[off + vaddr - va: off + vaddr - va + n]
        raise ValueError("address %x not in a data section" % vaddr)

    def u64(self, vaddr):
        return struct.unpack("<Q", self.read(vaddr, 8))[0]

    def disasm(self, start, stop):
        out = sh(["objdump", "-d", "--no-show-raw-insn", "--start-address=%d" % start, "--stop-address=%d" % stop, self.path])
        ins = []
        for ln in out.splitlines():
            m = re.match(r"\s+([0-9a-f]+):\s+(.*)$", ln)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        return ins

    def ifunc_fma_variant(self, sym):
        """Address of the first variant the IFUNC resolver of ``sym`` can return (the AVX2 + FMA one)."""
        for ln in sh(["nm", "-D", self.path]).splitlines():
            f = ln.split()
            if len(f) == 3 and f[1] == "i" and f[2].split("@")[0] == sym:
                addr = int(f[0], 16)
                for a, text in self.disasm(addr, addr + 0x60):
                    m = re.match(r"lea\s+.*\(%rip\),%rax\s+#\s+([0-9a-f]+)", text)
                    if m:
                        return int(m.group(1), 16)
        raise ValueError("no IFUNC / FMA variant for " + sym)


def function_body(lib, start):
    """Instructions from ``start`` to the function's __stack_chk_fail call (its last instruction)."""
    ins = lib.disasm(start, start + 0x1400)
    for k, (a, text) in enumerate(ins):
        if "__stack_chk_fail" in text:
            return ins[:k + 1]
    raise ValueError("function end not found")


class Translator:
    def __init__(self, lib, name, ins, tables, fallback):
        self.lib, self.name, self.ins, self.tables, self.fallback = lib, name, ins, tables, fallback
        self.lines = []
        self.slots = {}
        self.targets = set()
        self.flag = None   # ("comisd", b, a) | ("cmp", bits, b, a) | ("test", bits, expr)
        for a, text in ins:
            m = re.match(r"j\w+\s+([0-9a-f]+)", text)
            if m:
                self.targets.add(int(m.group(1), 16))

    # ---- operand helpers
    def slot(self, off, size):
        key = (off, size)
        if key not in self.slots:
            self.slots[key] = "s%s%x_%d" % ("m" if off < 0 else "p", abs(off), size)
        return self.slots[key]

    def gpr_read(self, r, ):
        r = r.lstrip("%")
        if r in REG64:
            return "%s" % r, 64
        if r in REG32:
            return "(uint32_t)%s" % REG32[r], 32
        if r in REG8:
            return "(uint8_t)%s" % REG8[r], 8
        if r in REG8H:
            return "(uint8_t)(%s >> 8)" % REG8H[r], 8
        raise ValueError("register " + r)

    def gpr_write(self, r, expr):
        r = r.lstrip("%")
        if r in REG64:
            return "%s = (uint64_t)(%s);" % (r, expr)
        if r in REG32:
            return "%s = (uint64_t)(uint32_t)(%s);" % (REG32[r], expr)
        if r in REG8:
            return "%s = (%s & ~0xffull) | (uint8_t)(%s);" % (REG8[r], REG8[r], expr)
        if r in REG8H:
            return "%s = (%s & ~0xff00ull) | ((uint64_t)(uint8_t)(%s) << 8);" % (REG8H[r], REG8H[r], expr)
        raise ValueError("register " + r)

    def mem(self, op, size, comment):
        """C expression reading ``size`` bytes (as an unsigned integer of that size) from a memory operand."""
        m = re.match(r"(-?0x[0-9a-f]+)?\(%rip\)", op)
        if m:
            addr = int(re.search(r"#\s+([0-9a-f]+)", comment).group(1), 16)
            raw = self.lib.read(addr, size)
            v = int.from_bytes(raw, "little")
            return ("0x%016xull" % v) if size == 8 else ("0x%08xu" % v)
        m = re.match(r"(-?0x[0-9a-f]+)?\(%rbp\)", op)
        if m:
            return self.slot(int(m.group(1) or "0", 16), size)
        m = re.match(r"%fs:0x28", op)
        if m:
            return "0ull"
        m = re.match(r"(-?0x[0-9a-f]+)?\(%(\w+),%(\w+),(\d)\)", op)
        if m and size == 8:
            disp = int(m.group(1) or "0", 16)
            return "ALI_GL_TAB(%s + %s * %s + (%d))" % (m.group(2), m.group(3), m.group(4), disp)
        m = re.match(r"(-?0x[0-9a-f]+)?\(%(\w+)\)", op)
        if m and size == 8:
            return "ALI_GL_TAB(%s + (%d))" % (m.group(2), int(m.group(1) or "0", 16))
        raise ValueError("memory operand " + op)

    def xsrc(self, op, comment):
        """64-bit pattern of an xmm register or memory operand."""
        if op.startswith("%xmm"):
            return "x" + op[4:]
        return self.mem(op, 8, comment)

    def emit(self, s):
        self.lines.append("    " + s)

    @staticmethod
    def split_ops(s):
        out, depth, cur = [], 0, ""
        for ch in s:
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            if ch == "," and depth == 0:
                out.append(cur)
                cur = ""
            else:
                cur += ch
        if cur:
            out.append(cur)
        return [o.strip() for o in out]

    def cond(self, cc):
        f = self.flag
        if f is None:
            raise ValueError("conditional without flags")
        if f[0] == "comisd":
            b, a = f[1], f[2]   # compares b with a (AT&T: vcomisd a, b); unordered sets ZF = PF = CF = 1
            table = {"a": "({b} > {a})", "ae": "({b} >= {a})", "b": "!({b} >= {a})", "be": "!({b} > {a})",
                     "e": "!({b} < {a} || {b} > {a})", "ne": "({b} < {a} || {b} > {a})",
                     "p": "({b} != {b} || {a} != {a})", "np": "!({b} != {b} || {a} != {a})"}
            return table[cc].format(a=a, b=b)
        if f[0] == "cmp":
            bits, b, a = f[1], f[2], f[3]   # flags of b - a
            st, ut = ("int%d_t" % bits), ("uint%d_t" % bits)
            table = {"e": "((%s)(%s) == (%s)(%s))" % (ut, b, ut, a), "ne": "((%s)(%s) != (%s)(%s))" % (ut, b, ut, a),
                     "g": "((%s)(%s) > (%s)(%s))" % (st, b, st, a), "ge": "((%s)(%s) >= (%s)(%s))" % (st, b, st, a),
                     "l": "((%s)(%s) < (%s)(%s))" % (st, b, st, a), "le": "((%s)(%s) <= (%s)(%s))" % (st, b, st, a),
                     "a": "((%s)(%s) > (%s)(%s))" % (ut, b, ut, a), "ae": "((%s)(%s) >= (%s)(%s))" % (ut, b, ut, a),
                     "b": "((%s)(%s) < (%s)(%s))" % (ut, b, ut, a), "be": "((%s)(%s) <= (%s)(%s))" % (ut, b, ut, a)}
            return table[cc]
        if f[0] == "test":
            bits, e = f[1], f[2]
            st = "int%d_t" % bits
            table = {"e": "((uint%d_t)(%s) == 0)" % (bits, e), "ne": "((uint%d_t)(%s) != 0)" % (bits, e),
                     "s": "((%s)(%s) < 0)" % (st, e), "ns": "((%s)(%s) >= 0)" % (st, e),
                     "le": "((%s)(%s) <= 0)" % (st, e), "g": "((%s)(%s) > 0)" % (st, e),
                     "l": "((%s)(%s) < 0)" % (st, e), "ge": "((%s)(%s) >= 0)" % (st, e)}
            return table[cc]
        raise ValueError(f)

    def int_src(self, op, bits, comment):
        if op.startswith("%fs:"):
            return "0ull", bits   # stack protector canary
        if op.startswith("$"):
            return "(%s)" % op[1:], bits
        if op.startswith("%"):
            return self.gpr_read(op)
        return self.mem(op, bits // 8, comment), bits

    def translate(self):
        D = "ALI_GL_D"
        B = "ALI_GL_B"
        for addr, text in self.ins:
            if addr in self.targets:
                self.lines.append("L_%x:;" % addr)
            comment = ""
            if "#" in text:
                text, comment = text.split("#", 1)
                comment = "# " + comment
            text = re.sub(r"<[^>]*>", "", text).strip()
            parts = text.split(None, 1)
            mn = parts[0]
            if mn in ("cs", "data16"):
                continue
            ops = self.split_ops(parts[1]) if len(parts) > 1 else []
            e = self.emit
            if mn in ("endbr64", "push", "pop", "leave", "nop", "nopl", "nopw", "xchg", "vldmxcsr"):
                continue
            if "%rsp" in text or (ops and ops[-1] == "%rbp"):
                continue   # frame set-up
            if "%fs:(" in text:
                continue   # errno = EDOM (infinite argument)
            if mn == "mov" and "(%rip)" in ops[0]:
                ga = int(re.search(r"#\s+([0-9a-f]+)", comment).group(1), 16)
                if any(nm == ".got" and va <= ga < va + sz for nm, va, off, sz in self.lib.sections):
                    continue   # TLS offset of errno
            if mn == "ret":
                e("return %s(x0);" % D)
            elif mn == "call":
                if "__stack_chk_fail" in "".join(ops) or "10290" in "".join(ops):
                    e("return %s(x0);   /* (stack protector: unreachable) */" % D)
                else:
                    e("return %s;   /* huge argument: glibc's __branred, not translated */" % self.fallback)
            elif mn == "jmp":
                e("goto L_%s;" % ops[0])
            elif mn.startswith("j"):
                e("if (%s) goto L_%s;" % (self.cond(mn[1:]), ops[0]))
            elif mn == "vstmxcsr":
                e("%s = 0x1f80u;   /* MXCSR: round to nearest, exceptions masked */" % self.mem(ops[0], 4, comment))
            elif mn in ("vmovsd", "vmovq", "vmovapd", "vmovaps"):
                if len(ops) == 3:
                    e("x%s = x%s;" % (ops[2][4:], ops[0][4:]))
                elif ops[1].startswith("%xmm"):
                    if ops[0].startswith("%xmm"):
                        e("x%s = x%s;" % (ops[1][4:], ops[0][4:]))
                    elif ops[0].startswith("%"):
                        e("x%s = %s;" % (ops[1][4:], self.gpr_read(ops[0])[0]))
                    else:
                        e("x%s = %s;" % (ops[1][4:], self.mem(ops[0], 8, comment)))
                elif ops[1].startswith("%"):
                    e(self.gpr_write(ops[1], "x" + ops[0][4:]))
                else:
                    e("%s = x%s;" % (self.mem(ops[1], 8, comment), ops[0][4:]))
            elif mn in ("vaddsd", "vsubsd", "vmulsd", "vdivsd"):
                opx = {"vaddsd": "+", "vsubsd": "-", "vmulsd": "*", "vdivsd": "/"}[mn]
                a, b, c = ops   # c = b op a
                e("x%s = %s(%s(x%s) %s %s(%s));" % (c[4:], B, D, b[4:], opx, D, self.xsrc(a, comment)))
            elif re.match(r"vf(n?)m(add|sub)(132|213|231)sd", mn):
                m = re.match(r"vf(n?)m(add|sub)(132|213|231)sd", mn)
                neg, kind, form = m.group(1) == "n", m.group(2), m.group(3)
                a, b, c = self.xsrc(ops[0], comment), "x" + ops[1][4:], "x" + ops[2][4:]
                if form == "132":
                    p1, p2, ad = c, a, b
                elif form == "213":
                    p1, p2, ad = b, c, a
                else:
                    p1, p2, ad = b, a, c
                p1e = "%s(%s)" % (D, p1)
                if neg:
                    p1e = "-" + p1e
                ade = "%s(%s)" % (D, ad)
                if kind == "sub":
                    ade = "-" + ade
                e("%s = %s(ALI_GL_FMA(%s, %s(%s), %s));" % (c, B, p1e, D, p2, ade))
            elif mn in ("vandpd", "vorpd", "vxorpd", "vandnpd", "vandps", "vxorps", "vorps"):
                a, b, c = self.xsrc(ops[0], comment), "x" + ops[1][4:], "x" + ops[2][4:]
                if mn.startswith("vandn"):
                    e("%s = ~%s & %s;" % (c, b, a))
                else:
                    opx = {"vand": "&", "vorp": "|", "vxor": "^"}[mn[:4]]
                    e("%s = %s %s %s;" % (c, b, opx, a))
            elif mn == "vblendvpd":
                mask, s2, s1, d = ops
                e("x%s = ((int64_t)x%s < 0) ? %s : x%s;" % (d[4:], mask[4:], self.xsrc(s2, comment), s1[4:]))
            elif re.match(r"vcmp(\w+)sd", mn):
                pred = re.match(r"vcmp(\w+)sd", mn).group(1)
                a, b, c = self.xsrc(ops[0], comment), "x" + ops[1][4:], "x" + ops[2][4:]
                ce = {"lt": "(%s(%s) < %s(%s))", "le": "(%s(%s) <= %s(%s))", "nlt": "!(%s(%s) < %s(%s))",
                      "nle": "!(%s(%s) <= %s(%s))", "eq": "(%s(%s) == %s(%s))", "neq": "!(%s(%s) == %s(%s))"}[pred] % (D, b, D, a)
                e("%s = %s ? ~0ull : 0ull;" % (c, ce))
            elif mn in ("vcomisd", "vucomisd"):
                a, b = self.xsrc(ops[0], comment), "x" + ops[1][4:]
                self.flag = ("comisd", "%s(%s)" % (D, b), "%s(%s)" % (D, a))
                # flags are consumed later; operands may be overwritten first -> snapshot
                e("fa = %s(%s); fb = %s(%s);" % (D, a, D, b))
                self.flag = ("comisd", "fb", "fa")
            elif mn == "vcvttsd2si":
                e(self.gpr_write(ops[1], "(int32_t)%s(%s)" % (D, self.xsrc(ops[0], comment))))
            elif mn == "vcvtsi2sd" or mn == "vcvtsi2sdl":
                src = self.int_src(ops[0], 32, comment)[0]
                e("x%s = %s((double)(int32_t)(%s));" % (ops[-1][4:], B, src))
            elif mn in ("mov", "movl", "movq"):
                src, dst = ops
                if dst.startswith("%"):
                    bits = self.gpr_read(dst)[1]
                    e(self.gpr_write(dst, self.int_src(src, bits, comment)[0]))
                else:
                    bits = 32 if mn == "movl" else (self.gpr_read(src)[1] if src.startswith("%") else 64)
                    e("%s = %s;" % (self.mem(dst, bits // 8, comment), self.int_src(src, bits, comment)[0]))
            elif mn == "movslq":
                e(self.gpr_write(ops[1], "(int64_t)(int32_t)(%s)" % self.int_src(ops[0], 32, comment)[0]))
            elif mn == "cltq":
                e("rax = (uint64_t)(int64_t)(int32_t)rax;")
            elif mn == "lea":
                src, dst = ops
                if "%rbp" in src:
                    continue   # address of a stack slot: argument of the untranslated __branred call
                m = re.match(r"(-?0x[0-9a-f]+)?\(%rip\)", src)
                if m:
                    addr = int(re.search(r"#\s+([0-9a-f]+)", comment).group(1), 16)
                    e(self.gpr_write(dst, "0x%xull" % addr))
                else:
                    m = re.match(r"(-?0x[0-9a-f]+)?\((?:%(\w+))?(?:,%(\w+),(\d))?\)", src)
                    disp = int(m.group(1) or "0", 16)
                    ex = "%s" % self.gpr_read("%" + m.group(2))[0].replace("(uint32_t)", "") if m.group(2) else "0ull"
                    if m.group(3):
                        ex += " + %s * %s" % (self.gpr_read("%" + m.group(3))[0].replace("(uint32_t)", ""), m.group(4))
                    e(self.gpr_write(dst, "%s + (%d)" % (ex, disp)))
            elif mn in ("add", "sub", "and", "or", "xor", "shl", "sar", "shr", "addl", "subl", "andl", "orl"):
                src, dst = ops
                base = mn.rstrip("l") if mn not in ("shl",) else mn
                if base == "sh":
                    base = "shl"
                bits = self.gpr_read(dst)[1] if dst.startswith("%") else 32
                s_e = self.int_src(src, bits, comment)[0]
                if dst.startswith("%"):
                    d_e = self.gpr_read(dst)[0]
                else:
                    d_e = self.mem(dst, bits // 8, comment)
                if base == "sar":
                    ex = "(int%d_t)(%s) >> (%s)" % (bits, d_e, s_e)
                elif base in ("shl", "shr"):
                    ex = "(uint%d_t)(%s) %s (%s)" % (bits, d_e, "<<" if base == "shl" else ">>", s_e)
                else:
                    opx = {"add": "+", "sub": "-", "and": "&", "or": "|", "xor": "^"}[base]
                    ex = "(uint%d_t)(%s) %s (uint%d_t)(%s)" % (bits, d_e, opx, bits, s_e)
                if dst.startswith("%"):
                    e("t64 = (uint64_t)(%s);" % ex)
                    e(self.gpr_write(dst, "t64"))
                else:
                    e("t64 = (uint64_t)(%s); %s = (uint%d_t)t64;" % (ex, d_e, bits))
                self.flag = ("test", bits, "t64f")
                e("t64f = t64;")
            elif mn == "not":
                e(self.gpr_write(ops[0], "~%s" % self.gpr_read(ops[0])[0]))
            elif mn in ("cmp", "cmpl"):
                a, b = ops   # flags of b - a
                bits = self.gpr_read(b)[1] if b.startswith("%") else 32
                e("ca = (uint64_t)(%s); cb = (uint64_t)(%s);" % (self.int_src(a, bits, comment)[0], self.int_src(b, bits, comment)[0]))
                self.flag = ("cmp", bits, "cb", "ca")
            elif mn in ("test", "testb", "testl"):
                a, b = ops
                bits = self.gpr_read(b)[1] if b.startswith("%") else (8 if mn == "testb" else 32)
                e("t64f = (uint64_t)((%s) & (%s));" % (self.int_src(a, bits, comment)[0], self.int_src(b, bits, comment)[0]))
                self.flag = ("test", bits, "t64f")
            else:
                raise ValueError("%x: unhandled instruction: %s %s" % (addr, mn, ops))
        return self.lines

    def function(self):
        body = self.translate()
        base, (tname, tcount) = self.table
        text = "\n".join(body)
        used_x = [k for k in range(1, 16) if re.search(r"\bx%d\b" % k, text)]
        used_r = [r for r in REG64 if re.search(r"\b%s\b" % r, text)]
        decl = ["// (tab: the routine's table -- %s -- or a copy of it in faster memory)" % tname,
                "#define ALI_GL_TAB(a) tab[((a) - 0x%xull) >> 3]" % base,
                "ALI_GL_DEV double %s_t(double x, const uint64_t *tab)" % self.name, "{",
                "    uint64_t x0 = ALI_GL_B(x), " + ", ".join("x%d = 0" % k for k in used_x) + ";",
                "    uint64_t " + ", ".join("%s = 0" % r for r in used_r) + ";",
                "    uint64_t t64 = 0, t64f = 0, ca = 0, cb = 0; double fa = 0.0, fb = 0.0;"]
        for (off, size), nm in sorted(self.slots.items()):
            decl.append("    uint%d_t %s = 0;" % (size * 8, nm))
        decl.append("    (void)t64; (void)t64f; (void)ca; (void)cb; (void)fa; (void)fb; (void)tab;" +
                    "".join(" (void)%s;" % nm for nm in self.slots.values()))
        wrap = "ALI_GL_DEV double %s(double x) { return %s_t(x, %s); }" % (self.name, self.name, tname)
        return "\n".join(decl + body + ["}", "#undef ALI_GL_TAB", wrap])


def table_c(lib, name, addr, count):
    vals = [lib.u64(addr + 8 * k) for k in range(count)]
    rows = []
    for k in range(0, count, 4):
        rows.append("    " + ", ".join("0x%016xull" % v for v in vals[k:k + 4]) + ",")
    return "ALI_GL_TABLE uint64_t %s[%d] = {\n%s\n};" % (name, count, "\n".join(rows))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--libm", default="/lib/x86_64-linux-gnu/libm.so.6")
    ap.add_argument("--out", default=OUT)
    args = ap.parse_args()
    lib = Lib(args.libm)
    version = sh(["ldd", "--version"]).splitlines()[0]
    funcs = {}
    table_refs = set()
    for sym in ("sin", "cos", "tan", "atan"):
        start = lib.ifunc_fma_variant(sym)
        ins = function_body(lib, start)
        for a, text in ins:
            m = re.match(r"lea\s+.*\(%rip\),%\w+\s+#\s+([0-9a-f]+)", text)
            if m:
                table_refs.add(int(m.group(1), 16))
        funcs[sym] = (start, ins)
    # the three look-up tables of the routines, by the order of their addresses in .rodata:
    # __sincostab (440 doubles), atan's cij (241 x 7), tan's xfg (186 x 4)
    refs = sorted(table_refs)
    if len(refs) != 3:
        raise SystemExit("expected 3 table bases, found %s" % [hex(r) for r in refs])
    sizes = {refs[0]: ("ali_gl_sincostab", 440), refs[1]: ("ali_gl_atan_cij", 241 * 7), refs[2]: ("ali_gl_tan_xfg", 186 * 4)}
    parts = []
    parts.append("// ali_glibcmath.cuh -- GENERATED by tools/gen_glibc_math.py; do not edit.\n//\n"
                 "// sin / cos / tan / atan that return exactly what the host's glibc returns (the libm the reference\n"
                 "// runs on): an instruction-by-instruction restatement of glibc's FMA variants of the IBM Accurate\n"
                 "// Mathematical Library routines (sysdeps/ieee754/dbl-64/s_sin.c, s_tan.c, s_atan.c; LGPL-2.1+),\n"
                 "// derived from %s, %s.\n"
                 "// Round-to-nearest only; arguments beyond glibc's medium range fall back to the platform's function.\n"
                 "#pragma once\n#include <stdint.h>\n#include <string.h>\n#include <math.h>\n\n"
                 "#if defined(__CUDACC__)\n#if defined(ALI_GL_NOINLINE)\n#define ALI_GL_DEV __device__ __noinline__\n#else\n#define ALI_GL_DEV __device__ __forceinline__\n#endif\n#define ALI_GL_INL __device__ __forceinline__\n"
                 "#define ALI_GL_TABLE static __device__ const\n"
                 "#define ALI_GL_FMA(a, b, c) __fma_rn((a), (b), (c))\n"
                 "#define ALI_GL_D(u) __longlong_as_double((long long)(u))\n#define ALI_GL_B(d) ((uint64_t)__double_as_longlong(d))\n"
                 "#else\n#define ALI_GL_DEV static inline\n#define ALI_GL_INL static inline\n#define ALI_GL_TABLE static const\n"
                 "#define ALI_GL_FMA(a, b, c) fma((a), (b), (c))\n"
                 "static inline double ALI_GL_D(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }\n"
                 "static inline uint64_t ALI_GL_B(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }\n#endif\n"
                 % (os.path.basename(args.libm), version))
    for addr in refs:
        parts.append(table_c(lib, sizes[addr][0], addr, sizes[addr][1]))
    parts.append("\n".join("#define %s_COUNT %d" % (sizes[a][0].upper(), sizes[a][1]) for a in refs))
    parts.append("// Table look-ups keep the address arithmetic of the machine code (base + 8 * index); each routine uses one table.")
    fall = {"sin": "sin(x)", "cos": "cos(x)", "tan": "tan(x)", "atan": "atan(x)"}
    for sym in ("sin", "cos", "tan", "atan"):
        start, ins = funcs[sym]
        tr = Translator(lib, "ali_glibc_" + sym, ins, sizes, fall[sym])
        mine = [a for a in refs if any(re.search(r"#\s+%x\b" % a, t) for _, t in ins)]
        assert len(mine) == 1, (sym, mine)
        tr.table = (mine[0], sizes[mine[0]])
        parts.append("// %s: %d instructions at %s+0x%x\n%s" % (sym, len(ins), os.path.basename(args.libm), start, tr.function()))
    open(args.out, "w").write("\n\n".join(parts) + "\n")
    print("wrote", args.out, {k: len(v[1]) for k, v in funcs.items()})


if __name__ == "__main__":
    main()
