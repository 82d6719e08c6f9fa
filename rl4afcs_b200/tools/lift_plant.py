"""BUILD TOOL -- static translation of the reference's plant binary into portable C.

The reference's nonlinear aircraft (`envs/nonlinear/<variant>/_citation.cp39-win_amd64.pyd`, called at
/root/reference/envs/nonlinear/env.py:210,288-291 through envs/nonlinear/citation.py:62-69) exists only as x86-64 machine
code.  This script TRANSLATES the model
code (Simulink entry points initialize 0x96f0, step 0x3720, terminate 0xe620 and everything they reach) instruction by
instruction into C: every x86 function becomes a C function over an explicit machine state (16 integer registers, 16 SSE
registers, 5 flags; include/rl4_lift_runtime.h) and a flat byte image of the DLL's sections + heap + stack.  Nothing of the model is interpreted or
approximated -- each instruction is replaced by its architectural semantics -- so the translation computes what the binary
computes, and it compiles for any target: nvcc (the `dasmat` plant of the CUDA kernels, csrc/dasmat_plant.cu) and gcc (the
CPU build the test suite checks bit for bit against the binary executing natively, tests/test_citation_lifted.py).

Nothing derived from the binary is committed: the generated sources and the image land in a git-ignored directory and are
produced at build time where /root/reference exists (the build container); the GPU box receives the compiled libraries.

Usage: python rl4afcs_b200/tools/lift_plant.py [--variant extended_input] --out DIR
"""
from __future__ import annotations

import argparse
import os
import re
import struct
import subprocess
import sys

BASE = 0x180000000
REF_DIR = "/root/reference/envs/nonlinear"
ENTRY = {"initialize": 0x96f0, "step": 0x3720, "terminate": 0xe620}
MODEL_END = 0xe840      # RVA below which .text is Simulink-generated model code (above: rt helpers, SWIG wrappers, CRT)

R64 = ["rax", "rcx", "rdx", "rbx", "rsp", "rbp", "rsi", "rdi", "r8", "r9", "r10", "r11", "r12", "r13", "r14", "r15"]
R32 = ["eax", "ecx", "edx", "ebx", "esp", "ebp", "esi", "edi"] + [f"r{i}d" for i in range(8, 16)]
R16 = ["ax", "cx", "dx", "bx", "sp", "bp", "si", "di"] + [f"r{i}w" for i in range(8, 16)]
R8 = ["al", "cl", "dl", "bl", "spl", "bpl", "sil", "dil"] + [f"r{i}b" for i in range(8, 16)]
R8H = {"ah": 0, "ch": 1, "dh": 2, "bh": 3}
REG = {}
for i in range(16):
    REG[R64[i]] = (i, 64)
    REG[R32[i]] = (i, 32)
    REG[R16[i]] = (i, 16)
    REG[R8[i]] = (i, 8)
SIZES = {"BYTE": 8, "WORD": 16, "DWORD": 32, "QWORD": 64, "XMMWORD": 128}
CC = {"e": "ZF", "z": "ZF", "ne": "!ZF", "nz": "!ZF", "a": "(!CF&&!ZF)", "nbe": "(!CF&&!ZF)", "ae": "!CF", "nb": "!CF", "nc": "!CF",
      "b": "CF", "c": "CF", "nae": "CF", "be": "(CF||ZF)", "na": "(CF||ZF)", "l": "(SF!=OF)", "nge": "(SF!=OF)", "ge": "(SF==OF)",
      "nl": "(SF==OF)", "le": "(ZF||SF!=OF)", "ng": "(ZF||SF!=OF)", "g": "(!ZF&&SF==OF)", "nle": "(!ZF&&SF==OF)", "s": "SF", "ns": "!SF",
      "p": "PF", "pe": "PF", "np": "!PF", "po": "!PF", "o": "OF", "no": "!OF"}
# C-runtime imports the model code may reach, with their argument shape
IMPORTS = {"cos": "d_d", "sin": "d_d", "tan": "d_d", "exp": "d_d", "floor": "d_d", "log10": "d_d", "sqrt": "d_d", "pow": "d_dd",
           "memcpy": "mem", "memset": "mem", "malloc": "mem", "free": "mem"}


class Ins:
    __slots__ = ("addr", "size", "mn", "ops", "raw")

    def __init__(self, addr, size, mn, ops, raw):
        self.addr, self.size, self.mn, self.ops, self.raw = addr, size, mn, ops, raw


def disassemble(path):
    txt = subprocess.run(["objdump", "-d", "-M", "intel", "--no-show-raw-insn", path], check=True, capture_output=True, text=True).stdout
    raw = subprocess.run(["objdump", "-d", "-M", "intel", path], check=True, capture_output=True, text=True).stdout
    # sizes from the raw listing (continuation lines carry only bytes)
    size = {}
    last = None
    for line in raw.splitlines():
        m = re.match(r"\s*([0-9a-f]+):\t([0-9a-f ]+?)\s*(\t.*)?$", line)
        if not m:
            continue
        a = int(m.group(1), 16)
        nb = len(m.group(2).split())
        if m.group(3) is not None and m.group(3).strip():
            size[a] = nb
            last = a
        elif last is not None:
            size[last] += nb
    ins = {}
    for line in txt.splitlines():
        m = re.match(r"\s*([0-9a-f]+):\t(.*)$", line)
        if not m:
            continue
        a = int(m.group(1), 16)
        body = m.group(2).strip()
        if not body or a not in size:
            continue
        body = re.sub(r"\s+#.*$", "", body)          # objdump's resolved-address comment (recomputed below)
        body = re.sub(r"<[^>]*>", "", body).strip()
        parts = body.split(None, 1)
        mn = parts[0]
        while mn in ("rex", "rex.W", "rex.WB", "rex.R", "rex.X", "rex.B", "rex.WR", "rex.WX", "rex.RB", "rex.XB", "rex.RX",
                     "data16", "lock", "bnd", "notrack", "cs", "ds", "es", "ss") and len(parts) > 1:
            parts = parts[1].split(None, 1)
            mn = parts[0]
        if mn == "rep" or mn == "repz" or mn == "repnz":
            mn = mn + " " + (parts[1] if len(parts) > 1 else "")
            ops = []
        else:
            ops = split_ops(parts[1]) if len(parts) > 1 else []
        ins[a] = Ins(a, size[a], mn, ops, body)
    return ins


def split_ops(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "[":
            depth += 1
        elif ch == "]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


class PE:
    def __init__(self, path):
        b = open(path, "rb").read()
        self.file = b
        nt = struct.unpack_from("<I", b, 0x3c)[0]
        assert b[nt:nt + 4] == b"PE\0\0"
        nsec = struct.unpack_from("<H", b, nt + 6)[0]
        optsz = struct.unpack_from("<H", b, nt + 20)[0]
        opt = nt + 24
        assert struct.unpack_from("<H", b, opt)[0] == 0x20b
        assert struct.unpack_from("<Q", b, opt + 24)[0] == BASE
        self.image_size = struct.unpack_from("<I", b, opt + 56)[0]
        hdr = struct.unpack_from("<I", b, opt + 60)[0]
        dirs = opt + 112
        img = bytearray(self.image_size)
        img[:hdr] = b[:hdr]
        self.sections = []
        sec = opt + optsz
        for i in range(nsec):
            name = b[sec:sec + 8].rstrip(b"\0").decode()
            vsz, va, rsz, ro = struct.unpack_from("<IIII", b, sec + 8)
            n = min(rsz, vsz)
            img[va:va + n] = b[ro:ro + n]
            self.sections.append((name, va, vsz))
            sec += 40
        self.img = img
        # imports: IAT slot VA -> name
        self.iat = {}
        irva = struct.unpack_from("<I", b, dirs + 8)[0]
        d = irva
        while struct.unpack_from("<I", img, d + 12)[0]:
            oft = struct.unpack_from("<I", img, d)[0] or struct.unpack_from("<I", img, d + 16)[0]
            ft = struct.unpack_from("<I", img, d + 16)[0]
            k = 0
            while True:
                ent = struct.unpack_from("<Q", img, oft + k)[0]
                if not ent:
                    break
                name = "?ordinal"
                if not ent >> 63:
                    p = (ent & 0xffffffff) + 2
                    name = bytes(img[p:img.index(b"\0", p)]).decode()
                self.iat[BASE + ft + k] = name
                k += 8
            d += 20
        # absolute pointers in the image (base relocation table): candidates for function pointers
        self.abs_ptrs = []
        rrva, rsz = struct.unpack_from("<II", b, dirs + 5 * 8)
        o = 0
        while o + 8 <= rsz:
            page, blk = struct.unpack_from("<II", img, rrva + o)
            if blk < 8:
                break
            for k in range(8, blk, 2):
                e = struct.unpack_from("<H", img, rrva + o + k)[0]
                if e >> 12 == 10:
                    self.abs_ptrs.append(page + (e & 0xfff))
            o += blk
        self.text = next((va, vsz) for n, va, vsz in self.sections if n == ".text")

    def in_text(self, va):
        return BASE + self.text[0] <= va < BASE + self.text[0] + self.text[1]


# ---------------------------------------------------------------------------------------------------------------------
ARG_REGS = [1, 2, 8, 9]          # rcx rdx r8 r9: integer arguments of the Windows x64 convention (xmm0-3: floating point)


class Lifter:
    """x86-64 -> C.  Inside a translated function the machine registers are C LOCALS (r0..r15, x0l/x0h..x15l/x15h, fl):
    the C compiler allocates them, removes dead flag computations and keeps values in registers across instructions.
    Registers cross function boundaries the way the Windows x64 calling convention says they do: rcx rdx r8 r9 xmm0-3 and rsp
    are written to the shared `cpu_t` before a call, rax and xmm0 are read back after it; callee-saved registers need no
    traffic (the callee's prologue / epilogue saves and restores whatever its own locals hold), volatile ones are dead."""

    STEP, SOLVER = BASE + 0x3720, BASE + 0x1c80
    COOKIE_CHECK = BASE + 0x11870          # MSVC's __security_check_cookie(rcx): compares with a constant of the image, no effect
    # Specialised S-function copies that call nothing but C-runtime imports run on a PRIVATE frame: a C array local to the
    # translated function, addressed from a compile-time-constant stack pointer.  The C compiler then sees every frame access
    # at a constant offset of a non-escaping array (after unrolling the 3-trip matrix loops) and keeps the x86 spill slots and
    # the 3x3 work matrices in registers instead of thread-local memory.  Virtual range: the never-used bottom of the stack region.
    LEAF_LO, LEAF_SIZE, LEAF_ENTRY_RSP = BASE + 0x40000, 0x800, BASE + 0x40800 - 0x48

    def __init__(self, pe, ins):
        self.pe, self.ins = pe, ins
        self.funcs = {}          # entry VA -> sorted list of instruction addresses
        self.unknown = {}
        self.frame_reg = None
        self.const_regs = {}
        self.region_hints = True
        self.leaf_private = False
        self.frozen = None          # second pass: the image after initialize() (bytes); loads from its constant part are folded
        self.clones = {}            # (function, rcx) -> name of the copy specialised for that first argument
        self.clone_work = []

    # ---- operands
    def is_stack_operand(self, s):
        inner = s[s.index("[") + 1:s.rindex("]")]
        base = re.split(r"[+\-*]", inner)[0].strip()
        if "*" in re.split(r"[+\-]", inner)[0]:
            return False
        return base == "rsp" or (self.frame_reg is not None and base == self.frame_reg) or self.const_regs.get(base) == "stk"

    def mem_addr(self, s, ins):
        """C expression of the effective address of a memory operand 'SIZE PTR [..]' or '[..]'."""
        inner = s[s.index("[") + 1:s.rindex("]")]
        terms = re.findall(r"([+-]?)\s*([^+-]+)", inner)
        parts = []
        for sign, t in terms:
            t = t.strip()
            if t == "rip":
                parts.append(f"0x{ins.addr + ins.size:x}ULL")
            elif "*" in t:
                r, sc = t.split("*")
                parts.append(f"{sign}(r{REG[r][0]}*{sc}ULL)" if sign == "-" else f"r{REG[r][0]}*{sc}ULL")
            elif t in REG:
                assert REG[t][1] == 64, s
                parts.append(f"{sign}r{REG[t][0]}")
            else:
                v = int(t, 16)
                parts.append(f"-0x{v:x}ULL" if sign == "-" else f"0x{v:x}ULL")
        expr = parts[0]
        for p in parts[1:]:
            expr += p if p.startswith("-") else "+" + p
        return "(" + expr + ")"

    def mem_a32(self, s, ins):
        """The LOW 32 BITS of the effective address, as a uint32_t C expression: every object of the emulated address space
        lies within 2 GB of the image base, so memory operands are addressed with 32-bit arithmetic (half the integer work
        on a 32-bit machine like the GPU).  A register the function holds constant (const_regs) contributes its value."""
        inner = s[s.index("[") + 1:s.rindex("]")]
        terms = re.findall(r"([+-]?)\s*([^+-]+)", inner)
        const, parts = 0, []
        for sign, t in terms:
            t = t.strip()
            sg = -1 if sign == "-" else 1
            if t == "rip":
                const += ins.addr + ins.size
            elif "*" in t:
                r, sc = t.split("*")
                assert sg == 1
                if isinstance(self.const_regs.get(r), int):
                    const += self.const_regs[r] * int(sc)
                else:
                    parts.append(f"(uint32_t)r{REG[r][0]}*{sc}u")
            elif t in REG:
                assert REG[t][1] == 64 and sg == 1, s
                if isinstance(self.const_regs.get(t), int):
                    const += self.const_regs[t]
                else:
                    parts.append(f"(uint32_t)r{REG[t][0]}")
            else:
                const += sg * int(t, 16)
        parts.append(f"0x{const & 0xffffffff:x}u")
        self.last_const, self.last_dynamic = const & 0xffffffffffffffff, len(parts) > 1
        return "(" + "+".join(parts) + ")"

    # The constant part of an address (rip-relative displacement, or a base register the analysis knows + displacement) names
    # the OBJECT the operand addresses; an index register only moves inside that object.  Objects never straddle the regions
    # of the memory model, so the region is known at translation time and the access needs no run-time decoding:
    #   S  the stack (rsp / frame-pointer based)            W  the writable window of .data (block signals, states, work arrays)
    #   I  the rest of the image (code constants, tables, parameters: read-only while stepping)      ''  unknown: decode
    # The CPU build checks every hinted access against the region it claims (tests/test_citation_lifted.py runs it).
    WINDOW = (0x3a000, 0x3c200)
    IMAGE_RO = (0x13000, 0x3a000)          # .rdata .. the start of the window, minus the few words step() writes there
    IMAGE_RW = (0x2eb00, 0x2ec00)
    FROZEN_HOLE = (0x2eb48, 0x2eb60)       # the model time inside the rtModel structure: the only word of that region step() writes
    # parts of the writable window that step() never writes (write tracking on the CPU build, all test scenarios): they hold the
    # port-pointer arrays of the S-function blocks, set by initialize().  Folding them makes the specialised S-function copies
    # fully static.  The memory models enforce the assumption: a store into a gap raises an error (CUDA) / traps (CPU build).
    FROZEN_GAPS = ((0x3a000, 0x3a078), (0x3a4b8, 0x3a540), (0x3a574, 0x3a5e8), (0x3a5f0, 0x3ab80))

    def frozen_value(self, addr, nbytes):
        """the value at `addr` if it lies in the part of the image that nothing writes after initialize(), else None"""
        if self.frozen is None:
            return None
        off = (addr & 0xffffffffffffffff) - BASE
        for lo, hi in self.FROZEN_GAPS:
            if lo <= off and off + nbytes <= hi:
                return int.from_bytes(self.frozen[off:off + nbytes], "little")
        if not (self.IMAGE_RO[0] <= off and off + nbytes <= self.IMAGE_RO[1]):
            return None
        if off + nbytes > self.FROZEN_HOLE[0] and off < self.FROZEN_HOLE[1]:
            return None
        return int.from_bytes(self.frozen[off:off + nbytes], "little")

    def frozen_operand(self, s, ins, nbytes):
        """value of a memory operand whose address is a translation-time constant inside the frozen image, else None"""
        if self.frozen is None or "[" not in s:
            return None
        a = self.fold_address(s, ins, self.const_regs)
        return None if a is None else self.frozen_value(a, nbytes)

    def leaf_eligible(self, body):
        depth = 0
        for a in body:
            i = self.ins[a]
            if i.mn == "call":
                t = self.direct_target(i)
                if t is None:
                    if not ("[rip" in i.ops[0] and self.const_addr(i.ops[0], i) in self.pe.iat):
                        return False
                elif not (self.is_import_thunk(t) or t == self.COOKIE_CHECK):
                    return False
            elif i.mn == "jmp":
                t = self.direct_target(i)
                if t is None or t not in set(body):
                    return False
            elif i.mn == "push":
                depth += 8
            elif i.mn == "sub" and i.ops[0] == "rsp":
                if not re.match(r"0x[0-9a-f]+$", i.ops[1]):
                    return False
                depth += int(i.ops[1], 16)
            elif i.mn.startswith("rep "):
                return False
        return depth + 0x100 < self.LEAF_SIZE - 0x48

    def hint(self, s):
        if self.is_stack_operand(s):
            return "F" if self.leaf_private else "S"
        if not self.region_hints:
            return ""
        inner = s[s.index("[") + 1:s.rindex("]")]
        terms = [t.strip() for _, t in re.findall(r"([+-]?)\s*([^+-]+)", inner)]
        unscaled = [t for t in terms if t in REG]
        known = [t for t in unscaled if isinstance(self.const_regs.get(t), int)]
        if len(known) != len(unscaled) or ("rip" not in inner and not known):
            return ""                       # a pointer the analysis does not know, or only a displacement: decode at run time
        off = self.last_const - BASE
        if self.WINDOW[0] <= off < self.WINDOW[1]:
            return "W"
        if self.IMAGE_RO[0] <= off < self.IMAGE_RO[1] and not (self.IMAGE_RW[0] - 0x40 <= off < self.IMAGE_RW[1]):
            return "I"
        return ""

    def ld(self, w, s, ins):
        a = self.mem_a32(s, ins)
        return f"LD{self.hint(s)}{w}({a})"

    def step_reach(self):
        if not hasattr(self, "_step_reach"):
            self._step_reach = self.reachable([self.STEP])
        return self._step_reach

    def check_store(self, s, ins):
        if self.frozen is None or self.last_dynamic:
            return
        off = self.last_const - BASE
        for lo, hi in self.FROZEN_GAPS:
            if lo <= off < hi and self.cur_fn in self.step_reach():
                raise RuntimeError(f"store at {ins.addr:#x} into a region assumed constant after initialize(): {ins.raw}")

    def st(self, w, s, ins, val):
        a = self.mem_a32(s, ins)
        self.check_store(s, ins)
        k = self.hint(s)
        return f"ST{'' if k == 'I' else k}{w}({a},({val}));"

    def const_addr(self, s, ins):
        """VA if the operand is rip-relative, else None."""
        if "[rip" not in s:
            return None
        inner = s[s.index("[") + 1:s.rindex("]")]
        m = re.match(r"rip([+-])0x([0-9a-f]+)$", inner)
        v = int(m.group(2), 16)
        return (ins.addr + ins.size + (v if m.group(1) == "+" else -v)) & 0xffffffffffffffff

    def op_size(self, s, default=None):
        if "PTR" in s:
            return SIZES[s.split()[0]]
        if s in REG:
            return REG[s][1]
        if s in R8H:
            return 8
        if s.startswith("xmm"):
            return 128
        return default

    def rd(self, s, ins, size=None):
        """C expression reading an integer operand (zero-extended into uint64_t)."""
        if s in REG:
            i, w = REG[s]
            return f"r{i}" if w == 64 else f"(uint64_t)(uint{w}_t)r{i}"
        if s in R8H:
            return f"((r{R8H[s]}>>8)&0xff)"
        if "[" in s:
            w = SIZES[s.split()[0]] if "PTR" in s else size
            fv = self.frozen_operand(s, ins, w // 8)
            if fv is not None:
                return f"0x{fv:x}ULL"
            return self.ld(w, s, ins)
        v = int(s, 16) if s.startswith(("0x", "-0x")) else int(s)
        return f"0x{v & 0xffffffffffffffff:x}ULL"

    def wr(self, s, ins, val, size=None):
        """C statement writing integer `val` (a uint64_t expression) to an operand."""
        if s in REG:
            i, w = REG[s]
            if w == 64:
                return f"r{i}=({val});"
            if w == 32:
                return f"r{i}=(uint32_t)({val});"
            mask = (1 << w) - 1
            return f"r{i}=(r{i}&~0x{mask:x}ULL)|(({val})&0x{mask:x}ULL);"
        if s in R8H:
            i = R8H[s]
            return f"r{i}=(r{i}&~0xff00ULL)|((({val})&0xff)<<8);"
        w = SIZES[s.split()[0]] if "PTR" in s else size
        return self.st(w, s, ins, val)

    def xr(self, s):
        return int(s[3:])

    def frozen_lane(self, s, ins, lane):
        if self.frozen is None or "[" not in s:
            return None
        a = self.fold_address(s, ins, self.const_regs)
        return None if a is None else self.frozen_value(a + 8 * lane, 8)

    def xd(self, s, ins, lane=0):
        """double-valued expression of lane `lane` of an xmm register or of a memory operand"""
        if s.startswith("xmm"):
            return f"U2D(x{self.xr(s)}{'lh'[lane]})"
        fv = self.frozen_lane(s, ins, lane)
        if fv is not None:
            return f"U2D(0x{fv:x}ULL)"
        a = self.mem_a32(s, ins)
        k = self.hint(s)
        return f"LD{k}D({a}+{8 * lane}u)" if lane else f"LD{k}D({a})"

    def xu(self, s, ins, lane=0):
        """uint64-valued expression of a lane"""
        if s.startswith("xmm"):
            return f"x{self.xr(s)}{'lh'[lane]}"
        fv = self.frozen_lane(s, ins, lane)
        if fv is not None:
            return f"0x{fv:x}ULL"
        a = self.mem_a32(s, ins)
        k = self.hint(s)
        return f"LD{k}64({a}+{8 * lane}u)" if lane else f"LD{k}64({a})"

    # ---- discovery
    def flow(self, entry):
        seen, work = set(), [entry]
        while work:
            a = work.pop()
            while a not in seen:
                if a not in self.ins:
                    self.unknown.setdefault(entry, []).append(a)
                    break
                seen.add(a)
                i = self.ins[a]
                mn = i.mn
                if mn == "ret" or mn == "int3" or mn == "ud2":
                    break
                if mn == "jmp":
                    t = self.direct_target(i)
                    if t is not None and t not in self.entries and not self.is_import_thunk(t):
                        work.append(t)
                    break
                if mn.startswith("j"):
                    t = self.direct_target(i)
                    if t is not None:
                        work.append(t)
                a += i.size
        return sorted(seen)

    def direct_target(self, i):
        if len(i.ops) == 1 and re.match(r"0x[0-9a-f]+$", i.ops[0]):
            return int(i.ops[0], 16)
        return None

    def is_import_thunk(self, a):
        i = self.ins.get(a)
        return i is not None and i.mn == "jmp" and i.ops and "[rip" in i.ops[0] and self.const_addr(i.ops[0], i) in self.pe.iat

    def thunk_name(self, a):
        i = self.ins[a]
        return self.pe.iat[self.const_addr(i.ops[0], i)]

    def discover(self, roots):
        self.entries = set(roots)
        # address-taken code: lea reg,[rip+X] into .text, and relocated absolute pointers into .text
        self.addr_taken = set()
        for off in self.pe.abs_ptrs:
            v = struct.unpack_from("<Q", self.pe.img, off)[0]
            if self.pe.in_text(v) and v in self.ins:
                self.addr_taken.add(v)
                if v < BASE + MODEL_END:        # function pointers of the model (S-function methods); the CRT's stay untranslated
                    self.entries.add(v)
        done = set()
        while True:
            todo = [e for e in self.entries if e not in done]
            if not todo:
                break
            for e in todo:
                done.add(e)
                if self.is_import_thunk(e):
                    continue
                body = self.flow(e)
                self.funcs[e] = body
                for a in body:
                    i = self.ins[a]
                    if i.mn == "call" or i.mn == "jmp":
                        t = self.direct_target(i)
                        if t is not None and (i.mn == "call" or t in self.entries or self.is_import_thunk(t)):
                            self.entries.add(t)
                    if i.mn == "lea" and "[rip" in i.ops[1]:
                        t = self.const_addr(i.ops[1], i)
                        if self.pe.in_text(t) and t in self.ins:
                            self.addr_taken.add(t)
                            self.entries.add(t)
        # a jump to an entry that was discovered later than the function containing the jump must become a tail call
        for e in list(self.funcs):
            self.funcs[e] = self.flow(e)

    # ---- emission
    def reachable(self, roots):
        """functions reachable from `roots` through direct calls / tail calls, plus every address-taken function if any
        indirect call is reachable"""
        seen, work, indirect = set(), list(roots), False
        while work:
            e = work.pop()
            if e in seen or e not in self.funcs:
                continue
            seen.add(e)
            for a in self.funcs[e]:
                i = self.ins[a]
                if i.mn in ("call", "jmp"):
                    t = self.direct_target(i)
                    if t is not None and t in self.funcs:
                        work.append(t)
                    elif t is None and not ("[rip" in i.ops[0] and self.const_addr(i.ops[0], i) in self.pe.iat):
                        if not indirect:
                            indirect = True
                            work.extend(self.addr_taken & set(self.funcs))
        return seen

    INLINE_MAX = 0          # measured: inlining the two table-lookup helpers (99 sites) 1.81e7 -> 1.74e7 plant steps/s: code growth costs more than the call protocol

    def inlinable(self, e, depth=0):
        """small helpers (the table-lookup routines at RVA 0xe840 / 0xe8d0: ~770 calls per step) are emitted `inline`: the call
        protocol through cpu_t disappears, and their pointer arguments are often constants of the call site"""
        body = self.funcs.get(e)
        if body is None or len(body) > self.INLINE_MAX or depth > 2 or e in (self.STEP, self.SOLVER) or e in self.addr_taken:
            return False
        for a in body:
            i = self.ins[a]
            if i.mn in ("call", "jmp"):
                t = self.direct_target(i)
                if t is None:
                    if not ("[rip" in i.ops[0] and self.const_addr(i.ops[0], i) in self.pe.iat):
                        return False
                elif t not in set(body) and not self.is_import_thunk(t) and not (t in self.funcs and t != e and self.inlinable(t, depth + 1)):
                    return False
        return True

    def qual(self, e):
        return "LIFT_FN_INLINE" if self.inlinable(e) else "LIFT_FN"

    def emit_all(self, roots=None):
        out = []
        names = sorted(self.funcs if roots is None else self.reachable(roots))
        self.emitting = set(names)
        for e in names:
            out.append(f"{self.qual(e)} LIFT_RET f_{e:x}(LIFT_PARAMS);")
        with_step = self.STEP in self.emitting
        if with_step:
            out.append(f"LIFT_FN LIFT_RET f_{self.STEP:x}_m(LIFT_PARAMS);")
        out.append("")
        # indirect-call dispatcher over address-taken functions
        out.append("LIFT_FN LIFT_RET lift_dispatch(LIFT_PARAMS, uint64_t target) {")
        out.append("  switch (target) {")
        for e in sorted(self.addr_taken & self.emitting):
            out.append(f"  case 0x{e:x}ULL: LIFT_FORWARD(f_{e:x});")
        out.append("  default: LIFT_TRAP(\"indirect call to an address that is not a translated function\", target);")
        out.append("  }\n}\n")
        self.clones, self.clone_work = {}, []
        bodies = []
        for e in names:
            bodies.extend(self.emit_func(e))
        if with_step:
            bodies.extend(self.emit_func(self.STEP, minor=True))
        while self.clone_work:
            t, rcx = self.clone_work.pop()
            bodies.extend(self.emit_func(t, entry_consts={"rcx": rcx}, name=self.clones[(t, rcx)]))
        self.n_clones = len(self.clones)
        out.extend(f"LIFT_FN LIFT_RET {n}(LIFT_PARAMS);" for n in sorted(self.clones.values()))
        out.extend(bodies)
        return "\n".join(out)

    def import_call(self, name):
        """statement performing a C-runtime import on the LOCAL registers"""
        kind = IMPORTS.get(name)
        if kind == "d_d":
            return f"x0l=D2U(lift_{name}(U2D(x0l)));"
        if kind == "d_dd":
            return f"x0l=D2U(lift_{name}(U2D(x0l),U2D(x1l)));"
        if name == "memcpy":
            return "lift_memcpy(c,r1,r2,r8); r0=r1;"
        if name == "memset":
            return "lift_memset(c,r1,(int)r2,r8); r0=r1;"
        if name == "malloc":
            return "r0=lift_malloc(c,r1);"
        if name == "free":
            return ";"
        return f"LIFT_TRAP(\"unbound import {name}\", 0);"

    PRECALL = "LIFT_PRECALL;"       # locals -> cpu_t: rcx rdx r8 r9 rsp xmm0-3
    POSTCALL = "LIFT_POSTCALL;"     # cpu_t -> locals: rax xmm0

    # ---- constant registers (flow-sensitive) ------------------------------------------------------------------------------
    NO_DEST = {"cmp", "test", "push", "bt", "comisd", "ucomisd", "call", "jmp", "ret", "nop", "int3", "ud2"}
    VOLATILE = ("rax", "rcx", "rdx", "r8", "r9", "r10", "r11")

    def reg64(self, name):
        return R64[REG[name][0]] if name in REG else (R64[R8H[name]] if name in R8H else None)

    def transfer(self, i, st):
        """constants known after instruction `i`, given those known before it (dict 64-bit register name -> value)"""
        mn, ops = i.mn, i.ops
        if mn == "call":
            st = {r: v for r, v in st.items() if r not in self.VOLATILE}
            return st
        if mn.startswith("rep "):
            return {r: v for r, v in st.items() if r not in ("rcx", "rdi", "rsi")}
        if mn in ("cdq", "cqo"):
            return {r: v for r, v in st.items() if r != "rdx"}
        if mn == "cdqe":
            return {r: v for r, v in st.items() if r != "rax"}
        if mn in self.NO_DEST or mn.startswith("j") or not ops:
            return st
        killed = []
        d = ops[0]
        dr = self.reg64(d)
        if mn == "xchg":
            killed = [self.reg64(o) for o in ops if self.reg64(o)]
        elif mn == "imul" and len(ops) == 1:
            killed = ["rax", "rdx"]
        elif dr is not None:
            killed = [dr]
        if not killed:
            return st
        new = {r: v for r, v in st.items() if r not in killed}
        if dr is None or mn == "xchg":
            return new
        w = REG[d][1] if d in REG else 8
        val = None
        if mn == "lea" and w == 64:
            val = self.fold_address(ops[1], i, st)
            if val is None:
                # an address formed from the stack pointer (or a register that already holds one) points into the stack
                inner = ops[1][ops[1].index("[") + 1:ops[1].rindex("]")]
                first = re.split(r"[+\-]", inner)[0].strip()
                if first == "rsp" or st.get(first) == "stk":
                    val = "stk"
        elif mn in ("mov", "movabs") and w >= 32 and len(ops) == 2:
            src = ops[1]
            if re.match(r"-?0x[0-9a-f]+$", src):
                v = int(src, 16)
                val = v & (0xffffffff if w == 32 else 0xffffffffffffffff)
            elif src == "rsp" and w == 64:
                val = "stk"
            elif src in REG and REG[src][1] == w and R64[REG[src][0]] in st:
                v = st[R64[REG[src][0]]]
                val = v if v == "stk" else (v & 0xffffffff if w == 32 else v)
                if val == "stk" and w != 64:
                    val = None
            elif "[" in src and self.frozen is not None:
                a = self.fold_address(src, i, st)
                if a is not None:
                    val = self.frozen_value(a, w // 8)
        elif mn == "movsxd" and w == 64 and "[" in ops[1] and self.frozen is not None:
            a = self.fold_address(ops[1], i, st)
            v = None if a is None else self.frozen_value(a, 4)
            if v is not None:
                val = (v - (1 << 32) if v >> 31 else v) & 0xffffffffffffffff
        elif mn == "xor" and len(ops) == 2 and ops[0] == ops[1] and w >= 32:
            val = 0
        elif mn in ("add", "sub") and w == 64 and dr in st and re.match(r"-?0x[0-9a-f]+$", ops[1]):
            v = int(ops[1], 16)
            val = "stk" if st[dr] == "stk" else (st[dr] + (v if mn == "add" else -v)) & 0xffffffffffffffff
        if val is not None:
            new[dr] = val
        return new

    def fold_address(self, s, ins, st):
        """value of a memory-operand address expression if every register in it is a known constant, else None"""
        inner = s[s.index("[") + 1:s.rindex("]")]
        total = 0
        for sign, t in re.findall(r"([+-]?)\s*([^+-]+)", inner):
            t = t.strip()
            sg = -1 if sign == "-" else 1
            if t == "rip":
                total += ins.addr + ins.size
            elif "*" in t:
                r, sc = t.split("*")
                if not isinstance(st.get(r), int):
                    return None
                total += sg * st[r] * int(sc)
            elif t in REG:
                if REG[t][1] != 64 or not isinstance(st.get(t), int):
                    return None
                total += sg * st[t]
            else:
                total += sg * int(t, 16)
        return total & 0xffffffffffffffff

    def spine_calls(self, body, inside):
        """Call instructions of the function that EVERY execution passes exactly once: they dominate every `ret` and lie in
        no loop.  All aircraft of a CTA reach them the same number of times whatever their data, so the CUDA build may put a
        CTA barrier in front of them (LIFT_SYNC) to keep the warps on the same stretch of the code."""
        idx = {a: k for k, a in enumerate(body)}
        succ = [[] for _ in body]
        for a in body:
            i = self.ins[a]
            if i.mn.startswith("j"):
                t = self.direct_target(i)
                if t in inside:
                    succ[idx[a]].append(idx[t])
            if not self.ends_flow(i) and a + i.size in inside:
                succ[idx[a]].append(idx[a + i.size])
        n = len(body)
        pred = [[] for _ in body]
        for u, vs in enumerate(succ):
            for v in vs:
                pred[v].append(u)
        # dominators (bit sets), reverse-post-order iteration
        order, seen, stack = [], [False] * n, [(0, 0)]
        seen[0] = True
        while stack:
            u, k = stack.pop()
            if k < len(succ[u]):
                stack.append((u, k + 1))
                v = succ[u][k]
                if not seen[v]:
                    seen[v] = True
                    stack.append((v, 0))
            else:
                order.append(u)
        order.reverse()
        full = (1 << n) - 1
        dom = [full] * n
        dom[0] = 1
        changed = True
        while changed:
            changed = False
            for u in order[1:]:
                d = full
                for q in pred[u]:
                    if seen[q]:
                        d &= dom[q]
                d |= 1 << u
                if d != dom[u]:
                    dom[u] = d
                    changed = True
        exits = [idx[a] for a in body if self.ins[a].mn == "ret" and seen[idx[a]]]
        if not exits:
            return set()
        spine = full
        for e in exits:
            spine &= dom[e]
        # instructions on a cycle: reachable from one of their own successors
        on_cycle = set()
        for u in range(n):
            if not (spine >> u) & 1 or self.ins[body[u]].mn != "call":
                continue
            st, vis = list(succ[u]), set()
            while st:
                v = st.pop()
                if v == u:
                    on_cycle.add(u)
                    break
                if v in vis:
                    continue
                vis.add(v)
                st.extend(succ[v])
        return {body[u] for u in range(n) if (spine >> u) & 1 and self.ins[body[u]].mn == "call" and u not in on_cycle}

    def analyse_constants(self, body, inside, entry=None):
        """in-state (known constants) of every instruction of the function: forward data flow, meet = agreement"""
        ins_state = {}
        work = [(body[0], dict(entry or {}))]
        while work:
            a, st = work.pop()
            if a in ins_state:
                old = ins_state[a]
                merged = {r: v for r, v in old.items() if st.get(r) == v}
                if merged == old:
                    continue
                st = merged
            ins_state[a] = st
            i = self.ins[a]
            out = self.transfer(i, st)
            if i.mn.startswith("j"):
                t = self.direct_target(i)
                if t in inside:
                    work.append((t, out))
            if not self.ends_flow(i) and a + i.size in inside:
                work.append((a + i.size, out))
        return ins_state

    def find_frame_reg(self, body):
        """rbp when the function uses it as a frame pointer: written once by `lea rbp,[rsp+-c]` (or from rax = rsp at entry),
        otherwise only pushed / popped."""
        first = self.ins[body[0]]
        rax_is_rsp = first.mn == "mov" and first.ops == ["rax", "rsp"]
        ok = False
        for n, a in enumerate(body):
            i = self.ins[a]
            if not i.ops:
                continue
            d = i.ops[0]
            if d in ("rbp", "ebp", "bp", "bpl") and i.mn not in ("push", "cmp", "test"):
                if i.mn == "pop":
                    continue
                if i.mn == "lea" and d == "rbp" and n < 16 and (i.ops[1].startswith("[rsp") or (rax_is_rsp and i.ops[1].startswith("[rax"))):
                    if ok:
                        return None
                    ok = True
                    continue
                return None
            if i.mn == "xchg" and "rbp" in i.ops:
                return None
        if ok and rax_is_rsp:
            # rax must still be rsp at the lea: no write to rax before it
            for a in body[1:16]:
                i = self.ins[a]
                if i.mn == "lea" and i.ops[0] == "rbp":
                    break
                if i.ops and i.ops[0] in ("rax", "eax", "ax", "al") and i.mn not in ("push", "cmp", "test"):
                    return None
                if i.mn == "call":
                    return None
        return "rbp" if ok else None

    def emit_func(self, e, minor=False, entry_consts=None, name=None):
        self.cur_fn, self.cur_minor = e, minor
        body = self.funcs[e]
        inside = set(body)
        self.frame_reg = self.find_frame_reg(body)
        self.consts_at = self.analyse_constants(body, inside, entry_consts or {})
        self.const_regs = {}
        targets = set()
        for a in body:
            i = self.ins[a]
            if i.mn.startswith("j"):
                t = self.direct_target(i)
                if t in inside:
                    targets.add(t)
        leaders, cur, prev = {}, None, None
        for a in body:
            i = self.ins[a]
            if cur is None or a in targets or prev is None or prev.addr + prev.size != a or prev.mn.startswith("j") or prev.mn in ("call", "ret"):
                cur = a
                leaders[cur] = 0
            leaders[cur] += 1
            prev = i
        self.leaf_private = bool(name) and self.frozen is not None and self.leaf_eligible(body)
        if self.leaf_private:
            self.n_leaf = getattr(self, "n_leaf", 0) + 1
            # the private frame changes which operands are 'stack' operands of THIS function only; constants were analysed above
        out = [f"{'LIFT_FN' if (name or minor) else self.qual(e)} LIFT_RET {name or f'f_{e:x}' + ('_m' if minor else '')}(LIFT_PARAMS) {{",
               "  LIFT_LOCALS; LIFT_ENTER_LEAF;" if self.leaf_private else "  LIFT_LOCALS; LIFT_ENTER;"]
        sync_at = set()
        if e == self.STEP:
            out.append("  LIFT_SYNC;        /* every aircraft calls step() the same number of times: a convergent point */")
            sync_at = self.spine_calls(body, inside)
            self.n_spine = len(sync_at)
        prev_end = None
        for a in body:
            i = self.ins[a]
            if a in targets or (prev_end is not None and prev_end != a):
                out.append(f"L_{a:x}: ;")
            if a in leaders:
                out.append(f"  LIFT_BB(0x{e:x}ULL, {leaders[a]});")
            self.const_regs = self.consts_at.get(a, {})
            if a in sync_at:
                # two classes, so that the build can choose the barrier density: calls through a function pointer (the
                # model's S-function blocks, ~25 per evaluation) and direct calls of the table-lookup helpers (~90)
                kind = "LIFT_SYNC_SFUN" if self.direct_target(i) is None else "LIFT_SYNC_HELPER"
                out.append(f"  {kind};        /* a call every execution of step() passes exactly once */")
            try:
                code = self.emit_ins(i, inside, e)
            except Exception as ex:  # noqa: BLE001 - report the instruction and keep going: unreachable CRT code may be odd
                code = f"LIFT_TRAP(\"untranslated instruction\", 0x{a:x}ULL); /* {ex!r} */"
                self.unknown.setdefault(e, []).append((a, i.raw))
            out.append(f"  {code}   /* {a:x}: {i.raw} */")
            prev_end = a + i.size
            if not self.ends_flow(i) and prev_end not in inside:
                out.append(f"  LIFT_TRAP(\"fell off the translated code\", 0x{prev_end:x}ULL);")
        out.append("}\n")
        return out

    def ends_flow(self, i):
        return i.mn in ("ret", "jmp", "int3", "ud2")

    def operand_const(self, s, ins):
        """translation-time value of an integer operand (a register the analysis knows, or a frozen memory location)"""
        if s in REG and REG[s][1] == 64:
            v = self.const_regs.get(s)
            return v if isinstance(v, int) else None
        if "[" in s:
            return self.frozen_operand(s, ins, 8)
        return None

    def call_fn(self, t, tail=False, rcx=None):
        """statement(s) calling translated function / import thunk `t` (direct)"""
        if self.is_import_thunk(t):
            return self.import_call(self.thunk_name(t)) + (" LIFT_RETURN;" if tail else "")
        if t == self.SOLVER and self.cur_minor:
            return "LIFT_TRAP(\"solver reached from a minor step\", 0);"
        name = f"f_{t:x}_m" if (t == self.STEP and self.cur_fn == self.SOLVER) else f"f_{t:x}"
        if rcx is not None and self.frozen is not None and t not in (self.STEP, self.SOLVER):
            key = (t, rcx)
            if key not in self.clones:
                self.clones[key] = f"f_{t:x}_s{rcx & 0xffffffff:x}"
                self.clone_work.append(key)
            name = self.clones[key]
        if tail:
            # the callee returns to OUR caller: its `ret` pops the return address our caller pushed
            return f"LIFT_TAILCALL({name});"
        return f"r4-=8; LIFT_CALL({name}); r4+=8;"

    def goto(self, t, inside):
        if t in inside:
            return f"goto L_{t:x};"
        if t in self.funcs or self.is_import_thunk(t):
            return "{ " + self.call_fn(t, tail=True) + " }"
        return f"LIFT_TRAP(\"jump out of the translated code\", 0x{t:x}ULL);"

    def emit_ins(self, i, inside, fentry):
        mn, ops = i.mn, i.ops
        rd, wr = self.rd, self.wr
        if mn.startswith("nop"):
            return ";"
        if mn in ("int3", "ud2"):
            return f"LIFT_TRAP(\"{mn}\", 0x{i.addr:x}ULL);"
        if mn == "ret":
            return "LIFT_RETURN;"
        if mn == "call":
            t = self.direct_target(i)
            if t is not None:
                if t == self.COOKIE_CHECK:
                    return ";   /* __security_check_cookie: compares rcx with a constant of the image; no effect on a correct run */"
                if t in self.funcs or self.is_import_thunk(t):
                    return self.call_fn(t)
                return f"LIFT_TRAP(\"call to untranslated code\", 0x{t:x}ULL);"
            if "[rip" in ops[0]:
                slot = self.const_addr(ops[0], i)
                if slot in self.pe.iat:
                    return self.import_call(self.pe.iat[slot])
            tv = self.operand_const(ops[0], i)
            if tv is not None and tv in self.funcs:
                # the target is a constant of the initialised model (an S-function method pointer); when the first argument
                # (the block's SimStruct) is constant too, call a copy of the method specialised for it
                rcx = self.const_regs.get("rcx")
                return self.call_fn(tv, rcx=rcx if isinstance(rcx, int) else None)
            return f"{{ uint64_t t_={rd(ops[0], i, 64)}; r4-=8; LIFT_CALL_DISPATCH(t_); r4+=8; }}"
        if mn == "jmp":
            t = self.direct_target(i)
            if t is not None:
                return self.goto(t, inside)
            if "[rip" in ops[0]:
                slot = self.const_addr(ops[0], i)
                if slot in self.pe.iat:
                    return f"{self.import_call(self.pe.iat[slot])} LIFT_RETURN;"
            return f"{{ uint64_t t_={rd(ops[0], i, 64)}; LIFT_TAILCALL_DISPATCH(t_); }}"
        if mn.startswith("j") and mn[1:] in CC:
            return f"if ({CC[mn[1:]]}) {self.goto(self.direct_target(i), inside)}"
        if mn.startswith("set") and mn[3:] in CC:
            return wr(ops[0], i, f"({CC[mn[3:]]})?1:0", 8)
        if mn.startswith("cmov") and mn[4:] in CC:
            w = self.op_size(ops[0])
            # a 32-bit cmov zero-extends the destination even when the condition is false
            return f"if ({CC[mn[4:]]}) {{ {wr(ops[0], i, rd(ops[1], i, w))} }} else {{ {wr(ops[0], i, rd(ops[0], i, w))} }}"
        if mn in ("mov", "movabs"):
            w = self.op_size(ops[0]) or self.op_size(ops[1])
            return wr(ops[0], i, rd(ops[1], i, w), w)
        if mn == "movzx":
            return wr(ops[0], i, rd(ops[1], i))
        if mn in ("movsx", "movsxd"):
            ws = self.op_size(ops[1])
            return wr(ops[0], i, f"(uint64_t)(int64_t)(int{ws}_t)({rd(ops[1], i)})")
        if mn == "lea":
            return wr(ops[0], i, self.mem_addr(ops[1], i))
        if mn == "cdqe":
            return "r0=(uint64_t)(int64_t)(int32_t)r0;"
        if mn == "cdq":
            return "r2=(uint32_t)(((int32_t)r0)>>31);"
        if mn == "cqo":
            return "r2=(uint64_t)(((int64_t)r0)>>63);"
        if mn == "push":
            return f"r4-=8; ST{'F' if self.leaf_private else 'S'}64((uint32_t)r4,{rd(ops[0], i, 64)});"
        if mn == "pop":
            return f"{{ uint64_t t_=LD{'F' if self.leaf_private else 'S'}64((uint32_t)r4); r4+=8; {wr(ops[0], i, 't_')} }}"
        if mn == "xchg":
            w = self.op_size(ops[0]) or self.op_size(ops[1])
            return f"{{ uint64_t a_={rd(ops[0], i, w)}, b_={rd(ops[1], i, w)}; {wr(ops[0], i, 'b_', w)} {wr(ops[1], i, 'a_', w)} }}"
        if mn in ("add", "sub", "and", "or", "xor", "cmp", "test", "adc", "sbb"):
            w = self.op_size(ops[0]) or self.op_size(ops[1])
            a, b = rd(ops[0], i, w), rd(ops[1], i, w)
            if mn == "xor" and ops[0] == ops[1] and ops[0] in REG:
                z = f"r{REG[ops[0]][0]}=0;" if REG[ops[0]][1] >= 32 else wr(ops[0], i, "0")
                return z + " ZF=1; SF=0; CF=0; OF=0; PF=1;"
            fn = {"add": "ADD", "sub": "SUB", "and": "AND", "or": "OR", "xor": "XOR", "cmp": "SUB", "test": "AND", "adc": "ADC", "sbb": "SBB"}[mn]
            call = f"lift_{fn}{w}(&fl,{a},{b})"
            if mn in ("cmp", "test"):
                return f"(void){call};"
            return wr(ops[0], i, call, w)
        if mn in ("inc", "dec", "neg", "not"):
            w = self.op_size(ops[0])
            a = rd(ops[0], i, w)
            if mn == "not":
                return wr(ops[0], i, f"~({a})", w)
            return wr(ops[0], i, f"lift_{mn.upper()}{w}(&fl,{a})", w)
        if mn in ("shl", "sal", "shr", "sar", "rol", "ror"):
            w = self.op_size(ops[0])
            cnt = rd(ops[1], i, 8) if len(ops) > 1 else "1"
            fn = {"shl": "SHL", "sal": "SHL", "shr": "SHR", "sar": "SAR", "rol": "ROL", "ror": "ROR"}[mn]
            return wr(ops[0], i, f"lift_{fn}{w}(&fl,{rd(ops[0], i, w)},{cnt})", w)
        if mn == "imul":
            w = self.op_size(ops[0])
            if len(ops) == 1:
                raise NotImplementedError("one-operand imul")
            a = rd(ops[1] if len(ops) == 3 else ops[0], i, w)
            b = rd(ops[2] if len(ops) == 3 else ops[1], i, w)
            return wr(ops[0], i, f"lift_IMUL{w}(&fl,{a},{b})", w)
        if mn == "bt":
            w = self.op_size(ops[0])
            return f"CF=(({rd(ops[0], i, w)})>>(({rd(ops[1], i, 8)})&{w - 1}))&1;"
        if mn in ("btr", "bts", "btc"):
            w = self.op_size(ops[0])
            op = {"btr": "&~", "bts": "|", "btc": "^"}[mn]
            return (f"{{ uint64_t v_={rd(ops[0], i, w)}; unsigned n_=(unsigned)({rd(ops[1], i, 8)})&{w - 1}; CF=(v_>>n_)&1; "
                    f"v_=v_{op}(1ULL<<n_); {wr(ops[0], i, 'v_', w)} }}")
        # ---- SSE (xmm N = locals xNl, xNh)
        xu, xd, st = self.xu, self.xd, self.st
        if mn in ("movsd", "movq"):
            d, s = ops
            if d.startswith("xmm") and s.startswith("xmm"):
                n = self.xr(d)
                return f"x{n}l=x{self.xr(s)}l;" + (f" x{n}h=0;" if mn == "movq" else "")
            if d.startswith("xmm"):
                n = self.xr(d)
                return f"x{n}l={rd(s, i) if s in REG else xu(s, i)}; x{n}h=0;"
            if d in REG:
                return wr(d, i, f"x{self.xr(s)}l")
            return st(64, d, i, f"x{self.xr(s)}l")
        if mn == "movd":
            d, s = ops
            if d.startswith("xmm"):
                n = self.xr(d)
                return f"x{n}l=(uint32_t)({rd(s, i, 32)}); x{n}h=0;"
            return wr(d, i, f"(uint32_t)x{self.xr(s)}l", 32)
        if mn in ("movaps", "movups", "movapd", "movupd", "movdqa", "movdqu"):
            d, s = ops
            if d.startswith("xmm"):
                n = self.xr(d)
                if s.startswith("xmm"):
                    return f"x{n}l=x{self.xr(s)}l; x{n}h=x{self.xr(s)}h;"
                return f"{{ uint64_t p_={xu(s, i, 0)}, q_={xu(s, i, 1)}; x{n}l=p_; x{n}h=q_; }}"
            a = self.mem_a32(d, i)
            k = self.hint(d)
            k = "" if k == "I" else k
            return f"{{ uint32_t a_={a}; ST{k}64(a_,x{self.xr(s)}l); ST{k}64(a_+8u,x{self.xr(s)}h); }}"
        if mn in ("movlpd", "movhpd", "movlps", "movhps"):
            k = "l" if mn[3] == "l" else "h"
            d, s = ops
            if d.startswith("xmm"):
                return f"x{self.xr(d)}{k}={xu(s, i)};"
            return st(64, d, i, f"x{self.xr(s)}{k}")
        sc = {"addsd": "F_ADD", "subsd": "F_SUB", "mulsd": "F_MUL", "divsd": "F_DIV", "maxsd": "F_MAX", "minsd": "F_MIN"}
        if mn in sc:
            d, s = ops
            n = self.xr(d)
            return f"x{n}l=D2U({sc[mn]}(U2D(x{n}l),{xd(s, i)}));"
        if mn == "sqrtsd":
            d, s = ops
            return f"x{self.xr(d)}l=D2U(F_SQRT({xd(s, i)}));"
        pk = {"addpd": "F_ADD", "subpd": "F_SUB", "mulpd": "F_MUL", "divpd": "F_DIV", "maxpd": "F_MAX", "minpd": "F_MIN"}
        if mn in pk:
            d, s = ops
            n = self.xr(d)
            return (f"{{ double p_={pk[mn]}(U2D(x{n}l),{xd(s, i, 0)}), q_={pk[mn]}(U2D(x{n}h),{xd(s, i, 1)}); "
                    f"x{n}l=D2U(p_); x{n}h=D2U(q_); }}")
        if mn == "sqrtpd":
            d, s = ops
            n = self.xr(d)
            return f"{{ double p_=F_SQRT({xd(s, i, 0)}), q_=F_SQRT({xd(s, i, 1)}); x{n}l=D2U(p_); x{n}h=D2U(q_); }}"
        if mn in ("unpcklpd", "movlhps"):
            d, s = ops
            return f"x{self.xr(d)}h={xu(s, i, 0)};"
        if mn == "unpckhpd":
            d, s = ops
            n = self.xr(d)
            return f"{{ uint64_t v_={xu(s, i, 1)}; x{n}l=x{n}h; x{n}h=v_; }}"
        if mn == "movhlps":
            d, s = ops
            return f"x{self.xr(d)}l={xu(s, i, 1)};"
        if mn == "shufpd":
            d, s, imm = ops
            k = int(imm, 16)
            n = self.xr(d)
            return f"{{ uint64_t p_=x{n}{'lh'[k & 1]}, q_={xu(s, i, (k >> 1) & 1)}; x{n}l=p_; x{n}h=q_; }}"
        bw = {"xorps": "^", "xorpd": "^", "pxor": "^", "andps": "&", "andpd": "&", "pand": "&", "orps": "|", "orpd": "|", "por": "|"}
        if mn in bw:
            d, s = ops
            n = self.xr(d)
            if s.startswith("xmm") and self.xr(s) == n and bw[mn] == "^":
                return f"x{n}l=0; x{n}h=0;"
            return f"{{ uint64_t p_={xu(s, i, 0)}, q_={xu(s, i, 1)}; x{n}l{bw[mn]}=p_; x{n}h{bw[mn]}=q_; }}"
        if mn in ("andnps", "andnpd", "pandn"):
            d, s = ops
            n = self.xr(d)
            return f"{{ uint64_t p_={xu(s, i, 0)}, q_={xu(s, i, 1)}; x{n}l=~x{n}l&p_; x{n}h=~x{n}h&q_; }}"
        if mn in ("comisd", "ucomisd"):
            d, s = ops
            return f"lift_COMISD(&fl,U2D(x{self.xr(d)}l),{xd(s, i)});"
        if mn == "cvtsi2sd":
            d, s = ops
            w = self.op_size(s)
            return f"x{self.xr(d)}l=D2U((double)(int{w}_t)({rd(s, i, w)}));"
        if mn in ("cvttsd2si", "cvtsd2si"):
            d, s = ops
            w = self.op_size(d)
            fn = "lift_CVTT" if mn == "cvttsd2si" else "lift_CVTR"
            return wr(d, i, f"{fn}{w}({xd(s, i)})", w)
        if mn == "cvtdq2pd":
            d, s = ops
            n = self.xr(d)
            return f"{{ uint64_t v_={xu(s, i, 0)}; x{n}l=D2U((double)(int32_t)(uint32_t)v_); x{n}h=D2U((double)(int32_t)(uint32_t)(v_>>32)); }}"
        if mn == "cvtps2pd":
            d, s = ops
            n = self.xr(d)
            return f"{{ uint64_t v_={xu(s, i, 0)}; x{n}l=D2U((double)lift_u2f((uint32_t)v_)); x{n}h=D2U((double)lift_u2f((uint32_t)(v_>>32))); }}"
        if mn == "cvtss2sd":
            d, s = ops
            src = f"x{self.xr(s)}l" if s.startswith("xmm") else self.ld(32, s, i)
            return f"x{self.xr(d)}l=D2U((double)lift_u2f((uint32_t)({src})));"
        if mn == "cvtsd2ss":
            d, s = ops
            n = self.xr(d)
            return f"x{n}l=(x{n}l&~0xffffffffULL)|lift_f2u((float)({xd(s, i)}));"
        if mn == "movss":
            d, s = ops
            if d.startswith("xmm") and s.startswith("xmm"):
                n = self.xr(d)
                return f"x{n}l=(x{n}l&~0xffffffffULL)|(x{self.xr(s)}l&0xffffffffULL);"
            if d.startswith("xmm"):
                n = self.xr(d)
                return f"x{n}l={self.ld(32, s, i)}; x{n}h=0;"
            return st(32, d, i, f"x{self.xr(s)}l")
        if mn.startswith("rep stos"):
            w = SIZES[mn.split()[2]]
            return f"lift_REPSTOS(c,&r1,&r7,r0,{w // 8});"
        if mn.startswith("rep movs"):
            w = SIZES[mn.split()[2]]
            return f"lift_REPMOVS(c,&r1,&r7,&r6,{w // 8});"
        raise NotImplementedError(mn)


PRELUDE = r"""/* GENERATED by rl4afcs_b200/tools/lift_plant.py from the reference's plant binary -- do not edit, do not commit. */
#ifndef LIFT_FN
#define LIFT_FN static
#endif
#ifndef LIFT_FN_INLINE
#define LIFT_FN_INLINE static inline
#endif
#ifndef LIFT_RESTRICT
#define LIFT_RESTRICT __restrict__
#endif
#ifndef LIFT_BB
#define LIFT_BB(fn, n)
#endif
#ifndef LIFT_MEM_CTX
#define LIFT_MEM_CTX
#endif
#ifndef LIFT_SYNC
#define LIFT_SYNC
#endif
#ifndef LIFT_SYNC_SFUN
#define LIFT_SYNC_SFUN
#endif
#ifndef LIFT_SYNC_HELPER
#define LIFT_SYNC_HELPER
#endif
#define ZF (fl.zf)
#define SF (fl.sf)
#define CF (fl.cf)
#define OF (fl.of)
#define PF (fl.pf)
#define LIFT_LOCALS \
    uint64_t r0=0,r1=0,r2=0,r3=0,r4=0,r5=0,r6=0,r7=0,r8=0,r9=0,r10=0,r11=0,r12=0,r13=0,r14=0,r15=0; \
    uint64_t x0l=0,x0h=0,x1l=0,x1h=0,x2l=0,x2h=0,x3l=0,x3h=0,x4l=0,x4h=0,x5l=0,x5h=0,x6l=0,x6h=0,x7l=0,x7h=0; \
    uint64_t x8l=0,x8h=0,x9l=0,x9h=0,x10l=0,x10h=0,x11l=0,x11h=0,x12l=0,x12h=0,x13l=0,x13h=0,x14l=0,x14h=0,x15l=0,x15h=0; \
    lift_flags fl = {0,0,0,0,0}; LIFT_MEM_CTX \
    (void)r0;(void)r1;(void)r2;(void)r3;(void)r5;(void)r6;(void)r7;(void)r8;(void)r9;(void)r10;(void)r11;(void)r12;(void)r13;(void)r14;(void)r15; \
    (void)x0h;(void)x1l;(void)x1h;(void)x2l;(void)x2h;(void)x3l;(void)x3h;(void)x4l;(void)x4h;(void)x5l;(void)x5h;(void)x6l;(void)x6h;(void)x7l;(void)x7h; \
    (void)x8l;(void)x8h;(void)x9l;(void)x9h;(void)x10l;(void)x10h;(void)x11l;(void)x11h;(void)x12l;(void)x12h;(void)x13l;(void)x13h;(void)x14l;(void)x14h;(void)x15l;(void)x15h;(void)fl
/* How registers cross a function boundary (Windows x64 convention: rcx rdx r8 r9 xmm0-3 rsp in, rax xmm0 out):
 * LIFT_REG_PROTOCOL 1 -- as C parameters / a two-word return value (they travel in machine registers);
 * LIFT_REG_PROTOCOL 0 -- through the shared cpu_t (memory). */
#ifndef LIFT_REG_PROTOCOL
#define LIFT_REG_PROTOCOL 1
#endif
/* private frame of a leaf function: a local array addressed from a constant stack pointer (tools/lift_plant.py: LEAF_*) */
#define LIFT_LEAF_LO32 0x80040000u
#define LIFT_LEAF_RSP 0x1800407b8ULL
#define LIFT_LEAF_DECL uint8_t lf_[0x800] __attribute__((aligned(16)))
#if LIFT_REG_PROTOCOL
#define LIFT_RET lift_ret
#define LIFT_PARAMS cpu_t* LIFT_RESTRICT c, uint64_t a1_, uint64_t a2_, uint64_t a8_, uint64_t a9_, uint64_t ax0_, uint64_t ax1_, uint64_t ax2_, uint64_t ax3_, uint64_t a4_
#define LIFT_ARGS c, r1, r2, r8, r9, x0l, x1l, x2l, x3l, r4
#define LIFT_ENTER r1=a1_; r2=a2_; r8=a8_; r9=a9_; r4=a4_; x0l=ax0_; x1l=ax1_; x2l=ax2_; x3l=ax3_
#define LIFT_ENTER_LEAF LIFT_LEAF_DECL; r1=a1_; r2=a2_; r8=a8_; r9=a9_; r4=LIFT_LEAF_RSP; (void)a4_; x0l=ax0_; x1l=ax1_; x2l=ax2_; x3l=ax3_
#define LIFT_CALL(f) do { const lift_ret t__ = f(LIFT_ARGS); r0 = t__.rax; x0l = t__.x0; } while (0)
#define LIFT_TAILCALL(f) return f(LIFT_ARGS)
#define LIFT_FORWARD(f) return f(c, a1_, a2_, a8_, a9_, ax0_, ax1_, ax2_, ax3_, a4_)
#define LIFT_CALL_DISPATCH(t) do { const lift_ret t__ = lift_dispatch(LIFT_ARGS, (t)); r0 = t__.rax; x0l = t__.x0; } while (0)
#define LIFT_TAILCALL_DISPATCH(t) return lift_dispatch(LIFT_ARGS, (t))
#define LIFT_RETURN return lift_mkret(r0, x0l)
#define LIFT_TRAP_RETURN return lift_mkret(0, 0)
#else
#define LIFT_RET void
#define LIFT_PARAMS cpu_t* LIFT_RESTRICT c
#define LIFT_ENTER \
    r1=c->r[1]; r2=c->r[2]; r8=c->r[8]; r9=c->r[9]; r4=c->r[4]; \
    x0l=c->x[0].u[0]; x1l=c->x[1].u[0]; x2l=c->x[2].u[0]; x3l=c->x[3].u[0]
#define LIFT_ENTER_LEAF \
    LIFT_LEAF_DECL; r1=c->r[1]; r2=c->r[2]; r8=c->r[8]; r9=c->r[9]; r4=LIFT_LEAF_RSP; \
    x0l=c->x[0].u[0]; x1l=c->x[1].u[0]; x2l=c->x[2].u[0]; x3l=c->x[3].u[0]
#define LIFT_PRECALL \
    c->r[1]=r1; c->r[2]=r2; c->r[8]=r8; c->r[9]=r9; c->r[4]=r4; \
    c->x[0].u[0]=x0l; c->x[1].u[0]=x1l; c->x[2].u[0]=x2l; c->x[3].u[0]=x3l
#define LIFT_CALL(f) do { LIFT_PRECALL; f(c); r0=c->r[0]; x0l=c->x[0].u[0]; } while (0)
#define LIFT_TAILCALL(f) do { LIFT_PRECALL; f(c); return; } while (0)
#define LIFT_FORWARD(f) do { f(c); return; } while (0)
#define LIFT_CALL_DISPATCH(t) do { LIFT_PRECALL; lift_dispatch(c, (t)); r0=c->r[0]; x0l=c->x[0].u[0]; } while (0)
#define LIFT_TAILCALL_DISPATCH(t) do { LIFT_PRECALL; lift_dispatch(c, (t)); return; } while (0)
#define LIFT_RETURN do { c->r[0]=r0; c->x[0].u[0]=x0l; return; } while (0)
#define LIFT_TRAP_RETURN return
#endif
#define LIFT_LF(T, a) (*(T*)(lf_ + LIFT_LF_OFF((uint32_t)((a) - LIFT_LEAF_LO32))))
#ifndef LIFT_LF_OFF
#define LIFT_LF_OFF(o) (o)          /* the CUDA build clamps (a NaN-driven index of a diverging aircraft must stay in bounds) */
#endif
#define LDF8(a) ((uint64_t)LIFT_LF(uint8_t, a))
#define LDF16(a) ((uint64_t)LIFT_LF(uint16_t, a))
#define LDF32(a) ((uint64_t)LIFT_LF(uint32_t, a))
#define LDF64(a) LIFT_LF(uint64_t, a)
#define LDFD(a) LIFT_LF(double, a)
#define STF8(a, v) (LIFT_LF(uint8_t, a) = (uint8_t)(v))
#define STF16(a, v) (LIFT_LF(uint16_t, a) = (uint16_t)(v))
#define STF32(a, v) (LIFT_LF(uint32_t, a) = (uint32_t)(v))
#define STF64(a, v) (LIFT_LF(uint64_t, a) = (uint64_t)(v))
"""

POSTLUDE = "\n" + "".join(f"#undef {m}\n" for m in (
    "ZF SF CF OF PF LIFT_LOCALS LIFT_LEAF_LO32 LIFT_LEAF_RSP LIFT_LEAF_DECL LIFT_RET LIFT_PARAMS LIFT_ARGS LIFT_ENTER LIFT_ENTER_LEAF "
    "LIFT_PRECALL LIFT_CALL LIFT_TAILCALL LIFT_FORWARD LIFT_CALL_DISPATCH LIFT_TAILCALL_DISPATCH LIFT_RETURN LIFT_TRAP_RETURN LIFT_LF "
    "LDF8 LDF16 LDF32 LDF64 LDFD STF8 STF16 STF32 STF64").split())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="extended_input")
    ap.add_argument("--out", required=True)
    ap.add_argument("--no-fold", action="store_true", help="single pass: do not fold loads from the initialised image")
    args = ap.parse_args()
    path = os.path.join(REF_DIR, args.variant, "_citation.cp39-win_amd64.pyd")
    pe = PE(path)
    ins = disassemble(path)
    L = Lifter(pe, ins)
    L.discover([BASE + v for v in ENTRY.values()])
    os.makedirs(args.out, exist_ok=True)
    # pass 1 -> run the translated initialize() on the host -> pass 2 folds loads from the part of the image that is constant
    # from then on (everything outside the regions step() writes)
    if not args.no_fold:
        import tempfile

        here = os.path.dirname(os.path.abspath(__file__))
        with tempfile.TemporaryDirectory() as tmp:
            inc, img, exe = os.path.join(tmp, "init.inc"), os.path.join(tmp, "image.bin"), os.path.join(tmp, "init_host")
            with open(inc, "w") as f:
                f.write(PRELUDE + L.emit_all([BASE + ENTRY["initialize"]]) + POSTLUDE)
            with open(img, "wb") as f:
                f.write(bytes(pe.img))
            subprocess.run(["gcc", "-O1", "-ffp-contract=off", "-w", f"-DLIFT_GENERATED_INC=\"{inc}\"", "-o", exe,
                            os.path.join(here, "lift_init_host.c"), "-lm"], check=True)
            L.frozen = subprocess.run([exe, img], check=True, capture_output=True).stdout
            assert len(L.frozen) == 0x40000
    # three files: everything (CPU library), what step() reaches, what initialize() / terminate() reach (the CUDA build
    # compiles the last two under different memory models)
    for tag, roots in (("code", None), ("code_step", [BASE + ENTRY["step"]]), ("code_init", [BASE + ENTRY["initialize"], BASE + ENTRY["terminate"]])):
        with open(os.path.join(args.out, f"citation_{args.variant}_{tag}.inc"), "w") as f:
            f.write(PRELUDE)
            f.write(L.emit_all(roots))
            f.write(POSTLUDE)
    with open(os.path.join(args.out, f"citation_{args.variant}_image.bin"), "wb") as f:
        f.write(bytes(pe.img))
    # the same bytes as a C initialiser list (up to the end of .data: the unwind / resource / relocation sections are not used)
    end = max(va + vsz for n, va, vsz in pe.sections if n in (".text", ".rdata", ".data"))
    with open(os.path.join(args.out, f"citation_{args.variant}_image.inc"), "w") as f:
        blob = bytes(pe.img[:end])
        for o in range(0, len(blob), 64):
            f.write(",".join(str(v) for v in blob[o:o + 64]) + ",\n")
    n_ins = sum(len(v) for v in L.funcs.values())
    print(f"{args.variant}: {len(L.funcs)} functions, {n_ins} instructions translated, image {pe.image_size} bytes, "
          f"{getattr(L, 'n_spine', 0)} convergent call sites in step(), {getattr(L, 'n_clones', 0)} specialised S-function copies, "
          f"{getattr(L, 'n_leaf', 0)} emitted with a private frame")
    for e, u in sorted(L.unknown.items()):
        print(f"  f_{e:x}: untranslated:", [(hex(x[0]), x[1]) if isinstance(x, tuple) else hex(x) for x in u[:6]])
    return 0


if __name__ == "__main__":
    sys.exit(main())
