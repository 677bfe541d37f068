"""A small interpreter for the three hot-path compute shaders the reference ships as DXBC (SM 5.0) blobs.

TEST INFRASTRUCTURE.  The reference (StarsX/FluidX12) cannot be built or run outside Windows + D3D12, and it ships no
tests or golden vectors.  What it does ship is the fxc-compiled bytecode of its shaders (``Bin/CSAdvect.cso``,
``Bin/CSProject3D.cso``, ``Bin/CSProject2D.cso``).  This module decodes the SHEX chunk of such a blob and EXECUTES it,
vectorised over all threads of a dispatch with numpy, so that the golden vectors in this directory
(``make_dxbc_golden.py``) are outputs of the reference's own compiled code rather than of a hand transliteration:
operation order, operand swizzles, folded constants and control flow all come from the bytecode.

What the bytecode does not contain, and is therefore restated here (decisions in App. D):
  * ``sample_l`` — D3D's LINEAR sampler: t = fma(coord, W, -0.5), taps floor(t) and floor(t)+1 under MIRROR (or CLAMP)
    addressing, fp32 weights, lerps in x, then y, then z as fma(f, b - a, a);
  * ``ld`` / ``ld_uav_typed`` / ``store_uav_typed`` format conversion — fp16 -> fp32 exact, fp32 -> fp16 round to
    nearest even, R32_FLOAT reads return (v, 0, 0, 1);
  * ``exp`` (2^x) — libm's exp2f;
  * thread scheduling of the unsynchronised relaxation loop over the ``globallycoherent`` pressure UAV — executed in
    lock-step: every active thread runs iteration k on the values iteration k-1 left (stores become visible at the
    ``sync``), a thread that breaks keeps its value, and the code after the loop runs when every thread has left it.
    This is the synchronous reading of SURVEY.md App. A.3; a real GPU interleaves the threads arbitrarily.

Token format: d3d11TokenizedProgramFormat.hpp (opcode / operand token layout, restated from memory of that header;
the disassembly this module prints is checked against SURVEY.md App. E by tests/test_dxbc_golden.py).
"""
from __future__ import annotations

import ctypes
import ctypes.util
import struct
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

F32, U32, I32 = np.float32, np.uint32, np.int32

OPCODES = {0: "add", 1: "and", 2: "break", 3: "breakc", 6: "case", 10: "default", 14: "div", 16: "dp3", 17: "dp4", 18: "else",
           23: "endswitch", 41: "ishl", 43: "itof", 76: "switch", 105: "dcl_indexable_temp", 21: "endif", 22: "endloop", 25: "exp",
           60: "or", 68: "rsq", 162: "dcl_resource_structured", 167: "ld_structured",
           29: "ge", 30: "iadd", 31: "if", 45: "ld", 48: "loop", 49: "lt", 50: "mad", 51: "min", 52: "max", 54: "mov",
           55: "movc", 56: "mul", 61: "resinfo", 62: "ret", 72: "sample_l", 80: "uge", 83: "umax", 84: "umin", 86: "utof",
           88: "dcl_resource", 89: "dcl_constant_buffer", 90: "dcl_sampler", 95: "dcl_input", 104: "dcl_temps",
           106: "dcl_global_flags", 155: "dcl_thread_group", 156: "dcl_uav_typed", 163: "ld_uav_typed",
           164: "store_uav_typed", 190: "sync"}
OPERAND_TYPES = {0: "r", 3: "x", 4: "l", 6: "s", 7: "t", 8: "cb", 30: "u", 32: "vThreadID"}
COMP = "xyzw"


@dataclass
class Operand:
    kind: str                    # r, l, s, t, cb, u, vThreadID
    index: List[int] = field(default_factory=list)
    ncomp: int = 4               # 0, 1 or 4
    mode: str = "mask"           # mask | swizzle | select1
    mask: int = 0xF
    swizzle: List[int] = field(default_factory=lambda: [0, 1, 2, 3])
    neg: bool = False
    abs: bool = False
    imm: Optional[List[int]] = None  # raw 32-bit words of an immediate

    def text(self) -> str:
        if self.kind == "l":
            vals = ", ".join("0x%08x" % v for v in self.imm)
            return "l(%s)" % vals
        s = self.kind
        if self.kind == "cb":
            s += "[%d][%d]" % (self.index[0], self.index[1])
        elif self.kind == "x":
            s += "%d[%d]" % (self.index[0], self.index[1])
        elif self.kind != "vThreadID":
            s += str(self.index[0])
        if self.ncomp == 4:
            if self.mode == "mask":
                s += "." + "".join(COMP[i] for i in range(4) if self.mask >> i & 1)
            elif self.mode == "swizzle":
                s += "." + "".join(COMP[i] for i in self.swizzle)
            else:
                s += "." + COMP[self.swizzle[0]]
        if self.abs:
            s = "|" + s + "|"
        if self.neg:
            s = "-" + s
        return s


@dataclass
class Instr:
    op: str
    sat: bool = False
    test_nz: bool = False
    resinfo_type: int = 0
    operands: List[Operand] = field(default_factory=list)
    raw0: int = 0
    offset: tuple = (0, 0, 0)    # aoffimmi of a sample instruction

    def text(self) -> str:
        name = self.op + ("_sat" if self.sat else "")
        if self.offset != (0, 0, 0):
            name += "_aoffimmi(%d,%d,%d)" % self.offset
        if self.op in ("if", "breakc"):
            name += "_nz" if self.test_nz else "_z"
        return (name + " " + ", ".join(o.text() for o in self.operands)).strip()


def shex_tokens(blob: bytes):
    assert blob[:4] == b"DXBC"
    n_chunks = struct.unpack_from("<I", blob, 28)[0]
    for off in struct.unpack_from("<%dI" % n_chunks, blob, 32):
        if blob[off:off + 4] in (b"SHEX", b"SHDR"):
            size = struct.unpack_from("<I", blob, off + 4)[0]
            return list(struct.unpack_from("<%dI" % (size // 4), blob, off + 8))
    raise ValueError("no SHEX chunk")


def _decode_operand(tok, pos):
    t0 = tok[pos]
    pos += 1
    o = Operand(kind=OPERAND_TYPES.get(t0 >> 12 & 0xFF, "?%d" % (t0 >> 12 & 0xFF)))
    nc = t0 & 3
    o.ncomp = {0: 0, 1: 1, 2: 4}[nc]
    if nc == 2:
        sel = t0 >> 2 & 3
        if sel == 0:
            o.mode, o.mask = "mask", t0 >> 4 & 0xF
        elif sel == 1:
            o.mode, o.swizzle = "swizzle", [t0 >> (4 + 2 * i) & 3 for i in range(4)]
        else:
            o.mode, o.swizzle = "select1", [t0 >> 4 & 3] * 4
    if t0 >> 31:
        ext = tok[pos]
        pos += 1
        assert ext & 0x3F == 1 and not ext >> 31, "only the operand-modifier extension is expected"
        mod = ext >> 6 & 0xFF
        o.neg, o.abs = mod in (1, 3), mod in (2, 3)
    dims = t0 >> 20 & 3
    for d in range(dims):
        rep = t0 >> (22 + 3 * d) & 7
        assert rep == 0, "only immediate32 operand indices are expected"
        o.index.append(tok[pos])
        pos += 1
    if o.kind == "l":
        n = 4 if o.ncomp == 4 else 1
        o.imm = list(tok[pos:pos + n])
        pos += n
    return o, pos


def decode(blob: bytes):
    """Returns (declarations, instructions) of the blob's SHEX chunk."""
    tok = shex_tokens(blob)
    assert tok[0] >> 16 == 5 and tok[0] & 0xFF == 0x50, "compute shader 5.0 expected"
    assert tok[1] == len(tok)
    pos, decls, code = 2, {"temps": 0, "group": None, "uav": {}, "srv": [], "glc": []}, []
    while pos < len(tok):
        t0 = tok[pos]
        opc, length = t0 & 0x7FF, t0 >> 24 & 0x7F
        assert length > 0 and opc in OPCODES, "unexpected opcode %d at token %d" % (opc, pos)
        name = OPCODES[opc]
        p = pos + 1
        ext = t0 >> 31
        offset = (0, 0, 0)
        while ext:  # extended opcode tokens: sample offsets are kept; resource dimension / return type are not needed
            e = tok[p]
            if e & 0x3F == 1:  # D3D10_SB_EXTENDED_OPCODE_SAMPLE_CONTROLS: three signed 4-bit texel offsets
                offset = tuple(((e >> sh & 0xF) ^ 8) - 8 for sh in (9, 13, 17))
            ext = e >> 31
            p += 1
        end = pos + length
        if name.startswith("dcl_"):
            if name == "dcl_temps":
                decls["temps"] = tok[p]
            elif name == "dcl_thread_group":
                decls["group"] = tuple(tok[p:p + 3])
            elif name == "dcl_uav_typed":
                o, _ = _decode_operand(tok, p)
                decls["uav"][o.index[0]] = {"glc": bool(t0 >> 16 & 1)}
            elif name in ("dcl_resource", "dcl_resource_structured"):
                o, _ = _decode_operand(tok, p)
                decls["srv"].append(o.index[0])
            pos = end
            continue
        ins = Instr(op=name, sat=bool(t0 >> 13 & 1), test_nz=bool(t0 >> 18 & 1), resinfo_type=t0 >> 11 & 3, raw0=t0,
                    offset=offset)
        while p < end:
            o, p = _decode_operand(tok, p)
            ins.operands.append(o)
        assert p == end
        code.append(ins)
        pos = end
    return decls, code


def disassemble(blob: bytes) -> List[str]:
    return [i.text() for i in decode(blob)[1]]


# ---- arithmetic helpers ---------------------------------------------------------------------------------------------
def fma32(a, b, c):
    """Correctly rounded fp32 fused multiply-add: the product is exact in fp64; the fp64 sum is fixed up with its
    rounding error when it lands exactly half-way between two fp32 neighbours (the double-rounding case)."""
    a, b, c = (np.asarray(v, F32).astype(np.float64) for v in (a, b, c))
    with np.errstate(invalid="ignore", over="ignore"):
        p = a * b
        t = p + c
        bp = t - p
        e = (p - (t - bp)) + (c - bp)  # TwoSum: t + e == p + c exactly
        tie = (t.view(np.uint64) & np.uint64((1 << 29) - 1)) == np.uint64(1 << 28)
        fix = tie & (e != 0) & np.isfinite(t)
        t = np.where(fix, np.nextafter(t, np.where(e > 0, np.inf, -np.inf)), t)
        return t.astype(F32)


_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.exp2f.restype = ctypes.c_float
_libm.exp2f.argtypes = [ctypes.c_float]


def exp2f(x):
    flat = np.asarray(x, F32).reshape(-1)
    return np.array([_libm.exp2f(float(v)) for v in flat], F32).reshape(np.shape(x))


def mirror_tap(i, w):
    m = np.mod(i, 2 * w)
    return np.where(m < w, m, 2 * w - 1 - m)


def sample_linear(tex_f32, coord, clamp=False, offset=(0, 0, 0)):
    """tex_f32: [nz, ny, nx, 4] fp32 texel values; coord: [N, 3] normalised (x, y, z).  Returns [N, 4] fp32.
    offset: the instruction's integer texel offsets (aoffimmi), added to both taps before the addressing mode."""
    nz, ny, nx, _ = tex_f32.shape
    idx, frac = [], []
    for axis, w in enumerate((nx, ny, nz)):
        t = fma32(coord[:, axis], F32(w), F32(-0.5))
        i0 = np.floor(t)
        frac.append((t - i0).astype(F32))
        i0 = i0.astype(np.int64) + int(offset[axis])
        tap = (lambda i, w=w: np.clip(i, 0, w - 1)) if clamp else (lambda i, w=w: mirror_tap(i, w))
        idx.append((tap(i0), tap(i0 + 1)))
    (x0, x1), (y0, y1), (z0, z1) = idx
    fx, fy, fz = (f[:, None] for f in frac)

    def lerp(a, b, f):
        return fma32(f, (b - a).astype(F32), a)

    x00 = lerp(tex_f32[z0, y0, x0], tex_f32[z0, y0, x1], fx)
    x10 = lerp(tex_f32[z0, y1, x0], tex_f32[z0, y1, x1], fx)
    x01 = lerp(tex_f32[z1, y0, x0], tex_f32[z1, y0, x1], fx)
    x11 = lerp(tex_f32[z1, y1, x0], tex_f32[z1, y1, x1], fx)
    return lerp(lerp(x00, x10, fy), lerp(x01, x11, fy), fz)


def pack_r11g11b10(rgb):
    """fp32 [N, 3] -> DXGI_FORMAT_R11G11B10_FLOAT words (R in bits 0-10, G 11-21, B 22-31; 5-bit exponent of bias 15,
    6 / 6 / 5 mantissa bits, no sign).  Restated conversion (the blob only holds `store_uav_typed`): negative values and
    -0 become 0, NaN becomes the all-ones NaN, values at or above 2^16 become the largest finite value's successor INF
    only for +INF itself (finite overflow clamps to the largest finite value), everything else is TRUNCATED toward
    zero, denormals included."""
    v = np.ascontiguousarray(rgb, F32).view(U32).astype(np.uint64)
    out = np.zeros(v.shape[0], np.uint64)
    for k, (mbits, shift) in enumerate(((6, 0), (6, 11), (5, 22))):
        w = v[:, k]
        sign, e, m = w >> np.uint64(31), (w >> np.uint64(23)) & np.uint64(0xFF), w & np.uint64(0x7FFFFF)
        drop = np.uint64(23 - mbits)
        maxfin = np.uint64((30 << mbits) | ((1 << mbits) - 1))
        normal = ((e - np.uint64(112)) << np.uint64(mbits)) | (m >> drop)      # valid for 113 <= e <= 142
        sh = np.minimum(np.uint64(113) - np.minimum(e, np.uint64(113)), np.uint64(24))
        den = ((m | np.uint64(0x800000)) >> sh) >> drop                          # e <= 112: denormal or zero
        r = np.where(e >= np.uint64(113), np.minimum(normal, maxfin), den)
        r = np.where(e >= np.uint64(143), maxfin, r)
        r = np.where((e == np.uint64(0xFF)) & (m == 0), np.uint64(31 << mbits), r)              # +INF
        r = np.where(sign != 0, np.uint64(0), r)
        r = np.where((e == np.uint64(0xFF)) & (m != 0), np.uint64((31 << mbits) | ((1 << mbits) - 1)), r)  # NaN
        out |= r << np.uint64(shift)
    return out.astype(U32)


def unpack_r11g11b10(words):
    """DXGI_FORMAT_R11G11B10_FLOAT words -> fp32 [..., 3] (exact: every such value is an fp32)."""
    w = np.asarray(words, U32)
    out = np.zeros(w.shape + (3,), F32)
    for k, (mb, sh) in enumerate(((6, 0), (6, 11), (5, 22))):
        f = (w >> U32(sh)) & U32((1 << (mb + 5)) - 1)
        e, m = (f >> U32(mb)).astype(np.int64), (f & U32((1 << mb) - 1)).astype(np.float64)
        v = np.where(e == 0, m * 2.0 ** (-14 - mb), (1 + m / (1 << mb)) * 2.0 ** (e - 15.0))
        v = np.where(e == 31, np.where(m == 0, np.inf, np.nan), v)
        out[..., k] = v.astype(F32)
    return out


def float_to_unorm8(v):
    """fp32 -> UNORM8 as D3D converts on a typed UAV store (restated): NaN -> 0, clamp to [0, 1], scale by 255, add
    0.5, truncate."""
    v = np.asarray(v, F32)
    c = np.where(np.isnan(v), F32(0), np.clip(v, F32(0), F32(1))).astype(F32)
    return ((c * F32(255.0)).astype(F32) + F32(0.5)).astype(F32).astype(np.uint8)


# ---- the machine ------------------------------------------------------------------------------------------------------
class Texture:
    """A 3D texture / typed UAV.  fmt 'rgba16f': data [nz, ny, nx, 4] float16; 'r32f': data [nz, ny, nx] float32;
    'r11g11b10f': data [nz, ny, nx] uint32; 'rgba8unorm': data [slices, ny, nx, 4] uint8 (a Texture2DArray UAV)."""

    def __init__(self, data, fmt):
        self.data, self.fmt = data, fmt

    @property
    def dims(self):
        return self.data.shape[2], self.data.shape[1], self.data.shape[0]

    def texels_f32(self):
        if self.fmt == "rgba16f":
            return self.data.astype(F32)
        if self.fmt == "r11g11b10f":
            out = np.ones(self.data.shape + (4,), F32)
            out[..., :3] = unpack_r11g11b10(self.data)
            return out
        out = np.zeros(self.data.shape + (4,), F32)
        out[..., 0] = self.data
        out[..., 3] = 1.0
        return out


class Machine:
    def __init__(self, blob: bytes, grid, cb0, srv, uav, clamp=False, threads=None):
        """grid = (nx, ny, nz): one thread per voxel.  cb0: 4 raw uint32 words of cb[0][0].  srv / uav: dicts slot ->
        Texture (UAV contents are modified in place).  threads: optional [n, 3] array of the (x, y, z) thread ids to run
        (default: the whole dispatch) — running thread groups one after the other models a GPU that does NOT execute
        the relaxation loop in lock-step (tools/dxbc_schedule_sensitivity.py)."""
        self.decls, self.code = decode(blob)
        nx, ny, nz = grid
        gx, gy, gz = self.decls["group"]
        assert nx % gx == 0 and ny % gy == 0 and nz % gz == 0, "fixtures use grids the dispatch covers exactly"
        z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
        ids = np.stack([x.reshape(-1), y.reshape(-1), z.reshape(-1)], 1) if threads is None else np.asarray(threads)
        self.n = len(ids)
        self.tid = np.concatenate([ids, np.zeros((self.n, 1), ids.dtype)], 1).astype(U32)
        self.r = np.zeros((self.decls["temps"], self.n, 4), U32)
        # cb0: the 4 words of cb[0][0], or {slot: [rows, 4] uint32} when a shader reads several constant buffers
        self.cb = {k: np.asarray(v, U32).reshape(-1, 4) for k, v in cb0.items()} if isinstance(cb0, dict) else \
            {0: np.asarray(cb0, U32).reshape(1, 4)}
        self.srv, self.uav, self.clamp = srv, uav, clamp
        self.retired = np.zeros(self.n, bool)  # threads that executed `ret` inside control flow
        self.x = {}             # indexable temp arrays x#: [element, thread, component]
        self.iterations = 0     # trips of the relaxation loop in which at least one thread was inside
        self.active_entering = []  # threads inside the loop at the start of each trip

    # -- operands
    def read(self, o: Operand) -> np.ndarray:
        if o.kind == "l":
            v = np.asarray(o.imm if len(o.imm) == 4 else o.imm * 4, U32)
            v = np.broadcast_to(v, (self.n, 4))
        else:
            if o.kind == "r":
                base = self.r[o.index[0]]
            elif o.kind == "x":  # indexable temp array x#[i], immediate index
                base = self.x.setdefault(o.index[0], np.zeros((32, self.n, 4), U32))[o.index[1]]
            elif o.kind == "vThreadID":
                base = self.tid
            elif o.kind == "cb":
                base = np.broadcast_to(self.cb[o.index[0]][o.index[1]], (self.n, 4))
            else:
                raise ValueError(o.kind)
            sw = o.swizzle if o.mode in ("swizzle", "select1") else [0, 1, 2, 3]
            v = base[:, sw]
        if o.abs:
            v = v & U32(0x7FFFFFFF)
        if o.neg:
            v = v ^ U32(0x80000000)
        return np.ascontiguousarray(v, U32)

    def read_int(self, o: Operand) -> np.ndarray:
        """Operand of an integer instruction: the `-` modifier is two's-complement negation there."""
        if not o.neg:
            return self.read(o)
        plain = Operand(kind=o.kind, index=o.index, ncomp=o.ncomp, mode=o.mode, mask=o.mask, swizzle=o.swizzle, imm=o.imm)
        return (~self.read(plain) + U32(1)).astype(U32)

    def write(self, o: Operand, value_u32, mask_threads, sat=False):
        assert o.kind in ("r", "x") and o.mode == "mask"
        if sat:
            f = value_u32.view(F32)
            value_u32 = np.where(np.isnan(f), F32(0), np.clip(f, F32(0), F32(1))).astype(F32).view(U32)
        reg = self.r[o.index[0]] if o.kind == "r" else \
            self.x.setdefault(o.index[0], np.zeros((32, self.n, 4), U32))[o.index[1]]
        for k in range(4):
            if o.mask >> k & 1:
                col = reg[:, k]
                col[mask_threads] = value_u32[mask_threads, k]

    # -- run
    def run(self):
        f = lambda a: a.view(F32)  # noqa: E731
        u = lambda a: np.ascontiguousarray(a, F32).view(U32)  # noqa: E731
        true_, false_ = U32(0xFFFFFFFF), U32(0)
        M = np.ones(self.n, bool)
        stack = []       # entries: ("if", saved mask) or ("loop", start pc, entry mask, broken)
        pc = 0
        pending = {}     # UAV slot -> array that receives stores until the next sync
        while pc < len(self.code):
            ins = self.code[pc]
            op, ops = ins.op, ins.operands
            pc += 1
            with np.errstate(all="ignore"):
                if op == "ret":
                    if not stack:
                        break
                    self.retired |= M
                    M = M & ~self.retired
                elif op == "switch":
                    # frame: kind, mask on entry, selector, threads that matched a case so far, threads that left
                    stack.append(["switch", M.copy(), self.read(ops[0])[:, 0].copy(), np.zeros(self.n, bool),
                                  np.zeros(self.n, bool)])
                    M = np.zeros(self.n, bool)
                elif op in ("case", "default"):
                    fr = stack[-1]
                    assert fr[0] == "switch"
                    if op == "case":
                        hit = fr[1] & ~fr[4] & (fr[2] == U32(ops[0].imm[0]))
                    else:
                        hit = fr[1] & ~fr[4] & ~fr[3]
                    fr[3] |= hit
                    M = (M | hit) & ~self.retired  # threads falling through from the previous case keep running
                elif op == "endswitch":
                    fr = stack.pop()
                    assert fr[0] == "switch"
                    M = fr[1] & ~self.retired
                elif op == "if":
                    c = self.read(ops[0])[:, 0] != 0
                    M0 = M.copy()
                    M = M & (c if ins.test_nz else ~c)
                    stack.append(("if", M0, M.copy()))
                elif op == "else":
                    kind, saved, taken = stack[-1]
                    assert kind == "if"
                    broken = np.zeros(self.n, bool)
                    for fr in stack:
                        if fr[0] == "loop":
                            broken = broken | fr[3]
                        elif fr[0] == "switch":
                            broken = broken | fr[4]
                    M = saved & ~taken & ~broken & ~self.retired
                elif op == "endif":
                    kind, saved, _ = stack.pop()
                    assert kind == "if"
                    broken = np.zeros(self.n, bool)
                    for fr in stack:
                        if fr[0] == "loop":
                            broken = broken | fr[3]
                        elif fr[0] == "switch":
                            broken = broken | fr[4]
                    M = saved & ~broken & ~self.retired
                elif op == "loop":
                    stack.append(["loop", pc, M.copy(), np.zeros(self.n, bool)])
                    self.active_entering.append(int(M.sum()))
                elif op in ("breakc", "break"):
                    fr = [s for s in stack if s[0] in ("loop", "switch")][-1]
                    c = np.ones(self.n, bool) if op == "break" else (self.read(ops[0])[:, 0] != 0) == ins.test_nz
                    fr[4 if fr[0] == "switch" else 3] |= M & c
                    M = M & ~c
                elif op == "endloop":
                    fr = stack[-1]
                    assert fr[0] == "loop"
                    for slot, arr in pending.items():  # a trip without a sync would still publish here
                        self.uav[slot].data[...] = arr
                    pending = {}
                    if M.any():
                        self.iterations += 1
                        self.active_entering.append(int(M.sum()))
                        pc = fr[1]
                    else:
                        stack.pop()
                        M = fr[2] & ~self.retired
                elif op == "sync":
                    for slot, arr in pending.items():
                        self.uav[slot].data[...] = arr
                    pending = {}
                elif op == "mov":
                    self.write(ops[0], self.read(ops[1]), M, ins.sat)
                elif op == "movc":
                    c, a, b = (self.read(o) for o in ops[1:4])
                    self.write(ops[0], np.where(c != 0, a, b), M, ins.sat)
                elif op in ("add", "mul", "div", "min", "max"):
                    a, b = f(self.read(ops[1])), f(self.read(ops[2]))
                    r = {"add": lambda: a + b, "mul": lambda: a * b, "div": lambda: a / b,
                         "min": lambda: np.fmin(a, b), "max": lambda: np.fmax(a, b)}[op]()
                    self.write(ops[0], u(r), M, ins.sat)
                elif op == "mad":
                    a, b, c = (f(self.read(o)) for o in ops[1:4])
                    self.write(ops[0], u(fma32(a, b, c)), M, ins.sat)
                elif op == "dp3":
                    a, b = f(self.read(ops[1])), f(self.read(ops[2]))
                    d = ((a[:, 0] * b[:, 0]).astype(F32) + (a[:, 1] * b[:, 1]).astype(F32)).astype(F32)
                    d = (d + (a[:, 2] * b[:, 2]).astype(F32)).astype(F32)
                    self.write(ops[0], u(np.repeat(d[:, None], 4, 1)), M, ins.sat)
                elif op == "exp":
                    self.write(ops[0], u(exp2f(f(self.read(ops[1])))), M, ins.sat)
                elif op in ("lt", "ge"):
                    a, b = f(self.read(ops[1])), f(self.read(ops[2]))
                    self.write(ops[0], np.where(a < b if op == "lt" else a >= b, true_, false_), M)
                elif op == "uge":
                    a, b = self.read(ops[1]), self.read(ops[2])
                    self.write(ops[0], np.where(a >= b, true_, false_), M)
                elif op in ("or", "and"):
                    a, b = self.read(ops[1]), self.read(ops[2])
                    self.write(ops[0], (a | b) if op == "or" else (a & b), M)
                elif op == "rsq":  # restated: 1 / sqrt(x), both correctly rounded (real hardware approximates)
                    a = f(self.read(ops[1]))
                    self.write(ops[0], u((F32(1.0) / np.sqrt(a).astype(F32)).astype(F32)), M, ins.sat)
                elif op == "ld_structured":
                    buf = self.srv[ops[3].index[0]]  # [elements, words] uint32
                    e = self.read(ops[1])[:, 0].astype(np.int64)
                    w0 = self.read(ops[2])[:, 0].astype(np.int64) // 4
                    sw = ops[3].swizzle if ops[3].mode != "mask" else [0, 1, 2, 3]
                    inside = e < buf.shape[0]
                    val = np.zeros((self.n, 4), U32)
                    for k in range(4):
                        w = w0 + sw[k]
                        ok = inside & (w < buf.shape[1])
                        val[ok, k] = buf[e[ok], w[ok]]
                    self.write(ops[0], val, M)
                elif op in ("umax", "umin"):
                    a, b = self.read(ops[1]), self.read(ops[2])
                    self.write(ops[0], np.maximum(a, b) if op == "umax" else np.minimum(a, b), M)
                elif op == "dp4":
                    a, b = f(self.read(ops[1])), f(self.read(ops[2]))
                    d = ((a[:, 0] * b[:, 0]).astype(F32) + (a[:, 1] * b[:, 1]).astype(F32)).astype(F32)
                    d = (d + (a[:, 2] * b[:, 2]).astype(F32)).astype(F32)
                    d = (d + (a[:, 3] * b[:, 3]).astype(F32)).astype(F32)
                    self.write(ops[0], u(np.repeat(d[:, None], 4, 1)), M, ins.sat)
                elif op == "ishl":
                    a, b = self.read(ops[1]), self.read(ops[2])
                    self.write(ops[0], (a.astype(np.uint64) << (b & U32(31)).astype(np.uint64)).astype(U32), M)
                elif op == "itof":
                    self.write(ops[0], u(self.read(ops[1]).view(I32).astype(F32)), M)
                elif op == "iadd":
                    a, b = (self.read_int(o) for o in ops[1:3])
                    self.write(ops[0], (a.astype(np.uint64) + b.astype(np.uint64)).astype(U32), M)
                elif op == "utof":
                    self.write(ops[0], u(self.read(ops[1]).astype(F32)), M)
                elif op == "resinfo":
                    w, h, d = (self.uav if ops[2].kind == "u" else self.srv)[ops[2].index[0]].dims
                    dims = np.array([w, h, d, 1], np.int64)
                    raw = dims.astype(U32) if ins.resinfo_type == 2 else dims.astype(F32).view(U32)
                    assert ins.resinfo_type in (0, 2)
                    v = np.broadcast_to(raw[ops[2].swizzle if ops[2].mode != "mask" else [0, 1, 2, 3]], (self.n, 4))
                    self.write(ops[0], np.ascontiguousarray(v), M)
                elif op in ("ld", "ld_uav_typed"):
                    tex = (self.srv if op == "ld" else self.uav)[ops[2].index[0]]
                    c = self.read(ops[1]).astype(np.int64)
                    texels = tex.texels_f32()[c[:, 2], c[:, 1], c[:, 0]]
                    sw = ops[2].swizzle if ops[2].mode != "mask" else [0, 1, 2, 3]
                    self.write(ops[0], u(texels[:, sw]), M)
                elif op == "sample_l":
                    coord = f(self.read(ops[1]))[:, :3]
                    tex = self.srv[ops[2].index[0]]
                    texels = sample_linear(tex.texels_f32(), coord, self.clamp, ins.offset)
                    sw = ops[2].swizzle if ops[2].mode != "mask" else [0, 1, 2, 3]
                    self.write(ops[0], u(texels[:, sw]), M, ins.sat)
                elif op == "store_uav_typed":
                    slot = ops[0].index[0]
                    tex = self.uav[slot]
                    c = self.read(ops[1]).astype(np.int64)[M]
                    v = f(self.read(ops[2]))[M]
                    in_loop = any(s[0] == "loop" for s in stack)
                    if in_loop:  # visible to the other threads at the next sync (lock-step schedule)
                        target = pending.setdefault(slot, tex.data.copy())
                    else:
                        target = tex.data
                    if tex.fmt == "r11g11b10f":
                        target[c[:, 2], c[:, 1], c[:, 0]] = pack_r11g11b10(v[:, :3])
                    elif tex.fmt == "rgba8unorm":
                        target[c[:, 2], c[:, 1], c[:, 0]] = float_to_unorm8(v)
                    elif tex.fmt == "rgba16f":
                        target[c[:, 2], c[:, 1], c[:, 0]] = v.astype(np.float16)  # round to nearest even
                    else:
                        target[c[:, 2], c[:, 1], c[:, 0]] = v[:, 0]
                else:
                    raise NotImplementedError(op)
        return self
