"""A minimal JVM bytecode interpreter, just large enough to run the reference's own compiled classes
(web/bin/plaac.jar: plaac, hmm, disorderreport, fastareader) end to end on small inputs.

Why: the image has no JVM, so plaac.jar cannot be executed -- but its BYTECODE can be interpreted.  This file executes
`plaac.main(String[])` exactly as compiled by the reference authors (class files inside the jar, untouched) and
records every value the program hands to System.out.format()/print(): full-precision doubles and ints of every output
cell.  tests/golden/make_jar_vectors.py turns those into fixtures that pin oracle/ and the CUDA path against the
reference itself.  The only things not taken from the jar are the JDK natives it calls: java.lang.Math
(log/exp/floor/abs/min/max/sqrt -> Python's math = this box's libm, specified to <= 1 ulp like Java's), string and
stream plumbing (String, StringBuffer, BufferedReader, PrintStream, HashMap, boxing).

Scope: ints, doubles, chars, booleans, references, arrays; no threads, no exceptions thrown, no floats/longs beyond
what the four classes use.  Not a general JVM.  TEST INFRASTRUCTURE ONLY (golden-vector generator).
"""
from __future__ import annotations

import math
import struct
import zipfile


# ----------------------------------------------------------------------------------------------- class files
class Method:
    __slots__ = ("cls", "name", "desc", "flags", "code", "max_locals", "nargs", "ret", "ins", "argcats")


class JClass:
    def __init__(self, data: bytes):
        self.data = data
        self.pos = 0
        self.statics = {}
        self.methods = {}
        self.initialised = False
        self._parse()

    def u1(self):
        v = self.data[self.pos]
        self.pos += 1
        return v

    def u2(self):
        v = struct.unpack_from(">H", self.data, self.pos)[0]
        self.pos += 2
        return v

    def u4(self):
        v = struct.unpack_from(">I", self.data, self.pos)[0]
        self.pos += 4
        return v

    def _parse(self):
        assert self.u4() == 0xCAFEBABE
        self.u2()
        self.u2()
        n = self.u2()
        cp = [None] * n
        i = 1
        while i < n:
            tag = self.u1()
            if tag == 1:
                ln = self.u2()
                cp[i] = ("utf8", self.data[self.pos:self.pos + ln].decode("utf-8", "replace"))
                self.pos += ln
            elif tag == 3:
                cp[i] = ("int", struct.unpack_from(">i", self.data, self.pos)[0])
                self.pos += 4
            elif tag == 4:
                cp[i] = ("float", struct.unpack_from(">f", self.data, self.pos)[0])
                self.pos += 4
            elif tag == 5:
                cp[i] = ("long", struct.unpack_from(">q", self.data, self.pos)[0])
                self.pos += 8
                i += 1
            elif tag == 6:
                cp[i] = ("double", struct.unpack_from(">d", self.data, self.pos)[0])
                self.pos += 8
                i += 1
            elif tag == 7:
                cp[i] = ("class", self.u2())
            elif tag == 8:
                cp[i] = ("string", self.u2())
            elif tag in (9, 10, 11):
                cp[i] = ({9: "field", 10: "method", 11: "imethod"}[tag], self.u2(), self.u2())
            elif tag == 12:
                cp[i] = ("nat", self.u2(), self.u2())
            elif tag == 15:
                cp[i] = ("mh", self.u1(), self.u2())
            elif tag == 16:
                cp[i] = ("mt", self.u2())
            elif tag == 18:
                cp[i] = ("indy", self.u2(), self.u2())
            else:
                raise ValueError(f"constant pool tag {tag}")
            i += 1
        self.cp = cp
        self.access = self.u2()
        self.name = self.cls_name(self.u2())
        sup = self.u2()
        self.super_name = self.cls_name(sup) if sup else None
        for _ in range(self.u2()):
            self.u2()
        self.fields = []
        for _ in range(self.u2()):
            flags, nm, desc = self.u2(), self.utf(self.u2()), self.utf(self.u2())
            const = None
            for _ in range(self.u2()):
                an, ln = self.utf(self.u2()), self.u4()
                if an == "ConstantValue":
                    const = self.const_value(struct.unpack_from(">H", self.data, self.pos)[0])
                self.pos += ln
            self.fields.append((flags, nm, desc))
            if flags & 0x0008:
                self.statics[nm] = const if const is not None else default_value(desc)
        for _ in range(self.u2()):
            m = Method()
            m.cls = self
            m.flags, m.name, m.desc = self.u2(), self.utf(self.u2()), self.utf(self.u2())
            m.code = None
            m.ins = None
            for _ in range(self.u2()):
                an, ln = self.utf(self.u2()), self.u4()
                if an == "Code":
                    p = self.pos
                    struct.unpack_from(">H", self.data, p)
                    m.max_locals = struct.unpack_from(">H", self.data, p + 2)[0]
                    clen = struct.unpack_from(">I", self.data, p + 4)[0]
                    m.code = self.data[p + 8:p + 8 + clen]
                self.pos += ln
            m.argcats, m.ret = parse_desc(m.desc)
            m.nargs = len(m.argcats) + (0 if m.flags & 0x0008 else 1)
            self.methods[(m.name, m.desc)] = m

    def utf(self, i):
        return self.cp[i][1]

    def cls_name(self, i):
        return self.utf(self.cp[i][1])

    def const_value(self, i):
        t = self.cp[i]
        if t[0] == "string":
            return self.utf(t[1])
        return t[1]

    def member(self, i):
        _, ci, nti = self.cp[i]
        nt = self.cp[nti]
        return self.cls_name(ci), self.utf(nt[1]), self.utf(nt[2])


def default_value(desc):
    if desc in ("D", "F"):
        return 0.0
    if desc in ("I", "Z", "B", "C", "S", "J"):
        return 0
    return None


def parse_desc(desc):
    """-> ([category of each argument: 'D' double / 'J' long / 'I' int-like / 'A' ref], return kind)"""
    assert desc[0] == "("
    i, cats = 1, []
    while desc[i] != ")":
        c = desc[i]
        if c in "DJ":
            cats.append(c)
            i += 1
        elif c in "IZBCSF":
            cats.append("I")
            i += 1
        elif c == "L":
            cats.append("A")
            i = desc.index(";", i) + 1
        elif c == "[":
            while desc[i] == "[":
                i += 1
            if desc[i] == "L":
                i = desc.index(";", i) + 1
            else:
                i += 1
            cats.append("A")
        else:
            raise ValueError(desc)
    return cats, desc[i + 1:]


# ----------------------------------------------------------------------------------------------- runtime objects
class JObject:
    __slots__ = ("cls", "f")

    def __init__(self, cls):
        self.cls = cls
        self.f = {}


class JArray:
    __slots__ = ("a", "t")

    def __init__(self, a, t):
        self.a = a
        self.t = t  # 'D','I','C','Z','B','A',...


class Boxed:
    """java.lang.Integer / Double / Character / Boolean as handed to format()."""
    __slots__ = ("kind", "v")

    def __init__(self, kind, v):
        self.kind = kind
        self.v = v


class JStringBuffer:
    def __init__(self, s=""):
        self.s = s


class JReader:
    """BufferedReader(FileReader(name)) with readLine's \\n / \\r / \\r\\n semantics."""

    def __init__(self, path):
        import re

        data = open(path, "rb").read().decode("latin-1")
        lines = re.split(r"\r\n|\n|\r", data)
        if lines and lines[-1] == "":
            lines.pop()
        self.lines = lines
        self.i = 0

    def readline(self):
        if self.i >= len(self.lines):
            return None
        self.i += 1
        return self.lines[self.i - 1]


class JTokenizer:
    def __init__(self, s):
        self.toks = s.split()
        self.i = 0


def i32(v):
    v &= 0xFFFFFFFF
    return v - 0x100000000 if v & 0x80000000 else v


def d2i(x):
    if x != x:
        return 0
    if x >= 2147483647.0:
        return 2147483647
    if x <= -2147483648.0:
        return -2147483648
    return int(x)


def jdiv(a, b):
    if b == 0.0:
        if a != a or a == 0.0:
            return float("nan")
        neg = (math.copysign(1.0, a) < 0) != (math.copysign(1.0, b) < 0)
        return float("-inf") if neg else float("inf")
    return a / b


def jlog(x):
    if x != x or x < 0:
        return float("nan")
    if x == 0:
        return float("-inf")
    return math.log(x)


def jexp(x):
    try:
        return math.exp(x)
    except OverflowError:
        return float("inf")


class Captured:
    """Everything the program printed: `events` holds ('print', str) and ('format', fmt, [python values])."""

    def __init__(self):
        self.events = []


# ----------------------------------------------------------------------------------------------- the machine
class VM:
    def __init__(self, jar_path):
        z = zipfile.ZipFile(jar_path)
        self.classes = {}
        for n in z.namelist():
            if n.endswith(".class"):
                c = JClass(z.read(n))
                self.classes[c.name] = c
        self.out = Captured()
        self.steps = 0
        self.stdout_obj = JObject(None)

    # ---- class init / lookup
    def cls(self, name):
        c = self.classes[name]
        if not c.initialised:
            c.initialised = True
            m = c.methods.get(("<clinit>", "()V"))
            if m:
                self.invoke(m, [])
        return c

    def find_method(self, cname, name, desc):
        c = self.classes.get(cname)
        while c is not None:
            m = c.methods.get((name, desc))
            if m:
                return m
            c = self.classes.get(c.super_name)
        return None

    # ---- natives
    def native(self, cname, name, desc, args):
        key = f"{cname}.{name}"
        a = args
        if cname == "java/lang/Math" or cname == "java/lang/StrictMath":
            if name == "log":
                return jlog(a[0])
            if name == "exp":
                return jexp(a[0])
            if name == "floor":
                return float(math.floor(a[0])) if math.isfinite(a[0]) else a[0]
            if name == "ceil":
                return float(math.ceil(a[0])) if math.isfinite(a[0]) else a[0]
            if name == "sqrt":
                return math.sqrt(a[0]) if a[0] >= 0 else float("nan")
            if name == "abs":
                return abs(a[0])
            if name == "min":
                if desc.startswith("(DD"):
                    return a[0] if (a[0] != a[0] or a[0] < a[1] or (a[0] == a[1] and math.copysign(1, a[0]) < 0)) else a[1]
                return min(a[0], a[1])
            if name == "max":
                if desc.startswith("(DD"):
                    return a[0] if (a[0] != a[0] or a[0] > a[1] or (a[0] == a[1] and math.copysign(1, a[0]) > 0)) else a[1]
                return max(a[0], a[1])
            if name == "pow":
                return math.pow(a[0], a[1])
            if name == "round":
                return int(math.floor(a[0] + 0.5))
        if cname == "java/lang/Object" and name == "<init>":
            return None
        if cname == "java/lang/Double":
            if name == "valueOf":
                return Boxed("D", a[0]) if isinstance(a[0], float) else Boxed("D", float(a[0]))
            if name == "isInfinite":
                return 1 if math.isinf(a[-1] if not isinstance(a[-1], Boxed) else a[-1].v) else 0
            if name == "isNaN":
                return 1 if a[-1] != a[-1] else 0
            if name == "parseDouble":
                return float(a[0])
            if name == "doubleValue":
                return a[0].v
        if cname == "java/lang/Integer":
            if name == "valueOf":
                return Boxed("I", a[0])
            if name == "parseInt":
                return int(a[0])
            if name == "intValue":
                return a[0].v
        if cname == "java/lang/Character" and name == "valueOf":
            return Boxed("C", a[0])
        if cname == "java/lang/Boolean" and name == "valueOf":
            return Boxed("Z", a[0])
        if cname == "java/lang/String":
            s = a[0]
            if name == "<init>":
                # `new String(x)`: the receiver placeholder is replaced by the caller (see op new/invokespecial)
                if len(a) == 1:
                    v = ""
                elif isinstance(a[1], JArray):
                    chars = a[1].a if len(a) == 2 else a[1].a[a[2]:a[2] + a[3]]
                    v = "".join(chr(c) for c in chars)
                elif isinstance(a[1], JStringBuffer):
                    v = a[1].s
                else:
                    v = a[1]
                return ("__string_init__", v)
            if name == "length":
                return len(s)
            if name == "charAt":
                return ord(s[a[1]])
            if name == "equals":
                return 1 if isinstance(a[1], str) and s == a[1] else 0
            if name == "substring":
                return s[a[1]:] if len(a) == 2 else s[a[1]:a[2]]
            if name == "trim":
                return s.strip("".join(chr(c) for c in range(33)))
            if name == "split":
                import re

                parts = re.split(a[1], s)
                while parts and parts[-1] == "":
                    parts.pop()
                return JArray(parts, "A")
            if name == "indexOf":
                return s.find(a[1] if isinstance(a[1], str) else chr(a[1]), *(a[2:3]))
            if name == "valueOf":
                return java_tostring(a[0], desc)
            if name == "format":
                return java_format(a[0], [(chr(x.v) if x.kind == "C" else x.v) if isinstance(x, Boxed) else x for x in a[1].a])
            if name == "toCharArray":
                return JArray([ord(ch) for ch in s], "C")
            if name == "startsWith":
                return 1 if s.startswith(a[1]) else 0
        if cname in ("java/lang/StringBuffer", "java/lang/StringBuilder"):
            sb = a[0]
            if name == "<init>":
                sb.s = a[1] if len(a) > 1 and isinstance(a[1], str) else ""
                return None
            if name == "append":
                v = a[1]
                ad = desc[1:desc.index(")")]
                if ad == "C":
                    sb.s += chr(v)
                elif ad == "Z":
                    sb.s += "true" if v else "false"
                else:
                    sb.s += java_tostring(v, "(" + ad + ")")
                return sb
            if name == "toString":
                return sb.s
            if name == "length":
                return len(sb.s)
            if name == "charAt":
                return ord(sb.s[a[1]])
            if name == "deleteCharAt":
                sb.s = sb.s[:a[1]] + sb.s[a[1] + 1:]
                return sb
            if name == "setCharAt":
                sb.s = sb.s[:a[1]] + chr(a[2]) + sb.s[a[1] + 1:]
                return None
            if name == "setLength":
                sb.s = sb.s[:a[1]] + "\0" * max(0, a[1] - len(sb.s))
                return None
            if name == "insert":
                sb.s = sb.s[:a[1]] + java_tostring(a[2], "(" + desc[desc.index("I") + 1:desc.index(")")] + ")") + sb.s[a[1]:]
                return sb
            if name == "reverse":
                sb.s = sb.s[::-1]
                return sb
        if cname == "java/io/FileReader" and name == "<init>":
            a[0].f["path"] = a[1]
            return None
        if cname == "java/io/BufferedReader":
            if name == "<init>":
                a[0].f["r"] = JReader(a[1].f["path"])
                return None
            if name == "readLine":
                return a[0].f["r"].readline()
            if name == "close":
                return None
        if cname == "java/io/PrintStream":
            if name in ("println", "print"):
                s = "" if len(a) == 1 else java_tostring(a[1], desc)
                self.out.events.append(("print", s + ("\n" if name == "println" else "")))
                return None
            if name in ("format", "printf"):
                self.out.events.append(("format", a[1], [(chr(x.v) if x.kind == "C" else x.v) if isinstance(x, Boxed) else x
                                                          for x in a[2].a]))
                return a[0]
        if cname == "java/util/HashMap":
            if name == "<init>":
                a[0].f["d"] = {}
                return None
            if name == "containsKey":
                return 1 if a[1] in a[0].f["d"] else 0
            if name == "get":
                return a[0].f["d"].get(a[1])
            if name == "put":
                old = a[0].f["d"].get(a[1])
                a[0].f["d"][a[1]] = a[2]
                return old
            if name == "size":
                return len(a[0].f["d"])
        if cname == "java/util/StringTokenizer":
            if name == "<init>":
                a[0].f["t"] = JTokenizer(a[1])
                return None
            if name == "countTokens":
                t = a[0].f["t"]
                return len(t.toks) - t.i
            if name == "nextToken":
                t = a[0].f["t"]
                t.i += 1
                return t.toks[t.i - 1]
        raise NotImplementedError(f"native {key}{desc}")

    # ---- execution
    def invoke(self, m: Method, args):
        if m.code is None:
            raise NotImplementedError(f"abstract/native {m.cls.name}.{m.name}")
        if m.ins is None:
            m.ins = decode(self, m)
        lo = [None] * (m.max_locals + 1)
        # arguments into locals: doubles/longs take two slots
        k = 0
        ai = 0
        if not (m.flags & 0x0008):
            lo[0] = args[0]
            k = 1
            ai = 1
        for c in m.argcats:
            lo[k] = args[ai]
            ai += 1
            k += 2 if c in "DJ" else 1
        st = []
        ins = m.ins
        pc = 0
        n = 0
        while pc is not None:
            fn, a, b = ins[pc]
            pc = fn(self, st, lo, a, b, pc)
            n += 1
        self.steps += n
        return st.pop() if m.ret != "V" else None


def java_tostring(v, desc=""):
    if v is None:
        return "null"
    if isinstance(v, str):
        return v
    if isinstance(v, Boxed):
        v, kind = v.v, v.kind
        if kind == "C":
            return chr(v)
        if kind == "Z":
            return "true" if v else "false"
    ad = desc[1:desc.index(")")] if desc.startswith("(") else ""
    if ad == "C":
        return chr(v)
    if ad == "Z":
        return "true" if v else "false"
    if isinstance(v, float):
        return java_double_tostring(v)
    if isinstance(v, int):
        return str(v)
    if isinstance(v, JStringBuffer):
        return v.s
    return f"<{type(v).__name__}>"


def java_fixed(x, d):
    """java.util.Formatter %.<d>f: half-up on the shortest round-trip digits."""
    from decimal import ROUND_HALF_UP, Decimal

    if x != x:
        return "NaN"
    if math.isinf(x):
        return "Infinity" if x > 0 else "-Infinity"
    q = Decimal(repr(abs(float(x)))).quantize(Decimal(1).scaleb(-d), rounding=ROUND_HALF_UP)
    return ("-" if math.copysign(1.0, x) < 0 else "") + f"{q:f}"


def java_format(fmt, args):
    import re

    out, k = [], 0
    pos = 0
    for mt in re.finditer(r"%(\.(\d+))?([sdfn%])", fmt):
        out.append(fmt[pos:mt.start()])
        pos = mt.end()
        c = mt.group(3)
        if c == "n":
            out.append("\n")
        elif c == "%":
            out.append("%")
        else:
            v = args[k]
            k += 1
            if c == "f":
                out.append(java_fixed(float(v), int(mt.group(2)) if mt.group(2) else 6))
            else:
                out.append(java_tostring(v))
    out.append(fmt[pos:])
    return "".join(out)


def java_double_tostring(x):
    if x != x:
        return "NaN"
    if math.isinf(x):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0:
        return "-0.0" if math.copysign(1, x) < 0 else "0.0"
    ax = abs(x)
    r = repr(ax)
    if 1e-3 <= ax < 1e7:
        if "e" in r:
            r = f"{ax:f}".rstrip("0")
            if r.endswith("."):
                r += "0"
        s = r if "." in r else r + ".0"
    else:
        m, e = f"{ax:.16e}".split("e")
        # shortest digits
        digits = repr(ax)
        mant, ex = (digits.split("e") + ["0"])[:2] if "e" in digits else (None, None)
        if mant is None:
            from decimal import Decimal

            t = Decimal(digits).as_tuple()
            ds = "".join(map(str, t.digits)).rstrip("0") or "0"
            ex10 = len(t.digits) + t.exponent - 1
            mant = ds[0] + "." + (ds[1:] or "0")
            ex = str(ex10)
        else:
            if "." not in mant:
                mant += ".0"
        s = f"{mant}E{int(ex)}"
    return ("-" if x < 0 else "") + s


# ----------------------------------------------------------------------------------------------- instruction set
def decode(vm: VM, m: Method):
    code = m.code
    cp = m.cls.cp
    cls = m.cls
    offs = []
    raw = []
    pc = 0
    n = len(code)

    def s1(p):
        v = code[p]
        return v - 256 if v > 127 else v

    def u2(p):
        return (code[p] << 8) | code[p + 1]

    def s2(p):
        v = u2(p)
        return v - 65536 if v > 32767 else v

    def s4(p):
        return struct.unpack_from(">i", code, p)[0]

    while pc < n:
        op = code[pc]
        start = pc
        a = b = None
        if op == 16:
            a = s1(pc + 1)
            pc += 2
        elif op == 17:
            a = s2(pc + 1)
            pc += 3
        elif op == 18:
            a = code[pc + 1]
            pc += 2
        elif op in (19, 20):
            a = u2(pc + 1)
            pc += 3
        elif 21 <= op <= 25 or 54 <= op <= 58 or op == 169:
            a = code[pc + 1]
            pc += 2
        elif op == 132:
            a, b = code[pc + 1], s1(pc + 2)
            pc += 3
        elif 153 <= op <= 168 or op in (198, 199):
            a = start + s2(pc + 1)
            pc += 3
        elif op == 170:
            p = (pc + 4) & ~3
            default, low, high = s4(p), s4(p + 4), s4(p + 8)
            table = [start + s4(p + 12 + 4 * k) for k in range(high - low + 1)]
            a, b = (low, high, start + default), table
            pc = p + 12 + 4 * (high - low + 1)
        elif op == 171:
            p = (pc + 4) & ~3
            default, npairs = s4(p), s4(p + 4)
            a = start + default
            b = {s4(p + 8 + 8 * k): start + s4(p + 12 + 8 * k) for k in range(npairs)}
            pc = p + 8 + 8 * npairs
        elif 178 <= op <= 184 or op in (187, 189, 192, 193):
            a = u2(pc + 1)
            pc += 3
        elif op == 185:
            a = u2(pc + 1)
            pc += 5
        elif op == 188:
            a = code[pc + 1]
            pc += 2
        elif op == 196:
            op2 = code[pc + 1]
            if op2 == 132:
                a, b = u2(pc + 2), s2(pc + 4)
                pc += 6
            else:
                a = u2(pc + 2)
                pc += 4
            op = op2
        elif op == 197:
            a, b = u2(pc + 1), code[pc + 3]
            pc += 4
        elif op in (200, 201):
            a = start + s4(pc + 1)
            pc += 5
        else:
            pc += 1
        offs.append(start)
        raw.append((op, a, b))
    index_of = {o: i for i, o in enumerate(offs)}
    ins = []
    for (op, a, b) in raw:
        ins.append(build(vm, cls, cp, op, a, b, index_of))
    return ins


def build(vm, cls, cp, op, a, b, index_of):
    T = index_of
    # constants
    if op == 0:
        return (o_nop, None, None)
    if op == 1:
        return (o_const, None, None)
    if 2 <= op <= 8:
        return (o_const, op - 3, None)
    if op in (9, 10):
        return (o_const, op - 9, None)
    if op in (14, 15):
        return (o_const, float(op - 14), None)
    if op in (11, 12, 13):
        return (o_const, float(op - 11), None)
    if op in (16, 17):
        return (o_const, a, None)
    if op in (18, 19, 20):
        t = cp[a]
        if t[0] == "string":
            return (o_const, cls.utf(t[1]), None)
        if t[0] in ("int", "double", "long", "float"):
            return (o_const, t[1], None)
        if t[0] == "class":
            return (o_const, ("class", cls.utf(t[1])), None)
        raise NotImplementedError(t)
    # loads / stores
    if 21 <= op <= 25:
        return (o_load, a, None)
    if 26 <= op <= 45:
        return (o_load, (op - 26) % 4, None)
    if 54 <= op <= 58:
        return (o_store, a, None)
    if 59 <= op <= 78:
        return (o_store, (op - 59) % 4, None)
    if op in (46, 47, 48, 49, 50, 51, 52, 53):
        return (o_aload, None, None)
    if op in (79, 83, 84):
        return (o_astore_any, None, None)
    if op == 85:
        return (o_castore, None, None)
    if op in (80, 81, 82, 86):
        return (o_astore_any, None, None)
    if op == 87:
        return (o_pop, None, None)
    if op == 88:
        return (o_pop2, None, None)
    if op == 89:
        return (o_dup, None, None)
    if op == 90:
        return (o_dup_x1, None, None)
    if op == 91:
        return (o_dup_x2, None, None)
    if op == 92:
        return (o_dup2, None, None)
    if op == 93:
        return (o_dup2_x1, None, None)
    if op == 94:
        return (o_dup2_x2, None, None)
    if op == 95:
        return (o_swap, None, None)
    arith = {96: o_iadd, 99: o_dadd, 100: o_isub, 103: o_dsub, 104: o_imul, 107: o_dmul, 108: o_idiv, 111: o_ddiv,
             112: o_irem, 115: o_drem, 116: o_ineg, 119: o_dneg, 120: o_ishl, 122: o_ishr, 124: o_iushr, 126: o_iand,
             128: o_ior, 130: o_ixor, 135: o_i2d, 142: o_d2i, 145: o_i2b, 146: o_i2c, 147: o_i2s, 151: o_dcmpl,
             152: o_dcmpg, 133: o_nop, 136: o_l2i, 138: o_i2d, 143: o_d2l, 141: o_nop, 144: o_d2f, 134: o_i2d,
             139: o_d2i, 97: o_ladd, 101: o_lsub, 105: o_lmul, 148: o_lcmp, 98: o_dadd, 102: o_dsub, 106: o_fmul,
             110: o_fdiv, 149: o_dcmpl, 150: o_dcmpg}
    if op in arith:
        return (arith[op], None, None)
    if op == 132:
        return (o_iinc, a, b)
    if 153 <= op <= 158:
        return (o_if, T[a], op - 153)
    if 159 <= op <= 164:
        return (o_if_icmp, T[a], op - 159)
    if op in (165, 166):
        return (o_if_acmp, T[a], op - 165)
    if op in (167, 200):
        return (o_goto, T[a], None)
    if op == 198:
        return (o_ifnull, T[a], 1)
    if op == 199:
        return (o_ifnull, T[a], 0)
    if op == 170:
        low, high, default = a
        return (o_tableswitch, (low, high, T[default]), [T[x] for x in b])
    if op == 171:
        return (o_lookupswitch, T[a], {k: T[v] for k, v in b.items()})
    if 172 <= op <= 176:
        return (o_return_value, None, None)
    if op == 177:
        return (o_return, None, None)
    if op in (178, 179, 180, 181):
        cn, fn, fd = cls.member(a)
        return ({178: o_getstatic, 179: o_putstatic, 180: o_getfield, 181: o_putfield}[op], (cn, fn, fd), None)
    if op in (182, 183, 184, 185):
        cn, mn, md = cls.member(a)
        cats, ret = parse_desc(md)
        nargs = len(cats) + (0 if op == 184 else 1)
        return (o_invoke, (cn, mn, md, nargs, ret, op), None)
    if op == 187:
        return (o_new, cls.cls_name(a), None)
    if op == 188:
        return (o_newarray, {4: "Z", 5: "C", 6: "F", 7: "D", 8: "B", 9: "S", 10: "I", 11: "J"}[a], None)
    if op == 189:
        return (o_newarray, "A", None)
    if op == 190:
        return (o_arraylength, None, None)
    if op == 192:
        return (o_nop, None, None)
    if op == 197:
        return (o_multianewarray, cls.cls_name(a), b)
    if op == 191:
        return (o_athrow, None, None)
    if op in (194, 195):
        return (o_pop, None, None)
    raise NotImplementedError(f"opcode {op}")


def o_nop(vm, st, lo, a, b, pc):
    return pc + 1


def o_const(vm, st, lo, a, b, pc):
    st.append(a)
    return pc + 1


def o_load(vm, st, lo, a, b, pc):
    st.append(lo[a])
    return pc + 1


def o_store(vm, st, lo, a, b, pc):
    lo[a] = st.pop()
    return pc + 1


def o_aload(vm, st, lo, a, b, pc):
    i = st.pop()
    arr = st.pop()
    st.append(arr.a[i] if i >= 0 else _oob(i))
    return pc + 1


def _oob(i):
    raise IndexError(f"java.lang.ArrayIndexOutOfBoundsException: {i}")


def o_astore_any(vm, st, lo, a, b, pc):
    v = st.pop()
    i = st.pop()
    arr = st.pop()
    if i < 0:
        _oob(i)
    arr.a[i] = v
    return pc + 1


def o_castore(vm, st, lo, a, b, pc):
    v = st.pop()
    i = st.pop()
    arr = st.pop()
    if i < 0:
        _oob(i)
    arr.a[i] = v & 0xFFFF
    return pc + 1


def o_pop(vm, st, lo, a, b, pc):
    st.pop()
    return pc + 1


def _cat2(v):
    return isinstance(v, float) or isinstance(v, JLong)


class JLong(int):
    pass


def o_pop2(vm, st, lo, a, b, pc):
    v = st.pop()
    if not _cat2(v):
        st.pop()
    return pc + 1


def o_dup(vm, st, lo, a, b, pc):
    st.append(st[-1])
    return pc + 1


def o_dup_x1(vm, st, lo, a, b, pc):
    v1 = st.pop()
    v2 = st.pop()
    st += [v1, v2, v1]
    return pc + 1


def o_dup_x2(vm, st, lo, a, b, pc):
    v1 = st.pop()
    v2 = st.pop()
    if _cat2(v2):
        st += [v1, v2, v1]
    else:
        v3 = st.pop()
        st += [v1, v3, v2, v1]
    return pc + 1


def o_dup2(vm, st, lo, a, b, pc):
    v1 = st[-1]
    if _cat2(v1):
        st.append(v1)
    else:
        st += [st[-2], v1]
    return pc + 1


def o_dup2_x1(vm, st, lo, a, b, pc):
    v1 = st.pop()
    if _cat2(v1):
        v2 = st.pop()
        st += [v1, v2, v1]
    else:
        v2 = st.pop()
        v3 = st.pop()
        st += [v2, v1, v3, v2, v1]
    return pc + 1


def o_dup2_x2(vm, st, lo, a, b, pc):
    v1 = st.pop()
    if _cat2(v1):
        v2 = st.pop()
        if _cat2(v2):
            st += [v1, v2, v1]
        else:
            v3 = st.pop()
            st += [v1, v3, v2, v1]
    else:
        v2 = st.pop()
        v3 = st.pop()
        if _cat2(v3):
            st += [v2, v1, v3, v2, v1]
        else:
            v4 = st.pop()
            st += [v2, v1, v4, v3, v2, v1]
    return pc + 1


def o_swap(vm, st, lo, a, b, pc):
    st[-1], st[-2] = st[-2], st[-1]
    return pc + 1


def o_iadd(vm, st, lo, a, b, pc):
    y = st.pop()
    st[-1] = i32(st[-1] + y)
    return pc + 1


def o_isub(vm, st, lo, a, b, pc):
    y = st.pop()
    st[-1] = i32(st[-1] - y)
    return pc + 1


def o_imul(vm, st, lo, a, b, pc):
    y = st.pop()
    st[-1] = i32(st[-1] * y)
    return pc + 1


def o_idiv(vm, st, lo, a, b, pc):
    y = st.pop()
    x = st[-1]
    if y == 0:
        raise ZeroDivisionError("java.lang.ArithmeticException: / by zero")
    q = abs(x) // abs(y)
    st[-1] = i32(q if (x >= 0) == (y >= 0) else -q)
    return pc + 1


def o_irem(vm, st, lo, a, b, pc):
    y = st.pop()
    x = st[-1]
    if y == 0:
        raise ZeroDivisionError("java.lang.ArithmeticException: % by zero")
    r = abs(x) % abs(y)
    st[-1] = r if x >= 0 else -r
    return pc + 1


def o_ineg(vm, st, lo, a, b, pc):
    st[-1] = i32(-st[-1])
    return pc + 1


def o_ishl(vm, st, lo, a, b, pc):
    y = st.pop()
    st[-1] = i32(st[-1] << (y & 31))
    return pc + 1


def o_ishr(vm, st, lo, a, b, pc):
    y = st.pop()
    st[-1] = st[-1] >> (y & 31)
    return pc + 1


def o_iushr(vm, st, lo, a, b, pc):
    y = st.pop()
    st[-1] = i32((st[-1] & 0xFFFFFFFF) >> (y & 31))
    return pc + 1


def o_iand(vm, st, lo, a, b, pc):
    y = st.pop()
    st[-1] = st[-1] & y
    return pc + 1


def o_ior(vm, st, lo, a, b, pc):
    y = st.pop()
    st[-1] = st[-1] | y
    return pc + 1


def o_ixor(vm, st, lo, a, b, pc):
    y = st.pop()
    st[-1] = st[-1] ^ y
    return pc + 1


def o_dadd(vm, st, lo, a, b, pc):
    y = st.pop()
    st[-1] = st[-1] + y
    return pc + 1


def o_dsub(vm, st, lo, a, b, pc):
    y = st.pop()
    st[-1] = st[-1] - y
    return pc + 1


def o_dmul(vm, st, lo, a, b, pc):
    y = st.pop()
    st[-1] = st[-1] * y
    return pc + 1


def o_ddiv(vm, st, lo, a, b, pc):
    y = st.pop()
    st[-1] = jdiv(st[-1], y)
    return pc + 1


def o_drem(vm, st, lo, a, b, pc):
    y = st.pop()
    x = st[-1]
    st[-1] = math.fmod(x, y) if (y != 0 and math.isfinite(x)) else float("nan")
    return pc + 1


def o_dneg(vm, st, lo, a, b, pc):
    st[-1] = -st[-1]
    return pc + 1


def o_fmul(vm, st, lo, a, b, pc):
    raise NotImplementedError("float arithmetic")


o_fdiv = o_fmul


def o_ladd(vm, st, lo, a, b, pc):
    raise NotImplementedError("long arithmetic")


o_lsub = o_lmul = o_lcmp = o_l2i = o_d2l = o_ladd


def o_i2d(vm, st, lo, a, b, pc):
    st[-1] = float(st[-1])
    return pc + 1


def o_d2i(vm, st, lo, a, b, pc):
    st[-1] = d2i(st[-1])
    return pc + 1


def o_d2f(vm, st, lo, a, b, pc):
    st[-1] = struct.unpack("f", struct.pack("f", st[-1]))[0]
    return pc + 1


def o_i2b(vm, st, lo, a, b, pc):
    v = st[-1] & 0xFF
    st[-1] = v - 256 if v > 127 else v
    return pc + 1


def o_i2c(vm, st, lo, a, b, pc):
    st[-1] = st[-1] & 0xFFFF
    return pc + 1


def o_i2s(vm, st, lo, a, b, pc):
    v = st[-1] & 0xFFFF
    st[-1] = v - 65536 if v > 32767 else v
    return pc + 1


def o_dcmpl(vm, st, lo, a, b, pc):
    y = st.pop()
    x = st[-1]
    st[-1] = -1 if (x != x or y != y) else (1 if x > y else (0 if x == y else -1))
    return pc + 1


def o_dcmpg(vm, st, lo, a, b, pc):
    y = st.pop()
    x = st[-1]
    st[-1] = 1 if (x != x or y != y) else (1 if x > y else (0 if x == y else -1))
    return pc + 1


def o_iinc(vm, st, lo, a, b, pc):
    lo[a] = i32(lo[a] + b)
    return pc + 1


def o_if(vm, st, lo, a, b, pc):
    v = st.pop()
    if b == 0:
        t = v == 0
    elif b == 1:
        t = v != 0
    elif b == 2:
        t = v < 0
    elif b == 3:
        t = v >= 0
    elif b == 4:
        t = v > 0
    else:
        t = v <= 0
    return a if t else pc + 1


def o_if_icmp(vm, st, lo, a, b, pc):
    y = st.pop()
    x = st.pop()
    if b == 0:
        t = x == y
    elif b == 1:
        t = x != y
    elif b == 2:
        t = x < y
    elif b == 3:
        t = x >= y
    elif b == 4:
        t = x > y
    else:
        t = x <= y
    return a if t else pc + 1


def o_if_acmp(vm, st, lo, a, b, pc):
    y = st.pop()
    x = st.pop()
    same = (x is y) or (isinstance(x, str) and isinstance(y, str) and x == y and len(x) < 8)  # interned literals
    return a if (same if b == 0 else not same) else pc + 1


def o_ifnull(vm, st, lo, a, b, pc):
    v = st.pop()
    return a if ((v is None) == bool(b)) else pc + 1


def o_goto(vm, st, lo, a, b, pc):
    return a


def o_tableswitch(vm, st, lo, a, b, pc):
    v = st.pop()
    low, high, default = a
    return b[v - low] if low <= v <= high else default


def o_lookupswitch(vm, st, lo, a, b, pc):
    return b.get(st.pop(), a)


def o_return_value(vm, st, lo, a, b, pc):
    v = st.pop()
    del st[:]
    st.append(v)
    return None


def o_return(vm, st, lo, a, b, pc):
    return None


def o_getstatic(vm, st, lo, a, b, pc):
    cn, fn, fd = a
    if cn == "java/lang/System":
        st.append(vm.stdout_obj)
    else:
        st.append(vm.cls(cn).statics[fn])
    return pc + 1


def o_putstatic(vm, st, lo, a, b, pc):
    cn, fn, fd = a
    vm.cls(cn).statics[fn] = st.pop()
    return pc + 1


def o_getfield(vm, st, lo, a, b, pc):
    obj = st.pop()
    cn, fn, fd = a
    f = obj.f
    st.append(f[fn] if fn in f else default_value(fd))
    return pc + 1


def o_putfield(vm, st, lo, a, b, pc):
    v = st.pop()
    obj = st.pop()
    obj.f[a[1]] = v
    return pc + 1


def o_new(vm, st, lo, a, b, pc):
    if a in vm.classes:
        st.append(JObject(vm.cls(a)))
    elif a in ("java/lang/StringBuffer", "java/lang/StringBuilder"):
        st.append(JStringBuffer())
    elif a == "java/lang/String":
        st.append(JStringBuffer())  # placeholder, replaced by the constructed str in o_invoke
    else:
        o = JObject(None)
        o.f["__class__"] = a
        st.append(o)
    return pc + 1


def o_newarray(vm, st, lo, a, b, pc):
    n = st.pop()
    if n < 0:
        raise ValueError("java.lang.NegativeArraySizeException")
    init = 0.0 if a in ("D", "F") else (None if a == "A" else 0)
    st.append(JArray([init] * n, a))
    return pc + 1


def o_multianewarray(vm, st, lo, a, b, pc):
    dims = [st.pop() for _ in range(b)][::-1]
    elem = a.lstrip("[")
    total_dims = len(a) - len(elem)

    def mk(level):
        n = dims[level]
        if level == len(dims) - 1:
            if total_dims == len(dims):
                init = 0.0 if elem in ("D", "F") else (None if elem.startswith("L") else 0)
                return JArray([init] * n, elem[0] if not elem.startswith("L") else "A")
            return JArray([None] * n, "A")
        return JArray([mk(level + 1) for _ in range(n)], "A")

    st.append(mk(0))
    return pc + 1


def o_arraylength(vm, st, lo, a, b, pc):
    st[-1] = len(st[-1].a)
    return pc + 1


def o_athrow(vm, st, lo, a, b, pc):
    raise RuntimeError("athrow")


def o_invoke(vm, st, lo, a, b, pc):
    cn, mn, md, nargs, ret, op = a
    args = st[len(st) - nargs:] if nargs else []
    if nargs:
        del st[len(st) - nargs:]
    if op == 182 and isinstance(args[0], JObject) and args[0].cls is not None:
        m = vm.find_method(args[0].cls.name, mn, md)
    elif cn in vm.classes:
        vm.cls(cn)
        m = vm.find_method(cn, mn, md)
    else:
        m = None
    if m is not None:
        r = vm.invoke(m, args)
    else:
        r = vm.native(cn, mn, md, args)
        if isinstance(r, tuple) and r and r[0] == "__string_init__":
            # `new String(x)`: replace the placeholder (pushed twice by new/dup) with the string itself
            ph = args[0]
            for k in range(len(st)):
                if st[k] is ph:
                    st[k] = r[1]
            for k in range(len(lo)):
                if lo[k] is ph:
                    lo[k] = r[1]
            return pc + 1
    if ret != "V":
        st.append(r)
    return pc + 1
