"""
jl2py -- a small mechanical Julia(0.6)-to-Python translator.  TEST INFRASTRUCTURE ONLY.

Purpose: the reference solver (`/root/reference/Weno5_1D.jl`, `Weno5_2D.jl`) is Julia 0.6 source and
no `julia` binary exists in this image, so the reference cannot be executed directly.  This module
translates the reference's functions *mechanically* (tokenizer + expression parser, no knowledge of
the numerics) into Python that runs on top of `oracle/jlrt.py`, a runtime that reproduces the Julia
semantics the files rely on (1-based column-major arrays, copying slices, linear indexing, `@inbounds`
reads that run off the end of a column, `x^2 == x*x`, ...).  `scripts/make_reference_golden.py` uses
it inside this container to run the UNMODIFIED reference functions and to write golden vectors under
`tests/golden/ref_*.npz`; the translated source is written to a scratch directory outside the repository and is never
committed.  Nothing in the product path imports this file.

Supported subset: function definitions (long and short form), for/while/if/elseif/else, return,
const/global/local, tuple and chained assignment, indexed assignment, op-assignment, `@.`, `@assert`,
ranges, array literals, ternaries, dotted operators and calls, parametric constructors
(`Array{Float64}(n,m)`), `end` inside indices, transposition.  Macros that only print or time are
turned into `pass`.
"""
import keyword
import re

KEYWORDS = {"function", "for", "while", "if", "elseif", "else", "end", "return", "const", "global", "local",
            "using", "export", "import", "module", "include", "break", "continue", "in", "true", "false"}
STRIP_MACROS = {"@inbounds", "@simd", "@fastmath", "@everywhere", "@sync", "@parallel", "@inline", "@noinline"}
NOOP_MACROS = {"@printf", "@show", "@time", "@sprintf", "@animate", "@assert_false"}
NOOP_CALLS = {"println", "print", "tic", "toc", "gc", "display", "flush"}

OPS = ["...", ".==", ".!=", ".<=", ".>=", "===", "!==", ".+=", ".-=", ".*=", "./=",
       "==", "!=", "<=", ">=", "&&", "||", "+=", "-=", "*=", "/=", "^=", ".*", "./", ".^", ".+", ".-", ".<", ".>",
       "::", "->", "=>", "<:",
       "+", "-", "*", "/", "^", "=", "<", ">", "!", "?", ":", ",", ";", "(", ")", "[", "]", "{", "}", ".", "%",
       "&", "|", "$", "~", "÷", "\\"]

ID_START = re.compile(r"[A-Za-z_¡-￿]")
ID_RE = re.compile(r"[A-Za-z_¡-￿][A-Za-z0-9_!¡-￿]*")
NUM_RE = re.compile(r"(0x[0-9a-fA-F]+|\d+\.\d*(?:[eE][+-]?\d+)?|\d+(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?)")


class Tok:
    __slots__ = ("kind", "val", "line", "sp")

    def __init__(self, kind, val, line, sp):
        self.kind, self.val, self.line, self.sp = kind, val, line, sp

    def __repr__(self):
        return f"{self.kind}:{self.val!r}@{self.line}"


class JlSyntaxError(Exception):
    pass


def tokenize(src):
    toks = []
    i, n, line = 0, len(src), 1
    sp = False
    while i < n:
        c = src[i]
        if c in " \t\r":
            i += 1
            sp = True
            continue
        if c == "\n":
            toks.append(Tok("NL", "\n", line, sp))
            line += 1
            i += 1
            sp = True
            continue
        if c == "#":
            if src.startswith("#=", i):
                depth, j = 1, i + 2
                while j < n and depth:
                    if src.startswith("#=", j):
                        depth += 1
                        j += 2
                    elif src.startswith("=#", j):
                        depth -= 1
                        j += 2
                    else:
                        if src[j] == "\n":
                            line += 1
                        j += 1
                i = j
            else:
                while i < n and src[i] != "\n":
                    i += 1
            continue
        if src.startswith('"""', i):
            j = src.index('"""', i + 3)
            s = src[i + 3:j]
            toks.append(Tok("STR", s, line, sp))
            line += s.count("\n")
            i = j + 3
            sp = False
            continue
        if c == '"':
            j = i + 1
            buf = []
            while src[j] != '"':
                if src[j] == "\\":
                    buf.append(src[j:j + 2])
                    j += 2
                else:
                    buf.append(src[j])
                    j += 1
            toks.append(Tok("STR", "".join(buf), line, sp))
            i = j + 1
            sp = False
            continue
        prev = toks[-1] if toks else None
        if c == "'":
            if prev is not None and not sp and (prev.kind in ("ID", "NUM") or prev.val in (")", "]", "'")):
                toks.append(Tok("OP", "'", line, sp))
                i += 1
                sp = False
                continue
            j = src.index("'", i + 1)
            toks.append(Tok("STR", src[i + 1:j], line, sp))
            i = j + 1
            sp = False
            continue
        if c == "@":
            if src.startswith("@.", i):
                toks.append(Tok("MACRO", "@.", line, sp))
                i += 2
            else:
                m = ID_RE.match(src, i + 1)
                toks.append(Tok("MACRO", "@" + m.group(0), line, sp))
                i = m.end()
            sp = False
            continue
        if c.isdigit() or (c == "." and i + 1 < n and src[i + 1].isdigit() and
                           not (prev is not None and not sp and (prev.kind in ("ID", "NUM") or prev.val in (")", "]")))):
            m = NUM_RE.match(src, i)
            s = m.group(0)
            # "1:3" / "1.:" fine; but "2.^x", "1./x": give the dot to the operator
            if s.endswith(".") and m.end() < n and src[m.end()] in "^*/+-=<>":
                s = s[:-1]
            toks.append(Tok("NUM", s, line, sp))
            i += len(s)
            sp = False
            continue
        if ID_START.match(c):
            m = ID_RE.match(src, i)
            s = m.group(0)
            while s.endswith("!") and src.startswith("=", i + len(s)) and not src.startswith("==", i + len(s)):
                s = s[:-1]
            toks.append(Tok("ID", s, line, sp))
            i += len(s)
            sp = False
            continue
        for op in OPS:
            if src.startswith(op, i):
                toks.append(Tok("OP", op, line, sp))
                i += len(op)
                break
        else:
            raise JlSyntaxError(f"line {line}: unexpected character {c!r}")
        sp = False
    toks.append(Tok("NL", "\n", line, True))
    toks.append(Tok("EOF", "", line, True))
    return toks


PY_RESERVED = set(keyword.kwlist) | {"print", "len", "list", "range", "type", "id", "input", "object", "str", "int", "float"}
RENAME = {"true": "True", "false": "False", "pi": "PI", "Inf": "INF", "NaN": "NAN", "nothing": "None", "π": "PI",
          "max": "jmax", "min": "jmin", "round": "jround", "sum": "jsum", "abs": "jabs", "zip": "jzip", "all": "jall",
          "any": "jany", "copy": "jcopy", "float": "jfloat", "Int": "jInt", "Int64": "jInt", "length": "jlength",
          "map": "jmap", "filter": "jfilter", "reverse": "jreverse", "sort": "jsort", "string": "jstring"}


def mangle(name):
    if name in RENAME:
        return RENAME[name]
    name = name.replace("!", "_b")
    name = "".join(ch if (ch.isascii()) else "u%04x" % ord(ch) for ch in name)
    if name in PY_RESERVED:
        name += "_"
    return name


CMP_OPS = {"==", "!=", "<", "<=", ">", ">=", ".==", ".!=", ".<", ".<=", ".>", ".>=", "===", "!=="}
ASSIGN_OPS = {"=", "+=", "-=", "*=", "/=", "^=", ".+=", ".-=", ".*=", "./="}


class Parser:
    def __init__(self, toks):
        self.t = toks
        self.p = 0
        self.depth = 0          # >0 inside () [] {}: newlines are whitespace
        self.in_index = 0
        self.tmp = 0

    # -- token helpers
    def peek(self, k=0):
        return self.t[min(self.p + k, len(self.t) - 1)]

    def next(self):
        tok = self.t[self.p]
        self.p += 1
        return tok

    def skip_nl(self):
        while self.peek().kind == "NL" or (self.peek().kind == "OP" and self.peek().val == ";" and self.depth == 0):
            self.p += 1

    def skip_nl_in(self):
        while self.peek().kind == "NL":
            self.p += 1

    def at(self, val, kind="OP"):
        tok = self.peek()
        return tok.kind == kind and tok.val == val

    def at_kw(self, val):
        tok = self.peek()
        return tok.kind == "ID" and tok.val == val

    def expect(self, val):
        tok = self.next()
        if tok.val != val:
            raise JlSyntaxError(f"line {tok.line}: expected {val!r}, got {tok.val!r}")
        return tok

    def skip_to_eol(self):
        """Skip the rest of a statement (balanced brackets may span lines)."""
        d = 0
        while True:
            tok = self.peek()
            if tok.kind == "EOF":
                return
            if tok.kind == "NL" and d <= 0:
                return
            if tok.kind == "OP" and tok.val in "([{":
                d += 1
            if tok.kind == "OP" and tok.val in ")]}":
                d -= 1
            self.p += 1

    # -- expressions: precedence climbing
    def expr(self):
        return self.ternary()

    def ternary(self):
        c = self.oror()
        if self.at("?"):
            self.next()
            self.skip_nl_in()
            saved, self._colon_is_ternary = self._colon_is_ternary, True
            a = self.ternary()
            self._colon_is_ternary = saved
            self.skip_nl_in()
            self.expect(":")
            self.skip_nl_in()
            b = self.ternary()
            return ("tern", c, a, b)
        return c

    def binop_after(self):
        self.next()
        self.skip_nl_in()

    def oror(self):
        l = self.andand()
        while self.at("||"):
            self.binop_after()
            l = ("bin", "or", l, self.andand())
        return l

    def andand(self):
        l = self.comparison()
        while self.at("&&"):
            self.binop_after()
            l = ("bin", "and", l, self.comparison())
        return l

    def comparison(self):
        l = self.range_()
        if self.peek().kind == "OP" and self.peek().val in CMP_OPS:
            operands, ops = [l], []
            while self.peek().kind == "OP" and self.peek().val in CMP_OPS:
                ops.append(self.next().val)
                self.skip_nl_in()
                operands.append(self.range_())
            return ("cmp", operands, ops)
        if self.at_kw("in") and False:
            pass
        return l

    def range_(self):
        # a:b or a:s:b ; a bare ':' inside an index is handled in atom()
        l = self.additive()
        if self.at(":") and not self._colon_is_ternary:
            self.next()
            m = self.additive()
            if self.at(":"):
                self.next()
                r = self.additive()
                return ("range", l, r, m)
            return ("range", l, m, None)
        return l

    _colon_is_ternary = False

    def additive(self):
        l = self.multiplicative()
        while self.peek().kind == "OP" and self.peek().val in ("+", "-", ".+", ".-", "|", "$"):
            # inside [ ] a space-separated "a -b" would be a matrix row; the reference does not use that
            op = self.next().val
            self.skip_nl_in()
            r = self.multiplicative()
            l = ("bin", op.lstrip(".") if op not in ("|", "$") else op, l, r)
        return l

    def multiplicative(self):
        l = self.unary()
        while self.peek().kind == "OP" and self.peek().val in ("*", "/", ".*", "./", "%", "÷", "&", "\\"):
            op = self.next().val
            self.skip_nl_in()
            r = self.unary()
            l = ("bin", op.lstrip(".") if len(op) > 1 else op, l, r)
        return l

    def unary(self):
        tok = self.peek()
        if tok.kind == "OP" and tok.val in ("-", "+", "!", "~"):
            self.next()
            x = self.unary()
            return ("un", tok.val, x)
        return self.juxtapose()

    def juxtapose(self):
        x = self.power()
        # numeric literal coefficient: 2pi, 4x, 2(a+b)
        if x[0] == "num" and not self.peek().sp and (self.peek().kind == "ID" and self.peek().val not in KEYWORDS
                                                      or self.at("(")):
            r = self.power()
            return ("bin", "*", x, r)
        return x

    def power(self):
        base = self.postfix()
        if self.peek().kind == "OP" and self.peek().val in ("^", ".^"):
            self.next()
            self.skip_nl_in()
            tok = self.peek()
            if tok.kind == "OP" and tok.val in ("-", "+"):
                self.next()
                e = ("un", tok.val, self.power_rhs())
            else:
                e = self.power_rhs()
            return ("pow", base, e)
        return base

    def power_rhs(self):
        # right associative
        return self.power()

    def postfix(self):
        x = self.atom()
        while True:
            tok = self.peek()
            if tok.kind != "OP":
                break
            if tok.val == "(" and not tok.sp:
                x = self.call(x, False)
            elif tok.val == "[" and not tok.sp:
                x = self.index(x)
            elif tok.val == "{" and not tok.sp:
                self.skip_curly()
                x = ("curly", x)
            elif tok.val == "." and self.peek(1).val == "(":
                self.next()
                x = self.call(x, True)
            elif tok.val == "." and self.peek(1).kind == "ID" and not tok.sp:
                self.next()
                x = ("field", x, self.next().val)
            elif tok.val == "'":
                self.next()
                x = ("transpose", x)
            else:
                break
        return x

    def skip_curly(self):
        self.expect("{")
        d = 1
        while d:
            tok = self.next()
            if tok.val == "{":
                d += 1
            elif tok.val == "}":
                d -= 1

    def call(self, f, dotted):
        self.expect("(")
        self.depth += 1
        saved = self.in_index
        self.in_index = 0
        args, kwargs = [], []
        self.skip_nl_in()
        while not self.at(")"):
            if self.at(";"):
                self.next()
                self.skip_nl_in()
                continue
            a = self.expr()
            if self.at("="):                      # keyword argument
                self.next()
                self.skip_nl_in()
                v = self.expr()
                kwargs.append((a, v))
            elif self.at("..."):
                self.next()
                args.append(("splat", a))
            else:
                args.append(a)
            self.skip_nl_in()
            if self.at(","):
                self.next()
                self.skip_nl_in()
        self.expect(")")
        self.depth -= 1
        self.in_index = saved
        return ("call", f, args, kwargs, dotted)

    def index(self, base):
        self.expect("[")
        self.depth += 1
        self.in_index += 1
        args = []
        self.skip_nl_in()
        while not self.at("]"):
            args.append(self.expr())
            self.skip_nl_in()
            if self.at(","):
                self.next()
                self.skip_nl_in()
        self.expect("]")
        self.in_index -= 1
        self.depth -= 1
        return ("index", base, args)

    def atom(self):
        tok = self.next()
        if tok.kind == "NUM":
            return ("num", tok.val)
        if tok.kind == "STR":
            return ("str", tok.val)
        if tok.kind == "MACRO":
            # macro call in expression position: not needed by the hot path
            raise JlSyntaxError(f"line {tok.line}: macro {tok.val} in expression")
        if tok.kind == "ID":
            if tok.val == "end" and self.in_index:
                return ("end",)
            if tok.val in KEYWORDS and tok.val not in ("true", "false"):
                raise JlSyntaxError(f"line {tok.line}: keyword {tok.val!r} in expression")
            return ("id", tok.val)
        if tok.kind == "OP":
            if tok.val == "(":
                self.depth += 1
                saved = self.in_index
                self.in_index = 0
                self.skip_nl_in()
                items = []
                is_tuple = False
                while not self.at(")"):
                    items.append(self.expr())
                    self.skip_nl_in()
                    if self.at(","):
                        is_tuple = True
                        self.next()
                        self.skip_nl_in()
                self.expect(")")
                self.depth -= 1
                self.in_index = saved
                if is_tuple or not items:
                    return ("tuple", items)
                return ("paren", items[0])
            if tok.val == "[":
                self.depth += 1
                saved = self.in_index
                self.in_index = 0
                self.skip_nl_in()
                items = []
                while not self.at("]"):
                    items.append(self.expr())
                    self.skip_nl_in()
                    if self.at_kw("for"):
                        self.next()
                        var = self.next().val
                        if not (self.at("=") or self.at_kw("in")):
                            raise JlSyntaxError(f"line {tok.line}: bad comprehension")
                        self.next()
                        it = self.expr()
                        self.skip_nl_in()
                        self.expect("]")
                        self.depth -= 1
                        self.in_index = saved
                        return ("comp", items[0], var, it)
                    if self.at(","):
                        self.next()
                        self.skip_nl_in()
                    elif not self.at("]"):
                        raise JlSyntaxError(f"line {tok.line}: matrix literal not supported")
                self.expect("]")
                self.depth -= 1
                self.in_index = saved
                return ("vect", items)
            if tok.val == ":" and self.in_index:
                return ("colon",)
            if tok.val == ":" and self.peek().kind == "ID":
                return ("str", self.next().val)          # symbol
        raise JlSyntaxError(f"line {tok.line}: unexpected token {tok.val!r}")

    # -- code generation for expressions
    def gen(self, n):
        k = n[0]
        if k == "num":
            s = n[1]
            return s + "0" if s.endswith(".") else s
        if k == "str":
            return repr(n[1])
        if k == "id":
            return mangle(n[1])
        if k == "paren":
            return "(" + self.gen(n[1]) + ")"
        if k == "tuple":
            return "(" + ", ".join(self.gen(x) for x in n[1]) + ("," if len(n[1]) == 1 else "") + ")"
        if k == "vect":
            return "jvect([" + ", ".join(self.gen(x) for x in n[1]) + "])"
        if k == "comp":
            return f"jvect([{self.gen(n[1])} for {mangle(n[2])} in {self.gen_iter(n[3])}])"
        if k == "colon":
            return "COLON"
        if k == "end":
            return "END"
        if k == "range":
            a, b, s = self.gen(n[1]), self.gen(n[2]), n[3]
            return f"jrange({a}, {b})" if s is None else f"jrange({a}, {b}, {self.gen(s)})"
        if k == "bin":
            op, l, r = n[1], self.gen(n[2]), self.gen(n[3])
            if op == "%":
                return f"jrem({l}, {r})"
            if op == "÷":
                return f"jdiv({l}, {r})"
            if op == "\\":
                return f"jldiv({l}, {r})"
            if op == "$":
                op = "^"
            return f"({l} {op} {r})"
        if k == "pow":
            return f"jpow({self.gen(n[1])}, {self.gen(n[2])})"
        if k == "un":
            if n[1] == "!":
                return f"jnot({self.gen(n[2])})"
            return f"({n[1]}{self.gen(n[2])})"
        if k == "cmp":
            if len(n[2]) == 1 and n[2][0] in ("==", "!=", "===", "!=="):
                # `a == b` on arrays is a single Bool in Julia (elementwise is `.==`)
                f = "jeq" if n[2][0] in ("==", "===") else "jne"
                return f"{f}({self.gen(n[1][0])}, {self.gen(n[1][1])})"
            parts = [self.gen(n[1][0])]
            for op, x in zip(n[2], n[1][1:]):
                op = {"===": "==", "!==": "!="}.get(op, op).lstrip(".")
                parts.append(op)
                parts.append(self.gen(x))
            return "(" + " ".join(parts) + ")"
        if k == "tern":
            return f"({self.gen(n[2])} if {self.gen(n[1])} else {self.gen(n[3])})"
        if k == "transpose":
            return f"transpose({self.gen(n[1])})"
        if k == "field":
            return f"{self.gen(n[1])}.{mangle(n[2])}"
        if k == "curly":
            return self.gen(n[1])
        if k == "splat":
            return "*" + self.gen(n[1])
        if k == "index":
            return f"jget({self.gen(n[1])}, ({', '.join(self.gen(x) for x in n[2])},))"
        if k == "call":
            f = self.gen(n[1])
            args = [self.gen(x) for x in n[2]] + [f"{self.gen(a)}={self.gen(v)}" for a, v in n[3]]
            return f"{f}({', '.join(args)})"
        raise JlSyntaxError(f"cannot generate {n!r}")

    def gen_iter(self, n):
        if n[0] == "range" and n[3] is None:
            return f"range({self.gen(n[1])}, ({self.gen(n[2])}) + 1)"
        return f"jiter({self.gen(n)})"

    # -- statements
    def block(self, out, ind, terminators=("end",), fn=None):
        """Parse statements until one of the terminator keywords; returns the terminator."""
        n0 = len(out)
        while True:
            self.skip_nl()
            tok = self.peek()
            if tok.kind == "EOF":
                if "EOF" in terminators:
                    break
                raise JlSyntaxError("unexpected end of file")
            if tok.kind == "ID" and tok.val in terminators:
                break
            self.statement(out, ind, fn)
        if len(out) == n0:
            out.append(ind + "pass")
        return self.peek().val

    def statement(self, out, ind, fn):
        tok = self.peek()
        if tok.kind == "MACRO":
            self.next()
            if tok.val in STRIP_MACROS:
                return self.statement(out, ind, fn)
            if tok.val == "@.":
                return self.assignment(out, ind, fn, broadcast=True)
            if tok.val == "@assert":
                cond = self.gen(self.expr())
                if self.peek().kind != "NL":
                    msg = self.gen(self.expr())
                    out.append(f"{ind}assert {cond}, {msg}")
                else:
                    out.append(f"{ind}assert {cond}")
                return
            # printing / timing / plotting macros
            self.skip_to_eol()
            out.append(ind + "pass")
            return
        if tok.kind == "STR":                       # docstring
            self.next()
            return
        if tok.kind == "ID":
            v = tok.val
            if v == "function":
                return self.function(out, ind)
            if v == "for":
                return self.for_(out, ind, fn)
            if v == "while":
                self.next()
                cond = self.gen(self.expr())
                out.append(f"{ind}while {cond}:")
                self.block(out, ind + "    ", fn=fn)
                self.expect("end")
                return
            if v == "if":
                return self.if_(out, ind, fn)
            if v == "return":
                self.next()
                if self.peek().kind == "NL" or self.at_kw("end"):
                    out.append(ind + "return")
                    return
                items = [self.expr()]
                while self.at(","):
                    self.next()
                    self.skip_nl_in()
                    items.append(self.expr())
                out.append(ind + "return " + ", ".join(self.gen(x) for x in items))
                return
            if v in ("const", "local"):
                self.next()
                return self.statement(out, ind, fn)
            if v == "global":
                self.next()
                names = []
                while True:
                    name = self.peek()
                    names.append(name.val)
                    if self.peek(1).val == ",":
                        self.p += 2
                        continue
                    break
                if fn is not None:
                    fn["globals"].update(mangle(x) for x in names)
                if self.peek(1).val == "=":
                    return self.assignment(out, ind, fn)
                self.next()
                return
            if v in ("using", "export", "import", "module", "include", "importall"):
                self.skip_to_eol()
                return
            if v in ("break", "continue"):
                self.next()
                out.append(ind + v)
                return
            if v in NOOP_CALLS and self.peek(1).val == "(":
                self.skip_to_eol()
                out.append(ind + "pass")
                return
        return self.assignment(out, ind, fn)

    def for_(self, out, ind, fn):
        self.expect("for")
        loops = []
        while True:
            var = self.expr_target()
            if not (self.at("=") or self.at_kw("in")):
                raise JlSyntaxError(f"line {self.peek().line}: bad for")
            self.next()
            it = self.expr()
            loops.append((var, it))
            if self.at(","):
                self.next()
                continue
            break
        cur = ind
        for var, it in loops:
            out.append(f"{cur}for {self.gen(var)} in {self.gen_iter(it)}:")
            cur += "    "
        self.block(out, cur, fn=fn)
        self.expect("end")

    def expr_target(self):
        tok = self.peek()
        if tok.kind == "OP" and tok.val == "(":
            return self.atom()
        return ("id", self.next().val)

    def if_(self, out, ind, fn):
        self.expect("if")
        cond = self.gen(self.expr())
        out.append(f"{ind}if {cond}:")
        while True:
            term = self.block(out, ind + "    ", terminators=("end", "elseif", "else"), fn=fn)
            self.next()
            if term == "end":
                return
            if term == "elseif":
                cond = self.gen(self.expr())
                out.append(f"{ind}elif {cond}:")
            else:
                if self.at_kw("if"):                # "else if": nested if that owns its own end
                    out.append(f"{ind}else:")
                    self.if_(out, ind + "    ", fn)
                    self.skip_nl()
                    self.expect("end")
                    return
                out.append(f"{ind}else:")

    def function(self, out, ind):
        self.expect("function")
        name = self.next().val
        if self.at("{"):
            self.skip_curly()
        params = self.params()
        fn = {"globals": set()}
        body = []
        self.block(body, ind + "    ", fn=fn)
        self.expect("end")
        out.append(f"{ind}def {mangle(name)}({', '.join(params)}):")
        if fn["globals"]:
            out.append(f"{ind}    global {', '.join(sorted(fn['globals']))}")
        out.extend(body)
        out.append("")

    def params(self):
        self.expect("(")
        self.depth += 1
        params = []
        self.skip_nl_in()
        while not self.at(")"):
            if self.at(";"):
                self.next()
                self.skip_nl_in()
                continue
            name = self.next().val
            if self.at("::"):
                self.next()
                self.type_()
            default = None
            if self.at("="):
                self.next()
                default = self.gen(self.expr())
            params.append(mangle(name) + (f"={default}" if default is not None else ""))
            self.skip_nl_in()
            if self.at(","):
                self.next()
                self.skip_nl_in()
        self.expect(")")
        self.depth -= 1
        if self.at("::"):
            self.next()
            self.type_()
        return params

    def type_(self):
        self.next()
        while self.at("."):
            self.next()
            self.next()
        if self.at("{"):
            self.skip_curly()

    def target_code(self, t, value, out, ind):
        k = t[0]
        if k == "id":
            out.append(f"{ind}{mangle(t[1])} = {value}")
        elif k == "index":
            out.append(f"{ind}jset({self.gen(t[1])}, ({', '.join(self.gen(x) for x in t[2])},), {value})")
        elif k == "tuple":
            if all(x[0] == "id" for x in t[1]):
                out.append(f"{ind}{', '.join(mangle(x[1]) for x in t[1])} = {value}")
            else:
                self.tmp += 1
                tmp = f"_t{self.tmp}"
                out.append(f"{ind}{tmp} = {value}")
                for i, x in enumerate(t[1]):
                    self.target_code(x, f"{tmp}[{i}]", out, ind)
        elif k == "field":
            out.append(f"{ind}{self.gen(t)} = {value}")
        elif k == "paren":
            self.target_code(t[1], value, out, ind)
        else:
            raise JlSyntaxError(f"cannot assign to {t!r}")

    def assignment(self, out, ind, fn, broadcast=False):
        line = self.peek().line
        first = self.expr()
        if self.at(","):                            # a, b, c = ...
            items = [first]
            while self.at(","):
                self.next()
                items.append(self.expr())
            first = ("tuple", items)
        tok = self.peek()
        if not (tok.kind == "OP" and tok.val in ASSIGN_OPS):
            out.append(ind + self.gen(first))       # expression statement
            return
        op = self.next().val
        self.skip_nl_in()
        if op != "=":
            rhs = self.gen(self.expr())
            bop = op[:-1].lstrip(".") if len(op) > 2 else op[0]
            if bop == "^":
                value = f"jpow({self.gen(first)}, {rhs})"
            else:
                value = f"({self.gen(first)} {bop} ({rhs}))"
            if op.startswith(".") and first[0] == "id":
                out.append(f"{ind}jbroadcast_assign({mangle(first[1])}, {value})")
            else:
                self.target_code(first, value, out, ind)
            return
        # short-form function definition f(x) = expr
        if first[0] == "call" and first[1][0] == "id" and not first[4]:
            body = self.gen(self.expr())
            args = [self.gen(a) for a in first[2]]
            out.append(f"{ind}def {mangle(first[1][1])}({', '.join(args)}):")
            out.append(f"{ind}    return {body}")
            return
        targets = [first]
        rhs = self.expr()
        if self.at(","):                            # x = a, b
            items = [rhs]
            while self.at(","):
                self.next()
                items.append(self.expr())
            rhs = ("tuple", items)
        while self.at("="):                         # chained a = b = c
            self.next()
            self.skip_nl_in()
            targets.append(rhs)
            rhs = self.expr()
        value = self.gen(rhs)
        if broadcast:
            if len(targets) != 1:
                raise JlSyntaxError(f"line {line}: chained @. assignment")
            t = targets[0]
            if t[0] == "id":
                out.append(f"{ind}jbroadcast_assign({mangle(t[1])}, {value})")
            else:
                self.target_code(t, value, out, ind)
            return
        if len(targets) == 1:
            self.target_code(targets[0], value, out, ind)
            return
        self.tmp += 1
        tmp = f"_t{self.tmp}"
        out.append(f"{ind}{tmp} = {value}")
        for t in reversed(targets):
            self.target_code(t, tmp, out, ind)


BLOCK_OPENERS = {"function", "for", "while", "if", "begin", "let", "try", "do", "type", "struct", "module", "quote",
                 "immutable", "macro"}


def split_top_level(toks):
    """Cut the token stream into top-level items (one per `function ... end`, `for ... end`, or line) by
    counting block openers against `end` (an `end` inside [ ] is an index, not a terminator), so that a
    construct the translator does not know (e.g. in the plotting driver) only loses its own function."""
    items = []
    i, n = 0, len(toks)
    while i < n and toks[i].kind != "EOF":
        if toks[i].kind == "NL":
            i += 1
            continue
        if toks[i].kind == "ID" and toks[i].val == "module":        # `module X` ... `end`: look inside
            while toks[i].kind != "NL":
                i += 1
            continue
        if toks[i].kind == "ID" and toks[i].val == "end":           # the module's closing `end`
            i += 1
            continue
        start = i
        depth = brack = 0
        while i < n and toks[i].kind != "EOF":
            tok = toks[i]
            if tok.kind == "OP" and tok.val == "[":
                brack += 1
            elif tok.kind == "OP" and tok.val == "]":
                brack -= 1
            elif tok.kind == "ID" and brack == 0:
                if tok.val in BLOCK_OPENERS:
                    depth += 1
                elif tok.val == "end":
                    depth -= 1
            elif tok.kind == "NL" and depth <= 0 and brack <= 0:
                break
            i += 1
        items.append(toks[start:i])
    return items


def translate(src, header="from oracle.jlrt import *\n", only=None):
    """Translate Julia source text; returns (python_source, report) where report lists the top-level items
    that could not be translated (line, first tokens, error).  `only`: keep just the named functions."""
    out = [header]
    failed = []
    for item in split_top_level(tokenize(src)):
        line0 = item[0].line
        if only is not None and not (item[0].val == "function" and len(item) > 1 and item[1].val in only):
            continue
        try:
            last = item[-1].line
            ps = Parser(item + [Tok("NL", "\n", last, True), Tok("EOF", "", last, True)])
            lines = []
            ps.block(lines, "", terminators=("EOF",))
            if lines == ["pass"]:
                continue
            code = "\n".join(lines)
            compile(code, "<jl2py>", "exec")
            out.append(f"# ---- reference line {line0}")
            out.append(code)
        except (JlSyntaxError, SyntaxError, ValueError, IndexError) as e:
            failed.append((line0, " ".join(str(t.val) for t in item[:6]), str(e)))
    return "\n".join(out) + "\n", failed
