"""oracle/vsim.py -- a small event-free simulator for the synthesizable Verilog subset that ZipCPU/cordic's
generator emits.  TEST INFRASTRUCTURE ONLY (see oracle/zc_oracle.h for the rules about oracle/).

Why it exists.  The reference holds no per-sample golden vectors for its cores and Verilator is not installable
here, so the C restatement in zc_oracle.c would otherwise be pinned to the RTL only by reading it.  This module
executes the reference's RTL *text* itself -- rtl/cordic.v, rtl/topolar.v, rtl/quadtbl.v, rtl/sintable.v,
rtl/quarterwav.v as checked in, and whatever oracle/_ref/gencordic (the reference's real generator) prints for
other command lines -- under the expression rules of IEEE 1364-2005 (section 5.4/5.5: context-determined widths,
an expression is signed only if all its operands are, part-selects and concatenations are unsigned, `>>>` is
arithmetic only in a signed context, non-blocking assignments read the old state).  tests/golden/make_rtl_vectors.py
runs it to produce tests/golden/rtl_vectors.json, against which the oracle (CPU tier) and the CUDA path (GPU tier)
are checked word for word.

Supported: module with #(parameter/localparam) header, ANSI ports, wire/reg [signed] [range] scalars and arrays,
localparam, assign (whole, bit, part-select), always @(posedge clk) / always @(*), initial (ignored except
$readmemh), begin/end, if/else, case/default, blocking and non-blocking assignment (incl. concatenation targets),
generate-for with a genvar, $signed, $readmemh, the operators the generator uses (+ - * ! ~ & | ^ && || == != < <= > >=
<< >> >>> ?: unary reductions), sized/unsized literals, concatenation and replication.
"""
import os
import re

# ------------------------------------------------------------------------------------------- tokenizer
_TOK = re.compile(r"""
    (?P<ws>\s+|//[^\n]*|/\*.*?\*/|`[^\n]*) |
    (?P<num>\d+\s*'\s*[sS]?[bBhHdDoO]\s*[0-9a-fA-F_xXzZ?]+|'[sS]?[bBhHdDoO]\s*[0-9a-fA-F_]+|\d[\d_]*) |
    (?P<id>[A-Za-z_$][A-Za-z0-9_$]*) |
    (?P<str>"[^"]*") |
    (?P<op>>>>|<<<|<=|>=|==|!=|&&|\|\||<<|>>|[-+*/%!~&|^<>=?:;,.#@(){}\[\]])
""", re.S | re.X)


def tokenize(text):
    out, pos = [], 0
    while pos < len(text):
        m = _TOK.match(text, pos)
        if not m:
            raise SyntaxError("vsim: cannot tokenize at %r" % text[pos:pos + 30])
        pos = m.end()
        if m.lastgroup != "ws":
            out.append((m.lastgroup, m.group(m.lastgroup)))
    out.append(("eof", ""))
    return out


# ------------------------------------------------------------------------------------------- parser
class Parser:
    def __init__(self, text):
        self.t = tokenize(text)
        self.i = 0

    def peek(self, k=0):
        return self.t[self.i + k][1]

    def kind(self):
        return self.t[self.i][0]

    def next(self):
        v = self.t[self.i][1]
        self.i += 1
        return v

    def accept(self, v):
        if self.peek() == v:
            self.i += 1
            return True
        return False

    def expect(self, v):
        if self.peek() != v:
            raise SyntaxError("vsim: expected %r, got %r (token %d)" % (v, self.peek(), self.i))
        self.i += 1

    # ---- expressions -----------------------------------------------------------------------
    BIN = [["||"], ["&&"], ["|"], ["^"], ["&"], ["==", "!="], ["<", "<=", ">", ">="], ["<<", ">>", ">>>", "<<<"],
           ["+", "-"], ["*", "/", "%"]]

    def expr(self, lvl=0, no_le=False):
        if lvl == 0:
            c = self.expr(1, no_le)
            if self.accept("?"):
                a = self.expr(0)
                self.expect(":")
                b = self.expr(0)
                return ("?:", c, a, b)
            return c
        if lvl > len(self.BIN):
            return self.unary()
        a = self.expr(lvl + 1, no_le)
        while self.peek() in self.BIN[lvl - 1] and not (no_le and self.peek() == "<="):
            op = self.next()
            a = ("bin", op, a, self.expr(lvl + 1, no_le))
        return a

    def unary(self):
        if self.peek() in ("+", "-", "!", "~", "&", "|", "^"):
            op = self.next()
            return ("un", op, self.unary())
        return self.primary()

    def primary(self):
        k, v = self.t[self.i]
        if v == "(":
            self.next()
            e = self.expr()
            self.expect(")")
            return e
        if v == "{":
            self.next()
            first = self.expr()
            if self.peek() == "{":            # replication {n{...}}
                self.next()
                items = [self.expr()]
                while self.accept(","):
                    items.append(self.expr())
                self.expect("}")
                self.expect("}")
                return ("rep", first, ("cat", items))
            items = [first]
            while self.accept(","):
                items.append(self.expr())
            self.expect("}")
            return ("cat", items)
        if k == "num":
            self.next()
            return ("num", v)
        if k == "id":
            self.next()
            if v == "$signed":
                self.expect("(")
                e = self.expr()
                self.expect(")")
                return ("signed", e)
            node = ("id", v)
            while self.peek() == "[":
                self.next()
                a = self.expr()
                if self.accept(":"):
                    b = self.expr()
                    self.expect("]")
                    node = ("part", node, a, b)
                else:
                    self.expect("]")
                    node = ("idx", node, a)
            return node
        raise SyntaxError("vsim: unexpected token %r" % v)

    # ---- statements ------------------------------------------------------------------------
    def stmt(self):
        v = self.peek()
        if v == "begin":
            self.next()
            if self.accept(":"):
                self.next()
            body = []
            while not self.accept("end"):
                body.append(self.stmt())
            return ("block", body)
        if v == "if":
            self.next()
            self.expect("(")
            c = self.expr()
            self.expect(")")
            t = self.stmt()
            e = self.stmt() if self.accept("else") else None
            return ("if", c, t, e)
        if v == "case":
            self.next()
            self.expect("(")
            sel = self.expr()
            self.expect(")")
            arms, default = [], None
            while not self.accept("endcase"):
                if self.accept("default"):
                    self.accept(":")
                    default = self.stmt()
                else:
                    labels = [self.expr()]
                    while self.accept(","):
                        labels.append(self.expr())
                    self.expect(":")
                    arms.append((labels, self.stmt()))
            return ("case", sel, arms, default)
        if v.startswith("$"):
            name = self.next()
            args = []
            if self.accept("("):
                while not self.accept(")"):
                    if self.kind() == "str":
                        args.append(("str", self.next()[1:-1]))
                    else:
                        args.append(self.expr())
                    self.accept(",")
            self.expect(";")
            return ("task", name, args)
        if v == ";":
            self.next()
            return ("block", [])
        lhs = self.lvalue()
        if self.accept("<="):
            kind = "nba"
        else:
            self.expect("=")
            kind = "ba"
        rhs = self.expr()
        self.expect(";")
        return (kind, lhs, rhs)

    def lvalue(self):
        if self.peek() == "{":
            self.next()
            items = [self.lvalue()]
            while self.accept(","):
                items.append(self.lvalue())
            self.expect("}")
            return ("cat", items)
        name = self.next()
        node = ("id", name)
        while self.peek() == "[":
            self.next()
            a = self.expr()
            if self.accept(":"):
                b = self.expr()
                self.expect("]")
                node = ("part", node, a, b)
            else:
                self.expect("]")
                node = ("idx", node, a)
        return node

    # ---- module ----------------------------------------------------------------------------
    def decl_tail(self, kind, items, ports=None, direction=None):
        """[signed] [range] name [array] {, name [array]} ;  -- appended to items as ('decl', ...)"""
        signed = self.accept("signed")
        rng = None
        if self.peek() == "[":
            self.next()
            a = self.expr()
            self.expect(":")
            b = self.expr()
            self.expect("]")
            rng = (a, b)
        while True:
            name = self.next()
            arr = None
            if self.peek() == "[":
                self.next()
                a = self.expr()
                self.expect(":")
                b = self.expr()
                self.expect("]")
                arr = (a, b)
            init = None
            if self.accept("="):
                init = self.expr()
            items.append(("decl", kind, signed, rng, name, arr, init))
            if ports is not None:
                ports.append((direction, name))
            if not self.accept(","):
                break
            if self.peek() in ("input", "output", "inout"):
                return False        # header list continues with a new direction
        return True

    def module(self):
        while self.peek() != "module":
            self.next()
        self.next()
        name = self.next()
        items, ports = [], []
        if self.accept("#"):
            self.expect("(")
            while not self.accept(")"):
                if self.peek() in ("parameter", "localparam"):
                    self.next()
                pname = self.next()
                self.expect("=")
                items.append(("param", pname, self.expr()))
                self.accept(",")
        self.expect("(")
        while not self.accept(")"):
            direction = self.next()
            assert direction in ("input", "output", "inout"), direction
            kind = "wire"
            if self.peek() in ("wire", "reg"):
                kind = self.next()
            self.decl_tail(kind, items, ports, direction)
            self.accept(",")
        self.expect(";")
        self.items(items, "endmodule")
        return name, items, ports

    def items(self, items, closer):
        while not self.accept(closer):
            v = self.peek()
            if v in ("localparam", "parameter"):
                self.next()
                while True:
                    pname = self.next()
                    self.expect("=")
                    items.append(("param", pname, self.expr()))
                    if not self.accept(","):
                        break
                self.expect(";")
            elif v in ("wire", "reg", "integer"):
                self.next()
                self.decl_tail(v, items)
                self.expect(";")
            elif v == "genvar":
                self.next()
                self.next()
                self.expect(";")
            elif v == "assign":
                self.next()
                lhs = self.lvalue()
                self.expect("=")
                items.append(("assign", lhs, self.expr()))
                self.expect(";")
            elif v == "initial":
                self.next()
                items.append(("initial", self.stmt()))
            elif v == "always":
                self.next()
                self.expect("@")
                edges = None
                if self.accept("("):
                    if self.accept("*"):
                        self.expect(")")
                    else:
                        edges = []
                        while not self.accept(")"):
                            e = self.next()
                            if e in ("posedge", "negedge"):
                                edges.append((e, self.next()))
                            self.accept(",")
                            if self.peek() == "or":
                                self.next()
                else:
                    self.expect("*")
                items.append(("always", edges, self.stmt()))
            elif v == "generate":
                self.next()
                self.items(items, "endgenerate")
            elif v == "for":
                self.next()
                self.expect("(")
                var = self.next()
                self.expect("=")
                init = self.expr()
                self.expect(";")
                cond = self.expr()
                self.expect(";")
                self.next()
                self.expect("=")
                step = self.expr()
                self.expect(")")
                self.expect("begin")
                if self.accept(":"):
                    self.next()
                body = []
                self.items(body, "end")
                items.append(("genfor", var, init, cond, step, body))
            else:
                raise SyntaxError("vsim: unsupported module item %r" % v)


# ------------------------------------------------------------------------------------------- elaboration
def _mask(w):
    return (1 << w) - 1


def _sext(v, w):
    v &= _mask(w)
    return v - (1 << w) if v >> (w - 1) else v


def parse_number(txt):
    """-> (value, width, signed)"""
    txt = txt.replace("_", "").replace(" ", "")
    if "'" not in txt:
        return int(txt), 32, True
    size, rest = txt.split("'")
    signed = rest[0] in "sS"
    if signed:
        rest = rest[1:]
    base = {"b": 2, "h": 16, "d": 10, "o": 8}[rest[0].lower()]
    val = int(rest[1:], base)
    width = int(size) if size else 32
    return val & _mask(width), width, signed


class Sig:
    __slots__ = ("name", "width", "signed", "depth", "lsb")

    def __init__(self, name, width, signed, depth, lsb=0):
        self.name, self.width, self.signed, self.depth, self.lsb = name, width, signed, depth, lsb


class Module:
    """Elaborated module: constants folded, generate loops unrolled, everything compiled to closures."""

    def __init__(self, path, overrides=None):
        self.dir = os.path.dirname(os.path.abspath(path))
        self.name, items, self.ports = Parser(open(path).read()).module()
        self.consts = dict(overrides or {})
        self.sigs = {}
        self.state = {}
        self.comb = []          # (kind, compiled) continuous assigns and always @(*) blocks
        self.seq = []           # compiled always @(posedge ...) bodies
        self.clock = None
        self.inits = []         # compiled `initial x = value;` statements, run once after elaboration
        self._elab(items, {})
        upd = []
        for f in self.inits:
            f(self.state, upd)
        self._commit(upd)
        self.settle()

    # ---- constant expressions ----------------------------------------------------------------
    def cval(self, e, env):
        k = e[0]
        if k == "num":
            v, w, s = parse_number(e[1])
            return _sext(v, w) if s else v
        if k == "id":
            if e[1] in env:
                return env[e[1]]
            return self.consts[e[1]]
        if k == "un":
            v = self.cval(e[2], env)
            return {"-": -v, "+": v, "!": int(not v), "~": ~v}[e[1]]
        if k == "bin":
            a, b = self.cval(e[2], env), self.cval(e[3], env)
            return {"+": a + b, "-": a - b, "*": a * b, "<<": a << b if e[1] == "<<" else 0, ">>": a >> b if e[1] == ">>" else 0,
                    "==": int(a == b), "!=": int(a != b), "<": int(a < b), "<=": int(a <= b), ">": int(a > b),
                    ">=": int(a >= b), "&&": int(bool(a) and bool(b)), "||": int(bool(a) or bool(b)),
                    "/": a // b if e[1] == "/" and b else 0, "%": a % b if e[1] == "%" and b else 0}[e[1]]
        if k == "?:":
            return self.cval(e[2], env) if self.cval(e[1], env) else self.cval(e[3], env)
        raise ValueError("vsim: not a constant expression: %r" % (e,))

    def is_const(self, e, env):
        try:
            self.cval(e, env)
            return True
        except (KeyError, ValueError):
            return False

    # ---- elaboration ---------------------------------------------------------------------------
    def _elab(self, items, env):
        for it in items:
            k = it[0]
            if k == "param":
                if it[1] not in self.consts:
                    self.consts[it[1]] = self.cval(it[2], env)
            elif k == "decl":
                _, kind, signed, rng, name, arr, init = it
                width = 32 if kind == "integer" else 1
                lsb = 0
                if rng:
                    hi, lo = self.cval(rng[0], env), self.cval(rng[1], env)
                    width, lsb = hi - lo + 1, lo
                depth = None
                if arr:
                    a, b = self.cval(arr[0], env), self.cval(arr[1], env)
                    depth = abs(b - a) + 1
                if name not in self.sigs:
                    self.sigs[name] = Sig(name, width, signed or kind == "integer", depth, lsb)
                    self.state[name] = [0] * depth if depth else 0
                if init is not None:            # net declaration assignment: wire [..] y = expr;
                    self.comb.append(self.c_assign(("ba", ("id", name), init), env))
            elif k == "assign":
                self.comb.append(self.c_assign(("ba", it[1], it[2]), env))
            elif k == "always":
                edges, body = it[1], it[2]
                if edges is None:
                    self.comb.append(self.c_stmt(body, env))
                else:
                    self.clock = edges[0][1]
                    self.seq.append(self.c_stmt(body, env))
            elif k == "initial":
                self._initial(it[1], env)
            elif k == "genfor":
                _, var, init, cond, step, body = it
                e2 = dict(env)
                e2[var] = self.cval(init, env)
                while self.cval(cond, e2):
                    self._elab(body, e2)
                    e2 = dict(e2)
                    e2[var] = self.cval(step, e2)

    def _initial(self, st, env):
        if st[0] == "block":
            for s in st[1]:
                self._initial(s, env)
        elif st[0] == "task" and st[1] == "$readmemh":
            fname, mem = st[2][0][1], st[2][1][1]
            words, addr = self.state[mem], 0
            for tok in open(os.path.join(self.dir, fname)).read().split():
                if tok.startswith("@"):
                    addr = int(tok[1:], 16)
                else:
                    words[addr] = int(tok, 16) & _mask(self.sigs[mem].width)
                    addr += 1
        elif st[0] in ("ba", "nba"):            # initial x = value;  (rtl/seqcordic.v:204-219, :230)
            self.inits.append(self.c_assign(st, env))

    # ---- expression typing (IEEE 1364-2005 table 5-22) ---------------------------------------
    def typ(self, e, env):
        """-> (self-determined width, signed)"""
        k = e[0]
        if k == "num":
            _, w, s = parse_number(e[1])
            return w, s
        if k == "id":
            if e[1] in env or e[1] in self.consts:
                return 32, True
            s = self.sigs[e[1]]
            return s.width, s.signed
        if k == "idx":
            base = e[1]
            if base[0] == "id" and base[1] in self.sigs and self.sigs[base[1]].depth:
                s = self.sigs[base[1]]
                return s.width, s.signed          # array element
            return 1, False                        # bit select
        if k == "part":
            return self.cval(e[2], env) - self.cval(e[3], env) + 1, False
        if k == "cat":
            return sum(self.typ(x, env)[0] for x in e[1]), False
        if k == "rep":
            return self.cval(e[1], env) * self.typ(e[2], env)[0], False
        if k == "signed":
            return self.typ(e[1], env)[0], True
        if k == "un":
            if e[1] in ("!", "&", "|", "^"):
                return 1, False
            return self.typ(e[2], env)
        if k == "bin":
            op = e[1]
            if op in ("==", "!=", "<", "<=", ">", ">=", "&&", "||"):
                return 1, False
            wa, sa = self.typ(e[2], env)
            if op in ("<<", ">>", ">>>", "<<<"):
                return wa, sa
            wb, sb = self.typ(e[3], env)
            return max(wa, wb), sa and sb
        if k == "?:":
            wa, sa = self.typ(e[2], env)
            wb, sb = self.typ(e[3], env)
            return max(wa, wb), sa and sb
        raise ValueError(e)

    # ---- expression compilation: returns f(state) -> unsigned int of W bits ------------------------
    def c_expr(self, e, W, S, env):
        k = e[0]
        ext = (lambda v, w: _sext(v, w) & _mask(W)) if S else (lambda v, w: v & _mask(w) & _mask(W))

        def leaf(fn, w):          # self-determined value of width w, extended to the context
            if S:
                return lambda st: _sext(fn(st), w) & _mask(W)
            mw = _mask(min(w, W))
            return lambda st: fn(st) & mw

        if k == "num":
            v, w, _ = parse_number(e[1])
            val = ext(v, w)
            return lambda st: val
        if k == "id":
            name = e[1]
            if name in env or name in self.consts:
                val = (env[name] if name in env else self.consts[name]) & _mask(32)
                val = ext(val, 32)
                return lambda st: val
            w = self.sigs[name].width
            return leaf(lambda st: st[name], w)
        if k in ("idx", "part", "cat", "rep", "signed") or (k == "un" and e[1] in ("!", "&", "|", "^")) or \
                (k == "bin" and e[1] in ("==", "!=", "<", "<=", ">", ">=", "&&", "||")):
            w, _ = self.typ(e, env)
            return leaf(self.c_self(e, env), w)
        if k == "un":
            a = self.c_expr(e[2], W, S, env)
            m = _mask(W)
            if e[1] == "-":
                return lambda st: (-a(st)) & m
            if e[1] == "~":
                return lambda st: (~a(st)) & m
            return a
        if k == "bin":
            op, m = e[1], _mask(W)
            a = self.c_expr(e[2], W, S, env)
            if op in ("<<", ">>", ">>>", "<<<"):
                wb, sb = self.typ(e[3], env)
                b = self.c_expr(e[3], wb, False, env)       # shift amount: self-determined, unsigned
                if op in ("<<", "<<<"):
                    return lambda st: (a(st) << b(st)) & m
                if op == ">>>" and S:
                    return lambda st: (_sext(a(st), W) >> b(st)) & m
                return lambda st: a(st) >> b(st)
            b = self.c_expr(e[3], W, S, env)
            if op == "+":
                return lambda st: (a(st) + b(st)) & m
            if op == "-":
                return lambda st: (a(st) - b(st)) & m
            if op == "*":
                return lambda st: (a(st) * b(st)) & m        # low W bits of the product are sign-agnostic
            if op == "&":
                return lambda st: a(st) & b(st)
            if op == "|":
                return lambda st: a(st) | b(st)
            if op == "^":
                return lambda st: a(st) ^ b(st)
            raise ValueError("vsim: operator %s" % op)
        if k == "?:":
            c = self.c_self(e[1], env)
            a, b = self.c_expr(e[2], W, S, env), self.c_expr(e[3], W, S, env)
            return lambda st: a(st) if c(st) else b(st)
        raise ValueError(e)

    def c_self(self, e, env):
        """Compile e in its self-determined context; returns f(state) -> unsigned int of its own width."""
        k = e[0]
        w, s = self.typ(e, env)
        if k == "idx":
            base, ie = e[1], e[2]
            if base[0] == "id" and base[1] in self.sigs and self.sigs[base[1]].depth:
                name = base[1]
                if self.is_const(ie, env):
                    i = self.cval(ie, env)
                    return lambda st: st[name][i]
                iw, _ = self.typ(ie, env)
                fi = self.c_expr(ie, iw, False, env)
                depth = self.sigs[name].depth
                return lambda st: (st[name][fi(st)] if fi(st) < depth else 0)
            fb = self.c_self(base, env)
            lsb = self.sigs[base[1]].lsb if base[0] == "id" and base[1] in self.sigs else 0
            if self.is_const(ie, env):
                i = self.cval(ie, env) - lsb
                return lambda st: (fb(st) >> i) & 1
            iw, _ = self.typ(ie, env)
            fi = self.c_expr(ie, iw, False, env)
            return lambda st: (fb(st) >> (fi(st) - lsb)) & 1
        if k == "part":
            fb = self.c_self(e[1], env)
            base = e[1]
            lsb0 = self.sigs[base[1]].lsb if base[0] == "id" and base[1] in self.sigs else 0
            lo = self.cval(e[3], env) - lsb0
            m = _mask(w)
            return lambda st: (fb(st) >> lo) & m
        if k == "cat":
            parts = [(self.c_self(x, env), self.typ(x, env)[0]) for x in e[1]]

            def cat(st):
                v = 0
                for f, pw in parts:
                    v = (v << pw) | (f(st) & _mask(pw))
                return v
            return cat
        if k == "rep":
            n = self.cval(e[1], env)
            f = self.c_self(e[2], env)
            pw = self.typ(e[2], env)[0]

            def rep(st):
                x, v = f(st) & _mask(pw), 0
                for _ in range(n):
                    v = (v << pw) | x
                return v
            return rep
        if k == "signed":
            return self.c_self(e[1], env)
        if k == "un" and e[1] in ("!", "&", "|", "^"):
            f = self.c_self(e[2], env)
            ow = self.typ(e[2], env)[0]
            if e[1] == "!":
                return lambda st: int(f(st) == 0)
            if e[1] == "&":
                full = _mask(ow)
                return lambda st: int(f(st) == full)
            if e[1] == "|":
                return lambda st: int(f(st) != 0)
            return lambda st: bin(f(st)).count("1") & 1
        if k == "bin" and e[1] in ("&&", "||"):
            a, b = self.c_self(e[2], env), self.c_self(e[3], env)
            if e[1] == "&&":
                return lambda st: int(bool(a(st)) and bool(b(st)))
            return lambda st: int(bool(a(st)) or bool(b(st)))
        if k == "bin" and e[1] in ("==", "!=", "<", "<=", ">", ">="):
            (wa, sa), (wb, sb) = self.typ(e[2], env), self.typ(e[3], env)
            cw, cs = max(wa, wb), sa and sb
            a, b = self.c_expr(e[2], cw, cs, env), self.c_expr(e[3], cw, cs, env)
            conv = (lambda v: _sext(v, cw)) if cs else (lambda v: v)
            op = e[1]
            return {"==": lambda st: int(a(st) == b(st)), "!=": lambda st: int(a(st) != b(st)),
                    "<": lambda st: int(conv(a(st)) < conv(b(st))), "<=": lambda st: int(conv(a(st)) <= conv(b(st))),
                    ">": lambda st: int(conv(a(st)) > conv(b(st))), ">=": lambda st: int(conv(a(st)) >= conv(b(st)))}[op]
        return self.c_expr(e, w, s, env)

    # ---- statements ------------------------------------------------------------------------------
    def lhs_width(self, l, env):
        if l[0] == "cat":
            return sum(self.lhs_width(x, env) for x in l[1])
        return self.typ(l, env)[0]

    def c_store(self, l, env):
        """-> store(state_read, updates, value): records the write into `updates` (list of closures)."""
        k = l[0]
        if k == "cat":
            parts = [(self.c_store(x, env), self.lhs_width(x, env)) for x in l[1]]
            total = sum(w for _, w in parts)

            def store(st, upd, v):
                sh = total
                for f, w in parts:
                    sh -= w
                    f(st, upd, (v >> sh) & _mask(w))
            return store
        if k == "id":
            name, m = l[1], _mask(self.sigs[l[1]].width)
            return lambda st, upd, v: upd.append((name, None, m, 0, v & m))
        if k == "idx":
            base, ie = l[1], l[2]
            name = base[1]
            sig = self.sigs[name]
            if sig.depth:                    # array element
                if self.is_const(ie, env):
                    i = self.cval(ie, env)
                    m = _mask(sig.width)
                    return lambda st, upd, v: upd.append((name, i, m, 0, v & m))
                fi = self.c_expr(ie, self.typ(ie, env)[0], False, env)
                m = _mask(sig.width)
                return lambda st, upd, v: upd.append((name, fi(st), m, 0, v & m))
            i = self.cval(ie, env) - sig.lsb   # bit of a vector
            return lambda st, upd, v: upd.append((name, None, 1, i, v & 1))
        if k == "part":
            base = l[1]
            hi, lo = self.cval(l[2], env), self.cval(l[3], env)
            m = _mask(hi - lo + 1)
            if base[0] == "id":
                name = base[1]
                lo -= self.sigs[name].lsb
                return lambda st, upd, v: upd.append((name, None, m, lo, v & m))
            name = base[1][1]                  # part-select of an array element
            i = self.cval(base[2], env)
            return lambda st, upd, v: upd.append((name, i, m, lo, v & m))
        raise ValueError(l)

    def c_assign(self, st_, env):
        _, lhs, rhs = st_
        lw = self.lhs_width(lhs, env)
        rw, rs = self.typ(rhs, env)
        W = max(lw, rw)
        f = self.c_expr(rhs, W, rs, env)
        store = self.c_store(lhs, env)
        m = _mask(lw)
        return lambda st, upd: store(st, upd, f(st) & m)

    def c_stmt(self, s, env):
        k = s[0]
        if k == "block":
            parts = [self.c_stmt(x, env) for x in s[1]]

            def block(st, upd):
                for p in parts:
                    p(st, upd)
            return block
        if k == "if":
            if self.is_const(s[1], env):         # generate-time constant conditions (e.g. i >= WW)
                if self.cval(s[1], env):
                    return self.c_stmt(s[2], env)
                return self.c_stmt(s[3], env) if s[3] else (lambda st, upd: None)
            c = self.c_self(s[1], env)
            t = self.c_stmt(s[2], env)
            e = self.c_stmt(s[3], env) if s[3] else None

            def if_(st, upd):
                if c(st):
                    t(st, upd)
                elif e:
                    e(st, upd)
            return if_
        if k == "case":
            sw, ss = self.typ(s[1], env)
            arms = []
            for labels, body in s[2]:
                lw = max([sw] + [self.typ(x, env)[0] for x in labels])
                arms.append(([(self.c_expr(x, lw, False, env)) for x in labels], lw, self.c_stmt(body, env)))
            default = self.c_stmt(s[3], env) if s[3] else None
            sels = {lw: self.c_expr(s[1], lw, False, env) for _, lw, _ in arms}

            def case(st, upd):
                for labels, lw, body in arms:
                    v = sels[lw](st)
                    if any(l(st) == v for l in labels):
                        body(st, upd)
                        return
                if default:
                    default(st, upd)
            return case
        if k in ("nba", "ba"):
            return self.c_assign(s, env)
        if k == "task":
            return lambda st, upd: None
        raise ValueError(s)

    # ---- simulation ------------------------------------------------------------------------------
    def _commit(self, upd):
        changed = False
        st = self.state
        for name, idx, m, sh, v in upd:
            if idx is None:
                old = st[name]
                new = (old & ~(m << sh)) | (v << sh)
                if new != old:
                    st[name] = new
                    changed = True
            else:
                arr = st[name]
                if 0 <= idx < len(arr):
                    old = arr[idx]
                    new = (old & ~(m << sh)) | (v << sh)
                    if new != old:
                        arr[idx] = new
                        changed = True
        return changed

    def settle(self):
        for _ in range(8):
            upd = []
            for c in self.comb:
                c(self.state, upd)
            if not self._commit(upd):
                return
        raise RuntimeError("vsim: combinational logic did not settle")

    def set(self, **kw):
        for k, v in kw.items():
            self.state[k] = v & _mask(self.sigs[k].width)
        self.settle()

    def get(self, name):
        return self.state[name]

    def gets(self, name):
        return _sext(self.state[name], self.sigs[name].width)

    def tick(self):
        """One rising edge of the clock: all non-blocking assignments read the pre-edge state."""
        upd = []
        for s in self.seq:
            s(self.state, upd)
        self._commit(upd)
        self.settle()


def run_pipeline(mod, inputs, outputs, ce="i_ce", aux_in="i_aux", aux_out="o_aux", reset="i_reset", max_flush=256):
    """Feeds one dict of port values per clock with i_ce = i_aux = 1 and collects `outputs` (unsigned port words)
    whenever o_aux is high, exactly as bench/cpp/cordic_tb.cpp:127-200 does; then flushes the pipe."""
    names = set(mod.sigs)
    if reset in names:
        mod.set(**{reset: 1, ce: 1})
        mod.tick()
        mod.set(**{reset: 0})
    mod.set(**{ce: 1})
    got = []
    for vec in inputs:
        mod.set(**vec, **{aux_in: 1})
        mod.tick()
        if mod.get(aux_out):
            got.append(tuple(mod.get(o) for o in outputs))
    mod.set(**{aux_in: 0})
    for _ in range(max_flush):
        if not mod.get(aux_out):
            break
        mod.tick()
        if mod.get(aux_out):
            got.append(tuple(mod.get(o) for o in outputs))
    return got


def run_handshake(mod, inputs, outputs, clocks_per_output, reset="i_reset", aux_in="i_aux", aux_out="o_aux"):
    """Drives a sequential core the way bench/cpp/cordic_tb.cpp:146-158 does when CLOCKS_PER_OUTPUT is defined:
    i_stb for one clock, CLOCKS_PER_OUTPUT ticks per sample, o_done low until the last tick and high on it.
    Returns the `outputs` port words per sample, or None as soon as a sample breaks that protocol (the core never
    finishes, or finishes early)."""
    names = set(mod.sigs)
    if reset in names:
        mod.set(**{reset: 1, "i_stb": 0})
        mod.tick()
        mod.set(**{reset: 0})
    got = []
    for vec in inputs:
        kw = dict(vec)
        kw["i_stb"] = 1
        if aux_in in names:
            kw[aux_in] = 1
        mod.set(**kw)
        for _ in range(clocks_per_output - 1):
            mod.tick()
            mod.set(i_stb=0)
            if mod.get("o_done"):
                return None
        mod.tick()
        if not mod.get("o_done") or (aux_out in names and not mod.get(aux_out)):
            return None
        got.append(tuple(mod.get(o) for o in outputs))
    return got
