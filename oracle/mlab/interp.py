"""A small MATLAB-subset interpreter.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Its single purpose is to
EXECUTE THE REFERENCE'S OWN, UNMODIFIED .m FILES inside this image (no MATLAB or
Octave exists here) so that the numpy oracle can be pinned against outputs of
the reference itself.  `tests/golden/make_golden.py` drives it; nothing in the
product imports it.

Supported: function files (sub-functions, varargin, nargin, multiple outputs,
`~` placeholders), scripts, if/elseif/else, for, while, switch/case (strings),
break/return, anonymous functions and handles with value capture, structs with
dynamic fields, cell indexing with {}, 1-based column-major indexing with
`end`/`:`/ranges and auto-growing assignment, matrix literals, the operator set
+ - * / ^ .* ./ .^ ' .' == ~= < <= > >= & | && || ~, and the builtins the hot
path uses (fft2/ifft2/conv2/norm/sum/mean/...).  Everything numeric is a 2-D
numpy array (float64 / complex128); a scalar is 1x1.
"""
import os
import re
import time

import numpy as np

# ---------------------------------------------------------------------------
# tokenizer
# ---------------------------------------------------------------------------
KEYWORDS = {"function", "if", "elseif", "else", "end", "for", "while", "switch", "case", "otherwise",
            "break", "return", "global", "continue", "try", "catch"}
TOKEN_RE = re.compile(r"""
    (?P<num>(\d+\.?\d*|\.\d+)([eE][+-]?\d+)?) |
    (?P<id>[A-Za-z_]\w*) |
    (?P<op>\.\*|\./|\.\^|\.'|==|~=|<=|>=|&&|\|\||[-+*/\\^<>=&|~:,;()\[\]{}@.'])
""", re.X)


class Tok:
    __slots__ = ("kind", "val", "sp", "line")

    def __init__(self, kind, val, sp, line):
        self.kind, self.val, self.sp, self.line = kind, val, sp, line

    def __repr__(self):
        return f"{self.kind}:{self.val!r}"


def tokenize(src):
    toks = []
    i, n, line = 0, len(src), 1
    sp = False
    depth = 0                       # [] / {} nesting (newlines are row separators there)
    while i < n:
        ch = src[i]
        if ch in " \t\r":
            sp = True; i += 1; continue
        if src.startswith("...", i):            # continuation: skip to end of line
            j = src.find("\n", i)
            i = n if j < 0 else j + 1
            line += 1; sp = True
            continue
        if ch == "%":
            j = src.find("\n", i)
            i = n if j < 0 else j
            continue
        if ch == "\n":
            toks.append(Tok("nl", "\n", sp, line)); line += 1; i += 1; sp = False
            continue
        if ch == "'":
            prev = toks[-1] if toks else None
            is_transpose = (prev is not None and not sp and
                            (prev.kind in ("num", "id", "str") or prev.val in (")", "]", "}", "'", ".'")
                             or (prev.kind == "kw" and prev.val == "end")))
            if prev is not None and sp and depth == 0 and (prev.kind in ("num", "id") or prev.val in (")", "]", "}")):
                is_transpose = True             # "a '" outside brackets
            if not is_transpose:
                j = i + 1; buf = []
                while j < n:
                    if src[j] == "'":
                        if j + 1 < n and src[j + 1] == "'":
                            buf.append("'"); j += 2; continue
                        break
                    buf.append(src[j]); j += 1
                toks.append(Tok("str", "".join(buf), sp, line)); i = j + 1; sp = False
                continue
        m = TOKEN_RE.match(src, i)
        if not m:
            raise SyntaxError(f"line {line}: cannot tokenize {src[i:i+20]!r}")
        if m.group("num"):
            txt = m.group("num")
            # "1./x", "2.*x", "3.^x": the dot belongs to the operator, not to the number
            if txt.endswith(".") and m.end() < n and src[m.end()] in "*/^'":
                txt = txt[:-1]
                toks.append(Tok("num", float(txt), sp, line))
                i = m.end() - 1; sp = False
                continue
            toks.append(Tok("num", float(txt), sp, line))
        elif m.group("id"):
            v = m.group("id")
            toks.append(Tok("kw" if v in KEYWORDS else "id", v, sp, line))
        else:
            v = m.group("op")
            if v in "[{":
                depth += 1
            elif v in "]}":
                depth -= 1
            toks.append(Tok("op", v, sp, line))
        i = m.end(); sp = False
    toks.append(Tok("nl", "\n", False, line))
    toks.append(Tok("eof", None, False, line))
    return toks


# ---------------------------------------------------------------------------
# parser -> tuples
# ---------------------------------------------------------------------------
class Parser:
    def __init__(self, toks):
        self.t = toks
        self.p = 0
        self.in_matrix = 0
        self.in_index = 0

    def peek(self, k=0):
        return self.t[self.p + k]

    def next(self):
        tk = self.t[self.p]; self.p += 1
        return tk

    def at(self, val):
        tk = self.peek()
        return tk.kind in ("op", "kw") and tk.val == val

    def expect(self, val):
        tk = self.next()
        if not (tk.kind in ("op", "kw") and tk.val == val):
            raise SyntaxError(f"line {tk.line}: expected {val!r}, got {tk}")
        return tk

    def skip_nl(self):
        while self.peek().kind == "nl" or self.at(";") or self.at(","):
            self.next()

    # ---- file level
    def parse_file(self):
        self.skip_nl()
        funcs, stmts = [], []
        if self.at("function"):
            while self.at("function"):
                funcs.append(self.parse_function())
                self.skip_nl()
            return ("file", funcs, None)
        stmts = self.parse_block(("eof",))
        return ("file", [], stmts)

    def parse_function(self):
        self.expect("function")
        outs = []
        # forms: function name(args) | function out = name(args) | function [o1,o2] = name(args)
        if self.at("["):
            self.next()
            while not self.at("]"):
                if self.at(","):
                    self.next(); continue
                outs.append(self.next().val)
            self.expect("]"); self.expect("=")
            name = self.next().val
        else:
            first = self.next().val
            if self.at("="):
                self.next(); outs = [first]; name = self.next().val
            else:
                name = first
        args = []
        if self.at("("):
            self.next()
            while not self.at(")"):
                if self.at(","):
                    self.next(); continue
                tk = self.next()
                args.append("~" if tk.val == "~" else tk.val)
            self.expect(")")
        body = self.parse_block(("function", "eof"), allow_end=True)
        return ("function", name, args, outs, body)

    def parse_block(self, stops, allow_end=False):
        stmts = []
        while True:
            self.skip_nl()
            tk = self.peek()
            if tk.kind == "eof":
                if "eof" in stops:
                    return stmts
                raise SyntaxError("unexpected end of file")
            if tk.kind == "kw" and tk.val in stops:
                return stmts
            if tk.kind == "kw" and tk.val == "end":
                if allow_end and "function" in stops:      # optional `end` closing a function
                    self.next()
                    return stmts
                if "end" in stops:
                    return stmts
                raise SyntaxError(f"line {tk.line}: unexpected end")
            stmts.append(self.parse_statement())

    def parse_statement(self):
        tk = self.peek()
        if tk.kind == "kw":
            if tk.val == "if":
                return self.parse_if()
            if tk.val == "for":
                self.next()
                paren = self.at("(")
                if paren:
                    self.next()
                var = self.next().val
                self.expect("=")
                e = self.parse_expr()
                if paren:
                    self.expect(")")
                body = self.parse_block(("end",)); self.expect("end")
                return ("for", var, e, body)
            if tk.val == "while":
                self.next()
                e = self.parse_expr()
                body = self.parse_block(("end",)); self.expect("end")
                return ("while", e, body)
            if tk.val == "switch":
                self.next()
                e = self.parse_expr()
                self.skip_nl()
                cases, default = [], None
                while self.at("case") or self.at("otherwise"):
                    if self.next().val == "case":
                        ce = self.parse_expr()
                        body = self.parse_block(("case", "otherwise", "end"))
                        cases.append((ce, body))
                    else:
                        default = self.parse_block(("case", "otherwise", "end"))
                self.expect("end")
                return ("switch", e, cases, default)
            if tk.val in ("break", "return", "continue"):
                self.next()
                return (tk.val,)
            if tk.val == "try":
                self.next()
                body = self.parse_block(("catch", "end"))
                handler = []
                if self.at("catch"):
                    self.next()
                    if self.peek().kind == "id" and self.peek(1).kind == "nl":
                        self.next()                 # catch ME
                    handler = self.parse_block(("end",))
                self.expect("end")
                return ("try", body, handler)
            if tk.val == "global":
                self.next()
                names = []
                while self.peek().kind == "id":
                    names.append(self.next().val)
                return ("global", names)
            raise SyntaxError(f"line {tk.line}: unexpected keyword {tk.val}")
        # multi-assignment  [a, b] = ...
        if self.at("["):
            save = self.p
            try:
                lhs = self.try_parse_multi_lhs()
            except SyntaxError:
                lhs = None
            if lhs is not None and self.at("="):
                self.next()
                rhs = self.parse_expr()
                return ("massign", lhs, rhs, self.end_stmt())
            self.p = save
        e = self.parse_expr()
        if self.at("=") :
            self.next()
            rhs = self.parse_expr()
            return ("assign", e, rhs, self.end_stmt())
        return ("expr", e, self.end_stmt())

    def end_stmt(self):
        """consume the statement terminator; returns True if output is suppressed"""
        tk = self.peek()
        if self.at(";"):
            self.next(); return True
        if self.at(",") or tk.kind == "nl":
            self.next(); return False
        return False

    def try_parse_multi_lhs(self):
        self.expect("[")
        items = []
        self.in_matrix += 1
        try:
            while not self.at("]"):
                if self.at(","):
                    self.next(); continue
                if self.at("~"):
                    self.next(); items.append(None); continue
                items.append(self.parse_postfix())
        finally:
            self.in_matrix -= 1
        self.expect("]")
        return items

    def parse_if(self):
        self.expect("if")
        branches = []
        cond = self.parse_expr()
        body = self.parse_block(("elseif", "else", "end"))
        branches.append((cond, body))
        other = None
        while True:
            if self.at("elseif"):
                self.next()
                cond = self.parse_expr()
                body = self.parse_block(("elseif", "else", "end"))
                branches.append((cond, body))
            elif self.at("else"):
                self.next()
                other = self.parse_block(("end",))
            else:
                break
        self.expect("end")
        return ("if", branches, other)

    # ---- expressions
    def parse_expr(self):
        return self.parse_oror()

    def _binary(self, sub, ops):
        left = sub()
        while True:
            tk = self.peek()
            if tk.kind == "op" and tk.val in ops:
                if self.in_matrix and tk.val in ("+", "-") and tk.sp and not self.peek(1).sp:
                    break           # "[a -b]": the sign starts a new element
                self.next()
                right = sub()
                left = ("bin", tk.val, left, right)
            else:
                return left
        return left

    def parse_oror(self):
        return self._binary(self.parse_andand, ("||",))

    def parse_andand(self):
        return self._binary(self.parse_or, ("&&",))

    def parse_or(self):
        return self._binary(self.parse_and, ("|",))

    def parse_and(self):
        return self._binary(self.parse_cmp, ("&",))

    def parse_cmp(self):
        return self._binary(self.parse_range, ("==", "~=", "<", "<=", ">", ">="))

    def parse_range(self):
        first = self.parse_add()
        if self.at(":") and not self._colon_is_bare():
            self.next()
            second = self.parse_add()
            if self.at(":") and not self._colon_is_bare():
                self.next()
                third = self.parse_add()
                return ("range", first, second, third)
            return ("range", first, None, second)
        return first

    def _colon_is_bare(self):
        nxt = self.peek(1)
        return nxt.kind == "op" and nxt.val in (")", ",")

    def parse_add(self):
        return self._binary(self.parse_mul, ("+", "-"))

    def parse_mul(self):
        return self._binary(self.parse_unary, ("*", "/", "\\", ".*", "./"))

    def parse_unary(self):
        tk = self.peek()
        if tk.kind == "op" and tk.val in ("-", "+", "~"):
            self.next()
            operand = self.parse_unary()
            return ("un", tk.val, operand)
        return self.parse_power()

    def parse_power(self):
        base = self.parse_postfix()
        while True:
            tk = self.peek()
            if tk.kind == "op" and tk.val in ("^", ".^"):
                self.next()
                # exponent may carry a unary sign:  2^-1
                if self.peek().kind == "op" and self.peek().val in ("-", "+"):
                    sign = self.next().val
                    ex = ("un", sign, self.parse_postfix())
                else:
                    ex = self.parse_postfix()
                base = ("bin", tk.val, base, ex)
            else:
                return base

    def parse_postfix(self):
        e = self.parse_primary()
        while True:
            tk = self.peek()
            if tk.kind != "op":
                return e
            if tk.val == "(" and not (self.in_matrix and tk.sp):
                self.next()
                args = self.parse_args(")")
                e = ("call", e, args)
            elif tk.val == "{" and not (self.in_matrix and tk.sp):
                self.next()
                args = self.parse_args("}")
                e = ("cell", e, args)
            elif tk.val == "." and self.peek(1).kind == "id" and not tk.sp:
                self.next()
                e = ("field", e, self.next().val)
            elif tk.val == "." and self.peek(1).kind == "op" and self.peek(1).val == "(" and not tk.sp:
                self.next(); self.next()
                saved = self.in_matrix
                self.in_matrix = 0
                name = self.parse_expr()
                self.in_matrix = saved
                self.expect(")")
                e = ("dynfield", e, name)
            elif tk.val in ("'", ".'"):
                self.next()
                e = ("un", tk.val, e)
            else:
                return e

    def parse_args(self, close):
        args = []
        self.in_index += 1
        saved = self.in_matrix
        self.in_matrix = 0
        try:
            while self.peek().kind == "nl":
                self.next()
            while not self.at(close):
                if self.at(","):
                    self.next(); continue
                if self.at(":") and self._colon_is_bare():
                    self.next(); args.append(("colon",)); continue
                args.append(self.parse_expr())
                while self.peek().kind == "nl":
                    self.next()
        finally:
            self.in_index -= 1
            self.in_matrix = saved
        self.expect(close)
        return args

    def parse_primary(self):
        tk = self.next()
        if tk.kind == "num":
            return ("num", tk.val)
        if tk.kind == "str":
            return ("str", tk.val)
        if tk.kind == "id":
            return ("id", tk.val)
        if tk.kind == "kw" and tk.val == "end" and self.in_index:
            return ("end",)
        if tk.kind == "op":
            if tk.val == "(":
                saved = self.in_matrix
                self.in_matrix = 0
                e = self.parse_expr()
                self.in_matrix = saved
                self.expect(")")
                return ("paren", e)
            if tk.val == "[":
                return self.parse_matrix()
            if tk.val == "{":
                items = []
                self.in_matrix += 1
                saved_idx = self.in_index
                self.in_index = 0
                try:
                    while not self.at("}"):
                        if self.at(",") or self.at(";") or self.peek().kind == "nl":
                            self.next(); continue
                        items.append(self.parse_expr())
                finally:
                    self.in_matrix -= 1
                    self.in_index = saved_idx
                self.expect("}")
                return ("cellarr", items)
            if tk.val == "@":
                if self.at("("):
                    self.next()
                    params = []
                    while not self.at(")"):
                        if self.at(","):
                            self.next(); continue
                        params.append(self.next().val)
                    self.expect(")")
                    saved = (self.in_matrix, self.in_index)
                    self.in_matrix = 0; self.in_index = 0
                    body = self.parse_expr()
                    self.in_matrix, self.in_index = saved
                    return ("anon", params, body)
                return ("handle", self.next().val)
            if tk.val == ":":
                return ("colon",)
        raise SyntaxError(f"line {tk.line}: unexpected token {tk}")

    def parse_matrix(self):
        rows, cur = [], []
        self.in_matrix += 1
        saved_idx = self.in_index
        self.in_index = 0
        try:
            while True:
                tk = self.peek()
                if tk.kind == "op" and tk.val == "]":
                    self.next(); break
                if (tk.kind == "op" and tk.val == ";") or tk.kind == "nl":
                    self.next()
                    if cur:
                        rows.append(cur); cur = []
                    continue
                if tk.kind == "op" and tk.val == ",":
                    self.next(); continue
                cur.append(self.parse_expr())
        finally:
            self.in_matrix -= 1
            self.in_index = saved_idx
        if cur:
            rows.append(cur)
        return ("matrix", rows)


# ---------------------------------------------------------------------------
# values
# ---------------------------------------------------------------------------
class MStruct(dict):
    pass


class MCell(list):
    pass


class MFunc:
    def __init__(self, fn, name="anon"):
        self.fn, self.name = fn, name

    def __call__(self, *a, **k):
        return self.fn(*a, **k)


def M(x):
    """to a 2-D numpy array"""
    if isinstance(x, np.ndarray):
        if x.ndim == 2:
            return x
        if x.ndim == 0:
            return x.reshape(1, 1)
        if x.ndim == 1:
            return x.reshape(1, -1)
    if isinstance(x, (bool, np.bool_)):
        return np.array([[float(x)]])
    if isinstance(x, (int, float, complex, np.number)):
        return np.array([[x]], dtype=np.complex128 if isinstance(x, complex) else np.float64)
    raise TypeError(f"not numeric: {type(x)}")


def scalar(x):
    a = M(x)
    if a.size != 1:
        raise ValueError("expected a scalar")
    v = a.flat[0]
    return v.real if np.iscomplexobj(v) and v.imag == 0 else v


def is_num(x):
    return isinstance(x, (np.ndarray, int, float, complex, np.number, bool, np.bool_))


def truth(x):
    if isinstance(x, str):
        return len(x) > 0
    a = M(x)
    return a.size > 0 and bool(np.all(a != 0))


class BreakEx(Exception):
    pass


class ContinueEx(Exception):
    pass


class ReturnEx(Exception):
    pass


class MatlabError(Exception):
    pass


# ---------------------------------------------------------------------------
# interpreter
# ---------------------------------------------------------------------------
class Interp:
    def __init__(self, path, randn=None, verbose=False):
        self.path = list(path)
        self.files = {}             # name -> (main function, {local functions})
        self.verbose = verbose
        self.randn = randn or (lambda shape: np.random.standard_normal(shape))
        self.tic = time.perf_counter()
        self.globals = {}
        self.builtins = self._make_builtins()

    # ---- loading
    def _find(self, name):
        if name in self.files:
            return self.files[name]
        for d in self.path:
            fn = os.path.join(d, name + ".m")
            if os.path.exists(fn):
                ast = Parser(tokenize(open(fn, encoding="utf-8", errors="replace").read())).parse_file()
                funcs = ast[1]
                if not funcs:
                    self.files[name] = None
                    return None
                local = {f[1]: f for f in funcs}
                self.files[name] = (funcs[0], local)
                return self.files[name]
        self.files[name] = None
        return None

    # ---- running code
    def run_source(self, src, scope):
        ast = Parser(tokenize(src)).parse_file()
        if ast[2] is None:
            raise ValueError("run_source expects script code")
        self.exec_block(ast[2], scope, {})
        return scope

    def call_function(self, name, args, nargout=1, local=None):
        if local and name in local:
            return self._call_user(local[name], local, args, nargout)
        ent = self._find(name)
        if ent is not None:
            return self._call_user(ent[0], ent[1], args, nargout)
        if name in self.builtins:
            return self.builtins[name](args, nargout)
        raise MatlabError(f"Undefined function or variable '{name}'")

    def _call_user(self, f, local, args, nargout):
        _, name, params, outs, body = f
        scope = {}
        np_ = len(params)
        if params and params[-1] == "varargin":
            fixed = params[:-1]
            for p, a in zip(fixed, args):
                scope[p] = a
            scope["varargin"] = MCell(args[len(fixed):])
        else:
            if len(args) > np_:
                raise MatlabError(f"{name}: too many input arguments")
            for p, a in zip(params, args):
                if p != "~":
                    scope[p] = a
        scope["nargin"] = M(float(len(args)))
        scope["nargout"] = M(float(nargout))
        try:
            self.exec_block(body, scope, local)
        except ReturnEx:
            pass
        res = []
        for o in outs[:max(nargout, 1)]:
            if o not in scope:
                if len(res) < nargout:
                    raise MatlabError(f"{name}: output argument '{o}' not assigned")
                break
            res.append(scope[o])
        return res

    def exec_block(self, stmts, scope, local):
        for st in stmts:
            self.exec_stmt(st, scope, local)

    def exec_stmt(self, st, scope, local):
        k = st[0]
        if k == "assign":
            val = self.eval(st[2], scope, local)
            self.assign(st[1], val, scope, local)
        elif k == "massign":
            vals = self.eval_multi(st[2], scope, local, len(st[1]))
            if len(vals) < len([t for t in st[1]]):
                needed = max(i for i, t in enumerate(st[1]) if t is not None) + 1 if any(t is not None for t in st[1]) else 0
                if len(vals) < needed:
                    raise MatlabError("not enough output values")
            for tgt, v in zip(st[1], vals):
                if tgt is not None:
                    self.assign(tgt, v, scope, local)
        elif k == "expr":
            e = st[1]
            # command-style calls we can ignore / handle:  "clear all", "clc", "tic"
            v = self.eval_multi(e, scope, local, 0)
            if v:
                scope["ans"] = v[0]
        elif k == "if":
            for cond, body in st[1]:
                if truth(self.eval(cond, scope, local)):
                    self.exec_block(body, scope, local)
                    return
            if st[2] is not None:
                self.exec_block(st[2], scope, local)
        elif k == "for":
            rng = self.eval(st[2], scope, local)
            a = M(rng) if is_num(rng) else rng
            ncol = a.shape[1] if a.size else 0
            for c in range(ncol):
                col = a[:, c]
                scope[st[1]] = M(col[0]) if col.size == 1 else col.reshape(-1, 1).copy()
                try:
                    self.exec_block(st[3], scope, local)
                except BreakEx:
                    break
                except ContinueEx:
                    continue
        elif k == "while":
            while truth(self.eval(st[1], scope, local)):
                try:
                    self.exec_block(st[2], scope, local)
                except BreakEx:
                    break
                except ContinueEx:
                    continue
        elif k == "switch":
            v = self.eval(st[1], scope, local)
            for ce, body in st[2]:
                cv = self.eval(ce, scope, local)
                if (isinstance(v, str) and isinstance(cv, str) and v == cv) or \
                   (not isinstance(v, str) and not isinstance(cv, str) and scalar(v) == scalar(cv)):
                    self.exec_block(body, scope, local)
                    return
            if st[3] is not None:
                self.exec_block(st[3], scope, local)
        elif k == "try":
            try:
                self.exec_block(st[1], scope, local)
            except MatlabError:
                self.exec_block(st[2], scope, local)
        elif k == "break":
            raise BreakEx()
        elif k == "continue":
            raise ContinueEx()
        elif k == "return":
            raise ReturnEx()
        elif k == "global":
            for n in st[1]:
                scope[n] = self.globals.setdefault(n, np.zeros((0, 0)))
                scope.setdefault("__globals__", set()).add(n)
        else:
            raise MatlabError(f"unknown statement {k}")

    # ---- assignment
    def assign(self, tgt, val, scope, local):
        k = tgt[0]
        if k == "id":
            scope[tgt[1]] = val
            if tgt[1] in scope.get("__globals__", ()):
                self.globals[tgt[1]] = val
        elif k == "field":
            base = self._lvalue_struct(tgt[1], scope, local)
            base[tgt[2]] = val
        elif k == "call":
            # x(idx) = val   or   s.f(idx) = val
            cur = self._get_or_none(tgt[1], scope, local)
            new = self.index_assign(cur, tgt[2], val, scope, local)
            self.assign(tgt[1], new, scope, local)
        elif k == "cell":
            cur = self._get_or_none(tgt[1], scope, local)
            if cur is None:
                cur = MCell()
            i = int(scalar(self.eval(tgt[2][0], scope, local))) - 1
            while len(cur) <= i:
                cur.append(M(np.zeros((0, 0))))
            cur[i] = val
            self.assign(tgt[1], cur, scope, local)
        else:
            raise MatlabError(f"cannot assign to {k}")

    def _lvalue_struct(self, e, scope, local):
        if e[0] == "id":
            v = scope.get(e[1])
            if v is None:
                v = MStruct(); scope[e[1]] = v
            if not isinstance(v, MStruct):
                raise MatlabError(f"{e[1]} is not a struct")
            return v
        if e[0] == "field":
            parent = self._lvalue_struct(e[1], scope, local)
            v = parent.get(e[2])
            if v is None:
                v = MStruct(); parent[e[2]] = v
            return v
        raise MatlabError("unsupported struct lvalue")

    def _get_or_none(self, e, scope, local):
        if e[0] == "id":
            return scope.get(e[1])
        if e[0] == "field":
            try:
                base = self._lvalue_struct(e[1], scope, local)
            except MatlabError:
                return None
            return base.get(e[2])
        raise MatlabError("unsupported indexed lvalue")

    def _resolve_index(self, ie, dimlen, scope, local):
        """-> numpy int array (0-based) or slice(None) for ':'"""
        if ie[0] == "colon":
            return None
        v = self.eval(ie, scope, local, end_val=dimlen)
        a = M(v)
        if a.dtype == bool:
            return np.flatnonzero(a.ravel(order="F"))
        idx = np.real(a).ravel(order="F")
        ii = np.rint(idx).astype(np.int64)
        if np.any(np.abs(idx - ii) > 0) or np.any(ii < 1):
            raise MatlabError("Subscript indices must either be real positive integers or logicals.")
        return ii - 1

    def index_assign(self, cur, idx_exprs, val, scope, local):
        val_a = M(val) if is_num(val) else None
        if val_a is None:
            raise MatlabError("indexed assignment of non-numeric values is not supported")
        if cur is None:
            cur = np.zeros((0, 0))
        cur = M(cur)
        if np.iscomplexobj(val_a) and not np.iscomplexobj(cur):
            cur = cur.astype(np.complex128)
        nidx = len(idx_exprs)
        if nidx == 1:
            n = cur.size
            ii = self._resolve_index(idx_exprs[0], n, scope, local)
            if ii is None:
                ii = np.arange(n)
            need = int(ii.max()) + 1 if ii.size else 0
            if need > n:
                if cur.size == 0:
                    cur = np.zeros((1, need), dtype=cur.dtype)
                elif cur.shape[0] == 1:
                    cur = np.concatenate([cur, np.zeros((1, need - n), dtype=cur.dtype)], axis=1)
                elif cur.shape[1] == 1:
                    cur = np.concatenate([cur, np.zeros((need - n, 1), dtype=cur.dtype)], axis=0)
                else:
                    raise MatlabError("cannot grow a matrix with a linear index")
            else:
                cur = cur.copy()
            flat = cur.reshape(-1, order="F")
            flat[ii] = val_a.ravel(order="F") if val_a.size > 1 else val_a.flat[0]
            return flat.reshape(cur.shape, order="F")
        if nidx == 2:
            r = self._resolve_index(idx_exprs[0], cur.shape[0], scope, local)
            c = self._resolve_index(idx_exprs[1], cur.shape[1], scope, local)
            if r is None:
                r = np.arange(cur.shape[0] if cur.shape[0] else val_a.shape[0])
            if c is None:
                c = np.arange(cur.shape[1] if cur.shape[1] else val_a.shape[1])
            nr = max(cur.shape[0], int(r.max()) + 1 if r.size else 0)
            nc = max(cur.shape[1], int(c.max()) + 1 if c.size else 0)
            if (nr, nc) != cur.shape:
                big = np.zeros((nr, nc), dtype=cur.dtype)
                big[:cur.shape[0], :cur.shape[1]] = cur
                cur = big
            else:
                cur = cur.copy()
            if val_a.size == 1:
                cur[np.ix_(r, c)] = val_a.flat[0]
            else:
                cur[np.ix_(r, c)] = val_a.reshape(len(r), len(c), order="F") if val_a.shape != (len(r), len(c)) else val_a
            return cur
        raise MatlabError("only 1-D and 2-D indexing is supported")

    def index_read(self, a, idx_exprs, scope, local):
        a = M(a)
        nidx = len(idx_exprs)
        if nidx == 0:
            return a
        if nidx == 1:
            ie = idx_exprs[0]
            if ie[0] == "colon":
                return a.reshape(-1, 1, order="F").copy()
            v = self.eval(ie, scope, local, end_val=a.size)
            va = M(v)
            if va.dtype == bool:
                ii = np.flatnonzero(va.ravel(order="F"))
                shape = (1, ii.size) if a.shape[0] == 1 else (ii.size, 1)
            else:
                idx = np.real(va)
                ii = np.rint(idx).astype(np.int64).ravel(order="F") - 1
                if np.any(ii < 0) or np.any(np.abs(idx.ravel(order="F") - (ii + 1)) > 0):
                    raise MatlabError("Subscript indices must either be real positive integers or logicals.")
                if np.any(ii >= a.size):
                    raise MatlabError("Index exceeds matrix dimensions.")
                if min(a.shape) == 1 and min(va.shape) == 1 and a.size > 0:
                    shape = (1, ii.size) if a.shape[0] == 1 else (ii.size, 1)      # orientation of the source vector
                else:
                    shape = va.shape
            out = a.reshape(-1, order="F")[ii]
            return out.reshape(shape, order="F")
        if nidx == 2:
            r = self._resolve_index(idx_exprs[0], a.shape[0], scope, local)
            c = self._resolve_index(idx_exprs[1], a.shape[1], scope, local)
            if r is None:
                r = np.arange(a.shape[0])
            if c is None:
                c = np.arange(a.shape[1])
            if (r.size and r.max() >= a.shape[0]) or (c.size and c.max() >= a.shape[1]):
                raise MatlabError("Index exceeds matrix dimensions.")
            return a[np.ix_(r, c)].copy()
        raise MatlabError("only 1-D and 2-D indexing is supported")

    # ---- evaluation
    def eval(self, e, scope, local, end_val=None):
        r = self.eval_multi(e, scope, local, 1, end_val)
        if not r:
            raise MatlabError("expression produced no value")
        return r[0]

    def eval_multi(self, e, scope, local, nargout, end_val=None):
        k = e[0]
        if k == "num":
            return [M(e[1])]
        if k == "str":
            return [e[1]]
        if k == "paren":
            return [self.eval(e[1], scope, local, end_val)]
        if k == "end":
            if end_val is None:
                raise MatlabError("'end' outside of an index")
            return [M(float(end_val))]
        if k == "colon":
            return [":"]
        if k == "id":
            name = e[1]
            if name in scope:
                if name in scope.get("__globals__", ()):
                    return [self.globals[name]]
                return [scope[name]]
            return self.call_function(name, [], nargout, local)
        if k == "field":
            base = self.eval(e[1], scope, local, end_val)
            if not isinstance(base, MStruct):
                raise MatlabError("field access on a non-struct")
            if e[2] not in base:
                raise MatlabError(f"Reference to non-existent field '{e[2]}'.")
            return [base[e[2]]]
        if k == "cell":
            base = self.eval(e[1], scope, local, end_val)
            i = int(scalar(self.eval(e[2][0], scope, local, end_val=len(base)))) - 1
            return [base[i]]
        if k == "call":
            return self.eval_call(e, scope, local, nargout, end_val)
        if k == "anon":
            params, body = e[1], e[2]
            captured = dict(scope)          # MATLAB captures variable VALUES at creation time

            def fn(args, nargout=1, _p=params, _b=body, _c=captured, _l=local):
                sc = dict(_c)
                for p, a in zip(_p, args):
                    sc[p] = a
                return self.eval_multi(_b, sc, _l, nargout)
            return [MFunc(fn)]
        if k == "handle":
            name = e[1]
            return [MFunc(lambda args, nargout=1, _n=name, _l=local: self.call_function(_n, args, nargout, _l), name)]
        if k == "matrix":
            return [self.build_matrix(e[1], scope, local, end_val)]
        if k == "emptycell":
            return [MCell()]
        if k == "cellarr":
            return [MCell([self.eval(x, scope, local, end_val) for x in e[1]])]
        if k == "dynfield":
            base = self.eval(e[1], scope, local, end_val)
            name = self.eval(e[2], scope, local, end_val)
            if not isinstance(base, MStruct) or name not in base:
                raise MatlabError(f"Reference to non-existent field '{name}'.")
            return [base[name]]
        if k == "range":
            a = scalar(self.eval(e[1], scope, local, end_val))
            b = scalar(self.eval(e[3], scope, local, end_val))
            s = 1.0 if e[2] is None else scalar(self.eval(e[2], scope, local, end_val))
            if s == 0 or (s > 0 and a > b) or (s < 0 and a < b):
                return [np.zeros((1, 0))]
            n = int(np.floor((b - a) / s * (1 + 1e-15) + 1e-10)) + 1
            return [(a + s * np.arange(n)).reshape(1, -1).astype(np.float64)]
        if k == "un":
            v = self.eval(e[2], scope, local, end_val)
            op = e[1]
            if op == "-":
                return [-M(v)]
            if op == "+":
                return [M(v)]
            if op == "~":
                return [(M(v) == 0).astype(np.float64)]
            if op == "'":
                return [np.conj(M(v)).T.copy()]
            if op == ".'":
                return [M(v).T.copy()]
        if k == "bin":
            op = e[1]
            if op == "&&":
                return [M(float(truth(self.eval(e[2], scope, local, end_val)) and truth(self.eval(e[3], scope, local, end_val))))]
            if op == "||":
                return [M(float(truth(self.eval(e[2], scope, local, end_val)) or truth(self.eval(e[3], scope, local, end_val))))]
            a = self.eval(e[2], scope, local, end_val)
            b = self.eval(e[3], scope, local, end_val)
            return [self.binop(op, a, b)]
        raise MatlabError(f"cannot evaluate {k}")

    def binop(self, op, a, b):
        if isinstance(a, str) or isinstance(b, str):
            if op == "==" and isinstance(a, str) and isinstance(b, str):
                return M(float(a == b))
            raise MatlabError("string arithmetic is not supported")
        a, b = M(a), M(b)
        with np.errstate(all="ignore"):
            if op == "+":
                return a + b
            if op == "-":
                return a - b
            if op == ".*":
                return a * b
            if op == "./":
                return a / b
            if op == ".^":
                return self._power(a, b)
            if op == "*":
                if a.size == 1 or b.size == 1:
                    return a * b
                return a @ b
            if op == "/":
                if b.size == 1:
                    return a / b
                return np.linalg.solve(b.T, a.T).T
            if op == "\\":
                if a.size == 1:
                    return b / a
                return np.linalg.solve(a, b)
            if op == "^":
                if a.size == 1 and b.size == 1:
                    return self._power(a, b)
                if b.size == 1 and float(np.real(b.flat[0])).is_integer():
                    return np.linalg.matrix_power(a, int(np.real(b.flat[0])))
                raise MatlabError("unsupported matrix power")
            if op in ("==", "~=", "<", "<=", ">", ">="):
                ar, br = np.real(a), np.real(b)
                r = {"==": a == b, "~=": a != b, "<": ar < br, "<=": ar <= br, ">": ar > br, ">=": ar >= br}[op]
                return r.astype(np.float64)
            if op == "&":
                return ((a != 0) & (b != 0)).astype(np.float64)
            if op == "|":
                return ((a != 0) | (b != 0)).astype(np.float64)
        raise MatlabError(f"unknown operator {op}")

    @staticmethod
    def _power(a, b):
        if not np.iscomplexobj(a) and not np.iscomplexobj(b):
            if np.any((a < 0) & (np.floor(b) != b)):
                return np.power(a.astype(np.complex128), b)
            return np.power(a, b)
        return np.power(a, b)

    def build_matrix(self, rows, scope, local, end_val):
        if not rows:
            return np.zeros((0, 0))
        out_rows = []
        for row in rows:
            vals = []
            for x in row:
                if x[0] == "cell" and len(x[2]) == 1 and x[2][0][0] == "colon":      # [c{:}]
                    vals.extend(list(self.eval(x[1], scope, local, end_val)))
                else:
                    vals.append(self.eval(x, scope, local, end_val))
            if all(isinstance(v, str) for v in vals):
                out_rows.append("".join(vals)); continue
            mats = [M(v) for v in vals if not (is_num(v) and M(v).size == 0)]
            if not mats:
                continue
            out_rows.append(np.concatenate(mats, axis=1))
        if not out_rows:
            return np.zeros((0, 0))
        if isinstance(out_rows[0], str):
            return out_rows[0]
        return np.concatenate(out_rows, axis=0)

    def eval_call(self, e, scope, local, nargout, end_val):
        target, arg_exprs = e[1], e[2]
        # variable indexing / handle call / struct-field handle call
        if target[0] == "id" and target[1] in scope:
            base = scope[target[1]]
        elif target[0] == "id":
            args = [self.eval(a, scope, local, end_val) for a in arg_exprs]
            if target[1] == "exist":                # exist('name','var') looks at the caller's workspace
                name = args[0]
                if len(args) > 1 and args[1] == "var":
                    return [M(float(name in scope))]
                return [M(1.0 if name in scope else (2.0 if self._find(name) is not None or name in self.builtins else 0.0))]
            return self.call_function(target[1], args, nargout, local)
        else:
            base = self.eval(target, scope, local, end_val)
        if isinstance(base, MFunc):
            args = [self.eval(a, scope, local, end_val) for a in arg_exprs]
            return base(args, nargout)
        if isinstance(base, MCell):
            i = int(scalar(self.eval(arg_exprs[0], scope, local, end_val=len(base)))) - 1
            return [MCell([base[i]])]
        if isinstance(base, str):
            idx = self._resolve_index(arg_exprs[0], len(base), scope, local)
            return ["".join(base[i] for i in (range(len(base)) if idx is None else idx))]
        return [self.index_read(base, arg_exprs, scope, local)]

    # ---- builtins
    def _make_builtins(self):
        B = {}

        def reg(name):
            def deco(f):
                B[name] = f
                return f
            return deco

        def shape_from(args):
            if len(args) == 0:
                return (1, 1)
            if len(args) == 1:
                a = M(args[0])
                if a.size == 1:
                    n = int(scalar(a)); return (n, n)
                v = a.ravel(order="F")
                return (int(v[0]), int(v[1]))
            return (int(scalar(args[0])), int(scalar(args[1])))

        B["zeros"] = lambda a, n: [np.zeros(shape_from(a))]
        B["ones"] = lambda a, n: [np.ones(shape_from(a))]
        B["eye"] = lambda a, n: [np.eye(*shape_from(a))]
        B["randn"] = lambda a, n: [np.asarray(self.randn(shape_from([x for x in a if not isinstance(x, str)])), dtype=np.float64)
                                    if not (a and isinstance(a[0], str)) else M(0.0)]
        B["pi"] = lambda a, n: [M(np.pi)]
        B["NaN"] = lambda a, n: [M(np.nan)]
        B["nan"] = B["NaN"]
        B["Inf"] = lambda a, n: [M(np.inf)]
        B["inf"] = B["Inf"]
        B["eps"] = lambda a, n: [M(np.finfo(float).eps)]
        B["true"] = lambda a, n: [M(1.0)]
        B["false"] = lambda a, n: [M(0.0)]

        def size_(a, n):
            x = a[0]
            shp = (1, len(x)) if isinstance(x, (str, MCell)) else M(x).shape
            if len(a) == 2:
                d = int(scalar(a[1]))
                return [M(float(shp[d - 1] if d <= 2 else 1))]
            if n <= 1:
                return [np.array([[float(shp[0]), float(shp[1])]])]
            return [M(float(shp[0])), M(float(shp[1]))][:n]
        B["size"] = size_
        B["numel"] = lambda a, n: [M(float(len(a[0]) if isinstance(a[0], (str, MCell)) else M(a[0]).size))]

        def length_(a, n):
            x = a[0]
            if isinstance(x, (str, MCell)):
                return [M(float(len(x)))]
            x = M(x)
            return [M(float(0 if x.size == 0 else max(x.shape)))]
        B["length"] = length_
        B["isempty"] = lambda a, n: [M(float((len(a[0]) == 0) if isinstance(a[0], (str, MCell)) else (M(a[0]).size == 0)))]
        B["isfield"] = lambda a, n: [M(float(isinstance(a[0], MStruct) and a[1] in a[0]))]
        B["not"] = lambda a, n: [(M(a[0]) == 0).astype(np.float64)]
        B["double"] = lambda a, n: [M(a[0]).astype(np.float64) if not np.iscomplexobj(M(a[0])) else M(a[0])]
        B["upper"] = lambda a, n: [a[0].upper()]
        B["lower"] = lambda a, n: [a[0].lower()]
        B["strcmp"] = lambda a, n: [M(float(isinstance(a[0], str) and isinstance(a[1], str) and a[0] == a[1]))]
        B["num2str"] = lambda a, n: [repr(scalar(a[0])) if is_num(a[0]) else str(a[0])]
        for nm, f in (("abs", np.abs), ("sqrt", None), ("exp", np.exp), ("log", None), ("log10", None), ("cos", np.cos),
                      ("sin", np.sin), ("tan", np.tan), ("real", np.real), ("imag", np.imag), ("conj", np.conj),
                      ("floor", np.floor), ("ceil", np.ceil), ("sign", np.sign)):
            if f is not None:
                B[nm] = (lambda ff: (lambda a, n: [np.asarray(ff(M(a[0])), dtype=None).copy()]))(f)

        def sqrt_(a, n):
            x = M(a[0])
            if not np.iscomplexobj(x) and np.any(x < 0):
                x = x.astype(np.complex128)
            return [np.sqrt(x)]
        B["sqrt"] = sqrt_

        def log_(a, n):
            x = M(a[0])
            with np.errstate(all="ignore"):
                if not np.iscomplexobj(x) and np.any(x < 0):
                    x = x.astype(np.complex128)
                return [np.log(x)]
        B["log"] = log_
        B["log10"] = lambda a, n: [np.log10(M(a[0]))]
        B["round"] = lambda a, n: [np.sign(M(a[0])) * np.floor(np.abs(M(a[0])) + 0.5)]
        B["mod"] = lambda a, n: [np.mod(M(a[0]), M(a[1]))]

        def reduce_(fn):
            def f(a, n):
                x = M(a[0])
                if len(a) > 1:
                    ax = int(scalar(a[1])) - 1
                elif x.shape[0] == 1:
                    ax = 1
                else:
                    ax = 0
                if x.size == 0 and len(a) == 1:
                    return [M(fn(np.zeros(0)))]
                return [M(fn(x, axis=ax, keepdims=True))]
            return f

        def seq_sum(x, axis=None, keepdims=False):
            """column/row sums accumulated sequentially like a plain loop"""
            if axis is None:
                return np.add.reduce(x.ravel()) if x.size else 0.0
            return np.add.reduce(x, axis=axis, keepdims=keepdims)
        B["sum"] = reduce_(seq_sum)

        def mean_fn(x, axis=None, keepdims=False):
            with np.errstate(all="ignore"):
                if x.size == 0:
                    return np.nan
                return np.mean(x, axis=axis, keepdims=keepdims)
        B["mean"] = reduce_(mean_fn)
        B["prod"] = reduce_(lambda x, axis=None, keepdims=False: np.prod(x, axis=axis, keepdims=keepdims))

        def minmax(fn_pair, fn_red):
            def f(a, n):
                if len(a) >= 2 and not (is_num(a[1]) and M(a[1]).size == 0):
                    return [fn_pair(M(a[0]), M(a[1]))]
                x = M(a[0])
                ax = 1 if x.shape[0] == 1 else 0
                return [M(fn_red(x, axis=ax, keepdims=True))]
            return f
        B["min"] = minmax(np.minimum, np.min)
        B["max"] = minmax(np.maximum, np.max)

        def norm_(a, n):
            x = M(a[0])
            kind = a[1] if len(a) > 1 else 2
            if isinstance(kind, str):
                if kind == "fro":
                    return [M(np.sqrt(np.sum(np.abs(x) ** 2)))]
                if kind == "inf":
                    kind = np.inf
            else:
                kind = scalar(kind)
            if min(x.shape) == 1:
                return [M(np.linalg.norm(x.ravel(), kind))]
            return [M(np.linalg.norm(x, kind))]
        B["norm"] = norm_
        B["fft2"] = lambda a, n: [np.fft.fft2(M(a[0]))]
        B["ifft2"] = lambda a, n: [np.fft.ifft2(M(a[0]))]

        def conv2_(a, n):
            from scipy.signal import convolve2d
            mode = a[2] if len(a) > 2 else "full"
            return [convolve2d(M(a[0]), M(a[1]), mode=mode)]
        B["conv2"] = conv2_

        def ndgrid_(a, n):
            x = M(a[0]).ravel(order="F")
            y = M(a[1]).ravel(order="F") if len(a) > 1 else x
            g1, g2 = np.meshgrid(x, y, indexing="ij")
            return [g1.copy(), g2.copy()][:max(n, 1)]
        B["ndgrid"] = ndgrid_

        def tic_(a, n):
            self.tic = time.perf_counter()
            return []
        B["tic"] = tic_
        B["toc"] = lambda a, n: [M(time.perf_counter() - self.tic)]
        B["cputime"] = lambda a, n: [M(time.process_time())]
        B["fprintf"] = lambda a, n: []
        B["disp"] = lambda a, n: []
        B["clc"] = lambda a, n: []

        def error_(a, n):
            raise MatlabError(a[0] if a else "error")
        B["error"] = error_
        B["isreal"] = lambda a, n: [M(float(not np.iscomplexobj(M(a[0]))))]

        def isa_(a, n):
            kind = a[1]
            v = a[0]
            if kind == "function_handle":
                return [M(float(isinstance(v, MFunc)))]
            if kind in ("double", "numeric", "float"):
                return [M(float(is_num(v)))]
            if kind == "struct":
                return [M(float(isinstance(v, MStruct)))]
            if kind == "char":
                return [M(float(isinstance(v, str)))]
            return [M(0.0)]
        B["isa"] = isa_
        B["xor"] = lambda a, n: [M(float(truth(a[0]) != truth(a[1])))]
        B["rem"] = lambda a, n: [np.fmod(M(a[0]), M(a[1]))]
        B["transpose"] = lambda a, n: [M(a[0]).T.copy()]
        B["inv"] = lambda a, n: [np.linalg.inv(M(a[0]))]
        B["any"] = lambda a, n: [M(float(np.any(M(a[0]) != 0)))]
        B["all"] = lambda a, n: [M(float(np.all(M(a[0]) != 0)))]
        B["svd"] = lambda a, n: [np.linalg.svd(M(a[0]), compute_uv=False).reshape(-1, 1)]
        return B


def to_py(v):
    """MATLAB value -> plain python / numpy for the fixtures"""
    if isinstance(v, MStruct):
        return {k: to_py(x) for k, x in v.items() if not isinstance(x, MFunc)}
    if isinstance(v, np.ndarray):
        if v.size == 1:
            x = v.flat[0]
            return float(x.real) if (np.iscomplexobj(x) and x.imag == 0) or not np.iscomplexobj(x) else complex(x)
        return v
    return v
