"""A minimal interpreter for the subset of the M language the reference's hot path uses.

TEST INFRASTRUCTURE.  Neither GNU Octave nor MATLAB exists in the build image, so the only
way to run the reference's own source text (fiber.m, fastexp.m, reset_all.m, create_field.m,
fastshift.m, nmod.m, checkfields.m, ampliflat.m under /root/reference) is to interpret it.
This module parses those files as they lie on disk -- nothing is copied -- and evaluates them
with numpy semantics chosen to match the interpreter the reference was written for:
column-major 2-D arrays, value semantics, 1-based indexing with `end`, `'` = conjugate
transpose, matrix `*`, short-circuit && ||, sub-functions, globals, nargin/nargout.
Builtins (fft, cos, rand, ...) are numpy/scipy; `rand`/`randn` draw from a Generator handed in
by the caller so that runs are reproducible.

It is deliberately small: no classes, no cellfun, no N-d arrays, no printing.  Unsupported
syntax raises MSyntaxError at parse time, unsupported builtins raise MError when reached.

oracle/make_golden.py uses it to produce tests/golden/*.npz.
"""
from __future__ import annotations

import math
import os
import re

import numpy as np
import scipy.fft as sfft
import scipy.special as sspecial


class MError(Exception):
    """error(...) raised by M code, or a run-time failure of the interpreter."""


class MSyntaxError(Exception):
    pass


class MStruct(dict):
    def copy(self):
        return MStruct(self)


class MCell(list):
    pass


# =============================================================================== lexer
KEYWORDS = {'function', 'if', 'elseif', 'else', 'end', 'while', 'for', 'switch', 'case', 'otherwise',
            'global', 'persistent', 'return', 'break', 'continue'}
_NUM = re.compile(r'(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?)[ij]?')
_ID = re.compile(r'[A-Za-z_]\w*')
_OPS = ['...', '.^', ".'", '.*', './', '.\\', '==', '~=', '<=', '>=', '&&', '||',
        '+', '-', '*', '/', '\\', '^', '<', '>', '=', '&', '|', '~', ':', ',', ';', '(', ')', '[', ']', '{', '}',
        '.', "'", '@']


class Tok:
    __slots__ = ('kind', 'val', 'line', 'ws')

    def __init__(self, kind, val, line, ws):
        self.kind, self.val, self.line, self.ws = kind, val, line, ws

    def __repr__(self):
        return '%s:%r@%d' % (self.kind, self.val, self.line)


def tokenize(src: str):
    toks = []
    i, n, line = 0, len(src), 1
    stack = []          # open brackets
    ws = False

    def ends_operand():
        if not toks:
            return False
        t = toks[-1]
        return t.kind in ('num', 'str', 'id') or (t.kind == 'op' and t.val in (')', ']', '}', "'", ".'")) or \
            (t.kind == 'kw' and t.val == 'end' and stack)

    while i < n:
        c = src[i]
        if c in ' \t':
            i += 1
            ws = True
            continue
        if src.startswith('...', i):                       # continuation: drop to end of line
            while i < n and src[i] != '\n':
                i += 1
            i += 1
            line += 1
            ws = True
            continue
        if c == '%' or c == '#':
            while i < n and src[i] != '\n':
                i += 1
            continue
        if c == '\r':
            i += 1
            continue
        if c == '\n':
            if stack and stack[-1] in '[{':
                toks.append(Tok('op', ';', line, ws))      # row separator inside a literal
            elif not stack:
                toks.append(Tok('nl', '\n', line, ws))
            i += 1
            line += 1
            ws = False
            continue
        in_lit = bool(stack) and stack[-1] in '[{'
        # element separation by white space inside [] / {}
        if in_lit and ws and ends_operand():
            nxt = src[i]
            sep = False
            if nxt.isalnum() or nxt in "_([{'~@" or (nxt == '.' and i + 1 < n and src[i + 1].isdigit()):
                sep = True
            elif nxt in '+-' and i + 1 < n and src[i + 1] not in ' \t':
                sep = True
            if sep:
                toks.append(Tok('op', ',', line, True))
        if c == "'":
            if ends_operand() and not ws:
                toks.append(Tok('op', "'", line, ws))
                i += 1
                ws = False
                continue
            if ends_operand() and ws and not in_lit:
                toks.append(Tok('op', "'", line, ws))
                i += 1
                ws = False
                continue
            j = i + 1
            out = []
            while True:
                if j >= n or src[j] == '\n':
                    raise MSyntaxError('unterminated string at line %d' % line)
                if src[j] == "'":
                    if j + 1 < n and src[j + 1] == "'":
                        out.append("'")
                        j += 2
                        continue
                    break
                out.append(src[j])
                j += 1
            toks.append(Tok('str', ''.join(out), line, ws))
            i = j + 1
            ws = False
            continue
        m = _NUM.match(src, i)
        if m and (c.isdigit() or (c == '.' and i + 1 < n and src[i + 1].isdigit())):
            text = m.group(0)
            # "1./x" is 1 ./ x : give the dot back
            if text.endswith('.') and m.end() < n and src[m.end()] in "*/\\^'":
                text = text[:-1]
            toks.append(Tok('num', text, line, ws))
            i += len(text)
            ws = False
            continue
        m = _ID.match(src, i)
        if m:
            w = m.group(0)
            kind = 'kw' if w in KEYWORDS else 'id'
            if w == 'end' and stack:
                kind = 'kw'
            toks.append(Tok(kind, w, line, ws))
            i = m.end()
            ws = False
            continue
        for op in _OPS:
            if src.startswith(op, i):
                if op in '([{':
                    stack.append(op)
                elif op in ')]}':
                    if not stack:
                        raise MSyntaxError('unbalanced %s at line %d' % (op, line))
                    stack.pop()
                toks.append(Tok('op', op, line, ws))
                i += len(op)
                ws = False
                break
        else:
            raise MSyntaxError('bad character %r at line %d' % (c, line))
    toks.append(Tok('nl', '\n', line, False))
    toks.append(Tok('eof', None, line, False))
    return toks


# =============================================================================== parser
class Parser:
    def __init__(self, toks, fname='?'):
        self.t, self.p, self.fname = toks, 0, fname
        self.depth = 0   # () / {} nesting: `end` is a value inside

    def peek(self, k=0):
        return self.t[self.p + k]

    def next(self):
        tok = self.t[self.p]
        self.p += 1
        return tok

    def is_op(self, v, k=0):
        tok = self.peek(k)
        return tok.kind == 'op' and tok.val == v

    def is_kw(self, v):
        tok = self.peek()
        return tok.kind == 'kw' and tok.val == v

    def expect_op(self, v):
        tok = self.next()
        if tok.kind != 'op' or tok.val != v:
            raise MSyntaxError('%s:%d: expected %r, got %r' % (self.fname, tok.line, v, tok.val))

    def skip_seps(self):
        while self.peek().kind == 'nl' or (self.peek().kind == 'op' and self.peek().val in (';', ',')):
            self.next()

    # ---- file = script or list of functions
    def parse_file(self):
        funcs = []
        self.skip_seps()
        if not self.is_kw('function'):
            body = self.block(('eof',))
            return [('function', '__script__', [], [], body)]
        while self.is_kw('function'):
            funcs.append(self.function())
            self.skip_seps()
        if self.peek().kind != 'eof':
            raise MSyntaxError('%s:%d: trailing text after functions' % (self.fname, self.peek().line))
        return funcs

    def function(self):
        self.next()
        outs = []
        # forms: function name / function name(args) / function out = name(args) / function [o1,o2] = name(args)
        if self.is_op('['):
            self.next()
            while not self.is_op(']'):
                if self.is_op(','):
                    self.next()
                    continue
                outs.append(self.next().val)
            self.next()
            self.expect_op('=')
            name = self.next().val
        else:
            name = self.next().val
            if self.is_op('='):
                self.next()
                outs = [name]
                name = self.next().val
        args = []
        if self.is_op('('):
            self.next()
            while not self.is_op(')'):
                if self.is_op(','):
                    self.next()
                    continue
                args.append(self.next().val)
            self.next()
        body = self.block(('function', 'eof'), func_level=True)
        return ('function', name, args, outs, body)

    def block(self, stop, func_level=False):
        stmts = []
        while True:
            self.skip_seps()
            tok = self.peek()
            if tok.kind == 'eof':
                if 'eof' in stop:
                    return stmts
                raise MSyntaxError('%s: unexpected end of file' % self.fname)
            if tok.kind == 'kw' and tok.val in stop:
                return stmts
            if tok.kind == 'kw' and tok.val == 'end' and func_level:
                self.next()      # optional `end` closing a function
                return stmts
            stmts.append(self.statement())

    def statement(self):
        tok = self.peek()
        ln = tok.line
        if tok.kind == 'kw':
            kw = tok.val
            if kw == 'if':
                self.next()
                clauses = []
                cond = self.expr()
                body = self.block(('elseif', 'else', 'end'))
                clauses.append((cond, body))
                els = None
                while True:
                    if self.is_kw('elseif'):
                        self.next()
                        cond = self.expr()
                        clauses.append((cond, self.block(('elseif', 'else', 'end'))))
                    elif self.is_kw('else'):
                        self.next()
                        els = self.block(('end',))
                    else:
                        self.next()
                        break
                return ('if', clauses, els, ln)
            if kw == 'while':
                self.next()
                cond = self.expr()
                body = self.block(('end',))
                self.next()
                return ('while', cond, body, ln)
            if kw == 'for':
                self.next()
                paren = self.is_op('(')
                if paren:
                    self.next()
                var = self.next().val
                self.expect_op('=')
                rng = self.expr()
                if paren:
                    self.expect_op(')')
                body = self.block(('end',))
                self.next()
                return ('for', var, rng, body, ln)
            if kw == 'switch':
                self.next()
                subj = self.expr()
                self.skip_seps()
                cases, default = [], None
                while True:
                    if self.is_kw('case'):
                        self.next()
                        val = self.expr()
                        cases.append((val, self.block(('case', 'otherwise', 'end'))))
                    elif self.is_kw('otherwise'):
                        self.next()
                        default = self.block(('case', 'otherwise', 'end'))
                    else:
                        self.next()
                        break
                return ('switch', subj, cases, default, ln)
            if kw in ('global', 'persistent'):
                self.next()
                names = []
                while self.peek().kind == 'id':
                    names.append(self.next().val)
                return (kw, names, ln)
            if kw in ('return', 'break', 'continue'):
                self.next()
                return (kw, ln)
            raise MSyntaxError('%s:%d: unexpected keyword %s' % (self.fname, ln, kw))
        # multi-assignment  [a,b] = f(...)
        if self.is_op('['):
            save = self.p
            depth, k = 0, self.p
            while True:
                tk = self.t[k]
                if tk.kind == 'op' and tk.val == '[':
                    depth += 1
                elif tk.kind == 'op' and tk.val == ']':
                    depth -= 1
                    if depth == 0:
                        break
                elif tk.kind in ('nl', 'eof'):
                    break
                k += 1
            if self.t[k].kind == 'op' and self.t[k + 1].kind == 'op' and self.t[k + 1].val == '=':
                self.next()
                lhs = []
                while not self.is_op(']'):
                    if self.is_op(','):
                        self.next()
                        continue
                    lhs.append(self.postfix())
                self.next()
                self.expect_op('=')
                rhs = self.expr()
                return ('massign', lhs, rhs, ln)
            self.p = save
        e = self.expr()
        if self.is_op('='):
            self.next()
            rhs = self.expr()
            return ('assign', e, rhs, ln)
        return ('expr', e, ln)

    # ---- expressions
    def expr(self):
        return self.oror()

    def _bin(self, sub, ops):
        left = sub()
        while self.peek().kind == 'op' and self.peek().val in ops:
            op = self.next().val
            right = sub()
            left = ('bin', op, left, right)
        return left

    def oror(self):
        return self._bin(self.andand, ('||',))

    def andand(self):
        return self._bin(self.elor, ('&&',))

    def elor(self):
        return self._bin(self.eland, ('|',))

    def eland(self):
        return self._bin(self.cmp, ('&',))

    def cmp(self):
        return self._bin(self.colon, ('==', '~=', '<', '<=', '>', '>='))

    def colon(self):
        # a:b or a:s:b ; a lone ':' inside an index is handled by the caller
        first = self.additive()
        if self.is_op(':') and not self._colon_is_index_all():
            self.next()
            second = self.additive()
            if self.is_op(':') and not self._colon_is_index_all():
                self.next()
                third = self.additive()
                return ('range', first, second, third)
            return ('range', first, None, second)
        return first

    def _colon_is_index_all(self):
        nxt = self.peek(1)
        return nxt.kind == 'op' and nxt.val in (',', ')')

    def additive(self):
        return self._bin(self.mult, ('+', '-'))

    def mult(self):
        return self._bin(self.unary, ('*', '/', '.*', './', '\\', '.\\'))

    def unary(self):
        if self.peek().kind == 'op' and self.peek().val in ('-', '+', '~'):
            op = self.next().val
            return ('un', op, self.unary())
        return self.power()

    def power(self):
        base = self.postfix()
        while self.peek().kind == 'op' and self.peek().val in ('^', '.^'):
            op = self.next().val
            # exponent may carry its own unary sign: 2^-1
            if self.peek().kind == 'op' and self.peek().val in ('-', '+', '~'):
                uop = self.next().val
                expo = ('un', uop, self.postfix())
            else:
                expo = self.postfix()
            base = ('bin', op, base, expo)
        return base

    def args(self, close):
        out = []
        self.depth += 1
        while not self.is_op(close):
            if self.is_op(','):
                self.next()
                continue
            if self.is_op(':') and self.peek(1).kind == 'op' and self.peek(1).val in (',', close):
                self.next()
                out.append(('all',))
                continue
            out.append(self.expr())
        self.next()
        self.depth -= 1
        return out

    def postfix(self):
        e = self.primary()
        while True:
            tok = self.peek()
            if tok.kind != 'op':
                break
            if tok.val == '(' and not (tok.ws and self._in_literal()):
                self.next()
                e = ('index', e, self.args(')'))
            elif tok.val == '{' and not (tok.ws and self._in_literal()):
                self.next()
                e = ('cellindex', e, self.args('}'))
            elif tok.val == '.' and self.peek(1).kind == 'id' and not tok.ws:
                self.next()
                e = ('field', e, self.next().val)
            elif tok.val == "'":
                self.next()
                e = ('ctranspose', e)
            elif tok.val == ".'":
                self.next()
                e = ('transpose', e)
            else:
                break
        return e

    _lit = 0

    def _in_literal(self):
        return self._lit > 0

    def primary(self):
        tok = self.next()
        if tok.kind == 'num':
            txt = tok.val
            if txt[-1] in 'ij':
                return ('const', np.array([[complex(0, float(txt[:-1]))]]))
            return ('const', np.array([[float(txt)]]))
        if tok.kind == 'str':
            return ('const', tok.val)
        if tok.kind == 'id':
            return ('name', tok.val)
        if tok.kind == 'kw' and tok.val == 'end':
            return ('end',)
        if tok.kind == 'op':
            if tok.val == '(':
                self.depth += 1
                save, self._lit = self._lit, 0
                e = self.expr()
                self._lit = save
                self.depth -= 1
                self.expect_op(')')
                return ('paren', e)
            if tok.val in ('[', '{'):
                close = ']' if tok.val == '[' else '}'
                self._lit += 1
                rows, row = [], []
                while not self.is_op(close):
                    if self.is_op(';'):
                        self.next()
                        rows.append(row)
                        row = []
                        continue
                    if self.is_op(','):
                        self.next()
                        continue
                    row.append(self.expr())
                self.next()
                self._lit -= 1
                rows.append(row)
                rows = [r for r in rows if r]
                return ('matrix' if close == ']' else 'cell', rows)
            if tok.val == ':':
                return ('all',)
        raise MSyntaxError('%s:%d: unexpected token %r' % (self.fname, tok.line, tok.val))


# =============================================================================== values
def arr(x):
    """Any numeric -> 2-D numpy array."""
    if isinstance(x, np.ndarray):
        if x.ndim == 2:
            return x
        if x.ndim == 0:
            return x.reshape(1, 1)
        if x.ndim == 1:
            return x.reshape(1, -1)
        if x.ndim == 3:        # (minimal 3-D support: element-wise operations, conj, unary minus, squeeze)
            return x
        raise MError('N-d arrays are not supported')
    if isinstance(x, (bool, np.bool_)):
        return np.array([[bool(x)]])
    if isinstance(x, (int, float, complex, np.number)):
        return np.array([[x]], dtype=np.complex128 if isinstance(x, complex) else np.float64)
    if isinstance(x, str):
        return np.array([[float(ord(ch)) for ch in x]]) if x else np.zeros((0, 0))
    raise MError('not numeric: %r' % type(x))


def num(x):
    a = arr(x)
    if a.dtype == np.bool_:
        return a.astype(np.float64)
    return a


def is_scalar(a):
    return isinstance(a, np.ndarray) and a.size == 1


def truth(v):
    if isinstance(v, str):
        return len(v) > 0
    a = arr(v)
    if a.size == 0:
        return False
    if np.iscomplexobj(a):
        a = a != 0
    return bool(np.all(a != 0))


def scalar(v):
    a = arr(v)
    if a.size != 1:
        raise MError('scalar expected, got size %s' % (a.shape,))
    x = a.flat[0]
    return x


def mround(x):
    return np.sign(x) * np.floor(np.abs(x) + 0.5)


def mrange(a, s, b):
    a, s, b = float(np.real(scalar(a))), float(np.real(scalar(s))), float(np.real(scalar(b)))
    if s == 0 or (s > 0 and a > b) or (s < 0 and a < b):
        return np.zeros((1, 0))
    n = int(math.floor((b - a) / s * (1 + 1e-15) + 1e-10)) + 1
    return (a + s * np.arange(n, dtype=np.float64)).reshape(1, -1)


# =============================================================================== interpreter
class _Return(Exception):
    pass


class _Break(Exception):
    pass


class _Continue(Exception):
    pass


class Interp:
    def __init__(self, path, rng=None):
        # path: one directory, or a list searched in order (an earlier directory shadows a later one, like the
        # interpreter's path); self.builtins: per-instance functions (e.g. a MEX gateway) looked up before BUILTINS
        self.path = path
        self.builtins = {}
        self.rng = rng if rng is not None else np.random.default_rng(0)
        self.globals = {}
        self.files = {}       # file base name -> {func name: ast}
        self.persist = {}
        self.warnings = []

    # ---- loading
    def load(self, name):
        if name in self.files:
            return self.files[name]
        fn = None
        for d in ([self.path] if isinstance(self.path, str) else list(self.path)):
            cand = os.path.join(d, name + '.m')
            if os.path.exists(cand):
                fn = cand
                break
        if fn is None:
            return None
        src = open(fn, encoding='latin-1').read()
        funcs = Parser(tokenize(src), name + '.m').parse_file()
        table = {}
        for f in funcs:
            table[f[1]] = f
        table['__main__'] = funcs[0]
        self.files[name] = table
        return table

    # ---- calling
    def call(self, name, args, nargout=1, local_funcs=None):
        if local_funcs and name in local_funcs and name != '__main__':
            return self.run_function(local_funcs[name], args, nargout, local_funcs)
        if name in self.builtins:
            return self.builtins[name](self, args, nargout)
        table = self.load(name)
        if table is not None:
            return self.run_function(table['__main__'], args, nargout, table)
        if name in BUILTINS:
            return BUILTINS[name](self, args, nargout)
        raise MError("undefined function or variable '%s'" % name)

    def run_function(self, f, args, nargout, table):
        _, fname, params, outs, body = f
        if len(args) > len(params):
            raise MError('%s: too many input arguments' % fname)
        ws = {'__funcs__': table, '__globals__': set(), '__fname__': fname}
        for p, a in zip(params, args):
            ws[p] = a
        ws['nargin'] = np.array([[float(len(args))]])
        ws['nargout'] = np.array([[float(nargout)]])
        try:
            self.exec_block(body, ws)
        except _Return:
            pass
        if outs and outs[-1] == 'varargout':      # function varargout = f(...): the cell's elements are the outputs
            va = ws.get('varargout')
            head = [ws.get(o) for o in outs[:-1]]
            tail = list(va) if isinstance(va, MCell) else []
            return (head + tail)[:max(nargout, 1)]
        res = []
        for k, o in enumerate(outs[:max(nargout, 1)]):
            if o in ws['__globals__']:
                val = self.globals.get(o)
            else:
                val = ws.get(o)
            if val is None:
                if k < nargout:
                    raise MError("%s: output '%s' not assigned" % (fname, o))
                break
            res.append(val)
        return res

    # ---- variables
    def getvar(self, ws, name):
        if name in ws['__globals__']:
            return self.globals.get(name)
        if name in ws.get('__persist__', ()):
            return self.persist[ws['__fname__']].get(name)
        return ws.get(name)

    def setvar(self, ws, name, val):
        if name in ws['__globals__']:
            self.globals[name] = val
        elif name in ws.get('__persist__', ()):
            self.persist[ws['__fname__']][name] = val
        else:
            ws[name] = val

    # ---- statements
    def exec_block(self, stmts, ws):
        for s in stmts:
            self.exec_stmt(s, ws)

    def exec_stmt(self, s, ws):
        kind = s[0]
        if kind == 'expr':
            e = s[1]
            if e[0] == 'name' and self.getvar(ws, e[1]) is None:
                self.eval_multi(e, ws, 0)       # command-style call, e.g. `return`-less procedure
            else:
                self.eval_multi(e, ws, 0)
        elif kind == 'assign':
            val = self.eval(s[2], ws)
            self.assign(s[1], val, ws)
        elif kind == 'massign':
            vals = self.eval_multi(s[2], ws, len(s[1]))
            if len(vals) < len(s[1]):
                raise MError('line %d: not enough outputs' % s[3])
            for lhs, v in zip(s[1], vals):
                self.assign(lhs, v, ws)
        elif kind == 'if':
            for cond, body in s[1]:
                if truth(self.eval(cond, ws)):
                    self.exec_block(body, ws)
                    return
            if s[2] is not None:
                self.exec_block(s[2], ws)
        elif kind == 'while':
            while truth(self.eval(s[1], ws)):
                try:
                    self.exec_block(s[2], ws)
                except _Break:
                    break
                except _Continue:
                    continue
        elif kind == 'for':
            rng = self.eval(s[2], ws)
            cols = [rng] if isinstance(rng, str) else [arr(rng)[:, k:k + 1] for k in range(arr(rng).shape[1])]
            for c in cols:
                self.setvar(ws, s[1], c if c.shape[0] > 1 else c.reshape(1, 1))
                try:
                    self.exec_block(s[3], ws)
                except _Break:
                    break
                except _Continue:
                    continue
        elif kind == 'switch':
            subj = self.eval(s[1], ws)
            for val, body in s[2]:
                v = self.eval(val, ws)
                cands = v if isinstance(v, MCell) else [v]
                hit = False
                for cv in cands:
                    if isinstance(subj, str) or isinstance(cv, str):
                        hit = isinstance(subj, str) and isinstance(cv, str) and subj == cv
                    else:
                        hit = arr(subj).size > 0 and bool(np.all(arr(subj) == arr(cv)))
                    if hit:
                        break
                if hit:
                    self.exec_block(body, ws)
                    return
            if s[3] is not None:
                self.exec_block(s[3], ws)
        elif kind == 'global':
            for nm in s[1]:
                ws['__globals__'].add(nm)
                ws.pop(nm, None)
        elif kind == 'persistent':
            store = self.persist.setdefault(ws['__fname__'], {})
            ws.setdefault('__persist__', set())
            for nm in s[1]:
                store.setdefault(nm, np.zeros((0, 0)))
                ws['__persist__'].add(nm)
                ws.pop(nm, None)
        elif kind == 'return':
            raise _Return()
        elif kind == 'break':
            raise _Break()
        elif kind == 'continue':
            raise _Continue()
        else:
            raise MError('unknown statement %s' % kind)

    # ---- assignment
    def assign(self, lhs, val, ws):
        k = lhs[0]
        if k == 'name':
            self.setvar(ws, lhs[1], val)
            return
        if k == 'paren':
            return self.assign(lhs[1], val, ws)
        # build the access chain root.name -> steps
        chain = []
        node = lhs
        while node[0] in ('field', 'index', 'cellindex'):
            chain.append(node)
            node = node[1]
        if node[0] != 'name':
            raise MError('bad assignment target')
        root = node[1]
        cur = self.getvar(ws, root)
        new = self._assign_chain(cur, list(reversed(chain)), val, ws)
        self.setvar(ws, root, new)

    def _assign_chain(self, cur, chain, val, ws):
        if not chain:
            return val
        step = chain[0]
        if step[0] == 'field':
            base = MStruct() if cur is None or (isinstance(cur, np.ndarray) and cur.size == 0) else cur
            if not isinstance(base, MStruct):
                raise MError('field assignment to a non-struct')
            base = base.copy()
            base[step[2]] = self._assign_chain(base.get(step[2]), chain[1:], val, ws)
            return base
        if step[0] == 'cellindex':
            base = MCell(cur) if cur is not None else MCell()
            idx = int(scalar(self.eval(step[2][0], ws, end_ctx=(base, 0, 1)))) - 1
            while len(base) <= idx:
                base.append(np.zeros((0, 0)))
            base[idx] = self._assign_chain(base[idx], chain[1:], val, ws)
            return base
        # C(k) = {v}: paren-indexed assignment of a cell into a cell (varargout(2) = {x})
        if isinstance(val, MCell) and len(chain) == 1 and len(step[2]) == 1 and (cur is None or isinstance(cur, MCell) or
                                                                                  (isinstance(cur, np.ndarray) and cur.size == 0)):
            base = MCell(cur) if isinstance(cur, MCell) else MCell()
            idx = np.real(arr(self.eval(step[2][0], ws, end_ctx=(base, 0, 1)))).astype(np.int64).flatten() - 1
            if len(val) not in (1, len(idx)):
                raise MError('C(I) = {..}: number of elements must agree')
            for n, i in enumerate(idx):
                while len(base) <= i:
                    base.append(np.zeros((0, 0)))
                base[int(i)] = val[n if len(val) > 1 else 0]
            return base
        # numeric indexed assignment
        if len(chain) > 1:
            raise MError('nested indexed assignment is not supported')
        base = np.zeros((0, 0)) if cur is None else cur
        if isinstance(base, str):
            base = arr(base)
        if (isinstance(base, np.ndarray) and base.ndim == 3) or len(step[2]) == 3:
            return self.index_assign(base, step[2], val, ws)
        return self.index_assign(arr(base), step[2], val, ws)

    def _sel3(self, a, idx_nodes, ws):
        sel = []
        for d, node in enumerate(idx_nodes):
            if node[0] == 'all':
                sel.append(np.arange(a.shape[d]))
            else:
                ix = arr(self.eval(node, ws, end_ctx=(a, d, 3)))
                sel.append(np.real(ix).astype(np.int64).flatten('F') - 1)
        return sel

    def index_assign(self, a, idx_nodes, val, ws):
        if len(idx_nodes) == 3:   # minimal 3-D support: A(i,j,k) = v on a 3-D array, or creating / growing one
            v = np.asarray(val) if isinstance(val, np.ndarray) and val.ndim == 3 else num(val)
            a = np.asarray(a)
            if a.ndim < 3:
                a = a.reshape(a.shape + (1,)) if a.size else np.zeros((0, 0, 0), dtype=a.dtype)
            out = a.astype(np.complex128) if np.iscomplexobj(v) and not np.iscomplexobj(a) else a.copy()
            sel = []
            for d, node in enumerate(idx_nodes):
                if node[0] == 'all':       # ':' on a dimension that does not exist yet takes its extent from the value
                    sel.append(np.arange(out.shape[d] if out.shape[d] else (v.size if v.ndim < 3 else v.shape[d])))
                else:
                    ix = arr(self.eval(node, ws, end_ctx=(out, d, 3)))
                    sel.append(np.real(ix).astype(np.int64).flatten('F') - 1)
            need = tuple(max(out.shape[d], int(sel[d].max()) + 1 if sel[d].size else 0) for d in range(3))
            if need != out.shape:
                grown = np.zeros(need, dtype=out.dtype)
                grown[:out.shape[0], :out.shape[1], :out.shape[2]] = out
                out = grown
            out[np.ix_(*sel)] = v.flat[0] if v.size == 1 else np.reshape(v, tuple(len(q) for q in sel), order='F')
            return out
        v = num(val) if not isinstance(val, str) else arr(val)
        if np.iscomplexobj(v) and not np.iscomplexobj(a):
            a = a.astype(np.complex128)
        elif a.dtype == np.bool_:
            a = a.astype(v.dtype if v.dtype != np.bool_ else np.float64)
        else:
            a = a.copy()
        if len(idx_nodes) == 1:
            node = idx_nodes[0]
            if node[0] == 'all':
                flat = a.flatten('F')
                flat[:] = v.flatten('F') if v.size > 1 else v.flat[0]
                return flat.reshape(a.shape, order='F')
            ix = self.eval(node, ws, end_ctx=(a, 0, 1))
            ixa = arr(ix)
            if ixa.dtype == np.bool_:
                pos = np.flatnonzero(ixa.flatten('F'))
            else:
                pos = np.real(ixa).astype(np.int64).flatten('F') - 1
            if pos.size and pos.min() < 0:
                raise MError('index must be positive')
            need = int(pos.max()) + 1 if pos.size else 0
            if need > a.size:
                if a.size == 0:
                    a = np.zeros((1, need), dtype=a.dtype)
                elif a.shape[0] == 1:
                    a = np.concatenate([a, np.zeros((1, need - a.shape[1]), dtype=a.dtype)], axis=1)
                elif a.shape[1] == 1:
                    a = np.concatenate([a, np.zeros((need - a.shape[0], 1), dtype=a.dtype)], axis=0)
                else:
                    raise MError('cannot grow a matrix with a linear index')
            flat = a.flatten('F')
            if v.size == 1:
                flat[pos] = v.flat[0]
            else:
                if v.size != pos.size:
                    raise MError('A(I) = B: number of elements must agree (%d vs %d)' % (pos.size, v.size))
                flat[pos] = v.flatten('F')
            return flat.reshape(a.shape, order='F')
        if len(idx_nodes) == 2:
            sel = []
            for d, node in enumerate(idx_nodes):
                if node[0] == 'all':
                    if a.size == 0 and v.size > 0:
                        sel.append(np.arange(v.shape[d]))
                    else:
                        sel.append(np.arange(a.shape[d]))
                else:
                    ix = arr(self.eval(node, ws, end_ctx=(a, d, 2)))
                    if ix.dtype == np.bool_:
                        sel.append(np.flatnonzero(ix.flatten('F')))
                    else:
                        sel.append(np.real(ix).astype(np.int64).flatten('F') - 1)
            nr = max(a.shape[0], int(sel[0].max()) + 1 if sel[0].size else 0)
            nc = max(a.shape[1], int(sel[1].max()) + 1 if sel[1].size else 0)
            if (nr, nc) != a.shape:
                b = np.zeros((nr, nc), dtype=a.dtype)
                b[:a.shape[0], :a.shape[1]] = a
                a = b
            if v.size == 1:
                a[np.ix_(sel[0], sel[1])] = v.flat[0]
            else:
                a[np.ix_(sel[0], sel[1])] = v.reshape(len(sel[0]), len(sel[1]), order='F') if v.shape != (len(sel[0]), len(sel[1])) else v
            return a
        raise MError('more than two subscripts are not supported')

    # ---- expressions
    def eval(self, e, ws, end_ctx=None):
        r = self.eval_multi(e, ws, 1, end_ctx)
        if not r:
            raise MError('expression produced no value')
        return r[0]

    def eval_multi(self, e, ws, nargout, end_ctx=None):
        k = e[0]
        if k == 'const':
            return [e[1]]
        if k == 'name':
            v = self.getvar(ws, e[1])
            if v is not None:
                return [v]
            return self.call(e[1], [], nargout, ws['__funcs__'])
        if k == 'paren':
            return [self.eval(e[1], ws, end_ctx)]
        if k == 'end':
            if end_ctx is None:
                raise MError("'end' outside an index expression")
            a, d, nd = end_ctx
            if isinstance(a, (MCell, str)):
                return [np.array([[float(len(a))]])]
            return [np.array([[float(a.size if nd == 1 else a.shape[d])]])]
        if k == 'all':
            raise MError("':' outside an index expression")
        if k == 'un':
            v = self.eval(e[2], ws, end_ctx)
            if e[1] == '-':
                return [-num(v)]
            if e[1] == '+':
                return [num(v)]
            a = arr(v)
            return [~(a != 0)]
        if k == 'bin':
            op = e[1]
            if op == '&&':
                return [np.array([[truth(self.eval(e[2], ws, end_ctx)) and truth(self.eval(e[3], ws, end_ctx))]])]
            if op == '||':
                return [np.array([[truth(self.eval(e[2], ws, end_ctx)) or truth(self.eval(e[3], ws, end_ctx))]])]
            return [self.binop(op, self.eval(e[2], ws, end_ctx), self.eval(e[3], ws, end_ctx))]
        if k == 'range':
            a = self.eval(e[1], ws, end_ctx)
            b = self.eval(e[3], ws, end_ctx)
            s = self.eval(e[2], ws, end_ctx) if e[2] is not None else 1.0
            return [mrange(a, s, b)]
        if k == 'matrix':
            rows = []
            for row in e[1]:
                vals = [self.eval(x, ws, end_ctx) for x in row]
                if all(isinstance(v, str) for v in vals):
                    rows.append(''.join(vals))
                    continue
                parts = [num(v) if not isinstance(v, str) else arr(v) for v in vals]
                parts = [p for p in parts if p.size > 0] or [np.zeros((0, 0))]
                rows.append(np.concatenate(parts, axis=1) if len(parts) > 1 else parts[0])
            if not rows:
                return [np.zeros((0, 0))]
            if all(isinstance(r, str) for r in rows):
                if len(rows) == 1:
                    return [rows[0]]
                raise MError('char matrices are not supported')
            rows = [r for r in rows if not isinstance(r, str) and r.size > 0] or [np.zeros((0, 0))]
            return [np.concatenate(rows, axis=0) if len(rows) > 1 else rows[0]]
        if k == 'cell':
            out = MCell()
            for row in e[1]:
                for x in row:
                    out.append(self.eval(x, ws, end_ctx))
            return [out]
        if k == 'ctranspose':
            v = self.eval(e[1], ws, end_ctx)
            return [np.conj(arr(v)).T]
        if k == 'transpose':
            return [arr(self.eval(e[1], ws, end_ctx)).T]
        if k == 'field':
            base = self.eval(e[1], ws, end_ctx)
            if not isinstance(base, MStruct):
                raise MError("field access '.%s' on a non-struct" % e[2])
            if e[2] not in base:
                raise MError("reference to non-existent field '%s'" % e[2])
            return [base[e[2]]]
        if k == 'cellindex':
            base = self.eval(e[1], ws, end_ctx)
            if not isinstance(base, MCell):
                raise MError('{} indexing of a non-cell')
            i = int(scalar(self.eval(e[2][0], ws, end_ctx=(base, 0, 1)))) - 1
            return [base[i]]
        if k == 'index':
            target = e[1]
            if target[0] == 'name' and self.getvar(ws, target[1]) is None:
                # function call: `end` inside the arguments still belongs to the enclosing index
                args = [self.eval(a, ws, end_ctx) for a in e[2]]
                return self.call(target[1], args, nargout, ws['__funcs__'])
            base = self.eval(target, ws, end_ctx)
            return [self.index(base, e[2], ws)]
        raise MError('cannot evaluate %s' % k)

    def index(self, base, idx_nodes, ws):
        if isinstance(base, np.ndarray) and base.ndim == 3 and len(idx_nodes) == 3:
            return base[np.ix_(*self._sel3(base, idx_nodes, ws))]
        if isinstance(base, MCell):
            i = arr(self.eval(idx_nodes[0], ws, end_ctx=(base, 0, 1)))
            return MCell([base[int(j) - 1] for j in i.flatten()])
        if isinstance(base, str):
            a = arr(base)
            sub = self.index(a, idx_nodes, ws)
            return ''.join(chr(int(c)) for c in sub.flatten())
        if isinstance(base, MStruct):
            if len(idx_nodes) == 1 and scalar(self.eval(idx_nodes[0], ws, end_ctx=(np.zeros((1, 1)), 0, 1))) == 1:
                return base
            raise MError('struct arrays are not supported')
        a = arr(base)
        if len(idx_nodes) == 0:
            return a
        if len(idx_nodes) == 1:
            node = idx_nodes[0]
            if node[0] == 'all':
                return a.flatten('F').reshape(-1, 1)
            ix = arr(self.eval(node, ws, end_ctx=(a, 0, 1)))
            flat = a.flatten('F')
            if ix.dtype == np.bool_:
                res = flat[np.flatnonzero(ix.flatten('F'))]
                return res.reshape(-1, 1) if a.shape[1] == 1 and a.shape[0] != 1 else res.reshape(1, -1)
            pos = np.real(ix).astype(np.int64) - 1
            if pos.size and (pos.min() < 0 or pos.max() >= flat.size):
                raise MError('index (%d) out of bound %d' % (int(pos.max()) + 1 if pos.max() >= flat.size else int(pos.min()) + 1, flat.size))
            res = flat[pos.flatten('F')]
            if ix.shape[0] == 1 or ix.shape[1] == 1:
                if a.shape[0] == 1 or a.shape[1] == 1:       # vector source keeps its orientation
                    return res.reshape(1, -1) if a.shape[0] == 1 else res.reshape(-1, 1)
                return res.reshape(ix.shape, order='F')
            return res.reshape(ix.shape, order='F')
        if len(idx_nodes) == 2:
            sel = []
            for d, node in enumerate(idx_nodes):
                if node[0] == 'all':
                    sel.append(np.arange(a.shape[d]))
                else:
                    ix = arr(self.eval(node, ws, end_ctx=(a, d, 2)))
                    if ix.dtype == np.bool_:
                        sel.append(np.flatnonzero(ix.flatten('F')))
                    else:
                        p = np.real(ix).astype(np.int64).flatten('F') - 1
                        if p.size and (p.min() < 0 or p.max() >= a.shape[d]):
                            raise MError('index (%d) out of bound %d' % (int(p.max()) + 1, a.shape[d]))
                        sel.append(p)
            return a[np.ix_(sel[0], sel[1])]
        raise MError('more than two subscripts are not supported')

    def binop(self, op, x, y):
        if op in ('==', '~=') and (isinstance(x, str) or isinstance(y, str)):
            a, b = arr(x), arr(y)
            if a.shape != b.shape and a.size != 1 and b.size != 1:
                raise MError('nonconformant arguments')
            return (a == b) if op == '==' else (a != b)
        a, b = num(x), num(y)
        if op == '+':
            return a + b
        if op == '-':
            return a - b
        if op == '.*':
            return a * b
        if op == './':
            with np.errstate(divide='ignore', invalid='ignore'):
                return a / b
        if op == '*':
            if a.size == 1 or b.size == 1:
                return a * b
            return a @ b
        if op == '/':
            if b.size == 1:
                with np.errstate(divide='ignore', invalid='ignore'):
                    return a / b
            raise MError('matrix right division is not supported')
        if op in ('.^', '^'):
            if op == '^' and not (a.size == 1 and b.size == 1):
                raise MError('matrix power is not supported')
            with np.errstate(divide='ignore', invalid='ignore'):
                if b.size == 1 and not np.iscomplexobj(b):
                    return a ** float(b.flat[0])
                return a ** b
        if op in ('<', '<=', '>', '>=', '==', '~='):
            ar, br = (np.real(a), np.real(b)) if op not in ('==', '~=') else (a, b)
            return {'<': np.less, '<=': np.less_equal, '>': np.greater, '>=': np.greater_equal,
                    '==': np.equal, '~=': np.not_equal}[op](ar, br)
        if op == '&':
            return (a != 0) & (b != 0)
        if op == '|':
            return (a != 0) | (b != 0)
        raise MError('operator %s is not supported' % op)


# =============================================================================== builtins
BUILTINS = {}


def builtin(name):
    def deco(fn):
        BUILTINS[name] = fn
        return fn
    return deco


def _simple(name, fn):
    BUILTINS[name] = lambda it, a, n, fn=fn: [fn(*a)]


def _shape_args(a):
    if len(a) == 0:
        return (1, 1)
    if len(a) == 1:
        v = arr(a[0])
        if v.size == 1:
            k = int(np.real(v.flat[0]))
            return (k, k)
        return tuple(int(np.real(q)) for q in v.flatten())
    return tuple(max(0, int(np.real(scalar(q)))) for q in a)


def _reduce(fn, cplx_abs=False):
    def f(it, a, nargout):
        x = num(a[0])
        if len(a) >= 2 and not (isinstance(a[1], np.ndarray) and a[1].size == 0):
            if len(a) == 2:                    # max(a, b)
                return [fn[1](x, num(a[1]))]
        dim = int(scalar(a[2])) - 1 if len(a) >= 3 else (1 if x.shape[0] == 1 else 0)
        if x.size == 0:
            return [np.zeros((0, 0))]
        key = np.abs(x) if np.iscomplexobj(x) else x
        idx = fn[2](key, axis=dim)
        val = np.take_along_axis(x, np.expand_dims(idx, dim), axis=dim)
        out = [val]
        if nargout >= 2:
            out.append(np.expand_dims(idx, dim).astype(np.float64) + 1)
        return out
    return f


BUILTINS['max'] = _reduce((None, np.maximum, np.argmax))
BUILTINS['min'] = _reduce((None, np.minimum, np.argmin))


def _sumlike(fn):
    def f(it, a, nargout):
        x = num(a[0])
        if x.size == 0:
            return [np.zeros((1, 1))]
        dim = int(scalar(a[1])) - 1 if len(a) >= 2 else (1 if x.shape[0] == 1 else 0)
        return [fn(x, axis=dim, keepdims=True)]
    return f


BUILTINS['sum'] = _sumlike(np.sum)
BUILTINS['mean'] = _sumlike(np.mean)
BUILTINS['prod'] = _sumlike(np.prod)


def _anyall(fn):
    def f(it, a, nargout):
        x = arr(a[0])
        if x.size == 0:
            return [np.array([[fn is np.all]])]
        x = x != 0
        if x.shape[0] == 1 or x.shape[1] == 1:
            return [np.array([[bool(fn(x))]])]
        return [fn(x, axis=0, keepdims=True)]
    return f


BUILTINS['any'] = _anyall(np.any)
BUILTINS['all'] = _anyall(np.all)

for _n, _f in [('cos', np.cos), ('sin', np.sin), ('tan', np.tan), ('exp', np.exp), ('abs', np.abs),
               ('real', np.real), ('imag', np.imag), ('conj', np.conj), ('ceil', np.ceil), ('floor', np.floor),
               ('fix', np.trunc), ('sign', np.sign), ('angle', np.angle), ('isnan', np.isnan), ('isinf', np.isinf)]:
    _simple(_n, lambda x, _f=_f: _f(num(x)))
_simple('round', lambda x: mround(num(x)))
_simple('asin', lambda x: np.arcsin(num(x)))
_simple('acos', lambda x: np.arccos(num(x)))
_simple('atan', lambda x: np.arctan(num(x)))
_simple('erfcinv', lambda x: sspecial.erfcinv(num(x)))
_simple('erfc', lambda x: sspecial.erfc(num(x)))


@builtin('sqrt')
def _sqrt(it, a, n):
    x = num(a[0])
    if not np.iscomplexobj(x) and np.any(x < 0):
        x = x.astype(np.complex128)
    return [np.sqrt(x)]


@builtin('log')
def _log(it, a, n):
    x = num(a[0])
    with np.errstate(divide='ignore', invalid='ignore'):
        if not np.iscomplexobj(x) and np.any(x < 0):
            x = x.astype(np.complex128)
        return [np.log(x)]


@builtin('log10')
def _log10(it, a, n):
    with np.errstate(divide='ignore'):
        return [np.log10(num(a[0]))]


_simple('mod', lambda x, y: np.where(num(y) == 0, num(x), np.mod(num(x), np.where(num(y) == 0, 1, num(y)))))
_simple('complex', lambda x, y: num(x).astype(np.float64) + 1j * num(y).astype(np.float64))
def _fftn(fn):
    # fft(x) / fft(x, n): along the first non-singleton dimension, zero-padded or truncated to n points
    def f(x, *r):
        a = num(x)
        n = int(scalar(r[0])) if r and arr(r[0]).size else None
        return fn(a, n=n, axis=(1 if a.shape[0] == 1 else 0))
    return f


_simple('fft', _fftn(sfft.fft))
_simple('ifft', _fftn(sfft.ifft))
_simple('unwrap', lambda x: np.unwrap(num(x), axis=(1 if num(x).shape[0] == 1 else 0)))
_simple('cumsum', lambda x: np.cumsum(num(x), axis=(1 if num(x).shape[0] == 1 else 0)))
_simple('fftshift', lambda x: np.fft.fftshift(num(x), axes=(1 if num(x).shape[0] == 1 else 0)))
_simple('ifftshift', lambda x: np.fft.ifftshift(num(x), axes=(1 if num(x).shape[0] == 1 else 0)))
_simple('flipud', lambda x: num(x)[::-1, :])
_simple('fliplr', lambda x: num(x)[:, ::-1])
_simple('isempty', lambda x: np.array([[(len(x) == 0) if isinstance(x, (str, MCell, MStruct)) else arr(x).size == 0]]))
_simple('isstruct', lambda x: np.array([[isinstance(x, MStruct)]]))
_simple('ischar', lambda x: np.array([[isinstance(x, str)]]))
_simple('iscell', lambda x: np.array([[isinstance(x, MCell)]]))
_simple('isnumeric', lambda x: np.array([[isinstance(x, np.ndarray) and x.dtype != np.bool_]]))
_simple('islogical', lambda x: np.array([[isinstance(x, np.ndarray) and x.dtype == np.bool_]]))
_simple('isreal', lambda x: np.array([[not np.iscomplexobj(arr(x))]]))
_simple('length', lambda x: np.array([[float(len(x) if isinstance(x, (str, MCell)) else (max(arr(x).shape) if arr(x).size else 0))]]))
_simple('numel', lambda x: np.array([[float(len(x) if isinstance(x, (str, MCell)) else arr(x).size)]]))
_simple('lower', lambda s: s.lower() if isinstance(s, str) else s)
_simple('upper', lambda s: s.upper() if isinstance(s, str) else s)
_simple('double', lambda x: num(x).astype(np.complex128 if np.iscomplexobj(arr(x)) else np.float64))
_simple('logical', lambda x: arr(x) != 0)
_simple('find', lambda x: (lambda p, a: (p.reshape(1, -1) if a.shape[0] == 1 else p.reshape(-1, 1)))(
    np.flatnonzero(arr(x).flatten('F') != 0).astype(np.float64) + 1, arr(x)))
_simple('repmat', lambda x, m, n=None: np.tile(num(x), (int(scalar(m)), int(scalar(n if n is not None else m)))))
def _reshape(x, *dims):
    a = np.asarray(x) if isinstance(x, np.ndarray) and x.ndim == 3 else num(x)
    if len(dims) == 1:
        dims = tuple(np.real(arr(dims[0])).astype(int).ravel())
    else:
        dims = tuple(int(scalar(d)) for d in dims)
    if len(dims) < 2 or len(dims) > 3 or int(np.prod(dims)) != a.size:
        raise MError('reshape: the number of elements must not change')
    return np.reshape(a, dims, order='F')


_simple('reshape', _reshape)
_simple('circshift', lambda x, k: np.roll(num(x), int(scalar(k)), axis=(1 if num(x).shape[0] == 1 else 0)))
_simple('fieldnames', lambda s: MCell(list(s.keys())))
def _rmfield(st, f):
    out = MStruct(st)
    if f not in out:
        raise MError("rmfield: no field '%s'" % f)
    del out[f]
    return out


_simple('rmfield', _rmfield)
# sinc is a toolbox function that is not in the reference tree (myfilter.m:80 calls it): sin(pi*x)/(pi*x), 1 at x = 0
_simple('sinc', lambda x: np.sinc(num(x)))
_simple('isfield', lambda s, f: np.array([[isinstance(s, MStruct) and f in s]]))
_simple('num2str', lambda x, *r: ('%g' % np.real(scalar(x))))
_simple('struct', lambda *a: MStruct({a[i]: a[i + 1] for i in range(0, len(a), 2)}))
_simple('nargchk', lambda *a: np.zeros((0, 0)))
def _squeeze(x):
    a = np.asarray(x)
    if a.ndim <= 2:
        return num(x)
    b = np.squeeze(a)
    return b if b.ndim == 2 else (b.reshape(-1, 1) if b.ndim == 1 else b.reshape(1, 1) if b.ndim == 0 else b)


_simple('squeeze', _squeeze)


def _sort(x, *r):
    a = num(x)
    out = np.sort(a, axis=(1 if a.shape[0] == 1 else 0))
    if r and isinstance(r[-1], str) and r[-1].lower() == 'descend':
        out = out[::-1] if a.shape[0] != 1 else out[:, ::-1]
    return out


_simple('sort', _sort)
_simple('tic', lambda: np.zeros((0, 0)))

for _n, _v in [('pi', math.pi), ('Inf', math.inf), ('inf', math.inf), ('NaN', math.nan), ('nan', math.nan),
               ('eps', np.finfo(np.float64).eps)]:
    BUILTINS[_n] = lambda it, a, n, _v=_v: [np.array([[_v]])]
BUILTINS['i'] = BUILTINS['j'] = lambda it, a, n: [np.array([[1j]])]
BUILTINS['true'] = lambda it, a, n: [np.ones(_shape_args(a), dtype=bool)]
BUILTINS['false'] = lambda it, a, n: [np.zeros(_shape_args(a), dtype=bool)]
BUILTINS['zeros'] = lambda it, a, n: [np.zeros(_shape_args(a))]
BUILTINS['ones'] = lambda it, a, n: [np.ones(_shape_args(a))]
BUILTINS['eye'] = lambda it, a, n: [np.eye(*_shape_args(a))]


@builtin('rand')
def _rand(it, a, n):
    a = [x for x in a if not isinstance(x, str)]
    shp = _shape_args(a)
    # column-major fill order, like the interpreter's generators
    return [it.rng.random(shp[0] * shp[1]).reshape(shp, order='F')]


@builtin('randn')
def _randn(it, a, n):
    a = [x for x in a if not isinstance(x, str)]
    shp = _shape_args(a)
    return [it.rng.standard_normal(shp[0] * shp[1]).reshape(shp, order='F')]


@builtin('size')
def _size(it, a, nargout):
    x = a[0]
    shp = (1, len(x)) if isinstance(x, (str, MCell)) else ((1, 1) if isinstance(x, MStruct) else arr(x).shape)
    if len(a) == 2:
        d = int(scalar(a[1])) - 1
        return [np.array([[float(shp[d] if d < 2 else 1)]])]
    if nargout <= 1:
        return [np.array([[float(shp[0]), float(shp[1])]])]
    return [np.array([[float(shp[0])]]), np.array([[float(shp[1])]])] + [np.array([[1.0]])] * (nargout - 2)


@builtin('strcmp')
def _strcmp(it, a, n, fold=False):
    x, y = a

    def eq(p, q):
        if not (isinstance(p, str) and isinstance(q, str)):
            return False
        return p.lower() == q.lower() if fold else p == q
    if isinstance(x, MCell) or isinstance(y, MCell):
        cell, other = (x, y) if isinstance(x, MCell) else (y, x)
        return [np.array([[eq(c, other) for c in cell]]).reshape(-1, 1) if False else np.array([[eq(c, other) for c in cell]], dtype=bool).reshape(len(cell), 1)]
    return [np.array([[eq(x, y)]])]


BUILTINS['strcmpi'] = lambda it, a, n: _strcmp(it, a, n, fold=True)


@builtin('exist')
def _exist(it, a, n):
    # only the workspace form exist('name','var') is meaningful here; handled in Interp.call via ws
    raise MError('exist() must be resolved by the caller')


@builtin('error')
def _error(it, a, n):
    if not a or (isinstance(a[0], np.ndarray) and a[0].size == 0):
        return [np.zeros((0, 0))]
    msg = a[0] if isinstance(a[0], str) else repr(a[0])
    try:
        if len(a) > 1:
            msg = msg % tuple(x if isinstance(x, str) else float(np.real(scalar(x))) for x in a[1:])
    except Exception:
        pass
    raise MError(msg)


@builtin('warning')
def _warning(it, a, n):
    it.warnings.append(' '.join(x for x in a if isinstance(x, str)))
    return [np.zeros((0, 0))]


for _n in ('fprintf', 'disp', 'fclose', 'drawnow', 'figure', 'plot', 'hold', 'grid', 'xlabel', 'ylabel', 'title',
           'mkdir', 'fflush'):
    BUILTINS[_n] = lambda it, a, n: [np.zeros((0, 0))]
BUILTINS['fopen'] = lambda it, a, n: [np.array([[-1.0]])]


def m_sprintf(fmt, args):
    """fprintf / sprintf formatting: escapes of the format, arguments flattened column-major, the format reused
    until the arguments are consumed (the interpreter's semantics; enough for the simul_out blocks of the reference)."""
    fmt = fmt.replace('\\n', '\n').replace('\\t', '\t')
    flat = []
    for v in args:
        if isinstance(v, str):
            flat.append(v)
        else:
            flat.extend(np.real(np.asarray(v, dtype=complex)).ravel(order='F').tolist())
    specs = re.findall(r'%[-+ 0#]*\d*(?:\.\d+)?([diuoxXfeEgGcs])', fmt.replace('%%', ''))
    if not specs:
        return fmt.replace('%%', '%')
    out, i = '', 0
    while True:
        vals = []
        for k, sp in enumerate(specs):
            v = flat[i + k] if i + k < len(flat) else ('' if sp == 's' else 0.0)
            if sp in 'diu' and not isinstance(v, str):
                v = int(round(v))
            vals.append(v)
        piece = fmt % tuple(vals)
        if any(isinstance(v, float) and not np.isfinite(v) for v in vals):    # the interpreter spells them Inf / NaN
            piece = re.sub(r'\binf\b', 'Inf', re.sub(r'\bnan\b', 'NaN', piece))
        out += piece
        i += len(specs)
        if i >= len(flat):
            return out


def _fprintf_capture(it, a, n):
    """With it.printed a list: text written through fprintf(fid, fmt, ...) is appended to it (fid itself is ignored)."""
    if getattr(it, 'printed', None) is None or not a:
        return [np.zeros((0, 0))]
    rest = a[1:] if not isinstance(a[0], str) else a
    if rest and isinstance(rest[0], str):
        it.printed.append(m_sprintf(rest[0], rest[1:]))
    return [np.zeros((0, 0))]


BUILTINS['fprintf'] = _fprintf_capture
BUILTINS['input'] = lambda it, a, n: (_ for _ in ()).throw(MError('input() is interactive: not available'))


# exist('x','var') needs the caller's workspace: patch eval for that one name
_orig_eval_multi = Interp.eval_multi


def _eval_multi_exist(self, e, ws, nargout, end_ctx=None):
    if e[0] == 'index' and e[1][0] == 'name' and e[1][1] == 'exist' and self.getvar(ws, 'exist') is None:
        args = [self.eval(a, ws, end_ctx) for a in e[2]]
        name = args[0]
        kind = args[1] if len(args) > 1 else None
        if kind in (None, 'var') and isinstance(name, str) and self.getvar(ws, name) is not None:
            return [np.array([[1.0]])]
        if kind in (None, 'builtin', 'file') and isinstance(name, str) and kind != 'var':
            if name in BUILTINS and name != 'OCTAVE_VERSION':
                return [np.array([[5.0]])]
            if os.path.exists(os.path.join(self.path, name + '.m')):
                return [np.array([[2.0]])]
        return [np.array([[0.0]])]
    return _orig_eval_multi(self, e, ws, nargout, end_ctx)


Interp.eval_multi = _eval_multi_exist


# =============================================================================== python <-> M
def to_m(v):
    """python / numpy -> interpreter value."""
    if isinstance(v, dict):
        return MStruct({k: to_m(x) for k, x in v.items()})
    if isinstance(v, (list, tuple)) and v and all(isinstance(x, str) for x in v):
        return MCell(v)
    if isinstance(v, str):
        return v
    if v is None:
        return np.zeros((0, 0))
    if isinstance(v, (bool, np.bool_)):
        return np.array([[bool(v)]])
    a = np.asarray(v)
    if a.dtype.kind in 'iu':
        a = a.astype(np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1) if a.size != 1 else a.reshape(1, 1)
    return arr(a)


def from_m(v):
    if isinstance(v, MStruct):
        return {k: from_m(x) for k, x in v.items()}
    if isinstance(v, MCell):
        return [from_m(x) for x in v]
    return v
