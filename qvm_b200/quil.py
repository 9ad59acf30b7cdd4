"""A minimal Quil reader for the hot-path corpus (host side).

The reference parses Quil with cl-quil (`quil:parse-quil`, third-party, not in
the reference tree).  This reader covers what the bench corpus and the gate
tests use: gate applications with numeric parameters and the DAGGER /
CONTROLLED / FORKED modifiers, DEFGATE (matrix, parametric matrix, AS
PERMUTATION), DECLARE, MEASURE, RESET, HALT/NOP/WAIT/PRAGMA.  Classical control
flow and classical arithmetic are outside the hot path.
"""
from __future__ import annotations

import cmath
import math
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import gates as G


# ----------------------------------------------------------------- expressions
_TOKEN = re.compile(r"\s*(?:(\d+\.?\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?)|(%[A-Za-z_][\w-]*)|([A-Za-z_]\w*)|(.))")
_FUNCS = {"sin": cmath.sin, "cos": cmath.cos, "sqrt": cmath.sqrt, "exp": cmath.exp,
          "cis": lambda x: cmath.exp(1j * x)}


class _Expr:
    """Tiny recursive-descent evaluator: + - * / ^, unary minus, functions, pi, i, %params."""

    def __init__(self, text: str, env: Dict[str, complex]):
        self.toks = []
        pos = 0
        text = text.strip()
        while pos < len(text):
            m = _TOKEN.match(text, pos)
            if not m:
                break
            pos = m.end()
            num, par, name, op = m.groups()
            if num is not None:
                self.toks.append(("num", float(num)))
            elif par is not None:
                self.toks.append(("par", par[1:]))
            elif name is not None:
                self.toks.append(("name", name))
            elif op.strip():
                self.toks.append(("op", op))
        self.i = 0
        self.env = env

    def _peek(self):
        return self.toks[self.i] if self.i < len(self.toks) else (None, None)

    def _next(self):
        t = self._peek()
        self.i += 1
        return t

    def parse(self) -> complex:
        v = self._sum()
        if self.i != len(self.toks):
            raise ValueError(f"trailing tokens in expression: {self.toks[self.i:]}")
        return v

    def _sum(self):
        v = self._prod()
        while self._peek() in (("op", "+"), ("op", "-")):
            op = self._next()[1]
            r = self._prod()
            v = v + r if op == "+" else v - r
        return v

    def _prod(self):
        v = self._unary()
        while self._peek() in (("op", "*"), ("op", "/")):
            op = self._next()[1]
            r = self._unary()
            v = v * r if op == "*" else v / r
        return v

    def _unary(self):
        if self._peek() == ("op", "-"):
            self._next()
            return -self._unary()
        if self._peek() == ("op", "+"):
            self._next()
            return self._unary()
        return self._power()

    def _power(self):
        b = self._atom()
        if self._peek() == ("op", "^"):
            self._next()
            e = self._unary()
            return b ** e
        return b

    def _atom(self):
        kind, val = self._next()
        if kind == "num":
            # imaginary literal such as 0.5i
            if self._peek() == ("name", "i"):
                self._next()
                return complex(0.0, val)
            return complex(val, 0.0)
        if kind == "par":
            if val not in self.env:
                raise ValueError(f"unbound parameter %{val}")
            return complex(self.env[val])
        if kind == "name":
            if val == "pi":
                return complex(math.pi, 0.0)
            if val == "i":
                return 1j
            if val in _FUNCS:
                if self._next() != ("op", "("):
                    raise ValueError(f"expected ( after {val}")
                a = self._sum()
                if self._next() != ("op", ")"):
                    raise ValueError("expected )")
                return _FUNCS[val](a)
            raise ValueError(f"unknown name {val!r} in expression")
        if (kind, val) == ("op", "("):
            a = self._sum()
            if self._next() != ("op", ")"):
                raise ValueError("expected )")
            return a
        raise ValueError(f"unexpected token {val!r}")


def evaluate(text: str, env: Optional[Dict[str, complex]] = None) -> complex:
    return _Expr(text, env or {}).parse()


def _split_top(text: str, sep: str = ",") -> List[str]:
    out, depth, cur = [], 0, []
    for ch in text:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == sep and depth == 0:
            out.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    if cur or out:
        out.append("".join(cur))
    return [s.strip() for s in out if s.strip() != ""]


# ----------------------------------------------------------------- program objects
@dataclass
class GateDef:
    name: str
    params: List[str]
    rows: List[List[str]] = field(default_factory=list)   # matrix entries as expression text
    permutation: Optional[List[int]] = None

    @property
    def n_qubits(self) -> int:
        n = len(self.permutation) if self.permutation is not None else len(self.rows)
        return n.bit_length() - 1

    def matrix(self, values: Sequence[float] = ()) -> np.ndarray:
        if self.permutation is not None:
            return G._perm(self.permutation)
        if len(values) != len(self.params):
            raise ValueError(f"gate {self.name} takes {len(self.params)} parameter(s)")
        env = {p: complex(v) for p, v in zip(self.params, values)}
        return np.array([[evaluate(e, env) for e in row] for row in self.rows], dtype=np.complex128)


@dataclass
class GateApp:
    name: str
    params: Tuple[float, ...]
    qubits: Tuple[int, ...]
    modifiers: Tuple[str, ...] = ()          # outermost first, e.g. ("CONTROLLED", "DAGGER")


@dataclass
class Measure:
    qubit: int
    target: Optional[Tuple[str, int]] = None  # (register, offset) or None for measure-discard


@dataclass
class Reset:
    qubit: Optional[int] = None


@dataclass
class Declare:
    name: str
    kind: str
    length: int


@dataclass
class Halt:
    pass


Instruction = Union[GateApp, Measure, Reset, Declare, Halt]


@dataclass
class Program:
    instructions: List[Instruction] = field(default_factory=list)
    gate_defs: Dict[str, GateDef] = field(default_factory=dict)

    def qubits_needed(self) -> int:
        m = -1
        for ins in self.instructions:
            if isinstance(ins, GateApp):
                m = max(m, *ins.qubits)
            elif isinstance(ins, Measure):
                m = max(m, ins.qubit)
            elif isinstance(ins, Reset) and ins.qubit is not None:
                m = max(m, ins.qubit)
        return m + 1

    def gate_matrix(self, app: GateApp) -> np.ndarray:
        """Matrix of a gate application, modifiers applied innermost-first.

        FORKED splits the parameter list in halves (tests/modifier-tests.lisp:60-130)."""
        return self._modified(app.name, list(app.modifiers), list(app.params))

    def _base(self, name: str, params: Sequence[float]) -> np.ndarray:
        if name in self.gate_defs:
            return self.gate_defs[name].matrix(params)
        if name in G.STANDARD_GATES:
            return G.gate_matrix(name, params)
        raise KeyError(f"unknown gate {name}")

    def _modified(self, name: str, mods: List[str], params: List[float]) -> np.ndarray:
        if not mods:
            return self._base(name, params)
        head, rest = mods[0], mods[1:]
        if head == "DAGGER":
            return G.dagger(self._modified(name, rest, params))
        if head == "CONTROLLED":
            return G.controlled(self._modified(name, rest, params))
        if head == "FORKED":
            half = len(params) // 2
            return G.forked(self._modified(name, rest, params[:half]), self._modified(name, rest, params[half:]))
        raise ValueError(f"unknown modifier {head}")


_MODS = ("DAGGER", "CONTROLLED", "FORKED")
_GATE_RE = re.compile(r"^([A-Za-z_][\w-]*)\s*(\((.*)\))?\s*(.*)$")


def parse_quil(text: str) -> Program:
    prog = Program()
    lines = text.replace(";", "\n").split("\n")
    i = 0
    while i < len(lines):
        raw = lines[i]
        i += 1
        line = raw.split("#", 1)[0].rstrip()
        if not line.strip():
            continue
        s = line.strip()
        head = s.split()[0]
        if head == "DEFGATE":
            m = re.match(r"^DEFGATE\s+([A-Za-z_][\w-]*)\s*(\(([^)]*)\))?\s*(AS\s+(\w+))?\s*:\s*$", s)
            if not m:
                raise ValueError(f"cannot parse {s!r}")
            name, _, plist, _, kind = m.groups()
            params = [p.strip().lstrip("%") for p in plist.split(",")] if plist else []
            gd = GateDef(name=name, params=params)
            body = []
            while i < len(lines) and (lines[i].startswith((" ", "\t")) or not lines[i].strip()):
                b = lines[i].split("#", 1)[0].strip()
                i += 1
                if b:
                    body.append(b)
            if kind and kind.upper() == "PERMUTATION":
                gd.permutation = [int(x) for x in _split_top(" ".join(body).replace(" ", ","))]
            elif kind and kind.upper() not in ("MATRIX",):
                raise ValueError(f"unsupported DEFGATE kind {kind}")
            else:
                gd.rows = [_split_top(b) for b in body]
                if any(len(r) != len(gd.rows) for r in gd.rows):
                    raise ValueError(f"DEFGATE {name} is not square")
            prog.gate_defs[name] = gd
            continue
        if head == "DECLARE":
            m = re.match(r"^DECLARE\s+(\S+)\s+(\w+)(\[(\d+)\])?", s)
            prog.instructions.append(Declare(m.group(1), m.group(2), int(m.group(4) or 1)))
            continue
        if head == "MEASURE":
            parts = s.split()
            tgt = None
            if len(parts) > 2:
                mm = re.match(r"^([\w-]+)(\[(\d+)\])?$", parts[2])
                tgt = (mm.group(1), int(mm.group(3) or 0))
            prog.instructions.append(Measure(int(parts[1]), tgt))
            continue
        if head == "RESET":
            parts = s.split()
            prog.instructions.append(Reset(int(parts[1]) if len(parts) > 1 else None))
            continue
        if head == "HALT":
            prog.instructions.append(Halt())
            continue
        if head in ("NOP", "WAIT", "PRAGMA"):
            continue
        if head in ("INCLUDE", "DEFCIRCUIT", "JUMP", "JUMP-WHEN", "JUMP-UNLESS", "LABEL"):
            raise ValueError(f"{head} is outside the supported hot-path subset")
        mods = []
        while s.split()[0] in _MODS:
            mods.append(s.split()[0])
            s = s.split(None, 1)[1]
        m = _GATE_RE.match(s)
        if not m:
            raise ValueError(f"cannot parse {s!r}")
        name, _, plist, rest = m.groups()
        params = tuple(float(evaluate(p).real) for p in _split_top(plist)) if plist else ()
        qubits = tuple(int(q) for q in rest.split())
        prog.instructions.append(GateApp(name, params, qubits, tuple(mods)))
    return prog


def parse_quil_file(path: str) -> Program:
    with open(path) as f:
        return parse_quil(f.read())
